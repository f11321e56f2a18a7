// debugging aid: LD_PRELOAD to print a native backtrace (module + offset) on SIGSEGV
#define _GNU_SOURCE
#include <execinfo.h>
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
static void on_segv(int sig, siginfo_t* si, void* uc)
{
    void* frames[64];
    int n = backtrace(frames, 64);
    char msg[128];
    int len = snprintf(msg, sizeof msg, "\n== SIGSEGV at address %p, native backtrace:\n", si->si_addr);
    int fd = open("gpurun_out/segv_bt.txt", O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) fd = 2;
    write(fd, msg, len);
    backtrace_symbols_fd(frames, n, fd);
    FILE* f = fopen("/proc/self/maps", "r");
    if (f) {
        char line[512];
        while (fgets(line, sizeof line, f))
            if (strstr(line, "librin") && strstr(line, "r-xp")) write(fd, line, strlen(line));
        fclose(f);
    }
    _exit(139);
}
__attribute__((constructor)) static void install(void)
{
    static char stack[1 << 16];
    stack_t ss = {.ss_sp = stack, .ss_size = sizeof stack, .ss_flags = 0};
    sigaltstack(&ss, 0);
    struct sigaction sa;
    memset(&sa, 0, sizeof sa);
    sa.sa_sigaction = on_segv;
    sa.sa_flags = SA_SIGINFO | SA_ONSTACK;
    sigaction(SIGSEGV, &sa, 0);
}
