#!/usr/bin/env python
"""Benchmark of the per-tetrahedron hot path (BASELINE.json metric: arranged tets/sec, end to end to mesh).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2|C3|C4|C5]
                  [--scaling weak|strong]

A step = one pass of the hot path (function evaluation -> signs / highest material -> filter -> per-tet
arrangement -> mesh extraction + xyz) over one synthetic tet5 grid (SURVEY section 8(d) generators).
  N = 1: BASELINE C2 (implicit arrangement, 128^3, 8 functions) unless --config says otherwise
         (C3 material interface 128^3, C4 32 dense functions 128^3, C5 256^3 on one GPU).
  N > 1: x-slab sharding, weak scaling R = round(128 N^(1/3)) (N = 8 is BASELINE C5, 256^3);
         --scaling strong keeps 256^3 for every N.
`value`: inputs (grid + function description) resident in HBM, device-side slab exchange included for N > 1.
`e2e`: the reference-shaped entry point with HOST buffers (points, size_t tets, row-major function values, all
pinned; for N > 1 every rank uploads its slab's slice) and the mesh copied back to the host (N > 1: every rank
DMA-writes its slice of the merged mesh into one pinned POSIX shared-memory segment that rank 0 reads).
The merged mesh is checked against the committed oracle digest of the same configuration
(tests/golden/fullsize_golden.json) outside the timed region.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "robust-implicit-surface-networks_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "arranged tets/sec (end-to-end to mesh)"
UNIT = "tets/s"
FUNCTION_SET = {"C2": "C2", "C5": "C2", "C3": "C3", "C4": "C4"}
MESH_KEYS = ("vert_tet", "vert_local", "vert_simplex_size", "vert_simplex", "vert_funcs", "vert_xyz", "face_offsets",
             "face_verts", "face_tet_offsets", "face_tets", "face_funcs")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i] == "Active" for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def grid_resolution(config, n_gpus, scaling):
    if n_gpus == 1:
        return 256 if config == "C5" else 128
    return 256 if scaling == "strong" else int(round(128 * n_gpus ** (1.0 / 3.0)))


def golden_name(config, R):
    """Committed oracle digest of this (function set, grid) combination, if there is one."""
    if config in ("C2", "C5"):
        return {128: "C2", 161: "W2", 203: "W4", 256: "C5"}.get(R)
    return config if R == 128 else None


def workload_text(config, R, mi):
    F = {"C2": 8, "C5": 8, "C3": 6, "C4": 32}[config]
    return "%s, generated tet5 grid %d^3 (%d tets), %d synthetic functions of BASELINE config %s (SURVEY 8(d) " \
           "generator), lookup tables on" % ("material interface" if mi else "implicit arrangement", R, 5 * R ** 3, F,
                                            config)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def mesh_digest(mesh, mi):
    ff = mesh["face_funcs"].astype(np.int64)
    ff[ff == 0xFFFFFFFF] = -1
    d = {"face_offsets": sha(mesh["face_offsets"].astype(np.int64)),
         "face_verts": sha(mesh["face_verts"].astype(np.int64)),
         "face_tets": sha(mesh["face_tets"].astype(np.int64).ravel()),
         "face_funcs_first": sha(ff[:, 0]), "vert_xyz": sha(mesh["vert_xyz"])}
    if mi:
        d["face_funcs"] = sha(ff.ravel())
    return d


def check_digest(mesh, config, R, mi):
    name = golden_name(config, R)
    p = os.path.join(ROOT, "tests", "golden", "fullsize_golden.json")
    if name is None or not os.path.exists(p):
        return None
    with open(p) as f:
        gold = json.load(f)
    if name not in gold:
        return None
    return {"golden": "tests/golden/fullsize_golden.json:" + name, "pinned_by": gold[name]["pinned_by"],
            "equal": mesh_digest(mesh, mi) == gold[name]["digest"]}


# ---------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation of the path (oracle/_ref = its sources
# compiled in place, else the oracle port), one thread (the reference is single-threaded by construction), on a
# bounded sample of the workload: the first x-slabs of the same grid with the same functions
# ---------------------------------------------------------------------------------------------------------------
# (R, slabs in the sample).  C4: the 32 spheres sit around the origin (radius ~0.5), the first 20 slabs hold no surface at
# all (and the reference crashes on an empty mesh): half the grid is the smallest sample with the whole grid's density
# C5: 96 of the 256 slabs starting at slab 30: a window whose active-tet density (11.1 k per slab) is within 2 % of the
# whole grid's (10.9 k; per-slab counts from the calibration pass of the 8-GPU run), so tets/s is representative
CPU_SLABS = {"C2": (128, 128), "C3": (128, 128), "C4": (128, 64), "C5": (256, 96)}
CPU_FIRST_SLAB = {"C5": 30}
CPU_LABELS = {"ia": ("func signs", "filter", "simp_arr(other)", "simp_arr(1 func)", "simp_arr(2 func)",
                     "simp_arr(>=3 func)", "extract mesh", "compute xyz"),
              "mi": ("highest func", "filter", "MI(other)", "MI(2 func)", "MI(3 func)", "MI(>=4 func)", "extract mesh",
                     "compute xyz")}


_CPU_SAMPLE = {}  # config -> (pts, tets) of the sample: built once, every step of the reference arm reuses it


def cpu_reference_run(config):
    from helpers import make_funcs, orc_eval, orc_grid, orc_run, ref_lib, ref_run, synthetic_functions
    R, slabs = CPU_SLABS[config]
    mode = "mi" if config == "C3" else "ia"
    funcs = make_funcs(synthetic_functions(FUNCTION_SET[config]))
    N = R + 1
    first = CPU_FIRST_SLAB.get(config, 0)
    n_t = slabs * 5 * R * R
    n_v = (slabs + 1) * N * N
    if config not in _CPU_SAMPLE:
        pts, tets = orc_grid(R)
        v0, t0 = first * N * N, first * 5 * R * R
        # vertex ids relative to the sample's first plane
        _CPU_SAMPLE[config] = (pts[v0:v0 + n_v].copy(), (tets[t0:t0 + n_t] - tets.dtype.type(v0)).copy())
        del pts, tets
    pts_s, tets_s = _CPU_SAMPLE[config]
    t0 = time.perf_counter()
    vals = orc_eval(funcs, pts_s)  # load_functions restatement (stage 1 of the metric)
    t_eval = time.perf_counter() - t0
    kind = "port"
    if ref_lib() is not None:
        kind = "reference"
        b = ref_run(mode, pts_s, tets_s, vals)
        lab = dict(zip(b.timing_labels, b["timings"].tolist()))
        hot = sum(lab.get(k, 0.0) for k in CPU_LABELS[mode])
    else:
        b = orc_run(mode, pts_s, tets_s, vals)
        hot = float(np.sum(b["timings"]))
    total = t_eval + hot
    return {"value": n_t / total, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "x-slabs [%d, %d) of %d of the %d^3 grid (%d tets, %d functions); stages: function evaluation + "
                      "%s + filter + per-tet arrangement + extract mesh + compute xyz; %.2f s CPU" % (
                          first, first + slabs, R, R, n_t, len(funcs), "highest func" if mode == "mi" else "func signs",
                          total),
            "seconds": total, "tets": n_t, "host_cores_total": os.cpu_count()}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    config = args.config
    if args.gpus > 1 and config == "C2" and (args.scaling == "strong" or args.gpus == 8):
        config = "C5"  # the N-GPU workload of the repo arm is the 256^3 grid
    secs, last = [], None
    for it in range(args.warmup + args.steps):
        last = cpu_reference_run(config)
        if it >= args.warmup:
            secs.append(last["seconds"])
    v = last["tets"] / (sum(secs) / len(secs))
    last["value"] = v
    R_ours = grid_resolution(args.config, args.gpus, args.scaling)
    R_ref = CPU_SLABS[config][0]
    note = ""
    if R_ours != R_ref:
        note = "; the repo arm at %d GPUs runs the %d^3 instance of the same generator, this arm samples the %d^3 " \
               "instance (tets/s is size-normalised)" % (args.gpus, R_ours, R_ref)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(config, R_ref, config == "C3"), "baseline_config": config,
                       "sample": last["sample"] + note},
            "cpu_baseline": last,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def pinned(shape, dtype):
    import torch
    tdt = {np.float64: torch.float64, np.uint32: torch.int32, np.uint64: torch.int64, np.uint8: torch.uint8}[dtype]
    return torch.empty(shape, dtype=tdt, pin_memory=True).numpy().view(dtype)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="C2")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--two-call-exchange", action="store_true",
                    help="N > 1: rin_run then rin_exchange_nccl (two synchronisations) instead of rin_run_exchange")
    ap.add_argument("--resolution", type=int, default=0, help="override the grid resolution (tests)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import rin_b200 as rin
    import sharding
    from helpers import make_funcs, synthetic_functions

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None
        torch.cuda.set_device(0)

    config = args.config
    R = args.resolution or grid_resolution(config, world, args.scaling)
    funcs = make_funcs(synthetic_functions(FUNCTION_SET[config]))
    F = len(funcs)
    mi = config == "C3"
    mode = rin.MODE_MI if mi else rin.MODE_IA
    flags = rin.FLAG_LOOKUP | rin.FLAG_SECONDARY
    ctx = rin.Context(local)
    ctx.generate_grid(R)
    ctx.set_functions(funcs)
    i0, i1 = R * rank // world, R * (rank + 1) // world  # this rank's x-planes of cubes [i0, i1)
    t_first, t_count = sharding.slab_of_planes(R, i0, i1)
    if world > 1:
        ctx.set_tet_range(t_first, t_count)
    T_total = 5 * R ** 3
    N1 = R + 1

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def new_uid():
        # NCCL communicator of the engine itself (device-side exchange): rank 0's id to everyone
        uid = torch.from_numpy(rin.nccl_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
        dist.broadcast(uid, 0)
        return uid.cpu().numpy()

    slab_plan = None
    if dist is not None:
        ctx.nccl_init(new_uid(), rank, world)
        # calibration pass (outside the timed region): where the surface is decides what a slab costs; slabs are
        # re-cut so that the estimated cost is even (sharding.balanced_slab_planes), every rank from the same data
        ctx.run(mode, flags)
        hist = torch.zeros(R, dtype=torch.float64, device="cuda")
        act = ctx.download_active_tets() // np.uint32(5 * R * R)
        hist += torch.from_numpy(np.bincount(act, minlength=R).astype(np.float64)).cuda()
        dist.all_reduce(hist)
        slab_plan = sharding.balanced_slab_planes(hist.cpu().numpy(), world)
        i0, i1 = slab_plan[rank], slab_plan[rank + 1]
        t_first, t_count = sharding.slab_of_planes(R, i0, i1)
        ctx.set_tet_range(t_first, t_count)
        # feedback rounds (still outside the timed region): the measured device time of every rank's pass, spread
        # evenly over its cube planes, replaces the cost model; the slabs are re-cut until the plan stops moving
        ctx.set_stage_timing(False)
        for _ in range(3):
            for _ in range(3):
                ctx.run(mode, flags)
                ctx.exchange_nccl()
            t = 0.0
            for _ in range(3):
                ctx.run(mode, flags)
                t += ctx.kernel_times()["total_ms"] / 3
                ctx.exchange_nccl()
            dens = torch.zeros(R, dtype=torch.float64, device="cuda")
            dens[i0:i1] = t / max(1, i1 - i0)
            dist.all_reduce(dens)
            new_plan = sharding.slabs_of_equal_cost(dens.cpu().numpy(), world)
            if new_plan == slab_plan:
                break
            slab_plan = new_plan
            i0, i1 = slab_plan[rank], slab_plan[rank + 1]
            t_first, t_count = sharding.slab_of_planes(R, i0, i1)
            ctx.set_tet_range(t_first, t_count)
        ctx.set_stage_timing(True)

    split = {"run": 0.0, "exchange": 0.0, "n": 0}

    def step():
        """One pass: hot path on this rank's slab, then (N > 1) the slab-boundary exchange on the device."""
        if dist is None:
            return ctx.run(mode, flags)
        if not args.two_call_exchange:
            return ctx.run_exchange(mode, flags)  # one synchronisation: the exchange rides behind the run
        ta = time.perf_counter()
        ctx.run(mode, flags)
        tb = time.perf_counter()
        r = ctx.exchange_nccl()
        split["run"] += tb - ta
        split["exchange"] += time.perf_counter() - tb
        split["n"] += 1
        return r

    # ---- value: inputs resident in HBM ------------------------------------------------------------------------
    ctx.set_stage_timing(False)  # the timed region records two events per pass, not a dozen
    sampler = ClockSampler(local)  # nvidia-smi takes ~50 ms per query: started with the warm-up passes (same load) so
    sampler.start()                # that a timed region of a few milliseconds still gets its samples
    for _ in range(args.warmup):
        step()
    # further untimed passes for about 0.4 s (the same number on every rank) so that the clock sampler sees the load
    tw = time.perf_counter()
    step()
    extra = torch.tensor([min(2000, int(0.4 / max(time.perf_counter() - tw, 1e-5)))], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.broadcast(extra, 0)
    extra_warmup = int(extra.item())
    for _ in range(extra_warmup):
        step()
    barrier()
    dev_ms = []
    split.update(run=0.0, exchange=0.0, n=0)
    t0 = time.perf_counter()
    x_ms = []
    for _ in range(args.steps):
        info = step()
        dev_ms.append(ctx.kernel_times()["total_ms"])
        if dist is not None:
            x_ms.append(ctx.exchange_time())
    barrier()
    wall = allmax(time.perf_counter() - t0)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    cnt = ctx.counts()
    split_ms = None
    if dist is not None and split["n"]:  # host wall time of the two calls on this rank (the exchange includes waiting for the peers)
        split_ms = {"run_ms": 1e3 * split["run"] / split["n"], "exchange_ms": 1e3 * split["exchange"] / split["n"]}
        both = torch.tensor([split_ms["run_ms"], split_ms["exchange_ms"]], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(both) for _ in range(world)]
        dist.all_gather(allr, both)
        split_ms = {"per_rank_run_ms": [float(x[0]) for x in allr], "per_rank_exchange_ms": [float(x[1]) for x in allr]}
    per_rank = None
    if dist is not None:  # device times of every rank: the run's kernels and the forked exchange chain
        both = torch.tensor([float(np.mean(dev_ms)), float(np.mean(x_ms)), float(cnt.num_intersecting_tet), float(t_count)],
                            dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(both) for _ in range(world)]
        dist.all_gather(allr, both)
        per_rank = {"run_device_ms": [round(float(x[0]), 4) for x in allr],
                    "exchange_chain_device_ms": [round(float(x[1]), 4) for x in allr],
                    "active_tets": [int(x[2]) for x in allr], "tets": [int(x[3]) for x in allr]}
    launches_per_step = ctx.launch_count()
    ms_per_step = 1e3 * wall / args.steps
    value = T_total / (wall / args.steps)

    # separate profiled passes (per-stage events on) for the stage table and the kernel brackets
    ctx.set_stage_timing(True)
    eval_ms, filt_ms, stage_acc = [], [], None
    x_parts = None
    for _ in range(5):
        step()
        if dist is not None:
            p = np.array(ctx.exchange_parts())
            x_parts = p if x_parts is None else x_parts + p
        kt = ctx.kernel_times()
        eval_ms.append(kt["eval_ms"])
        filt_ms.append(kt["filter_ms"])
        st = ctx.stage_times()
        stage_acc = st if stage_acc is None else {k: stage_acc[k] + v for k, v in st.items()}
    stage = {k: v / 5 for k, v in stage_acc.items()}
    ctx.set_stage_timing(False)

    # ---- the merged mesh on rank 0's host (also the destination of the e2e legs) ------------------------------
    if dist is None:
        info = {"vert_offset": 0, "face_offset": 0, "fv_offset": 0, "ft_offset": 0, "n_verts_total": cnt.num_verts,
                "n_faces_total": cnt.num_faces, "n_fv_total": cnt.num_face_verts, "n_ft_total": cnt.num_face_tets}
    layout = sharding.merged_layout(info)
    sizes = {k: int(np.prod(sh)) * np.dtype(dt).itemsize for k, (sh, dt) in layout.items()}
    total_bytes = max(4096, sum((s + 255) & ~255 for s in sizes.values()))
    shm = None
    if dist is None:
        hostbuf = pinned((total_bytes,), np.uint8)
    else:
        from multiprocessing import shared_memory
        name = [None]
        if rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=total_bytes)
            name[0] = shm.name
        dist.broadcast_object_list(name, 0)
        if rank != 0:
            shm = shared_memory.SharedMemory(name=name[0])
        hostbuf = np.ndarray((total_bytes,), np.uint8, buffer=shm.buf)
        # page-lock the shared segment so that every rank's D2H copy is a plain DMA into "rank 0's" host memory
        rc = torch.cuda.cudart().cudaHostRegister(hostbuf.ctypes.data, total_bytes, 0)
        assert int(rc) == 0, "cudaHostRegister failed: %s" % rc
    merged, off = {}, 0
    for k, (sh, dt) in layout.items():
        merged[k] = np.ndarray(sh, dt, buffer=hostbuf, offset=off)
        off += (sizes[k] + 255) & ~255
    views = sharding.slice_views(merged, info, cnt)
    ctx.download_mesh(views)
    barrier()
    verified = check_digest(merged, config, R, mi) if rank == 0 else None

    # ---- e2e: host buffers in, merged mesh on the host out ----------------------------------------------------
    e2e = None
    d2h = sum(views[k].nbytes for k in MESH_KEYS)
    if not args.no_e2e:
        # this rank's slice of the reference-shaped inputs: the vertex planes its slab touches
        v_first, v_count = i0 * N1 * N1, (i1 - i0 + 1) * N1 * N1
        gp, gt = ctx.download_grid(N1 ** 3, T_total)
        pts_h = pinned((v_count, 3), np.float64)
        tets_h = pinned((t_count, 4), np.uint64)
        vals_h = pinned((v_count, F), np.float64)
        pts_h[:] = gp[v_first:v_first + v_count]
        tets_h[:] = gt[t_first:t_first + t_count]
        del gp, gt
        vals_h[:] = ctx.download_values()[v_first:v_first + v_count]
        h2d = pts_h.nbytes + tets_h.nbytes + vals_h.nbytes
        ctx2 = rin.Context(local)
        ctx2.set_stage_timing(False)
        if dist is not None:
            ctx2.nccl_init(new_uid(), rank, world)

        def e2e_step(upload_mesh=True):
            if dist is None:
                if upload_mesh:
                    ctx2.run_host(mode, flags, pts_h, tets_h, vals_h)
                else:
                    ctx2.set_values(vals_h)
                    ctx2.run(mode, flags)
            else:
                if upload_mesh:
                    ctx2.set_mesh_range(N1 ** 3, T_total, pts_h, v_first, tets_h, t_first)
                ctx2.set_values_range(vals_h, v_first)
                ctx2.run_exchange(mode, flags)
            ctx2.download_mesh(views)

        def timed(fn, k):
            for _ in range(2):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(k):
                fn()
            barrier()
            return allmax(time.perf_counter() - t0) / k

        k = max(3, min(args.steps, 10))
        if dist is None:
            ctx2.set_mesh(pts_h, tets_h)  # shapes for the "resident" variant exist before its first call
        e_wall = timed(e2e_step, k)
        c2 = ctx2.counts()
        assert c2.num_verts == cnt.num_verts and c2.num_faces == cnt.num_faces
        tot = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tot)
        e2e = {"value": T_total / e_wall, "unit": UNIT, "h2d_bytes_per_step": int(tot[0].item()),
               "d2h_bytes_per_step": int(tot[1].item()), "ms_per_step": 1e3 * e_wall,
               "api": "rin_run_host (pts, size_t tets, row-major funcVals) + rin_download_mesh" if dist is None else
                      "per rank: rin_set_mesh_host_range + rin_set_values_host_range (its slab's slice) + rin_run + "
                      "rin_exchange_nccl + rin_download_mesh into its slice of the merged mesh in a pinned "
                      "shared-memory segment read by rank 0"}
        barrier()
        if rank == 0:
            e2e["verified"] = check_digest(merged, config, R, mi)
        # the caller keeps the tet mesh on the device between calls (several function sets on one grid): only the
        # V x F values go up per step.  Reported beside the headline, not as it.
        r_wall = timed(lambda: e2e_step(False), k)
        e2e["resident_mesh"] = {"value": T_total / r_wall, "unit": UNIT, "ms_per_step": 1e3 * r_wall,
                                "h2d_bytes_per_step": int(vals_h.nbytes) * world,
                                "d2h_bytes_per_step": e2e["d2h_bytes_per_step"],
                                "api": "mesh uploaded once; per step the V x F values up, run, mesh down"}
        # the application-level call (impl_arrangement config.json): grid resolution + function parameters in,
        # mesh on the host out - the same inputs the reference arm starts from
        if dist is None:
            def app_step():
                ctx2.generate_grid(R)
                ctx2.set_functions(funcs)
                ctx2.run(mode, flags)
                ctx2.download_mesh(views)
            a_wall = timed(app_step, k)
            e2e["app_level"] = {"value": T_total / a_wall, "unit": UNIT, "ms_per_step": 1e3 * a_wall,
                                "h2d_bytes_per_step": int(funcs.nbytes), "d2h_bytes_per_step": int(d2h),
                                "api": "rin_generate_grid + rin_set_functions + rin_run + rin_download_mesh "
                                       "(what the app does for a config with gridResolution + funcFile)"}
        ctx2.close()

    # ---- roofline of the dominant HBM kernel -------------------------------------------------------------------
    # IA / MI on a generated grid: the evaluation kernel writes the SoA values (8F bytes per vertex) and the sign
    # masks; nothing is read (coordinates come from three axis tables).  SURVEY 8(d) K1 with an implicit grid:
    # V * 8F (+ the masks this implementation adds).  The filter no longer streams the index records.
    peak, peak_src = measured_peak()
    V_rank = N1 ** 3 if world == 1 else (i1 - i0 + 1) * N1 * N1
    mask_bytes = 4.0 if (F <= 16 and not mi) else 8.0 * ((F + 31) // 32)
    evl, filt = float(np.mean(eval_ms)), float(np.mean(filt_ms))
    alg_eval = (8.0 * F + mask_bytes) * V_rank
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")  # dram bytes per launch from the committed ncu capture
    if os.path.exists(tp) and world == 1 and not args.resolution:
        with open(tp) as f:
            t = json.load(f).get(config, {}).get("eval_mi_kernel" if (mi and F <= 8) else "eval_kernel")
        if t:
            traffic, traffic_src = t["dram_bytes"], t["source"]
    # achieved = ALGORITHMIC bytes / kernel time.  The ncu DRAM count of this write-only kernel is below the algorithmic
    # bytes only because ncu flushes L2 before the launch and part of the written lines is still dirty in the 126 MB L2
    # when the kernel ends (they reach HBM during the next kernel); in the timed loop the L2 is full of the previous
    # pass's lines, every written byte displaces one.  Both fractions are reported.
    achieved = alg_eval / (evl * 1e-3) / 1e9
    frac_ncu = None if traffic is None else min(alg_eval, traffic) / (evl * 1e-3) / 1e9 / peak
    survey_bytes = (20.0 + (24.0 + 16.0 * F) / 5.0) * t_count
    dev = float(np.mean(dev_ms))
    roofline = {"bound": "hbm", "kernel": ("eval_mi_kernel (evaluation fused with the highest-material loop)" if F <= 8
                                           else "eval_kernel+highest_material_kernel") if mi else "eval_kernel",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_eval,
                "kernel_ms": evl,
                "frac_with_ncu_dram_bytes": frac_ncu,
                "bytes_model": "V * (8F values + %d B sign masks) written, nothing read (generated grid); achieved = "
                               "algorithmic bytes / kernel time (CUDA events in this run); frac_with_ncu_dram_bytes uses "
                               "min(algorithmic, ncu DRAM bytes of one cold-L2 launch), see the comment in bench.py" % int(mask_bytes),
                "filter_kernel": {"name": "filter_mi_grid_kernel" if mi else "filter_classify_kernel", "ms": filt,
                                  "note": "cube-structured: reads the per-vertex masks once per cube corner through "
                                          "L1/L2, never the 16 B/tet index stream; instruction-bound, not HBM-bound"},
                "eval_plus_filter": {"ms": evl + filt, "bytes_survey": survey_bytes,
                                     "frac_survey": survey_bytes / ((evl + filt) * 1e-3) / 1e9 / peak},
                "whole_step": {"ms": dev, "frac_survey": survey_bytes / (dev * 1e-3) / 1e9 / peak,
                               "note": "SURVEY 8(d) eval+filter bytes (reference-shaped formulation) over the device "
                                       "time of the whole pass"}}

    # ---- CPU baseline on the host cores of this box (rank 0, N = 1 only) --------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline and not args.resolution:
        try:
            cpu = cpu_reference_run(config)
        except Exception as ex:  # the baseline is reported, never required
            cpu = {"value": None, "error": str(ex)}

    if rank == 0:
        # kernels of the neighbour exchange: counted by the library when it rides behind the run (rin_run_exchange)
        x_launches = 10 if (world > 1 and args.two_call_exchange) else 0
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "warmup_extra_for_clock_sampling": extra_warmup + 1,
                "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_text(config, R, mi),
                           "baseline_config": config if world == 1 else ("C5" if R == 256 else "C2-weak"),
                           "sharding": ("x-slabs, one contiguous tet range per GPU, cut at the cube planes %s so that the "
                                        "cost is even (calibration pass + up to three feedback rounds on the measured per-rank device time, all outside "
                                        "the timed region)" % slab_plan)
                           if world > 1 else "none",
                           "cache": "outputs of every stage (values %.0f MB, candidates, mesh) exceed the 126 MB L2 "
                                    "or are produced by the previous kernel" % (8.0 * F * N1 ** 3 / 1e6)},
                "device_ms_per_step": dev, "stage_ms": stage,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": (launches_per_step + x_launches) * args.steps,
                "launches_per_step": launches_per_step + x_launches,
                "clocks": sampler.summary(), "counts": cnt.as_dict(), "verified": verified,
                "exchange": None if dist is None else {
                    "what": "slab-boundary vertex keys: ncclSend/ncclRecv with the neighbour ranks + one ncclAllGather of "
                            "the counts and own indices; " + ("rin_run then rin_exchange_nccl (two synchronisations)"
                                                               if args.two_call_exchange else
                                                               "enqueued behind the run's kernels, counts read from "
                                                               "device memory, ONE host synchronisation per step "
                                                               "(rin_run_exchange)"),
                    "n_verts_total": info["n_verts_total"], "n_faces_total": info["n_faces_total"],
                    "host_wall": split_ms, "per_rank": per_rank,
                    "chain_kernels_ms_rank0": None if x_parts is None else dict(zip(
                        ("send_keys", "recv_insert(waits for r-1)", "mark_scan_publish",
                         "finish(waits for all records; ids + owned vertices)"),
                        [round(float(x) / 5, 4) for x in x_parts[:4]]))}}
        print(json.dumps(line))
    if dist is not None:
        barrier()
        torch.cuda.cudart().cudaHostUnregister(hostbuf.ctypes.data)
        del merged, views, hostbuf
        shm.close()
        if rank == 0:
            shm.unlink()
        dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        # a rank that fails must not leave its peers waiting inside a collective
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
