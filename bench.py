#!/usr/bin/env python
"""Benchmark of the per-tetrahedron hot path (BASELINE.json metric: arranged tets/sec, end to end
to mesh).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2|C3|C4]

A step = one pass of the hot path (function evaluation -> signs -> filter -> per-tet arrangement
-> mesh extraction + xyz) over one synthetic tet5 grid.  Workload: implicit
arrangement, 8 random spheres + planes (seed 1), grid resolution R = round(128 * N^(1/3)) so that
every GPU owns ~10.5 M tets (N=1: BASELINE C2, 128^3; N=8: BASELINE C5, 256^3), x-slab sharded.
Inputs (grid + function description) are resident in HBM for `value`; `e2e` goes through the
legacy host-array entry point (rin_run_host: pts, size_t tets, row-major funcVals from pinned
host memory, mesh arrays copied back), host<->device copies inside the timed region.
"""
import argparse
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


def grid_resolution(n_gpus):
    return int(round(128 * n_gpus ** (1.0 / 3.0)))


def slab_range(R, rank, world):
    import sharding
    return sharding.slab_range(R, rank, world)


def cpu_reference_run(config, seconds_budget=20.0):
    """Times the reference's CPU implementation of the path on a bounded sample of the workload:
    the first x-slabs of the SAME grid (same vertices, same functions), single-threaded like the
    reference.  Uses oracle/_ref (reference sources compiled in place) when present, else the port."""
    from helpers import make_funcs, orc_eval, orc_grid, orc_run, ref_lib, ref_run, synthetic_functions
    R = 128
    funcs = make_funcs(synthetic_functions(config))
    slabs = 128  # the whole 128^3 grid: 10.5 M tets, ~2.5 s of single-core CPU work per step
    N = R + 1
    pts, tets = orc_grid(R)
    n_t = slabs * 5 * R * R
    n_v = (slabs + 1) * N * N
    pts_s, tets_s = pts[:n_v].copy(), tets[:n_t].copy()
    del pts, tets
    t0 = time.perf_counter()
    vals = orc_eval(funcs, pts_s)  # load_functions restatement (stage 1 of the metric)
    t_eval = time.perf_counter() - t0
    kind = "port"
    if ref_lib() is not None:
        kind = "reference"
        b = ref_run("ia", pts_s, tets_s, vals)
        lab = dict(zip(b.timing_labels, b["timings"].tolist()))
        hot = sum(lab.get(k, 0.0) for k in ("func signs", "filter", "simp_arr(other)", "simp_arr(1 func)",
                                            "simp_arr(2 func)", "simp_arr(>=3 func)", "extract mesh",
                                            "compute xyz"))
    else:
        b = orc_run("ia", pts_s, tets_s, vals)
        hot = float(np.sum(b["timings"]))
    total = t_eval + hot
    return {"value": n_t / total, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "first %d of %d x-slabs of the 128^3 grid (%d tets, %d functions); stages: function "
                      "evaluation + func signs + filter + simp_arr + extract mesh + compute xyz; "
                      "%.2f s CPU" % (slabs, R, n_t, len(funcs), total),
            "seconds": total, "host_cores_total": os.cpu_count()}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    last = None
    for it in range(args.warmup + args.steps):
        last = cpu_reference_run(args.config)
        if it >= args.warmup:
            vals.append(last["seconds"])
    n_t = 128 * 5 * 128 * 128
    v = n_t / (sum(vals) / len(vals))
    last["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(vals) / len(vals),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "implicit arrangement, generated tet5 grid 128^3 (10485760 tets), 8 synthetic functions "
                                   "of BASELINE config C2 (SURVEY 8(d) generator), lookup tables on",
                       "baseline_config": "C2",
                       "sample": last["sample"] + ("" if args.gpus == 1 else
                                                   "; bounded sample of the %d-GPU weak-scaling workload (grid %d^3): its "
                                                   "single-GPU instance" % (args.gpus, grid_resolution(args.gpus)))},
            "cpu_baseline": last,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="C2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--resolution", type=int, default=0, help="override the grid resolution (tests)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import rin_b200 as rin
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

    R = args.resolution or grid_resolution(world)
    funcs = make_funcs(synthetic_functions(args.config))
    F = len(funcs)
    ctx = rin.Context(local)
    ctx.generate_grid(R)
    ctx.set_functions(funcs)
    t_first, t_count = slab_range(R, rank, world)
    if world > 1:
        ctx.set_tet_range(t_first, t_count)
    T_total = 5 * R ** 3
    mode = rin.MODE_MI if args.config == "C3" else rin.MODE_IA  # C3 is the material-interface configuration
    flags = rin.FLAG_LOOKUP | rin.FLAG_SECONDARY

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    import sharding
    gather = sharding.torch_gather(dist, torch.device("cuda", local)) if dist is not None else None
    xstate = sharding.ExchangeState()
    if dist is not None:
        # NCCL communicator of the engine itself (device-side exchange): rank 0's id to everyone
        uid = torch.from_numpy(rin.nccl_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
        dist.broadcast(uid, 0)
        ctx.nccl_init(uid.cpu().numpy(), rank, world)

    def step():
        """One pass: local hot path on this rank's slab, then (N > 1) the slab-boundary key
        exchange over NCCL and the rewrite to global vertex ids."""
        c = ctx.run(mode, flags)
        if dist is not None:
            ctx.exchange_nccl()
        return c

    # ---- value: inputs resident in HBM ---------------------------------------------------------
    ctx.set_stage_timing(False)  # the timed region records two events per pass, not a dozen
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    dev_ms, eval_ms, filt_ms = [], [], []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cnt = step()
        dev_ms.append(ctx.kernel_times()["total_ms"])
    barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches_per_step = ctx.launch_count()
    # separate profiled passes (per-stage events on) for the stage table and the kernel brackets
    ctx.set_stage_timing(True)
    stage_acc = None
    for _ in range(5):
        step()
        kt = ctx.kernel_times()
        eval_ms.append(kt["eval_ms"])
        filt_ms.append(kt["filter_ms"])
        st = ctx.stage_times()
        stage_acc = st if stage_acc is None else {k: stage_acc[k] + v for k, v in st.items()}
    stage = {k: v / 5 for k, v in stage_acc.items()}
    ctx.set_stage_timing(False)
    wall_t = torch.tensor([wall], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(wall_t, op=dist.ReduceOp.MAX)
    wall = float(wall_t.item())
    ms_per_step = 1e3 * wall / args.steps
    value = T_total / (wall / args.steps)

    # ---- N > 1: the device-side NCCL exchange against the host-driven protocol of sharding.py -----
    exchange_verified = None
    if dist is not None:
        ctx.run(mode, flags)
        info_a = ctx.exchange_nccl()
        mesh_a = ctx.download_mesh()
        c_b = ctx.run(mode, flags)
        info_b = sharding.exchange(ctx, rank, world, gather, c_b.num_faces, xstate)
        mesh_b = ctx.download_mesh()
        same = all(np.array_equal(mesh_a[k], mesh_b[k]) for k in mesh_a) and \
            info_a["vert_offset"] == info_b["vert_offset"] and info_a["n_verts_total"] == info_b["n_verts_total"]
        ok = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        exchange_verified = bool(ok.item())
        n_verts_total = info_a["n_verts_total"]

    # ---- roofline of the dominant streaming kernel (filter: reads every tet's index record) -----
    peak, peak_src = measured_peak()
    V_rank = (R + 1) ** 3 if world == 1 else ((R * (rank + 1) // world - R * rank // world) + 1) * (R + 1) ** 2
    # algorithmic bytes per launch: 16 B index record + 4 B result per tet, and each
    # function value once (8F per vertex); this implementation reads 8 B of sign masks per vertex
    # instead of the values, so the bytes it can possibly move are the smaller figure below.
    alg_survey = 20.0 * t_count + 8.0 * F * V_rank
    mask_bytes = 4.0 if F <= 16 else 8.0 * ((F + 31) // 32)  # packed P|N word when F <= 16
    alg_masks = 16.0 * t_count + mask_bytes * V_rank + 8.0 * cnt.num_intersecting_tet
    filt = float(np.mean(filt_ms))
    evl = float(np.mean(eval_ms))
    achieved = min(alg_survey, alg_masks) / (filt * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram bytes per launch from the committed ncu capture
    if os.path.exists(tp) and world == 1 and args.config == "C2" and not args.resolution:
        with open(tp) as f:
            traffic = json.load(f).get("filter_tiles_kernel", {}).get("dram_bytes")
    roofline = {"bound": "hbm", "kernel": "filter_tiles_kernel", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": min(alg_survey, alg_masks),
                "algorithmic_bytes_survey_formula": alg_survey, "kernel_ms": filt,
                "eval_kernel": {"ms": evl, "bytes": (24.0 + 8.0 * F + 8.0) * V_rank,
                                "achieved": (24.0 + 8.0 * F + 8.0) * V_rank / (evl * 1e-3) / 1e9,
                                "frac": (24.0 + 8.0 * F + 8.0) * V_rank / (evl * 1e-3) / 1e9 / peak},
                "eval_plus_filter": {"ms": evl + filt,
                                     "bytes_survey": (20.0 + (24.0 + 16.0 * F) / 5.0) * t_count,
                                     "frac_survey": (20.0 + (24.0 + 16.0 * F) / 5.0) * t_count /
                                     ((evl + filt) * 1e-3) / 1e9 / peak}}

    # ---- e2e: legacy host-array entry point, pinned host buffers, copies inside the timed region --
    e2e = None
    if world == 1 and not args.no_e2e and mode == rin.MODE_IA:
        V = (R + 1) ** 3
        pts_h = torch.empty((V, 3), dtype=torch.float64, pin_memory=True).numpy()
        tets_h = torch.empty((T_total, 4), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
        vals_h = torch.empty((V, F), dtype=torch.float64, pin_memory=True).numpy()
        gp, gt = ctx.download_grid(V, T_total)
        pts_h[:] = gp
        tets_h[:] = gt
        vals_h[:] = ctx.download_values()
        del gp, gt
        n = ctx.counts()
        cap = lambda x: int(x * 1.05) + 16
        out = {
            "vert_tet": torch.empty(cap(n.num_verts), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32),
            "vert_local": torch.empty(cap(n.num_verts), dtype=torch.uint8, pin_memory=True).numpy(),
            "vert_simplex_size": torch.empty(cap(n.num_verts), dtype=torch.uint8, pin_memory=True).numpy(),
            "vert_simplex": torch.empty((cap(n.num_verts), 4), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32),
            "vert_funcs": torch.empty((cap(n.num_verts), 4), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32),
            "vert_xyz": torch.empty((cap(n.num_verts), 3), dtype=torch.float64, pin_memory=True).numpy(),
            "face_offsets": torch.empty(cap(n.num_faces) + 1, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32),
            "face_verts": torch.empty(cap(n.num_face_verts), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32),
            "face_tet_offsets": torch.empty(cap(n.num_faces) + 1, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32),
            "face_tets": torch.empty((cap(n.num_face_tets), 2), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32),
            "face_funcs": torch.empty((cap(n.num_faces), 2), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32),
        }
        ctx2 = rin.Context(local)
        for _ in range(2):
            ctx2.run_host(mode, flags, pts_h, tets_h, vals_h)
            ctx2.download_mesh(out)
        torch.cuda.synchronize()
        k = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(k):
            c2 = ctx2.run_host(mode, flags, pts_h, tets_h, vals_h)
            ctx2.download_mesh(out)
        torch.cuda.synchronize()
        e_wall = (time.perf_counter() - t0) / k
        assert c2.num_verts == cnt.num_verts and c2.num_faces == cnt.num_faces
        h2d = pts_h.nbytes + tets_h.nbytes + vals_h.nbytes
        d2h = (n.num_verts * (4 + 1 + 1 + 16 + 16 + 24) + (n.num_faces + 1) * 8 + n.num_face_verts * 4 +
               n.num_face_tets * 8 + n.num_faces * 8)
        e2e = {"value": T_total / e_wall, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e_wall,
               "api": "rin_run_host (pts, size_t tets, row-major funcVals) + rin_download_mesh"}
        # same call sequence when the caller keeps the tet mesh on the device between calls (several function
        # sets on one grid): only the V x F values go up per step.  Reported beside the headline, not as it.
        ctx2.set_mesh(pts_h, tets_h)
        for _ in range(2):
            ctx2.set_values(vals_h)
            ctx2.run(mode, flags)
            ctx2.download_mesh(out)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k):
            ctx2.set_values(vals_h)
            c3 = ctx2.run(mode, flags)
            ctx2.download_mesh(out)
        torch.cuda.synchronize()
        r_wall = (time.perf_counter() - t0) / k
        assert c3.num_verts == cnt.num_verts and c3.num_faces == cnt.num_faces
        e2e["resident_mesh"] = {"value": T_total / r_wall, "unit": UNIT, "h2d_bytes_per_step": int(vals_h.nbytes),
                                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * r_wall,
                                "api": "rin_set_mesh_host once; per step rin_set_values_host + rin_run + rin_download_mesh"}
        ctx2.close()

    # ---- CPU baseline on the host cores of this box (rank 0, N=1 only) ---------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline and args.config == "C2":
        try:
            cpu = cpu_reference_run(args.config)
        except Exception as ex:  # the baseline is reported, never required
            cpu = {"value": None, "error": str(ex)}

    if rank == 0:
        # kernels launched by this library per step (counted from the orchestrator in rin_capi.cu):
        # eval, filter, classify, general small, general big, count_scan, emit, hash_insert, rank_reps,
        # write_verts, remap_face_verts, write_faces
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "%s, generated tet5 grid %d^3 (%d tets), %d synthetic functions of BASELINE "
                                       "config %s (SURVEY 8(d) generator), lookup tables on" % (
                                           "material interface" if mode == rin.MODE_MI else "implicit arrangement",
                                           R, T_total, F, args.config),
                           "baseline_config": args.config if world == 1 else ("C5" if world == 8 else "C2-weak"),
                           "sharding": "x-slabs, one contiguous tet range per GPU" if world > 1 else "none",
                           "cache": "inputs (%.0f MB) + intermediates exceed the 126 MB L2" %
                                    ((16.0 * T_total + 88.0 * (R + 1) ** 3) / 1e6)},
                "device_ms_per_step": float(np.mean(dev_ms)), "stage_ms": stage,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": (launches_per_step + (13 if world > 1 else 0)) * args.steps,
                "launches_per_step": launches_per_step,
                "clocks": sampler.summary(), "counts": cnt.as_dict(),
                "exchange": None if dist is None else {
                    "what": "slab-boundary vertex keys, 2 ncclAllGather per step (device-side, rin_exchange_nccl)",
                    "verified_against_host_protocol": exchange_verified, "n_verts_total": n_verts_total}}
        print(json.dumps(line))
    if dist is not None:
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
