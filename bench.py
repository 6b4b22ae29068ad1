#!/usr/bin/env python
"""Headline benchmark of the sculpt-stroke hot path (BASELINE.json: vertex-dabs/sec and ms/dab,
% of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--grid 4096]

N = 1 workload (config.workload): C3 -- draw brush + normal recompute + BB refit on the 16,777,216-
vertex height-field grid, radius sweep 1-50 % of the bounding-box diagonal, 32 dabs per radius.
One *step* = a device-to-device rollback of the mesh to its rest state + one pass of the whole 224-dab
stroke script (so every step does identical work).  `value` = vertex-dabs / second with the
mesh resident in HBM (CUDA events around the K strokes); `e2e` = the same strokes driven through
the reference-named host API with host buffers: per dab the descriptor goes host->device, at stroke
end positions, normals, node boxes and flags come back device->host, all inside the timed region.

Prints ONE JSON line on rank 0.  --impl reference times the CPU oracle (OpenMP, all host threads)
on a bounded sample of the same workload.
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
sys.path.insert(0, ROOT)

METRIC = "vertex_dabs_per_sec"
UNIT = "vertex-dabs/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(args):
    from dune_sculpt_b200 import meshgen, stroke
    t0 = time.time()
    if args.config == "c5":
        # C5: multires cube, `c5_base`^2 base quads per side, level `c5_level`; draw stroke, r = 8 % of the diagonal
        from dune_sculpt_b200 import capi
        mesh = meshgen.multires_cube_n(args.c5_base, args.c5_level)
        diag = mesh.bbox_diag()
        rng = np.random.default_rng(5)
        bs = stroke._strength(capi.TOOL_DRAW, 0.5)
        dabs = []
        # C5: a smooth stroke, then a draw stroke, same radius
        bsm = stroke._strength(capi.TOOL_SMOOTH, 0.75)
        for i in range(args.c5_smooth_dabs):
            p = rng.normal(size=3)
            p /= np.linalg.norm(p)
            dabs.append(capi.make_dab(capi.TOOL_SMOOTH, p.astype(np.float32), diag * 0.08, bstrength=bsm, view_normal=tuple(p),
                                      flags=capi.DAB_FIRST_STEP if i == 0 else 0))
        for i in range(args.c5_dabs):
            p = rng.normal(size=3)
            p /= np.linalg.norm(p)
            dabs.append(capi.make_dab(capi.TOOL_DRAW, p.astype(np.float32), diag * 0.08, bstrength=bs, view_normal=tuple(p),
                                      flags=capi.DAB_FIRST_STEP if i == 0 else 0))
        log("[bench] multires cube %d^2 x 6 base quads, level %d: %d grids of %d^2 = %d elements, diag=%.4f, %d dabs/stroke (%.1fs)" %
            (args.c5_base, args.c5_level, mesh.totgrid, mesh.grid_size, mesh.totelem, diag, len(dabs), time.time() - t0))
        return mesh, diag, dabs
    mesh = meshgen.grid(args.grid)
    diag = mesh.bbox_diag()
    dabs = stroke.c3_radius_sweep(diag, dabs_per_radius=args.dabs_per_radius)
    log("[bench] mesh grid %d^2: V=%d polys=%d diag=%.4f, %d dabs/stroke (%.1fs)" %
        (args.grid, mesh.totvert, mesh.totpoly, diag, len(dabs), time.time() - t0))
    return mesh, diag, dabs


def run_reference(args, rank):
    """CPU arm: the oracle (a port -- the reference itself does not compile here) with
    OpenMP over hit nodes, the decomposition the reference uses (lib/intern/task_range.cc:89-127)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from dune_sculpt_b200 import build as b
    b.build_oracle()
    from oracle_py import GridOracle, Oracle
    mesh, diag, dabs = build_workload(args)
    cores = os.cpu_count() or 1
    t0 = time.time()
    orc = GridOracle(mesh, threads=cores) if args.config == "c5" else Oracle(mesh, threads=cores)
    log("[bench] oracle PBVH build %.1fs, %d nodes, %d threads" % (time.time() - t0, orc.totnode, cores))
    # bounded sample: `sample_per_radius` dabs of every radius of the sweep per step
    per = args.dabs_per_radius
    nrad = len(dabs) // per
    if args.config == "c5":
        sample = dabs[:args.c5_cpu_dabs]
        nrad = 1
    else:
        sample = [dabs[r * per + k] for r in range(nrad) for k in range(args.cpu_sample_per_radius)]
    orc.stroke_begin()
    for _ in range(args.warmup_ref):
        for d in sample[:2]:
            orc.dab(d)
    vd0 = orc.vertex_dabs()
    t0 = time.perf_counter()
    for _ in range(args.steps_ref):
        for d in sample:
            orc.dab(d)
    dt = time.perf_counter() - t0
    vd = orc.vertex_dabs() - vd0
    orc.stroke_end()
    value = vd / dt
    sample_desc = "%d dabs per radius x %d radii of the C3 sweep per step (%d dabs), %d steps" % (
        args.cpu_sample_per_radius, nrad, len(sample), args.steps_ref)
    if args.config == "c5":
        sample_desc = "the first %d dabs of the C5 stroke per step, %d steps" % (len(sample), args.steps_ref)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps_ref,
        "warmup": args.warmup_ref, "ms_per_step": 1e3 * dt / args.steps_ref, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, mesh, len(dabs)),
        "ms_per_dab": 1e3 * dt / (args.steps_ref * len(sample)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)


def workload_config(args, mesh, ndabs):
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.config == "c5":
        return {"workload": "C5 multires grids: cube %d^2 x 6 base quads, level %d (%d grids of %d^2 = %d elements), %d smooth dabs (3 iterations) "
                            "then %d draw dabs, each + stitch + CCG normals + BB, r = 8%% bbox diag, %d dabs/stroke" %
                            (args.c5_base, args.c5_level, mesh.totgrid, mesh.grid_size, mesh.totelem, args.c5_smooth_dabs,
                             args.c5_dabs, ndabs),
                "parallelism": "single GPU" if world == 1 else
                               "grids PBVH partitioned spatially over %d GPUs (same mesh: strong scaling); per dab one NCCL all-reduce "
                               "(area sums + hit mask), the rim positions after the brush / every smoothing iteration and the rim "
                               "normals after the CCG normal pass exchanged with the neighbouring ranks" % world,
                "verts": mesh.totelem, "dabs_per_step": ndabs,
                "brush": "smooth (alpha 0.75) then draw (alpha 0.5), SMOOTH falloff, area-normal direction",
                "l2": "inputs larger than L2 (resident element arrays > 2 GB)" if mesh.totelem > 8000000 else "small mesh: L2 resident",
                "step": "device-to-device rollback to the rest state + one %d-dab stroke" % ndabs}
    return {"workload": "C3 draw+normals+BB radius sweep 1-50%% bbox diag, grid %d^2 (V=%d), %d dabs/stroke" %
                        (args.grid, mesh.totvert, ndabs) +
                        ("" if world == 1 else "; PBVH partitioned spatially over %d GPUs (same mesh: strong scaling), per dab one "
                         "NCCL all-reduce (area sums + hit mask) and one one-ring halo exchange" % world),
            "parallelism": "single GPU" if world == 1 else "pbvh-partition x%d" % world,
            "verts": mesh.totvert, "dabs_per_step": ndabs, "brush": "draw, SMOOTH falloff, area-normal direction",
            "l2": "inputs larger than L2 (resident mesh arrays > 2 GB; every stroke sweeps all of them)",
            "step": "device-to-device rollback to the rest state + one %d-dab stroke" % ndabs}


def analysis_pass(ses, dabs, na, grids=False):
    """One untimed stroke with a sync after every dab: per-dab U/A/T/M and per-stage device times,
    for the roofline object."""
    uniq, face, totprim = na["uniq_verts"].astype(np.int64), na["face_verts"].astype(np.int64), na["totprim"].astype(np.int64)
    n = ses.totnode
    parent = np.full(n, -1, dtype=np.int64)
    inner = np.nonzero((na["flag"] & 1) == 0)[0]
    parent[na["children_offset"][inner]] = inner
    parent[na["children_offset"][inner] + 1] = inner
    ses.rollback()
    ses.stage_timing(True)
    ses.stroke_begin()
    moved_prev = 0
    touched = np.zeros(n, dtype=bool)
    tot = {"U": 0, "A": 0, "T": 0, "M": 0, "first_A": 0, "hits": 0}
    stage_bytes = {"gather": 0, "area_normal": 0, "brush": 0, "smooth": 0, "normals_bb": 0, "bb_refit": 0}
    nleaf = int((na["flag"] & 1).sum())
    for d in dabs:
        ses.dab(d)
        h = ses.hits()
        st = ses.stats()
        M = st["moved_verts"] - moved_prev
        moved_prev = st["moved_verts"]
        U = int(uniq[h].sum())
        A = U + int(face[h].sum())
        T = int(totprim[h].sum())
        new = h[~touched[h]]
        touched[h] = True
        first_A = int(uniq[new].sum() + face[new].sum())
        anc = 0
        f = np.unique(parent[h])
        f = f[f >= 0]
        seen = np.zeros(n, dtype=bool)
        while f.size:
            f = f[~seen[f]]
            seen[f] = True
            anc += f.size
            f = np.unique(parent[f])
            f = f[f >= 0]
        tot["U"] += U; tot["A"] += A; tot["T"] += T; tot["M"] += M; tot["first_A"] += first_A; tot["hits"] += h.size
        stage_bytes["gather"] += 48 * nleaf
        if d.tool == 2:
            # SURVEY.md 8d smooth, per iteration U * (12 [+ 4 mask]) + M * (deg * (4 + 12) + 8 + 12); M here is already the
            # sum over the dab's iterations; grids: deg = 4 and the neighbours need no index (element's place in its grid)
            bs = min(max(float(d.bstrength), 0.0), 1.0)
            iters = int(bs * 4) + (0 if (int(bs * 4) > 0 and 4.0 * (bs - int(bs * 4) * 0.25) == 0.0) else 1)
            stage_bytes["smooth"] += iters * U * 12 + M * ((4 * 12 + 12) if grids else (6 * 16 + 8 + 12)) + first_A * 24
        else:
            stage_bytes["area_normal"] += U * 12 + M * 12
            stage_bytes["brush"] += U * 12 + M * 12 + first_A * 24
        if grids:
            # stitch + CCG normals: positions of the gathered leaves' grids read, normals written; leaf boxes read them again
            stage_bytes["normals_bb"] += U * 12 + U * 12 + U * 12 + 24 * h.size
        else:
            stage_bytes["normals_bb"] += T * 12 + A * 12 + M * 12 + 24 * h.size
        stage_bytes["bb_refit"] += 72 * anc
    ses.stroke_end()
    times = ses.stage_times()
    ses.stage_timing(False)
    return tot, stage_bytes, times


def run_ours(args, rank, world):
    from dune_sculpt_b200 import build as b
    b.build_cuda()
    b.build_host()
    from dune_sculpt_b200 import capi
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    mesh, diag, dabs = build_workload(args)
    t0 = time.time()
    dist_arg = None
    if world > 1:
        # NCCL id of the library's own communicator: made on rank 0, broadcast through torch.distributed
        import torch
        import torch.distributed as dist
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda:%d" % local_rank)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        dist_arg = (world, rank, bytes(idt.cpu().numpy().tobytes()))
    if args.config == "c5":
        ses = capi.GridSession(mesh, device=local_rank, dist=dist_arg)
    else:
        ses = capi.SculptSession(mesh, device=local_rank, dist=dist_arg)  # fails loudly without a device / the .so
    na = ses.node_arrays()
    log("[bench] host PBVH build + device upload %.1fs, %d nodes (%d leaves)" %
        (time.time() - t0, ses.totnode, int((na["flag"] & 1).sum())))
    D, ctx = ses.D, ses.ctx
    import ctypes as C

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    dab_arr = (capi.DscDab * len(dabs))(*dabs)  # the stroke script as one C array

    # every step starts from the same rest state: a device-to-device rollback to the checkpoint, timed as
    # part of the step (without it the draw strokes pile up and later steps sweep a different surface)
    ses.checkpoint()

    def device_stroke():
        ses._chk(D.dsc_state_restore(ctx))
        ses._chk(D.dsc_stroke_begin(ctx, None))
        ses._chk(D.dsc_dabs(ctx, dab_arr, len(dabs)))
        ses._chk(D.dsc_stroke_end(ctx))

    # ---- device-resident timing: `value`
    for _ in range(args.warmup):
        device_stroke()
    ses.synchronize()
    barrier()
    log("[bench] rank %d: warm-up done (%.1fs since upload)" % (rank, time.time() - t0))
    sampler = ClockSampler(local_rank)
    sampler.start()
    vd = 0
    launches = 0
    ses.timer_start()
    for _ in range(args.steps):
        device_stroke()
        st = ses.stats()  # small D2H of the counters, once per stroke
        vd += st["vertex_dabs"]
        launches += st["kernel_launches"]
    ms = ses.timer_stop()
    clocks = sampler.stop()
    barrier()
    log("[bench] rank %d: timed strokes done, %.3f ms/step" % (rank, ms / args.steps))

    # ---- end-to-end timing through the host API: `e2e`
    H = ses.H
    h2d = len(dabs) * C.sizeof(capi.DscDab)
    d2h = mesh.totvert * 24 + ses.totnode * (48 + 4) + 8
    for _ in range(1):
        ses.rollback(); ses.stroke_begin(); ses.dabs(dab_arr, len(dabs)); ses.stroke_end()
    ses.synchronize()
    barrier()
    vd_e = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ses.rollback()
        ses.stroke_begin()
        ses.dabs(dab_arr, len(dabs))  # host descriptors cross the ABI dab by dab inside the C loop
        vd_e += ses.stats()["vertex_dabs"]
        ses.stroke_end()  # flush + download co / no / boxes / flags into the host PBVH
    ses.synchronize()
    dt_e = time.perf_counter() - t0
    barrier()
    log("[bench] rank %d: end-to-end strokes done, %.3f ms/step" % (rank, 1e3 * dt_e / args.steps))

    # ---- max over ranks
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([ms, dt_e], dtype=torch.float64, device="cuda:%d" % local_rank)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s = torch.tensor([float(vd), float(vd_e), float(launches)], dtype=torch.float64, device="cuda:%d" % local_rank)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        ms, dt_e = float(t[0]), float(t[1])
        vd, vd_e, launches = int(s[0]), int(s[1]), int(s[2])

    # ---- the sweep radius by radius (untimed for `value`: one more stroke, CUDA events around each radius group)
    per = args.dabs_per_radius if args.config == "c3" else len(dabs)
    sweep = []
    ses._chk(D.dsc_state_restore(ctx))
    ses._chk(D.dsc_stroke_begin(ctx, None))
    vd_prev = 0
    for g in range(len(dabs) // per):
        grp = (capi.DscDab * per)(*dabs[g * per:(g + 1) * per])
        ses.timer_start()
        ses._chk(D.dsc_dabs(ctx, grp, per))
        g_ms = ses.timer_stop()
        vd_now = ses.stats()["vertex_dabs"]
        sweep.append({"radius_pct_diag": round(100.0 * float(dabs[g * per].radius) / diag, 2), "dabs": per,
                      "us_per_dab": round(1e3 * g_ms / per, 2), "vertex_dabs_per_dab": (vd_now - vd_prev) // per,
                      "gvd_per_s": round((vd_now - vd_prev) / (g_ms * 1e-3) / 1e9, 2) if g_ms > 0 else None})
        vd_prev = vd_now
    ses._chk(D.dsc_stroke_end(ctx))
    ses.synchronize()

    log("[bench] rank %d: radius sweep done" % rank)
    # ---- roofline of the dominant kernel (untimed analysis stroke, CUDA events per stage)
    tot, stage_bytes, times = analysis_pass(ses, dabs, na, grids=args.config == "c5")
    log("[bench] rank %d: analysis stroke done" % rank)
    peak, peak_src = measured_peaks()
    dom = max((k for k in stage_bytes), key=lambda k: times[k][0])
    dom_ms, dom_launches = times[dom]
    achieved = stage_bytes[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    total_bytes = sum(stage_bytes.values())
    total_ms = ms / args.steps  # the timed strokes themselves: side-stream refit overlapped, no per-stage events
    stages = {k: {"ms": round(times[k][0], 4), "launches": times[k][1], "alg_bytes": int(stage_bytes.get(k, 0)),
                  "gbs": round(stage_bytes.get(k, 0) / (times[k][0] * 1e-3) / 1e9, 1) if times[k][0] > 0 else None}
              for k in times}
    # DRAM traffic of the dominant kernel: one `ncu --set full` capture (profiles/r1_traffic.json) gives measured
    # bytes / algorithmic bytes for one launch; scaled to the average launch the `achieved` figure is quoted on
    traffic, traffic_note = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if dom == "normals_bb" and args.config == "c3":
            traffic = int(stage_bytes[dom] / max(dom_launches, 1) * float(tr["traffic_over_algorithmic"]))
            traffic_note = ("ncu dram__bytes_read+write of %s = %.3f x its algorithmic bytes on the captured launch (%s); "
                            "scaled to the average launch" % (tr["kernel"], tr["traffic_over_algorithmic"], "profiles/r1_traffic.json"))
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                "alg_bytes_per_launch": int(stage_bytes[dom] / max(dom_launches, 1)),
                "whole_path": {"achieved": round(total_bytes / (total_ms * 1e-3) / 1e9, 1) if total_ms > 0 else None,
                               "frac": round(total_bytes / (total_ms * 1e-3) / 1e9 / peak, 4) if total_ms > 0 else None,
                               "frac_of_8TBs_nominal": round(total_bytes / (total_ms * 1e-3) / 1e9 / 8000.0, 4) if total_ms > 0 else None,
                               "bytes_per_vertex_dab": round(total_bytes / max(tot["U"], 1), 2),
                               "how": "algorithmic bytes of one stroke / device time of one timed stroke"},
                "stages": stages, "radius_sweep": sweep}

    if rank != 0:
        ses.close()
        return

    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle, OpenMP, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, mesh, dabs)

    ndabs = len(dabs) * args.steps
    out = {
        "metric": METRIC, "value": vd / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, mesh, len(dabs)),
        "ms_per_dab": ms / ndabs, "vertex_dabs_per_step": vd // args.steps,
        "e2e": {"value": vd_e / dt_e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * dt_e / args.steps},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    if cpu:
        out["cpu_baseline"] = cpu
    emit(out)
    ses.close()


def cpu_baseline(args, mesh, dabs):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from dune_sculpt_b200 import build as b
    b.build_oracle()
    from oracle_py import GridOracle, Oracle
    cores = os.cpu_count() or 1
    t0 = time.time()
    if args.config == "c5":
        orc = GridOracle(mesh, threads=cores)
        log("[bench] cpu_baseline: grids oracle build %.1fs" % (time.time() - t0))
        sample = dabs[:args.c5_cpu_dabs]
        orc.stroke_begin()
        orc.dab(sample[0])
        vd0 = orc.vertex_dabs()
        t0 = time.perf_counter()
        for d in sample:
            orc.dab(d)
        dt = time.perf_counter() - t0
        vd = orc.vertex_dabs() - vd0
        orc.close()
        return {"value": vd / dt, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "the first %d dabs of the same stroke (%.1f s), OpenMP over faces / edges / nodes" % (len(sample), dt),
                "ms_per_dab": 1e3 * dt / len(sample)}
    orc = Oracle(mesh, threads=cores)
    log("[bench] cpu_baseline: oracle PBVH build %.1fs" % (time.time() - t0))
    per = args.dabs_per_radius
    nrad = len(dabs) // per
    sample = [dabs[r * per + k] for r in range(nrad) for k in range(args.cpu_sample_per_radius)]
    orc.stroke_begin()
    orc.dab(sample[0])
    vd0 = orc.vertex_dabs()
    t0 = time.perf_counter()
    for d in sample:
        orc.dab(d)
    dt = time.perf_counter() - t0
    vd = orc.vertex_dabs() - vd0
    orc.close()
    return {"value": vd / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d dabs per radius x %d radii of the same sweep (%d dabs, %.1f s), OpenMP over hit nodes" %
                      (args.cpu_sample_per_radius, nrad, len(sample), dt),
            "ms_per_dab": 1e3 * dt / len(sample)}


_REAL_STDOUT = None


def emit(obj):
    """the one JSON line, on the process's original stdout"""
    line = json.dumps(obj) + "\n"
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line.encode())
    else:
        sys.stdout.write(line)
        sys.stdout.flush()


def main():
    # libraries (NCCL's version banner, ...) may write to fd 1: route it to stderr and keep the real stdout for the JSON line
    global _REAL_STDOUT
    try:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
    except OSError:
        _REAL_STDOUT = None
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--dabs-per-radius", type=int, default=32)
    ap.add_argument("--cpu-sample-per-radius", type=int, default=32,
                    help="CPU legs: dabs of every radius of the sweep (32 = the whole stroke, ~ 13 s on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c3", choices=["c3", "c5"],
                    help="c3 (default): the headline 16.7M-vertex draw sweep; c5: multires grids, draw stroke")
    ap.add_argument("--c5-base", type=int, default=25)
    ap.add_argument("--c5-level", type=int, default=7)
    ap.add_argument("--c5-dabs", type=int, default=100, help="draw dabs of the C5 stroke")
    ap.add_argument("--c5-smooth-dabs", type=int, default=100, help="smooth dabs ahead of them")
    ap.add_argument("--c5-cpu-dabs", type=int, default=0, help="CPU legs of c5: the first N dabs of the stroke (0 = all of it, ~ 8 s)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # the CPU arm's steps are bounded samples: cap them so the run ends within minutes
    if args.c5_cpu_dabs <= 0:
        args.c5_cpu_dabs = args.c5_dabs + args.c5_smooth_dabs
    args.cpu_sample_per_radius = max(1, min(args.cpu_sample_per_radius, args.dabs_per_radius))
    args.steps_ref = max(1, min(args.steps, 3))
    args.warmup_ref = max(0, min(args.warmup, 1))

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl")
    run_ours(args, rank, world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
