#!/usr/bin/env python
"""Headline benchmark of the sculpt-stroke hot path (BASELINE.json: vertex-dabs/sec and ms/dab at 1M / 16M verts,
% of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c3|c1|c2|c4|c5]

Headline workload (config.workload): C3 -- draw brush + normal recompute + BB refit on the 16,777,216-vertex
height-field grid, radius sweep 1-50 % of the bounding-box diagonal, 32 dabs per radius.  One *step* = a
device-to-device rollback of the mesh to its rest state + one pass of the whole stroke script (every step does identical
work).  `value` = vertex-dabs / second with the mesh resident in HBM (CUDA events around the K steps); `e2e` = the same
steps driven through the reference-named host API with host buffers: descriptors go host->device, at stroke end
positions, normals, node boxes and flags come back device->host, all inside the timed region.

`roofline`: the dominant kernel's algorithmic bytes (SURVEY.md 8d formulas over device-counted U / A / T / M / U' / M')
divided by its duration INSIDE the replayed graphs (event-record nodes, dsc_stage_timing mode 2: the path that is
timed), against the measured copy bandwidth; per radius too.  `whole_path` = SURVEY.md 8d's headline formula
(U 12 + M 12 + T 12 + A 12 + D 12, area pass excluded) / device time of a timed step.
`cpu_baseline` + `parity_fullsize`: the CPU oracle (OpenMP) runs the very same full-size stroke; its result is
compared with the device's -- per-dab node-hit lists, undo-node membership, final positions / normals / boxes.

At N = 1 the line also carries `configs`: the same measurements for C1, C2, C4, C5 (and `c5` again at top level, which
N > 1 runs carry too: the partitioned multires config the north star names for scaling).
--impl reference times the CPU oracle (OpenMP, all host threads) on a bounded sample of the headline workload.
Prints ONE JSON line on rank 0.
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
sys.path.insert(0, ROOT)

METRIC = "vertex_dabs_per_sec"
UNIT = "vertex-dabs/s"
T0 = time.time()


def log(*a):
    print("[bench %6.1fs]" % (time.time() - T0), *a, file=sys.stderr, flush=True)


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


# ------------------------------------------------------------------------------------------ workloads
class Workload:
    """a mesh + the strokes of one step.  stroke = dict(label, dabs, automask, mask_on); groups = dabs per timed group
    of the sweep (C3: one radius), 0 = the whole stroke"""

    def __init__(self, name, mesh, grids, strokes, desc, brush, mask=None, group=0):
        self.name, self.mesh, self.grids, self.strokes, self.desc, self.brush = name, mesh, grids, strokes, desc, brush
        self.mask, self.group = mask, group
        self.diag = mesh.bbox_diag()
        self.verts = mesh.totelem if grids else mesh.totvert

    @property
    def ndabs(self):
        return sum(len(s["dabs"]) for s in self.strokes)


def build_workload(name, args):
    from dune_sculpt_b200 import capi, meshgen, stroke
    t0 = time.time()
    if name == "c1":
        mesh = meshgen.cube(args.c1_levels)
        w = Workload("c1", mesh, False, [dict(label="draw line", dabs=stroke.c1_draw_stroke(), automask=None, mask_on=False)],
                     "C1 draw, 100-dab straight stroke across the +Z face, r = 0.15, cube subdivided x%d (V=%d)" % (args.c1_levels, mesh.totvert),
                     "draw (alpha 0.5), SMOOTH falloff, area-normal direction")
    elif name == "c2":
        mesh = meshgen.icosphere(args.c2_freq, noise=0.002)
        w = Workload("c2", mesh, False, [dict(label="smooth arc", dabs=stroke.c2_smooth_stroke(), automask=None, mask_on=False)],
                     "C2 smooth (CSR one-ring averaging, 3 iterations / dab), 200-dab great-circle stroke, r = 0.2, icosphere f=%d (V=%d)" %
                     (args.c2_freq, mesh.totvert), "smooth (alpha 0.75), SMOOTH falloff")
    elif name == "c3":
        mesh = meshgen.grid(args.grid)
        dabs = stroke.c3_radius_sweep(mesh.bbox_diag(), dabs_per_radius=args.dabs_per_radius)
        w = Workload("c3", mesh, False, [dict(label="radius sweep", dabs=dabs, automask=None, mask_on=False)],
                     "C3 draw+normals+BB radius sweep 1-50%% bbox diag, grid %d^2 (V=%d), %d dabs/stroke" % (args.grid, mesh.totvert, len(dabs)),
                     "draw (alpha 0.5), SMOOTH falloff, area-normal direction", group=args.dabs_per_radius)
    elif name == "c4":
        mesh = meshgen.grid(args.c4_grid)
        diag = mesh.bbox_diag()
        mask = meshgen.low_freq_mask(mesh)
        # automask factors come from the host library's session helpers (boundary-edge propagation, 1 step; topology flood fill
        # from the vertex nearest the first dab)
        ses = capi.SculptSession(mesh)
        auto_b = np.zeros(mesh.totvert, dtype=np.float32)
        capi.host_lib().DUNE_sculpt_automask_boundary_edges(ses.pbvh, 1, capi.fptr(auto_b))
        strokes = []
        tools = (("draw", capi.TOOL_DRAW), ("inflate", capi.TOOL_INFLATE), ("clay strips", capi.TOOL_CLAY_STRIPS), ("grab", capi.TOOL_GRAB))
        for tname, tool in tools:
            dabs = stroke.c4_tool_stroke(tool, diag, dabs=args.c4_dabs)
            strokes.append(dict(label=tname + ", mask off", dabs=dabs, automask=None, mask_on=False))
        for tname, tool in tools:
            dabs = stroke.c4_tool_stroke(tool, diag, dabs=args.c4_dabs)
            strokes.append(dict(label=tname + ", mask + boundary automask", dabs=dabs, automask=auto_b, mask_on=True))
        dabs = stroke.c4_tool_stroke(capi.TOOL_DRAW, diag, dabs=args.c4_dabs)
        loc = np.array(dabs[0].location[:], dtype=np.float32)
        seed = int(np.argmin(((np.asarray(mesh.co) - loc) ** 2).sum(axis=1)))
        auto_t = np.zeros(mesh.totvert, dtype=np.float32)
        capi.host_lib().DUNE_sculpt_automask_topology(ses.pbvh, seed, capi.fptr(loc), C_float(0.0), capi.fptr(auto_t))
        ses.close()
        strokes.append(dict(label="draw, mask + topology automask", dabs=dabs, automask=auto_t, mask_on=True))
        w = Workload("c4", mesh, False, strokes,
                     "C4 tool sweep draw / inflate / clay strips / grab x {mask off, mask layer + boundary automask} + draw with topology "
                     "automask, r = 10%% bbox diag, %d dabs each, grid %d^2 (V=%d)" % (args.c4_dabs, args.c4_grid, mesh.totvert),
                     "draw / inflate / clay strips / grab (alpha 0.5)", mask=mask)
    elif name == "c5":
        mesh = meshgen.multires_cube_n(args.c5_base, args.c5_level)
        diag = mesh.bbox_diag()
        rng = np.random.default_rng(5)
        bs = stroke._strength(capi.TOOL_DRAW, 0.5)
        bsm = stroke._strength(capi.TOOL_SMOOTH, 0.75)
        dabs = []
        for tool, n, b in ((capi.TOOL_SMOOTH, args.c5_smooth_dabs, bsm), (capi.TOOL_DRAW, args.c5_dabs, bs)):
            for i in range(n):
                p = rng.normal(size=3)
                p /= np.linalg.norm(p)
                dabs.append(capi.make_dab(tool, p.astype(np.float32), diag * 0.08, bstrength=b, view_normal=tuple(p),
                                          flags=capi.DAB_FIRST_STEP if not dabs else 0))
        w = Workload("c5", mesh, True, [dict(label="smooth then draw", dabs=dabs, automask=None, mask_on=False)],
                     "C5 multires grids: cube %d^2 x 6 base quads, level %d (%d grids of %d^2 = %d elements), %d smooth dabs (3 iterations) then "
                     "%d draw dabs, each + stitch + CCG normals + BB, r = 8%% bbox diag" %
                     (args.c5_base, args.c5_level, mesh.totgrid, mesh.grid_size, mesh.totelem, args.c5_smooth_dabs, args.c5_dabs),
                     "smooth (alpha 0.75) then draw (alpha 0.5), SMOOTH falloff, area-normal direction")
    else:
        raise ValueError(name)
    log("%s: %s -- built in %.1fs" % (name, w.desc, time.time() - t0))
    return w


def C_float(x):
    import ctypes
    return ctypes.c_float(x)


def config_of(w, world):
    par = "single GPU"
    if world > 1:
        par = ("PBVH partitioned spatially over %d GPUs (the same mesh and stroke: strong scaling); a dab runs on, and is exchanged over "
               "NVLink peer memory among, only the ranks it can reach -- the others skip it and run ahead; stroke end brings every rank's "
               "own runs back to its host" % world)
    return {"workload": w.desc, "parallelism": par, "verts": w.verts, "dabs_per_step": w.ndabs, "strokes_per_step": len(w.strokes),
            "brush": w.brush,
            "l2": ("inputs larger than L2 (resident arrays > 1 GB; every stroke sweeps them)" if w.verts > 4000000 else
                   "flushed between steps by the rollback: a device-to-device copy of every position / normal array (> L2 for C2; C1's "
                   "arrays fit in L2 -- an L2-resident config by the north star's own sizing)"),
            "step": "device-to-device rollback to the rest state + %d stroke%s (%d dabs)" % (len(w.strokes), "" if len(w.strokes) == 1 else "s", w.ndabs)}


# --------------------------------------------------------------------------------- byte accounting
def smooth_iterations(bstrength):
    bs = min(max(float(bstrength), 0.0), 1.0)
    full = int(bs * 4)
    return full + (0 if (full > 0 and 4.0 * (bs - full * 0.25) == 0.0) else 1)


def stage_bytes_of(st, dabs, nleaf, grids, mask_on, automask_on):
    """Algorithmic bytes per stage of a run of dabs of ONE tool, from the device's counters `st` (deltas
    over the run): U vertex_dabs, A all_verts, T prims, M moved_verts, first_A first_touch_verts, U' area_verts,
    M' area_inside, hits, refit nodes."""
    U, A, T, M = st["vertex_dabs"], st["all_verts"], st["prims"], st["moved_verts"]
    fa, Ua, Ma, hits = st["first_touch_verts"], st["area_verts"], st["area_inside"], st["node_hits"]
    n = len(dabs)
    tool = dabs[0].tool
    per_u = 12 + (4 if mask_on else 0) + (4 if automask_on else 0)
    b = {"gather": 48 * nleaf * n, "area_normal": 0, "brush": 0, "smooth": 0, "normals_bb": 0, "bb_refit": 72 * st["refit_nodes"] + 24 * hits}
    if tool == 2:
        it = smooth_iterations(dabs[0].bstrength)
        # M counts every iteration's moved verts; the verts whose normals are recomputed are those of one iteration
        b["smooth"] = it * U * per_u + M * ((4 * 12 + 12) if grids else (6 * 16 + 8 + 12)) + fa * 24
        D = M // max(it, 1)
    else:
        if tool == 4 or (dabs[0].flags & 1):
            per_u += 12
        b["brush"] = U * per_u + M * 12 + fa * 24
        b["area_normal"] = Ua * 12 + Ma * 12  # zero when the tool samples no plane: the counters stay put
        D = M
    if grids:
        b["normals_bb"] = U * 36 + 24 * hits  # stitch + CCG normals: positions twice, normals once
    else:
        b["normals_bb"] = T * 12 + A * 12 + D * 12 + 24 * hits
    b["headline_8d"] = 0 if grids else U * 12 + M * 12 + T * 12 + A * 12 + D * 12
    return b


def tool_runs(dabs):
    """maximal runs of equal tool"""
    runs, i = [], 0
    while i < len(dabs):
        j = i
        while j < len(dabs) and dabs[j].tool == dabs[i].tool:
            j += 1
        runs.append((i, j))
        i = j
    return runs


# --------------------------------------------------------------------------------------- measurement
class Runner:
    def __init__(self, w, args, rank, world, local_rank):
        from dune_sculpt_b200 import capi
        import ctypes as C
        self.w, self.args, self.rank, self.world, self.capi, self.C = w, args, rank, world, capi, C
        t0 = time.time()
        dist_arg = None
        if world > 1:
            import torch
            import torch.distributed as dist
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda:%d" % local_rank)
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            dist_arg = (world, rank, bytes(idt.cpu().numpy().tobytes()))
        if w.grids:
            self.ses = capi.GridSession(w.mesh, device=local_rank, dist=dist_arg)
        else:
            self.ses = capi.SculptSession(w.mesh, mask=w.mask, device=local_rank, dist=dist_arg)  # fails loudly without a device / the .so
        self.na = self.ses.node_arrays()
        self.nleaf = int((self.na["flag"] & 1).sum())
        self.session_start_s = time.time() - t0
        log("%s rank %d: host PBVH build + device upload %.1fs, %d nodes (%d leaves)" % (w.name, rank, self.session_start_s, self.ses.totnode, self.nleaf))
        self.arrs = [(capi.DscDab * len(s["dabs"]))(*s["dabs"]) for s in w.strokes]
        self.mask_state = w.mask is not None
        # the per-stroke layers (mask, automask factors) go up from page-locked memory, as a host application would keep them
        seen = set()
        self.pinned = []
        for a in [w.mask] + [s["automask"] for s in w.strokes]:
            if a is not None and id(a) not in seen:
                seen.add(id(a))
                if self.ses.D.dsc_host_register(self.ses.ctx, a.ctypes.data, a.nbytes) == 0:  # best effort: pageable works too
                    self.pinned.append(a)
        self.ses.checkpoint()

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()

    def _set_mask(self, on):
        if self.w.mask is None or on == self.mask_state:
            return 0
        ses = self.ses
        ses._chk(ses.D.dsc_set_mask(ses.ctx, self.capi.fptr(self.w.mask) if on else None))
        self.mask_state = on
        return self.w.mask.nbytes if on else 0

    def device_step(self):
        ses, D, ctx = self.ses, self.ses.D, self.ses.ctx
        for s, arr in zip(self.w.strokes, self.arrs):
            ses._chk(D.dsc_state_restore(ctx))
            self._set_mask(s["mask_on"])
            a = s["automask"]
            ses._chk(D.dsc_stroke_begin(ctx, None if a is None else self.capi.fptr(a)))
            ses._chk(D.dsc_dabs(ctx, arr, len(arr)))
            ses._chk(D.dsc_stroke_end(ctx))

    def host_step(self):
        """the same through the reference-named host API: stroke end brings the mesh back to host memory"""
        ses = self.ses
        h2d = 0
        for s, arr in zip(self.w.strokes, self.arrs):
            ses.rollback()
            h2d += self._set_mask(s["mask_on"])
            ses.stroke_begin(s["automask"])
            if s["automask"] is not None:
                h2d += s["automask"].nbytes
            ses.dabs(arr, len(arr))
            ses.stroke_end()
            h2d += len(arr) * self.C.sizeof(self.capi.DscDab)
        return h2d

    def run(self):
        args, ses, w = self.args, self.ses, self.w
        steps, warmup = args.steps, args.warmup
        if w.name != args.config:  # side configs: bounded
            steps, warmup = min(steps, args.side_steps), min(warmup, 3)
        for _ in range(warmup):
            self.device_step()
        ses.synchronize()
        self.barrier()
        sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", self.rank)))
        sampler.start()
        ses.timer_start()
        for _ in range(steps):
            self.device_step()
        ms = ses.timer_stop()
        clocks = sampler.stop()
        self.barrier()
        # vertex-dabs of one step: the strokes' counters reset at stroke begin, so count them stroke by stroke once
        vd_step, launches_step = self.count_step()
        vd = vd_step * steps
        launches = launches_step * steps
        log("%s rank %d: timed steps done, %.3f ms/step" % (w.name, self.rank, ms / steps))

        # ---- end to end through the host API
        self.host_step()
        ses.synchronize()
        self.barrier()
        t0 = time.perf_counter()
        h2d = 0
        for _ in range(steps):
            h2d = self.host_step()
        ses.synchronize()
        dt_e = time.perf_counter() - t0
        self.barrier()
        # whole MVert records + normals (mesh) or whole CCGElem records (grids), node boxes and flags, once per stroke
        d2h = len(w.strokes) * (w.verts * 28 + ses.totnode * (48 + 4))
        log("%s rank %d: end-to-end steps done, %.3f ms/step" % (w.name, self.rank, 1e3 * dt_e / steps))

        if self.world > 1:
            import torch
            import torch.distributed as dist
            dev = "cuda:%d" % int(os.environ.get("LOCAL_RANK", self.rank))
            t = torch.tensor([ms, dt_e], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, dt_e = float(t[0]), float(t[1])
            s_ = torch.tensor([float(vd_step), float(launches_step)], dtype=torch.float64, device=dev)
            dist.all_reduce(s_, op=dist.ReduceOp.SUM)
            vd_step, launches_step = int(s_[0]), int(s_[1])
            vd, launches = vd_step * steps, launches_step * steps

        out = {"value": vd / (ms * 1e-3), "unit": UNIT, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
               "ms_per_dab": ms / (steps * w.ndabs), "vertex_dabs_per_step": vd_step, "gpu_launches": launches,
               "session_start_s": round(self.session_start_s, 2),
               "e2e": {"value": vd / dt_e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "ms_per_step": 1e3 * dt_e / steps}}
        if self.world == 1:
            out["roofline"] = self.roofline(ms / steps)
        return out, clocks

    def count_step(self):
        ses, D, ctx = self.ses, self.ses.D, self.ses.ctx
        vd = launches = 0
        for s, arr in zip(self.w.strokes, self.arrs):
            ses._chk(D.dsc_state_restore(ctx))
            self._set_mask(s["mask_on"])
            a = s["automask"]
            ses._chk(D.dsc_stroke_begin(ctx, None if a is None else self.capi.fptr(a)))
            ses._chk(D.dsc_dabs(ctx, arr, len(arr)))
            st = ses.stats()
            ses._chk(D.dsc_stroke_end(ctx))
            vd += st["vertex_dabs"]
            launches += st["kernel_launches"]
        ses.synchronize()
        return vd, launches

    # ---- roofline: device counters + kernel durations inside the replayed graphs
    def roofline(self, step_ms):
        ses, D, ctx, w, capi = self.ses, self.ses.D, self.ses.ctx, self.w, self.capi
        peak, peak_src = measured_peaks()
        keys = ("vertex_dabs", "all_verts", "prims", "moved_verts", "first_touch_verts", "area_verts", "area_inside", "node_hits", "refit_nodes")
        # (1) the sweep group by group, un-instrumented (CUDA events around each group's graph replays)
        groups = []
        for si, (s, arr) in enumerate(zip(w.strokes, self.arrs)):
            dabs = s["dabs"]
            cuts = [(a, b) for (a, b) in tool_runs(dabs)]
            if w.group:
                cuts = [(a, min(a + w.group, len(dabs))) for a in range(0, len(dabs), w.group)]
            ses._chk(D.dsc_state_restore(ctx))
            self._set_mask(s["mask_on"])
            a_ = s["automask"]
            ses._chk(D.dsc_stroke_begin(ctx, None if a_ is None else capi.fptr(a_)))
            prev = {k: 0 for k in keys}
            for (a, b) in cuts:
                grp = (capi.DscDab * (b - a))(*dabs[a:b])
                ses.timer_start()
                ses._chk(D.dsc_dabs(ctx, grp, b - a))
                g_ms = ses.timer_stop()
                st = ses.stats()
                delta = {k: st[k] - prev[k] for k in keys}
                prev = {k: st[k] for k in keys}
                groups.append({"stroke": si, "a": a, "b": b, "ms": g_ms, "st": delta,
                               "bytes": stage_bytes_of(delta, dabs[a:b], self.nleaf, w.grids, s["mask_on"], a_ is not None)})
            ses._chk(D.dsc_stroke_end(ctx))
        ses.synchronize()
        # (2) the same again with the event-record nodes: kernel durations per group and stage
        # kernel durations: CUPTI's hardware timestamps of the untouched replay path; event-record nodes (which also measure
        # each node's launch latency) when CUPTI cannot trace
        timing_how = "CUPTI activity records (hardware timestamps of every kernel of the replayed graphs)"
        try:
            ses.stage_timing(3)
        except Exception as e:
            log("CUPTI tracing unavailable (%s): event-record nodes instead" % e)
            timing_how = "event-record nodes inside the replayed graphs (each pair also measures its node's launch latency)"
            ses.stage_timing(2)
        gi = 0
        for si, (s, arr) in enumerate(zip(w.strokes, self.arrs)):
            dabs = s["dabs"]
            ses._chk(D.dsc_state_restore(ctx))
            self._set_mask(s["mask_on"])
            a_ = s["automask"]
            ses._chk(D.dsc_stroke_begin(ctx, None if a_ is None else capi.fptr(a_)))
            prev_t = {k: v[0] for k, v in ses.stage_times().items()}
            while gi < len(groups) and groups[gi]["stroke"] == si:
                g = groups[gi]
                grp = (capi.DscDab * (g["b"] - g["a"]))(*dabs[g["a"]:g["b"]])
                ses._chk(D.dsc_dabs(ctx, grp, g["b"] - g["a"]))
                now = ses.stage_times()
                g["stage_ms"] = {k: now[k][0] - prev_t.get(k, 0.0) for k in now}
                prev_t = {k: v[0] for k, v in now.items()}
                gi += 1
            ses._chk(D.dsc_stroke_end(ctx))
        times = ses.stage_times()
        ses.stage_timing(0)
        ses.synchronize()

        stage_names = [k for k in times if k not in ("other", "leaf_bb")]
        tot_b = {k: sum(g["bytes"].get(k, 0) for g in groups) for k in stage_names}
        tot_ms = {k: sum(g["stage_ms"].get(k, 0.0) for g in groups) for k in stage_names}
        dom = max(stage_names, key=lambda k: tot_ms[k])
        dom_launches = max(times[dom][1], 1)
        ach = tot_b[dom] / (tot_ms[dom] * 1e-3) / 1e9 if tot_ms[dom] > 0 else 0.0
        stages = {k: {"ms": round(tot_ms[k], 4), "launches": times[k][1], "alg_bytes": int(tot_b[k]),
                      "gbs": round(tot_b[k] / (tot_ms[k] * 1e-3) / 1e9, 1) if tot_ms[k] > 0 else None} for k in stage_names}
        total_bytes = sum(tot_b[k] for k in stage_names)
        head = sum(g["bytes"]["headline_8d"] for g in groups)
        kernels_ms = sum(tot_ms[k] for k in stage_names if k != "bb_refit")  # the refit overlaps on the side stream
        U = sum(g["st"]["vertex_dabs"] for g in groups)
        M = sum(g["st"]["moved_verts"] for g in groups)
        sweep = []
        for g in groups:
            d0 = w.strokes[g["stroke"]]["dabs"][g["a"]]
            n = g["b"] - g["a"]
            dm = g["stage_ms"].get(dom, 0.0)
            sweep.append({"stroke": w.strokes[g["stroke"]]["label"], "tool": int(d0.tool), "radius_pct_diag": round(100.0 * float(d0.radius) / w.diag, 2),
                          "dabs": n, "us_per_dab": round(1e3 * g["ms"] / n, 2), "vertex_dabs_per_dab": g["st"]["vertex_dabs"] // n,
                          "gvd_per_s": round(g["st"]["vertex_dabs"] / (g["ms"] * 1e-3) / 1e9, 2) if g["ms"] > 0 else None,
                          "moved_frac": round(g["st"]["moved_verts"] / max(g["st"]["vertex_dabs"], 1), 3),
                          "dominant_kernel_us_per_dab": round(1e3 * dm / n, 2),
                          "dominant_kernel_frac": round(g["bytes"].get(dom, 0) / (dm * 1e-3) / 1e9 / peak, 4) if dm > 0 else None,
                          "whole_path_frac": round(sum(g["bytes"][k] for k in stage_names) / (g["ms"] * 1e-3) / 1e9 / peak, 4) if g["ms"] > 0 else None})
        rl = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
              "traffic": None, "peak_source": peak_src, "alg_bytes_per_launch": int(tot_b[dom] / dom_launches),
              "how": "algorithmic bytes of the stage (SURVEY.md 8d over device-counted U/A/T/M/U'/M') / its kernels' durations in one more step of "
                     "the same replayed graphs: " + timing_how,
              "whole_path": {"achieved": round(total_bytes / (step_ms * 1e-3) / 1e9, 1), "frac": round(total_bytes / (step_ms * 1e-3) / 1e9 / peak, 4),
                             "frac_of_8TBs_nominal": round(total_bytes / (step_ms * 1e-3) / 1e9 / 8000.0, 4),
                             "bytes_per_vertex_dab": round(total_bytes / max(U, 1), 2),
                             "headline_8d": None if w.grids else {"bytes": int(head), "bytes_per_vertex_dab": round(head / max(U, 1), 2),
                                                                   "frac": round(head / (step_ms * 1e-3) / 1e9 / peak, 4),
                                                                   "formula": "U 12 + M 12 + T 12 + A 12 + D 12 (area pass excluded)"},
                             "moved_frac": round(M / max(U, 1), 3),
                             "how": "algorithmic bytes of one step (all stages, area pass charged with U' 12 + M' 12) / device time of a timed step"},
              "kernel_time_in_step": {"kernels_ms": round(kernels_ms, 4), "step_ms": round(step_ms, 4),
                                      "note": "sum of the main-stream kernels' durations inside the graphs vs the timed step (the rest is launch "
                                              "gaps and the rollback copy)"},
              "stages": stages, "groups": sweep}
        # DRAM traffic of the dominant kernel from the committed ncu capture, if one was made for this kernel
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            if tr.get("stage") == dom and tr.get("config") == w.name:
                rl["traffic"] = int(rl["alg_bytes_per_launch"] * float(tr["traffic_over_algorithmic"]))
                rl["traffic_note"] = ("ncu dram__bytes_read+write of %s = %.3f x its algorithmic bytes on the captured launch (profiles/r2_traffic.json), "
                                      "scaled to the average launch" % (tr["kernel"], tr["traffic_over_algorithmic"]))
                if "l2_hit_rate" in tr:
                    rl["l2_hit_rate"] = tr["l2_hit_rate"]
        except Exception:
            pass
        return rl

    # ---- the device result of one full-size step, dab by dab, for the parity leg
    def parity_device(self):
        ses, w = self.ses, self.w
        out = []
        for s in w.strokes:
            ses.rollback()
            self._set_mask(s["mask_on"])
            ses.stroke_begin(s["automask"])
            h = hashlib.sha1()
            for d in s["dabs"]:
                ses.dab(d)
                h.update(ses.hits().tobytes())
                h.update(b"|")
            touched = ses.touched()
            ses.stroke_end()
            bb, obb = ses.node_bb()
            out.append({"hits": h.hexdigest(), "touched": touched, "co": ses.co(), "no": ses.no(), "bb": bb, "vd": ses.stats()["vertex_dabs"]})
        return out

    def close(self):
        for a in self.pinned:
            self.ses.D.dsc_host_unregister(self.ses.ctx, a.ctypes.data)
        self.pinned = []
        self.ses.close()


def cpu_leg(w, args, device_results, sample=None):
    """The CPU oracle (OpenMP over hit nodes) over the same full-size strokes: timed (`cpu_baseline`) and, when the device
    results are given, compared with them (`parity_fullsize`)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from dune_sculpt_b200 import build as b
    b.build_oracle()
    from oracle_py import GridOracle, Oracle
    cores = os.cpu_count() or 1
    t0 = time.time()
    # threaded, but vertex normals summed in the serial loop's order: bit-identical to the single-threaded restatement
    from oracle_py import lib as oracle_lib
    oracle_lib().or_set_ordered_normals(1 if device_results is not None else 0)
    orc = GridOracle(w.mesh, threads=cores) if w.grids else Oracle(w.mesh, threads=cores, mask=w.mask)
    log("%s cpu: oracle PBVH build %.1fs" % (w.name, time.time() - t0))
    co0, no0 = orc.co(), orc.no()
    na0 = orc.node_arrays()
    vd = 0
    dt = 0.0
    ndabs = 0
    par = {"strokes": [], "tolerance": "1e-5 x bbox diagonal (BASELINE.json north star); bit equality reported beside it"}
    tol = 1e-5 * w.diag
    ok = True
    for si, s in enumerate(w.strokes):
        if si:
            # back to the rest state: coordinates, normals, boxes, flags
            orc.set_co(co0)
            np.ctypeslib.as_array(orc.L.or_pbvh_no(orc.p), shape=(orc.totvert * 3,))[:] = no0.reshape(-1)
            for n in np.nonzero(na0["flag"] & 1)[0]:
                orc.L.or_node_mark_update(orc.p, int(n))
            orc.update_bounds(4 | 8)
            for n in np.nonzero(na0["flag"] & 1)[0]:
                orc.set_node_flag(int(n), 2 | 4 | 8 | 16 | 32, False)
        if w.mask is not None:
            m = np.ctypeslib.as_array(orc.L.or_pbvh_mask(orc.p), shape=(orc.totvert,))
            m[:] = w.mask if s["mask_on"] else 0.0
        orc.stroke_begin(s["automask"])
        h = hashlib.sha1()
        t0 = time.perf_counter()
        for d in s["dabs"]:
            orc.dab(d)
            if device_results is not None:
                h.update(orc.hits().tobytes())
                h.update(b"|")
        dt += time.perf_counter() - t0
        ndabs += len(s["dabs"])
        vd += orc.vertex_dabs()
        touched = orc.touched()
        orc.stroke_end()
        if device_results is not None:
            dv = device_results[si]
            co, no = orc.co(), orc.no()
            dco, dno = float(np.abs(co - dv["co"]).max()), float(np.abs(no - dv["no"]).max())
            dbb = float(np.abs(orc.node_arrays()["vb"] - dv["bb"]).max())
            rec = {"stroke": s["label"], "dabs": len(s["dabs"]), "hit_lists_equal": h.hexdigest() == dv["hits"],
                   "touched_equal": bool(np.array_equal(touched, dv["touched"])), "vertex_dabs_equal": int(orc.vertex_dabs()) == int(dv["vd"]),
                   "max_dco": dco, "max_dno": dno, "max_dbb": dbb, "within_tolerance": bool(dco <= tol and dno <= tol and dbb <= tol),
                   "bit_exact": bool(np.array_equal(co, dv["co"]) and np.array_equal(no, dv["no"]))}
            ok = ok and rec["hit_lists_equal"] and rec["touched_equal"] and rec["vertex_dabs_equal"] and rec["within_tolerance"]
            par["strokes"].append(rec)
    orc.L.or_set_ordered_normals(0)
    orc.close()
    par["ok"] = bool(ok)
    par["bit_exact"] = bool(all(r["bit_exact"] for r in par["strokes"])) if par["strokes"] else None
    cpu = {"value": vd / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "the whole step: %d stroke%s, %d dabs (%.1f s), OpenMP over hit nodes" % (len(w.strokes), "" if len(w.strokes) == 1 else "s", ndabs, dt),
           "ms_per_dab": 1e3 * dt / max(ndabs, 1)}
    return cpu, (par if device_results is not None else None)


def measure(name, args, rank, world, local_rank, with_cpu):
    w = build_workload(name, args)
    r = Runner(w, args, rank, world, local_rank)
    res, clocks = r.run()
    res["config"] = config_of(w, world)
    if with_cpu and rank == 0 and world == 1:
        dev = r.parity_device()
        log("%s: device parity stroke done" % name)
        cpu, par = cpu_leg(w, args, dev)
        res["cpu_baseline"] = cpu
        res["parity_fullsize"] = par
        log("%s: cpu leg done, parity ok = %s" % (name, par["ok"]))
    elif world > 1 and args.verify_partition:
        res["parity_vs_1gpu"] = partition_digest(w, r, rank, local_rank)
    r.close()
    return res, clocks


def partition_digest(w, r, rank, local_rank):
    """N > 1: the partitioned stroke-end mesh on rank 0 against the same stroke on one GPU (a second, unpartitioned session)"""
    ses = r.ses
    res = None
    s = w.strokes[0]
    ses.rollback()
    ses.stroke_begin(s["automask"])
    ses.dabs(r.arrs[0], len(r.arrs[0]))
    ses.stroke_end()
    counts = ses.dist_dab_counts()
    ses.gather()  # the owners' vertex data to every rank: whole replicas for the comparison
    co, no = ses.co(), ses.no()
    r.barrier()
    if rank == 0:
        capi = r.capi
        one = capi.GridSession(w.mesh, device=local_rank) if w.grids else capi.SculptSession(w.mesh, mask=w.mask, device=local_rank)
        one.stroke_begin(s["automask"])
        one.dabs(r.arrs[0], len(r.arrs[0]))
        one.stroke_end()
        res = {"co_bit_equal": bool(np.array_equal(co, one.co())), "no_bit_equal": bool(np.array_equal(no, one.no())),
               "sha1_co": hashlib.sha1(co.tobytes()).hexdigest()[:16],
               "rank0_dabs": counts, "how": "rank 0's replica after one partitioned stroke + gather vs the same stroke on one GPU"}
        one.close()
    r.barrier()
    return res


# ---------------------------------------------------------------------------------- reference arm
def run_reference(args, rank):
    """CPU arm: the oracle (a port, pinned to the reference's own functions by tests/test_ref_pin.py) with OpenMP over hit
    nodes -- the decomposition the reference uses (lib/intern/task_range.cc:89-127) -- on the headline workload, every step a
    bounded sample of it sized so that the K + W steps end within a few minutes."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from dune_sculpt_b200 import build as b
    b.build_oracle()
    from oracle_py import GridOracle, Oracle
    w = build_workload(args.config, args)
    dabs = w.strokes[0]["dabs"]
    cores = os.cpu_count() or 1
    t0 = time.time()
    orc = GridOracle(w.mesh, threads=cores) if w.grids else Oracle(w.mesh, threads=cores, mask=w.mask)
    log("reference arm: oracle PBVH build %.1fs, %d nodes, %d threads" % (time.time() - t0, orc.totnode, cores))
    per = w.group or len(dabs)
    nrad = max(len(dabs) // per, 1)
    orc.stroke_begin(w.strokes[0]["automask"])
    # calibrate: one dab of every group, then size the per-step sample for the time budget
    t0 = time.perf_counter()
    for r_ in range(nrad):
        orc.dab(dabs[r_ * per])
    one = time.perf_counter() - t0
    budget = args.ref_budget_s / max(args.steps + args.warmup, 1)
    k = int(max(1, min(per, budget / max(one, 1e-6))))
    sample = [dabs[r_ * per + j] for r_ in range(nrad) for j in range(k)]
    for _ in range(args.warmup):
        for d in sample:
            orc.dab(d)
    vd0 = orc.vertex_dabs()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for d in sample:
            orc.dab(d)
    dt = time.perf_counter() - t0
    vd = orc.vertex_dabs() - vd0
    orc.stroke_end()
    value = vd / dt
    desc = "%d dabs of each of the %d groups of the stroke per step (%d dabs), %d steps after %d warm-up steps" % (k, nrad, len(sample), args.steps, args.warmup)
    emit({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
          "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
          "data": "synthetic", "config": config_of(w, 1), "ms_per_dab": 1e3 * dt / (args.steps * len(sample)),
          "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


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
    args = make_parser().parse_args()
    run(args)


def make_parser():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c1", "c2", "c3", "c4", "c5"], help="the headline workload of the line (default c3)")
    ap.add_argument("--side-configs", default=None, help="comma list of the other configs measured into `configs` (default: all at N=1, c5 at N>1; 'none')")
    ap.add_argument("--side-steps", type=int, default=3)
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--dabs-per-radius", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify-partition", dest="verify_partition", action="store_false")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="--impl reference: wall-clock budget of all its steps")
    ap.add_argument("--c1-levels", type=int, default=8)
    ap.add_argument("--c2-freq", type=int, default=316)
    ap.add_argument("--c4-grid", type=int, default=2048)
    ap.add_argument("--c4-dabs", type=int, default=50)
    ap.add_argument("--c5-base", type=int, default=25)
    ap.add_argument("--c5-level", type=int, default=7)
    ap.add_argument("--c5-dabs", type=int, default=100, help="draw dabs of the C5 stroke")
    ap.add_argument("--c5-smooth-dabs", type=int, default=100, help="smooth dabs ahead of them")
    return ap


def share_host_threads(world):
    """torchrun starts every rank with OMP_NUM_THREADS=1, which makes the host PBVH build and the upload's table passes serial:
    give each rank its share of the cores instead (session start only; nothing on the timed path runs on host threads)"""
    try:
        import ctypes
        n = max(1, (os.cpu_count() or 1) // max(world, 1))
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
        return n
    except OSError:
        return 0


def run(args):
    if args.impl == "ours":
        args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    from dune_sculpt_b200 import build as b
    b.build_cuda()
    b.build_host()
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
        share_host_threads(world)
    with_cpu = not args.no_cpu_baseline
    res, clocks = measure(args.config, args, rank, world, local_rank, with_cpu)
    if args.side_configs is None:
        side = [c for c in ("c1", "c2", "c4", "c5") if c != args.config] if world == 1 else ([] if args.config == "c5" else ["c5"])
    else:
        side = [c for c in args.side_configs.split(",") if c and c != "none"]
    sides = {}
    for c in side:
        try:
            sides[c], _ = measure(c, args, rank, world, local_rank, with_cpu)
        except Exception as e:  # a side config must not take the headline down with it
            log("side config %s failed: %r" % (c, e))
            sides[c] = {"error": repr(e)}
    if rank == 0:
        out = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": res["steps"], "warmup": res["warmup"],
               "ms_per_step": res["ms_per_step"], "higher_is_better": True,
               # one stroke on one mesh: the work is fixed as GPUs are added
               "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": res["config"],
               "ms_per_dab": res["ms_per_dab"], "vertex_dabs_per_step": res["vertex_dabs_per_step"], "e2e": res["e2e"],
               "gpu_launches": res["gpu_launches"], "clocks": clocks, "session_start_s": res["session_start_s"]}
        for k in ("roofline", "cpu_baseline", "parity_fullsize", "parity_vs_1gpu"):
            if k in res:
                out[k] = res[k]
        if sides:
            out["configs"] = [dict(name=c, **v) for c, v in sides.items()]
            if "c5" in sides:
                out["c5"] = sides["c5"]
        emit(out)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
