"""Short driver for ncu: one warm-up stroke and one measured stroke of the bench workload
(no CPU baseline, no end-to-end leg).  Usage: python tools/profile_stroke.py [--grid N] [--per-radius K]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dune_sculpt_b200 import capi, meshgen, stroke  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=4096)
ap.add_argument("--per-radius", type=int, default=4)
ap.add_argument("--strokes", type=int, default=2)
ap.add_argument("--tool", default="draw")
a = ap.parse_args()
mesh = meshgen.grid(a.grid)
dabs = stroke.c3_radius_sweep(mesh.bbox_diag(), dabs_per_radius=a.per_radius)
ses = capi.SculptSession(mesh, device=0)
for _ in range(a.strokes):
    ses._chk(ses.D.dsc_stroke_begin(ses.ctx, None))
    for d in dabs:
        ses._chk(ses.D.dsc_dab(ses.ctx, C.byref(d)))
    ses._chk(ses.D.dsc_stroke_end(ses.ctx))
ses.synchronize()
print("stats", ses.stats())
ses.close()
