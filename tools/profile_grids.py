"""Short driver for ncu on the grids path: a few smooth / draw dabs on a multires cube.
Usage: python tools/profile_grids.py [--base 25] [--level 7] [--dabs 6]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dune_sculpt_b200 import capi, meshgen, stroke  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--base", type=int, default=25)
ap.add_argument("--level", type=int, default=7)
ap.add_argument("--dabs", type=int, default=6)
ap.add_argument("--smooth", type=int, default=0, help="smooth dabs ahead of the draw dabs")
a = ap.parse_args()
mr = meshgen.multires_cube_n(a.base, a.level)
ses = capi.GridSession(mr, device=0)
rng = np.random.default_rng(5)
bs = stroke._strength(capi.TOOL_DRAW, 0.5)
ses.stroke_begin()
for i in range(a.smooth):
    p = rng.normal(size=3)
    p /= np.linalg.norm(p)
    ses.dab(capi.make_dab(capi.TOOL_SMOOTH, p.astype(np.float32), mr.bbox_diag() * 0.08,
                          bstrength=stroke._strength(capi.TOOL_SMOOTH, 0.75), view_normal=tuple(p)))
for i in range(a.dabs):
    p = rng.normal(size=3)
    p /= np.linalg.norm(p)
    ses.dab(capi.make_dab(capi.TOOL_DRAW, p.astype(np.float32), mr.bbox_diag() * 0.08, bstrength=bs, view_normal=tuple(p)))
ses.stroke_end()
print("stats", ses.stats())
ses.close()
