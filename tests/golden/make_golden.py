"""Generates tests/golden/oracle_strokes.json: digests of small scripted strokes run through the CPU
oracle.  These pin the ORACLE against accidental change; they are not reference outputs -- the
reference has no golden vectors, fixtures or known-answer tests for this path.

    python tests/golden/make_golden.py          # rewrite the fixture
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from dune_sculpt_b200 import capi, meshgen, stroke  # noqa: E402
from oracle_py import GridOracle, Oracle  # noqa: E402


def _digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def compute():
    out = {}
    cases = {
        "cube5_draw": (meshgen.cube(5), 150, lambda m: stroke.c1_draw_stroke(dabs=10, radius=0.3)),
        "ico20_smooth": (meshgen.icosphere(20, noise=0.004), 100, lambda m: stroke.c2_smooth_stroke(dabs=6, radius=0.45)),
        "grid97_inflate": (meshgen.grid(97), 200, lambda m: stroke.c4_tool_stroke(capi.TOOL_INFLATE, m.bbox_diag(), dabs=6)),
        "grid97_clay": (meshgen.grid(97), 200, lambda m: stroke.c4_tool_stroke(capi.TOOL_CLAY_STRIPS, m.bbox_diag(), dabs=6)),
        "grid97_grab": (meshgen.grid(97), 200, lambda m: stroke.c4_tool_stroke(capi.TOOL_GRAB, m.bbox_diag(), dabs=6)),
    }
    for name, (mesh, ll, mk) in cases.items():
        o = Oracle(mesh, leaf_limit=ll)
        na = o.node_arrays()
        rec = {"totnode": int(o.totnode), "tree": _digest(na["children_offset"], na["prim_offset"], na["totprim"],
                                                         na["uniq_verts"], na["face_verts"], o.prim_indices())}
        o.stroke_begin()
        hits, moved = [], []
        for d in mk(mesh):
            o.dab(d)
            hits.append(o.hits())
            moved.append(np.sort(o.moved()))
        o.stroke_end()
        rec["hits"] = _digest(*hits)
        rec["moved"] = _digest(*moved)
        rec["co"] = _digest(o.co())
        rec["no"] = _digest(o.no())
        rec["vb"] = _digest(o.node_arrays()["vb"] + np.float32(0.0))  # +0 folds -0 into +0
        rec["touched"] = _digest(o.touched())
        rec["vertex_dabs"] = int(o.vertex_dabs())
        out[name] = rec
        o.close()
    # multires grids: the element-neighbour lists (closed cube and open sheet), a smooth + draw stroke with the mask layer
    for name, mr in (("mr_cube_neighbours", meshgen.multires_cube(1, 3)), ("mr_plane_neighbours", meshgen.multires_plane(3, 3))):
        o = GridOracle(mr, leaf_limit=4)
        nb = [o.neighbors(e) for e in range(mr.totelem)]
        out[name] = {"counts": _digest(np.array([len(x) for x in nb], dtype=np.int32)), "lists": _digest(*nb),
                     "boundary": _digest(np.array([o.is_boundary(e) for e in range(mr.totelem)], dtype=np.uint8))}
        o.close()
    mr = meshgen.multires_cube(2, 4, noise=0.03, freq=17.0, with_mask=True)
    o = GridOracle(mr, leaf_limit=6)
    o.stroke_begin(None)
    rng = np.random.default_rng(5)
    hits = []
    for i in range(6):
        p = rng.normal(size=3)
        p = (p / np.linalg.norm(p)).astype(np.float32)
        tool = capi.TOOL_SMOOTH if i < 3 else capi.TOOL_DRAW
        o.dab(capi.make_dab(tool, p, mr.bbox_diag() * 0.15, bstrength=stroke._strength(tool, 0.6), view_normal=tuple(p)))
        hits.append(o.hits())
    o.stroke_end()
    out["mr_cube_smooth_draw"] = {"hits": _digest(*hits), "co": _digest(o.co()), "no": _digest(o.no() + np.float32(0.0)),
                                  "mask": _digest(o.mask()), "vb": _digest(o.node_arrays()["vb"] + np.float32(0.0)),
                                  "vertex_dabs": int(o.vertex_dabs())}
    o.close()
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "oracle_strokes.json"), "w") as f:
        json.dump(compute(), f, indent=1, sort_keys=True)
    print("wrote oracle_strokes.json")
