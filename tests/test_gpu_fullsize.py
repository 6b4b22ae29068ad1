"""Full-size parity of the BASELINE.json configurations that fit a test run: the same strokes bench.py times, through the device
and through the CPU oracle (threaded, normals summed in the serial loop's order), compared bit for bit.  C3 / C5 at full size are
checked inside bench.py itself (`parity_fullsize` of the bench line); here C1, C2 and C4 -- the smooth brush on the 1M-vertex
icosphere and every tool x {mask off, mask + boundary automask} + the topology automask on the 4.2M-vertex grid."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _fullsize(name, extra=()):
    import bench
    args = bench.make_parser().parse_args(list(extra))
    w = bench.build_workload(name, args)
    r = bench.Runner(w, args, 0, 1, 0)
    try:
        dev = r.parity_device()
    finally:
        r.close()
    cpu, par = bench.cpu_leg(w, args, dev)
    assert len(par["strokes"]) == len(w.strokes)
    for rec in par["strokes"]:
        assert rec["hit_lists_equal"] and rec["touched_equal"] and rec["vertex_dabs_equal"], rec
        assert rec["within_tolerance"], rec
        assert rec["bit_exact"], rec
    return par


def test_c1_full_size():
    _fullsize("c1")


def test_c2_smooth_icosphere_full_size():
    par = _fullsize("c2")
    assert par["strokes"][0]["dabs"] == 200


def test_c4_every_tool_mask_and_automask_full_size():
    # 20 dabs per stroke instead of the bench's 50 keeps the nine CPU strokes inside a test budget; the mesh is the full 2048^2 grid
    par = _fullsize("c4", ["--c4-dabs", "20"])
    labels = [r["stroke"] for r in par["strokes"]]
    assert len(labels) == 9 and any("topology" in x for x in labels) and any("grab" in x for x in labels)
