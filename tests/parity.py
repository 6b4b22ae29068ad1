"""Shared driver of the parity tests: run the same stroke script through the CPU oracle and through
the CUDA path (via the reference-named host API -> C ABI) and compare dab by dab.

Bars (BASELINE.json north star):
  * node-hit list (order included), moved-vertex set, undo-node membership: bit-exact;
  * positions and normals after the stroke: within 1e-5 of the bounding-box diagonal.
The CUDA path is built to do better than the second bar -- same IEEE operations in the same order,
order-free fixed-point reductions -- so the tests also assert bit-identical positions, normals,
boxes and area normals (`exact=True`).
"""
import numpy as np

from dune_sculpt_b200 import capi
from oracle_py import Oracle

TOL_FRACTION = 1e-5  # of the bounding-box diagonal


def run_parity(mesh, dabs, mask=None, automask=None, leaf_limit=0, exact=True, check_every=1, curve=None,
               pre=None, vert_flag=None):
    diag = mesh.bbox_diag()
    tol = TOL_FRACTION * diag
    orc = Oracle(mesh, mask=mask, leaf_limit=leaf_limit, vert_flag=vert_flag)
    ses = capi.SculptSession(mesh, mask=mask, leaf_limit=leaf_limit, device=0, vert_flag=vert_flag)
    try:
        assert orc.totnode == ses.totnode
        if curve is not None:
            orc.set_custom_curve(curve)
            ses.set_custom_curve(curve)
        if pre is not None:
            pre(orc, ses)
        # initial normals: device-computed vs oracle
        n0_o, n0_g = orc.no(), ses.no()
        assert np.abs(n0_o - n0_g).max() <= tol
        if exact:
            assert np.array_equal(n0_o, n0_g), "initial normals differ in bits"
        ses.capture(True)
        orc.stroke_begin(automask)
        ses.stroke_begin(automask)
        for i, d in enumerate(dabs):
            orc.dab(d)
            ses.dab(d)
            if i % check_every == 0 or i == len(dabs) - 1:
                ho, hg = orc.hits(), ses.hits()
                assert np.array_equal(ho, hg), "dab %d: node-hit list differs (%d vs %d nodes)" % (i, ho.size, hg.size)
                mo, mg = np.sort(orc.moved()), ses.moved()
                assert np.array_equal(mo, mg), "dab %d: moved-vertex set differs (%d vs %d)" % (i, mo.size, mg.size)
                ano, aco = orc.last_area()
                bno, bco = ses.last_area()
                assert np.abs(ano - bno).max() <= 1e-5 and np.abs(aco - bco).max() <= tol, "dab %d: area normal/centre" % i
                if exact:
                    assert np.array_equal(ano, bno) and np.array_equal(aco, bco), "dab %d: area normal bits" % i
        # undo membership while the stroke is still open
        assert np.array_equal(orc.touched(), ses.touched()), "undo-node membership differs"
        orc.stroke_end()
        ses.stroke_end()
        st = ses.stats()
        assert st["vertex_dabs"] == orc.vertex_dabs()
        co_o, co_g = orc.co(), ses.co()
        no_o, no_g = orc.no(), ses.no()
        dco = float(np.abs(co_o - co_g).max())
        dno = float(np.abs(no_o - no_g).max())
        assert dco <= tol, "positions differ by %g (> %g)" % (dco, tol)
        assert dno <= tol, "normals differ by %g (> %g)" % (dno, tol)
        na = orc.node_arrays()
        bb, obb = ses.node_bb()
        assert np.abs(na["vb"] - bb).max() <= tol and np.abs(na["orig_vb"] - obb).max() <= tol
        oo, og = orc.orig_co(), ses.orig_co()
        assert np.abs(oo - og).max() <= tol, "undo snapshot differs"
        assert np.abs(orc.orig_no() - ses.orig_no()).max() <= tol
        flags_g = ses.node_flags()
        keep = capi.PBVH_Leaf | capi.PBVH_UpdateNormals | capi.PBVH_UpdateBB | capi.PBVH_UpdateOriginalBB
        assert np.array_equal(na["flag"] & keep, flags_g & keep), "node update flags differ"
        if exact:
            assert np.array_equal(co_o, co_g), "positions differ in bits (max %g)" % dco
            assert np.array_equal(no_o, no_g), "normals differ in bits (max %g)" % dno
            assert np.array_equal(na["vb"], bb) and np.array_equal(na["orig_vb"], obb), "boxes differ in bits"
            assert np.array_equal(oo, og)
        # the host-side PBVH was refreshed by stroke end (device is a cache of host truth at stroke ends)
        hb = ses.node_arrays()
        assert np.array_equal(hb["vb"], bb) and np.array_equal(hb["orig_vb"], obb)
        return {"max_dco": dco, "max_dno": dno, "vertex_dabs": st["vertex_dabs"], "moved": st["moved_verts"],
                "co": co_g, "no": no_g}
    finally:
        ses.close()
        orc.close()
