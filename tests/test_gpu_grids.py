"""Multires grids on the device vs the CPU oracle: BKE_pbvh_build_grids -> dab (gather, brush, stitch,
CCG normals, bounds) through the host API and the C ABI, bit for bit."""
import numpy as np
import pytest

from dune_sculpt_b200 import capi, meshgen, stroke
from oracle_py import GridOracle

pytestmark = pytest.mark.gpu


def _grid_parity(mr, dabs, leaf_limit=0, automask=None, hidden=None):
    orc = GridOracle(mr, leaf_limit=leaf_limit, hidden=hidden)
    ses = capi.GridSession(mr, leaf_limit=leaf_limit, device=0, hidden=hidden)
    try:
        assert orc.totnode == ses.totnode
        assert np.array_equal(orc.no(), ses.no()), "initial CCG normals differ in bits"
        assert np.array_equal(orc.co(), ses.co())
        orc.stroke_begin(automask)
        ses.stroke_begin(automask)
        ses.capture(True)
        for i, d in enumerate(dabs):
            orc.dab(d)
            ses.dab(d)
            assert np.array_equal(orc.hits(), ses.hits()), "dab %d: node-hit list" % i
            assert np.array_equal(np.sort(orc.moved()), ses.moved()), "dab %d: moved elements" % i
            ano, aco = orc.last_area()
            bno, bco = ses.last_area()
            assert np.array_equal(ano, bno) and np.array_equal(aco, bco), "dab %d: area normal" % i
        assert np.array_equal(orc.touched(), ses.touched())
        orc.stroke_end()
        ses.stroke_end()
        assert ses.stats()["vertex_dabs"] == orc.vertex_dabs()
        assert np.array_equal(orc.co(), ses.co()), "positions differ in bits"
        assert np.array_equal(orc.no(), ses.no()), "normals differ in bits"
        if mr.mask is not None:
            assert np.array_equal(orc.mask(), ses.mask()), "mask layer differs in bits"
        na = orc.node_arrays()
        bb, obb = ses.node_bb()
        assert np.array_equal(na["vb"], bb) and np.array_equal(na["orig_vb"], obb), "boxes differ in bits"
        assert np.array_equal(orc.orig_co(), ses.orig_co()) and np.array_equal(orc.orig_no(), ses.orig_no())
        keep = capi.PBVH_Leaf | capi.PBVH_UpdateNormals | capi.PBVH_UpdateBB | capi.PBVH_UpdateOriginalBB
        assert np.array_equal(na["flag"] & keep, ses.node_flags() & keep)
        # the host's CCGElem storage was refreshed by stroke end
        hco, hno, hmask = ses.host_elements()
        assert np.array_equal(hco, orc.co()) and np.array_equal(hno, orc.no())
        if mr.mask is not None:
            assert np.array_equal(hmask, orc.mask())
        # multires write-back (multires_reshape_ccg.c:10-70): straight from the device into the MDisps / GridPaintMask arrays
        disps, masks = ses.multires_write_back()
        assert np.array_equal(disps.reshape(-1, 3), orc.co())
        if masks is not None:
            assert np.array_equal(masks.reshape(-1), orc.mask())
        if hidden is not None:
            # (elements on a grid's rim are stitched with their duplicates in other grids whether hidden or not)
            gs = mr.grid_size
            yy, xx = np.divmod(np.arange(gs * gs), gs)
            inner = np.tile((xx > 0) & (xx < gs - 1) & (yy > 0) & (yy < gs - 1), mr.totgrid)
            hid = (np.asarray(hidden).reshape(-1) != 0) & inner
            assert np.array_equal(ses.co()[hid], np.asarray(mr.co).reshape(-1, 3)[hid]), "a hidden element moved"
        return ses.stats()
    finally:
        ses.close()
        orc.close()


def _sweep(mr, per=2, tool=None, radii=(4.0, 10.0, 25.0, 45.0), seed=3):
    """dabs at seeded points of the unit sphere, radius in % of the diagonal"""
    tool = capi.TOOL_DRAW if tool is None else tool
    rng = np.random.default_rng(seed)
    diag = mr.bbox_diag()
    bs = stroke._strength(tool, 0.5)
    out = []
    for pct in radii:
        for _ in range(per):
            p = rng.normal(size=3)
            p /= np.linalg.norm(p)
            kw = dict(bstrength=bs, view_normal=tuple(p), flags=capi.DAB_FIRST_STEP if not out else 0)
            if tool in (capi.TOOL_GRAB, capi.TOOL_CLAY_STRIPS):
                kw["grab_delta"] = tuple(0.05 * rng.normal(size=3))
            out.append(capi.make_dab(tool, p.astype(np.float32), diag * pct / 100.0, **kw))
    return out


def test_grids_draw_radius_sweep():
    mr = meshgen.multires_cube(2, 4)           # 96 faces, 384 grids of 9 x 9
    st = _grid_parity(mr, _sweep(mr), leaf_limit=6)
    assert st["moved_verts"] > 0


def test_grids_with_mask_single_grid_leaves_level5():
    mr = meshgen.multires_cube(1, 5, with_mask=True)   # 96 grids of 17 x 17, one grid per leaf
    st = _grid_parity(mr, _sweep(mr, per=2, radii=(6.0, 20.0, 50.0)), leaf_limit=1)
    assert st["moved_verts"] > 0


@pytest.mark.parametrize("tool", [capi.TOOL_INFLATE, capi.TOOL_GRAB, capi.TOOL_CLAY_STRIPS])
def test_grids_other_tools(tool):
    mr = meshgen.multires_cube(1, 4)
    _grid_parity(mr, _sweep(mr, per=2, tool=tool, radii=(15.0, 35.0)), leaf_limit=4)


def test_grids_default_leaf_limit_c5_shape_reduced():
    """C5's shape at 1/100 of its size: 6 x 5 x 5 base quads, level 6 (33 x 33 per grid, 653,400
    elements), default leaf limit (10000 / 1089 = 9 grids per leaf)"""
    mr = meshgen.multires_cube_n(5, 6)
    assert mr.totelem == 600 * 33 * 33
    st = _grid_parity(mr, _sweep(mr, per=2, radii=(8.0, 30.0)))
    assert st["moved_verts"] > 0


def test_grids_reject_what_the_path_does_not_cover():
    mr = meshgen.multires_cube(1, 3)
    mr.edge_verts = None   # no coarse topology -> no rim neighbour table -> no smooth brush
    ses = capi.GridSession(mr, device=0)
    try:
        ses.stroke_begin()
        with pytest.raises(capi.DeviceError):
            ses.dab(capi.make_dab(capi.TOOL_SMOOTH, (0, 0, 1), 0.5, bstrength=0.5))
        with pytest.raises(capi.DeviceError):
            ses.dab(capi.make_dab(capi.TOOL_DRAW, (0, 0, 1), 0.5, bstrength=0.5, flags=capi.DAB_NO_NORMALS))
        ses.stroke_end()
    finally:
        ses.close()


def test_grids_fused_cooperative_kernel_is_bit_identical(monkeypatch):
    """DSC_GRID_FUSED=1: the nine launches after the brush as one cooperative kernel with grid barriers
    (an experiment that measured slower; kept honest here)"""
    monkeypatch.setenv("DSC_GRID_FUSED", "1")
    mr = meshgen.multires_cube(2, 4, with_mask=True)
    st = _grid_parity(mr, _sweep(mr, per=2, radii=(6.0, 20.0, 45.0)), leaf_limit=6)
    assert st["moved_verts"] > 0


def _smooth_dabs(mr, n=4, pct=(12.0, 30.0), alpha=0.75, seed=5, around=None):
    rng = np.random.default_rng(seed)
    diag = mr.bbox_diag()
    out = []
    for r in pct:
        for _ in range(n):
            if around is None:
                p = rng.normal(size=3)
                p /= np.linalg.norm(p)
            else:
                p = np.asarray(around, dtype=np.float64) + 0.3 * rng.normal(size=3) * np.array([1.0, 1.0, 0.0])
            out.append(capi.make_dab(capi.TOOL_SMOOTH, p.astype(np.float32), diag * r / 100.0,
                                     bstrength=stroke._strength(capi.TOOL_SMOOTH, alpha)))
    return out


@pytest.mark.parametrize("alpha", [0.75, 0.5, 0.2])
def test_grids_smooth_brush(alpha):
    """smooth on grids: neighbours from the element's place in its grid or from the rim table
    (KERNEL_subdiv_ccg_neighbor_coords_get, subdiv_ccg.c:1882-1909); alpha 0.75 = 3 full iterations,
    0.5 = 2, 0.2 = a partial one only"""
    mr = meshgen.multires_cube(2, 4, noise=0.03, freq=17.0, with_mask=True)
    st = _grid_parity(mr, _smooth_dabs(mr, n=2, alpha=alpha), leaf_limit=6)
    assert st["moved_verts"] > 0


def test_grids_smooth_open_base_boundary_elements():
    """open base mesh: coarse boundary edges / vertices -> boundary elements average boundary neighbours only,
    corner elements of the sheet (two neighbours) stay"""
    mr = meshgen.multires_plane(4, 4, noise=0.05, freq=11.0)
    dabs = _smooth_dabs(mr, n=3, pct=(15.0, 45.0), around=(0.0, 0.0, 0.0))
    dabs.append(capi.make_dab(capi.TOOL_SMOOTH, mr.co[np.argmax(mr.co[:, 0] + mr.co[:, 1])], mr.bbox_diag() * 0.2, bstrength=0.75))
    st = _grid_parity(mr, dabs, leaf_limit=3)
    assert st["moved_verts"] > 0


def test_grids_smooth_then_draw_c5_shape_reduced():
    """C5's stroke pair (smooth, then draw) on its shape at 1/100 of the size, default leaf limit"""
    mr = meshgen.multires_cube_n(5, 6, noise=0.02, freq=23.0)
    dabs = _smooth_dabs(mr, n=2, pct=(8.0,)) + _sweep(mr, per=2, radii=(8.0,))
    st = _grid_parity(mr, dabs)
    assert st["moved_verts"] > 0


def test_grids_batched_dabs_through_cuda_graphs():
    """dsc_dabs on grids: runs of one launch sequence (smooth with its iteration count, draw) replayed as CUDA graphs;
    the face / edge / vertex stamps of a dab come from its ring position on the device.  Two strokes: the second
    replays the cached graphs on stamps the first left behind."""
    mr = meshgen.multires_cube(2, 4, noise=0.03, freq=17.0, with_mask=True)
    dabs = _smooth_dabs(mr, n=9, pct=(10.0,)) + _sweep(mr, per=11, radii=(9.0, 22.0))   # 9 smooth, 22 draw
    orc = GridOracle(mr, leaf_limit=6)
    ses = capi.GridSession(mr, leaf_limit=6, device=0)
    try:
        arr = (capi.DscDab * len(dabs))(*dabs)
        for stroke_no in range(2):
            orc.stroke_begin(None)
            for d in dabs:
                orc.dab(d)
            orc.stroke_end()
            ses.stroke_begin(None)
            ses.dabs(arr, len(dabs))
            st = ses.stats()
            ses.stroke_end()
            assert st["dabs"] == len(dabs)
            assert np.array_equal(orc.co(), ses.co()), "stroke %d: positions differ in bits" % stroke_no
            assert np.array_equal(orc.no(), ses.no()), "stroke %d: normals differ in bits" % stroke_no
            assert np.array_equal(orc.mask(), ses.mask())
            na = orc.node_arrays()
            bb, obb = ses.node_bb()
            assert np.array_equal(na["vb"], bb) and np.array_equal(na["orig_vb"], obb)
        assert ses.stats()["kernel_launches"] > 0
    finally:
        ses.close()
        orc.close()


def test_grids_element_parallel_normal_pass_is_bit_identical(monkeypatch):
    """DSC_GRID_NORMALS_FLAT=1: one thread per element instead of one CTA per grid (an experiment that measured slower)"""
    monkeypatch.setenv("DSC_GRID_NORMALS_FLAT", "1")
    mr = meshgen.multires_cube(2, 4, with_mask=True)
    st = _grid_parity(mr, _sweep(mr, per=2, radii=(6.0, 20.0, 45.0)), leaf_limit=6)
    assert st["moved_verts"] > 0


@pytest.mark.parametrize("smooth", [True, False])
def test_grids_draw_buffers_from_the_device(smooth):
    """gpu_pbvh_grid_buffers_update (gpu_buffers.c:548-725) on the device: after a stroke the flagged leaves' vertex
    records, smooth (per element) or flat (four per quad), byte for byte.  Like pbvh_update_draw_buffers
    (pbvh.c:3169-3285) only the flagged leaves are refilled: an unflagged leaf keeps the records of its last
    fill even when the stitch moved some of its rim elements (the reference's buffers are stale in the same way)."""
    mr = meshgen.multires_cube(1, 4, with_mask=True)
    orc = GridOracle(mr, leaf_limit=3)
    ses = capi.GridSession(mr, leaf_limit=3, device=0, draw_buffers=True)
    try:
        na = orc.node_arrays()
        leaves = np.nonzero(na["flag"] & 1)[0]
        ses.update_draw_buffers(smooth=smooth, show_mask=True)     # every leaf starts flagged (build_grid_leaf_node)
        for n in leaves:
            assert np.array_equal(orc.draw_buffer(int(n), int(na["totprim"][n]), smooth=smooth), ses.draw_buffer(int(n))), n
        orc.stroke_begin(None)
        ses.stroke_begin(None)
        for d in _sweep(mr, per=1, radii=(12.0, 30.0)):
            orc.dab(d)
            ses.dab(d)
        orc.stroke_end()
        ses.stroke_end()
        flagged = set(int(n) for n in np.nonzero(orc.node_arrays()["flag"] & capi.PBVH_UpdateDrawBuffers)[0])
        assert 0 < len(flagged) < leaves.size
        before = {int(n): ses.draw_buffer(int(n)) for n in leaves}
        ses.update_draw_buffers(smooth=smooth, show_mask=True)
        for n in leaves:
            want = orc.draw_buffer(int(n), int(na["totprim"][n]), smooth=smooth) if int(n) in flagged else before[int(n)]
            assert np.array_equal(want, ses.draw_buffer(int(n))), n
    finally:
        ses.close()
        orc.close()


# ---- rows a10 / a11 / a19 on grids: grid_hidden, tube falloff, clipping, grab normal weight -------------------------

def _hidden_elems(mr, seed=11, frac=0.25):
    rng = np.random.default_rng(seed)
    h = (rng.random(mr.totelem) < frac).astype(np.uint8).reshape(mr.totgrid, -1)
    h[::7] = 1   # whole grids too: their quads all vanish
    return h.reshape(-1)


@pytest.mark.parametrize("tool", [capi.TOOL_DRAW, capi.TOOL_SMOOTH, capi.TOOL_GRAB, capi.TOOL_CLAY_STRIPS])
def test_grids_hidden_elements_are_skipped(tool):
    mr = meshgen.multires_cube(2, 4)
    hid = _hidden_elems(mr)
    dabs = _smooth_dabs(mr) if tool == capi.TOOL_SMOOTH else _sweep(mr, per=2, tool=tool, radii=(10.0, 30.0))
    st = _grid_parity(mr, dabs, leaf_limit=6, hidden=hid)
    assert st["moved_verts"] > 0


@pytest.mark.parametrize("tool", [capi.TOOL_DRAW, capi.TOOL_SMOOTH, capi.TOOL_INFLATE])
def test_grids_tube_falloff_and_clipping(tool):
    mr = meshgen.multires_cube(2, 4)
    base = _smooth_dabs(mr) if tool == capi.TOOL_SMOOTH else _sweep(mr, per=2, tool=tool, radii=(8.0, 20.0))
    for d in base:
        d.falloff_shape = capi.FALLOFF_TUBE
        d.clip_flags = capi.CLIP_X | capi.LOCK_Z
        d.clip_tolerance[0] = 0.05
        if tool == capi.TOOL_SMOOTH:
            d.view_normal[:] = [0.0, 0.6, 0.8]
    st = _grid_parity(mr, base, leaf_limit=6, hidden=_hidden_elems(mr, seed=2, frac=0.1))
    assert st["moved_verts"] > 0


def test_grids_grab_normal_weight():
    mr = meshgen.multires_cube(1, 4)
    dabs = _sweep(mr, per=2, tool=capi.TOOL_GRAB, radii=(15.0, 35.0))
    for d in dabs:
        d.normal_weight = 0.6
    _grid_parity(mr, dabs, leaf_limit=4)


def test_grids_draw_buffers_take_their_shading_per_leaf_from_grid_flag_mats():
    """gpu_buffers.c:574: smooth or flat by ME_SMOOTH of the leaf's first grid; flat leaves hold 4 (gs - 1)^2 records per grid,
    smooth ones gs^2, so the leaves' runs in the buffer have different lengths"""
    mr = meshgen.multires_cube(2, 3, with_mask=True)
    grid_flag = np.zeros(mr.totgrid, np.uint8)
    grid_flag[mr.totgrid // 3:] = 1
    grid_mat = np.zeros(mr.totgrid, np.int16)
    orc = GridOracle(mr, leaf_limit=4, grid_mat=grid_mat, grid_flag=grid_flag)
    ses = capi.GridSession(mr, leaf_limit=4, device=0, draw_buffers=True, grid_mat=grid_mat, grid_flag=grid_flag)
    try:
        na = orc.node_arrays()
        prim = orc.prim_indices()
        leaves = [int(n) for n in np.nonzero(na["flag"] & 1)[0]]
        smooth_of = {n: bool(grid_flag[prim[na["prim_offset"][n]]] & 1) for n in leaves}
        assert any(smooth_of.values()) and not all(smooth_of.values())
        ses.update_draw_buffers(smooth=-1, show_mask=True)
        lens = set()
        for n in leaves:
            ref = orc.draw_buffer(n, int(na["totprim"][n]), smooth=smooth_of[n])
            got = ses.draw_buffer(n)
            assert got.shape == ref.shape and np.array_equal(got, ref), n
            lens.add(got.shape[0] // int(na["totprim"][n]))
        gs = mr.grid_size
        assert lens == {gs * gs, 4 * (gs - 1) * (gs - 1)}
    finally:
        ses.close()
        orc.close()


@pytest.mark.parametrize("fused", [False, True])
def test_grids_dab_that_gathers_nothing_does_not_stitch(fused, monkeypatch):
    """the stroke step returns before multires_stitch_grids when the gather is empty: no all-coarse-vertex averaging for that dab
    (a mean of equal floats is not always that float, so running it anyway shows up in the bits)"""
    if fused:
        monkeypatch.setenv("DSC_GRID_FUSED", "1")
    mr = meshgen.multires_cube(2, 4)
    dabs = _sweep(mr, per=2, radii=(10.0, 25.0))
    far = [capi.make_dab(capi.TOOL_DRAW, (9.0, 9.0, 9.0), 0.05, bstrength=0.3) for _ in range(3)]
    dabs = [dabs[0], far[0], far[1], dabs[1], dabs[2], far[2], dabs[3]]
    st = _grid_parity(mr, dabs, leaf_limit=6)
    assert st["moved_verts"] > 0


# ---- cursor pick on grids: pbvh_grids_node_raycast -------------------------------------------

def _grid_rays(mr, rng, count):
    co = np.asarray(mr.co, np.float32).reshape(-1, 3)
    lo, hi = co.min(axis=0), co.max(axis=0)
    ctr, ext = 0.5 * (lo + hi), float(np.linalg.norm(hi - lo))
    gs = mr.grid_size
    rays = []
    for i in range(count):
        kind = i % 5
        if kind == 0:
            target = lo + rng.random(3).astype(np.float32) * (hi - lo)
        elif kind == 1:       # exactly at an element: the quads (and grids, and leaves) around it tie
            target = co[rng.integers(mr.totelem)]
        elif kind == 2:       # on a quad's edge or its diagonal: both triangles of a quad answer
            g, y, x = rng.integers(mr.totgrid), rng.integers(gs - 1), rng.integers(gs - 1)
            a = co[g * gs * gs + y * gs + x]
            b = co[g * gs * gs + (y + 1) * gs + x + 1] if i % 2 else co[g * gs * gs + y * gs + x + 1]
            target = (0.5 * (a + b)).astype(np.float32)
        elif kind == 3:
            target = co[rng.integers(mr.totelem)]
        else:
            target = ctr + (hi - lo) * 3.0 * np.sign(rng.normal(size=3)).astype(np.float32)  # a miss
        if kind == 3:
            d = np.zeros(3, np.float32)
            d[rng.integers(3)] = 1.0 if rng.random() < 0.5 else -1.0
        else:
            d = rng.normal(size=3).astype(np.float32)
            d /= np.float32(np.linalg.norm(d))
        start = (np.asarray(target, np.float32) - d * np.float32(2.0 * ext)).astype(np.float32)
        rays.append((start, d.astype(np.float32)))
    return rays


def _same_hit(ref, got, what):
    assert (ref is None) == (got is None), (what, ref, got)
    if ref is None:
        return 0
    assert np.float32(ref["depth"]).tobytes() == np.float32(got["depth"]).tobytes(), (what, ref, got)
    assert ref["face"] == got["face"] and ref["vertex"] == got["vertex"] and ref["node"] == got["node"], (what, ref, got)
    assert np.array_equal(ref["normal"].view(np.uint32), got["normal"].view(np.uint32)), what
    return 1


@pytest.mark.parametrize("hidden", [False, True])
def test_grids_raycast_matches_pbvh_grids_node_raycast(hidden):
    """depth, active grid, nearest element, quad normal and leaf of the nearest hit, bit-equal to BKE_pbvh_raycast over
    pbvh_grids_node_raycast (pbvh.c:4102-4200): a quad is two triangles of which the second is only looked at when the first
    is not a nearer hit, so the answer is a fold over the touched quads in (leaf by entry distance, grid, y, x) order --
    the device lists the touched quads, the host folds.  Before a stroke (flat faces: rays through vertices, edges and
    diagonals tie), mid-stroke with original / current coordinates (bumpy, non-planar quads), and after it."""
    mr = meshgen.multires_cube(2, 4)
    hid = _hidden_elems(mr, seed=4, frac=0.05) if hidden else None
    orc = GridOracle(mr, leaf_limit=5, hidden=hid)
    ses = capi.GridSession(mr, leaf_limit=5, device=0, hidden=hid, raycast=True)
    try:
        rng = np.random.default_rng(23)
        rays = _grid_rays(mr, rng, 150)
        hits = 0
        for i, (s, d) in enumerate(rays):
            hits += _same_hit(orc.raycast(s, d), ses.raycast(s, d), "ray %d" % i)
        assert 60 < hits < len(rays)
        s, d = rays[0]
        full = orc.raycast(s, d)
        if full is not None:
            for md in (float(full["depth"]) * 0.5, float(full["depth"]), float(full["depth"]) * 1.5):
                _same_hit(orc.raycast(s, d, max_depth=md), ses.raycast(s, d, max_depth=md), "max_depth %g" % md)
        orc.stroke_begin(None)
        ses.stroke_begin(None)
        for dab in _sweep(mr, per=2, radii=(12.0, 30.0)) + _sweep(mr, per=1, tool=capi.TOOL_INFLATE, radii=(20.0,), seed=9):
            dab.flags &= ~capi.DAB_FIRST_STEP
            orc.dab(dab)
            ses.dab(dab)
        differ = 0
        for i, (s, d) in enumerate(rays):
            a = orc.raycast(s, d, original=True)
            b = orc.raycast(s, d, original=False)
            _same_hit(a, ses.raycast(s, d, original=True), "mid-stroke original ray %d" % i)
            _same_hit(b, ses.raycast(s, d, original=False), "mid-stroke current ray %d" % i)
            differ += (a is not None and b is not None and a["depth"] != b["depth"])
        assert differ > 0
        orc.stroke_end()
        ses.stroke_end()
        for i, (s, d) in enumerate(rays[:60]):
            _same_hit(orc.raycast(s, d), ses.raycast(s, d), "after stroke ray %d" % i)
    finally:
        ses.close()
        orc.close()
