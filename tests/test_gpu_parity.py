"""CUDA path vs CPU oracle on identical meshes and stroke scripts (the parity tests proper).
Every test goes through the reference-named host API and the C ABI; nothing here touches torch."""
import ctypes as C

import numpy as np
import pytest

from dune_sculpt_b200 import capi, meshgen, stroke
from parity import run_parity

pytestmark = pytest.mark.gpu


def _line_dabs(tool, p0, p1, radius, n, **kw):
    pts = stroke.line_points(p0, p1, radius, count=n)
    bs = stroke._strength(tool, kw.pop("alpha", 0.6), invert=kw.pop("invert", False))
    out = []
    prev = None
    for i, p in enumerate(pts):
        k = dict(kw)
        k.setdefault("view_normal", (0, 0, 1))
        k["flags"] = k.get("flags", 0) | (capi.DAB_FIRST_STEP if i == 0 else 0)
        if tool == capi.TOOL_CLAY_STRIPS:
            k["grab_delta"] = (0, 0, 0) if prev is None else (p - prev)
        out.append(capi.make_dab(tool, p, radius, bstrength=bs, **k))
        prev = p
    return out


@pytest.mark.parametrize("preset", [capi.CURVE_SMOOTH, capi.CURVE_SPHERE, capi.CURVE_ROOT, capi.CURVE_SHARP,
                                    capi.CURVE_LIN, capi.CURVE_POW4, capi.CURVE_INVSQUARE, capi.CURVE_CONSTANT,
                                    capi.CURVE_SMOOTHER])
def test_draw_presets_small_grid(preset):
    m = meshgen.grid(129)
    dabs = _line_dabs(capi.TOOL_DRAW, (-0.6, -0.3, 0.0), (0.6, 0.4, 0.0), 0.3, 12, curve_preset=preset)
    r = run_parity(m, dabs, leaf_limit=300)
    assert r["moved"] > 0


def test_draw_custom_curve_and_hardness():
    m = meshgen.grid(129)
    t = np.linspace(0.0, 1.0, 257, dtype=np.float32)
    table = (1.0 - t) ** 1.7
    dabs = _line_dabs(capi.TOOL_DRAW, (-0.6, -0.3, 0.0), (0.6, 0.4, 0.0), 0.3, 10, curve_preset=capi.CURVE_CUSTOM,
                      hardness=0.35)
    run_parity(m, dabs, leaf_limit=300, curve=table)


@pytest.mark.parametrize("plane", [capi.DIR_AREA, capi.DIR_VIEW, capi.DIR_X, capi.DIR_Z])
def test_draw_sculpt_plane(plane):
    m = meshgen.cube(5)
    dabs = _line_dabs(capi.TOOL_DRAW, (-0.7, 0.1, 1.0), (0.7, -0.2, 1.0), 0.35, 10, sculpt_plane=plane,
                      view_normal=(0.1, 0.2, 0.97))
    run_parity(m, dabs, leaf_limit=150)


def test_draw_frontface_mask_automask_invert():
    m = meshgen.icosphere(24)
    mask = meshgen.low_freq_mask(m, seed=3)
    rng = np.random.default_rng(5)
    automask = rng.uniform(0.0, 1.0, size=m.totvert).astype(np.float32)
    dabs = _line_dabs(capi.TOOL_DRAW, (0.0, -0.5, 0.86), (0.3, 0.5, 0.81), 0.4, 10, flags=capi.DAB_FRONTFACE,
                      view_normal=(0.0, 0.0, 1.0), invert=True)
    run_parity(m, dabs, mask=mask, automask=automask, leaf_limit=120)


def test_inflate_icosphere():
    m = meshgen.icosphere(20, noise=0.002)
    dabs = _line_dabs(capi.TOOL_INFLATE, (0.0, 0.0, 1.0), (0.8, 0.0, 0.6), 0.35, 12)
    run_parity(m, dabs, leaf_limit=100)


def test_inflate_closed_form():
    # sphere + inflate with CONSTANT falloff: |co| grows by fade * r * bstrength along the (radial) normal
    m = meshgen.icosphere(16)
    d = capi.make_dab(capi.TOOL_INFLATE, (0, 0, 1), 0.5, curve_preset=capi.CURVE_CONSTANT, bstrength=0.2)
    r = run_parity(m, [d], leaf_limit=100)
    before = np.linalg.norm(m.co.astype(np.float64), axis=1)
    after = np.linalg.norm(r["co"].astype(np.float64), axis=1)
    inside = np.linalg.norm(m.co - np.array([0, 0, 1], np.float32), axis=1) <= 0.5
    assert np.allclose((after - before)[inside], 0.2 * 0.5, atol=2e-3)  # vertex normal ~ radial on a coarse sphere
    assert np.array_equal(r["co"][~inside], m.co[~inside])


def test_draw_closed_form_flat_grid():
    # flat grid + draw along +Z with CONSTANT falloff => exact dz = radius * bstrength
    m = meshgen.grid(65, height=0.0)
    d = capi.make_dab(capi.TOOL_DRAW, (0, 0, 0), 0.4, curve_preset=capi.CURVE_CONSTANT, sculpt_plane=capi.DIR_Z,
                      bstrength=0.25)
    r = run_parity(m, [d], leaf_limit=100)
    inside = (m.co[:, 0] ** 2 + m.co[:, 1] ** 2) <= np.float32(0.4) ** 2
    assert np.array_equal(r["co"][inside, 2], np.full(inside.sum(), np.float32(0.4) * np.float32(0.25), np.float32))
    assert np.array_equal(r["co"][~inside], m.co[~inside])


def test_grab_uses_original_boxes_and_coords():
    m = meshgen.grid(97)
    loc = (0.1, -0.2, 0.0)
    bs = stroke._strength(capi.TOOL_GRAB, 0.8)
    dabs = [capi.make_dab(capi.TOOL_GRAB, loc, 0.35, bstrength=bs, grab_delta=np.array([0.2, 0.1, 0.5]) * (i + 1) / 10,
                          flags=capi.DAB_FIRST_STEP if i == 0 else 0) for i in range(10)]
    run_parity(m, dabs, leaf_limit=200)


@pytest.mark.parametrize("invert", [False, True])
def test_clay_strips(invert):
    m = meshgen.grid(129)
    dabs = _line_dabs(capi.TOOL_CLAY_STRIPS, (-0.6, -0.3, 0.0), (0.6, 0.4, 0.0), 0.3, 14, flags=capi.DAB_PLANE_TRIM,
                      tip_roundness=0.18, invert=invert, alpha=1.0)
    r = run_parity(m, dabs, leaf_limit=300)
    assert r["moved"] > 0


def test_clay_strips_view_plane_with_mask():
    m = meshgen.grid(129)
    mask = meshgen.low_freq_mask(m, seed=9)
    dabs = _line_dabs(capi.TOOL_CLAY_STRIPS, (-0.5, 0.3, 0.0), (0.5, -0.4, 0.0), 0.25, 12, sculpt_plane=capi.DIR_VIEW,
                      plane_offset=0.1, alpha=1.0)
    run_parity(m, dabs, mask=mask, leaf_limit=300)


@pytest.mark.parametrize("alpha", [0.3, 0.75, 1.0])
def test_smooth_icosphere(alpha):
    m = meshgen.icosphere(24, noise=0.004)
    dabs = stroke.c2_smooth_stroke(dabs=12, radius=0.45, alpha=alpha)
    r = run_parity(m, dabs, leaf_limit=120)
    assert r["moved"] > 0


def test_smooth_grid_boundary_rules():
    # open mesh: boundary verts average boundary neighbours only, corners stay
    m = meshgen.grid(65)
    dabs = _line_dabs(capi.TOOL_SMOOTH, (-0.9, -0.9, 0.0), (0.9, -0.8, 0.0), 0.4, 8, alpha=1.0)
    r = run_parity(m, dabs, leaf_limit=150)
    assert np.array_equal(r["co"][0], m.co[0])  # corner vertex has two neighbours


def test_fully_masked_and_hidden_nodes_are_skipped():
    m = meshgen.grid(97)

    def pre(orc, ses):
        na = orc.node_arrays()
        leaves = np.nonzero(na["flag"] & 1)[0]
        for k, n in enumerate(leaves[::3]):
            f = capi.PBVH_FullyMasked if k % 2 else capi.PBVH_FullyHidden
            orc.set_node_flag(int(n), f, True)
            ses.set_node_flag(int(n), f, True)

    dabs = _line_dabs(capi.TOOL_DRAW, (-0.6, -0.3, 0.0), (0.6, 0.4, 0.0), 0.3, 8)
    run_parity(m, dabs, leaf_limit=200, pre=pre)


def test_skip_normals_and_bounds_flags_persist():
    m = meshgen.grid(97)
    dabs = _line_dabs(capi.TOOL_DRAW, (-0.6, -0.3, 0.0), (0.6, 0.4, 0.0), 0.3, 9)
    for i, d in enumerate(dabs):
        if i % 3 != 2:
            d.flags |= capi.DAB_NO_NORMALS
        if i % 2 == 0:
            d.flags |= capi.DAB_NO_BOUNDS
    run_parity(m, dabs, leaf_limit=200)


def test_triangle_mesh_default_leaf_limit_single_leaf():
    m = meshgen.icosphere(10)  # 2000 tris: the whole mesh is one leaf
    dabs = _line_dabs(capi.TOOL_DRAW, (0.0, 0.0, 1.0), (0.5, 0.2, 0.8), 0.5, 5, view_normal=(0, 0, 1))
    run_parity(m, dabs)


def test_search_gather_through_host_api():
    """BKE_pbvh_search_gather(SCULPT_search_sphere_cb) on a device-attached PBVH == oracle DFS"""
    import ctypes as C
    from oracle_py import Oracle
    m = meshgen.cube(5)
    orc = Oracle(m, leaf_limit=100)
    ses = capi.SculptSession(m, leaf_limit=100, device=0)
    H = capi.host_lib()
    rng = np.random.default_rng(2)
    for _ in range(20):
        c = rng.uniform(-1.2, 1.2, size=3).astype(np.float32)
        rsq = float(rng.uniform(0.01, 1.5))
        for original in (False, True):
            data = capi.SculptSearchSphereData(capi.fptr(c), rsq, original, True)
            arr = C.POINTER(C.POINTER(capi.PBVHNode))()
            tot = C.c_int(0)
            H.BKE_pbvh_search_gather(ses.pbvh, C.cast(H.SCULPT_search_sphere_cb, C.c_void_p), C.byref(data),
                                     C.byref(arr), C.byref(tot))
            base = C.addressof(ses.pbvh.contents.nodes.contents)
            got = [(C.addressof(arr[i].contents) - base) // C.sizeof(capi.PBVHNode) for i in range(tot.value)]
            if tot.value:
                H.MEM_freeN(arr)
            else:
                assert not arr  # NULL, 0 when nothing is found (pbvh.c:2760-2766)
            assert got == list(orc.gather_sphere(c, rsq, original))
    ses.close()
    orc.close()


def test_vert_coords_apply_roundtrip():
    """BKE_pbvh_vert_coords_apply -> device normals + boxes == oracle recompute"""
    from oracle_py import Oracle
    m = meshgen.grid(65)
    ses = capi.SculptSession(m, leaf_limit=150, device=0)
    rng = np.random.default_rng(4)
    co = m.co.copy()
    sel = rng.uniform(size=m.totvert) < 0.3
    co[sel] += rng.normal(scale=0.01, size=(int(sel.sum()), 3)).astype(np.float32)
    H = capi.host_lib()
    H.BKE_pbvh_vert_coords_apply(ses.pbvh, capi.fptr(co), m.totvert)
    m2 = meshgen.Mesh(co, m.poly_start, m.poly_len, m.loop_v)
    # oracle: same topology (built from the ORIGINAL coordinates), new coordinates, everything dirty
    orc = Oracle(m, leaf_limit=150)
    orc.set_co(co)
    orc.L.or_recalc_all_normals(orc.p)
    for n in np.nonzero(orc.node_arrays()["flag"] & 1)[0]:
        orc.L.or_node_mark_update(orc.p, int(n))
    orc.update_bounds(capi.PBVH_UpdateBB | capi.PBVH_UpdateOriginalBB)
    assert np.array_equal(ses.co(), co)
    # verts that did not move keep their old normal on the device (only changed verts are marked,
    # pbvh.c:4731-4736) unless a neighbour moved -- compare the moved ones
    no_g, no_o = ses.no(), orc.no()
    assert np.array_equal(no_g[sel], no_o[sel])
    bb, obb = ses.node_bb()
    na = orc.node_arrays()
    assert np.array_equal(bb, na["vb"]) and np.array_equal(obb, na["orig_vb"])
    out = H.BKE_pbvh_vert_coords_alloc(ses.pbvh)
    back = np.ctypeslib.as_array(out, shape=(m.totvert * 3,)).reshape(-1, 3).copy()
    H.MEM_freeN(out)
    assert np.array_equal(back, co)
    assert m2.totvert == m.totvert
    ses.close()
    orc.close()


def test_c1_cube_draw_stroke_full_size():
    """config C1: 393,218-vertex cube, 100-dab draw stroke, default leaf limit"""
    m = meshgen.cube(8)
    assert m.totvert == 393218
    r = run_parity(m, stroke.c1_draw_stroke(), check_every=5)
    assert r["moved"] > 0


def test_c2_icosphere_smooth_stroke_reduced():
    """config C2 at f=158 (249,642 verts), 40 dabs: same script shape, oracle-sized"""
    m = meshgen.icosphere(158, noise=0.002)
    r = run_parity(m, stroke.c2_smooth_stroke(dabs=40), check_every=4)
    assert r["moved"] > 0


def test_c4_tools_reduced_grid():
    """config C4 on a 512^2 grid: each tool with mask and automask"""
    m = meshgen.grid(512)
    mask = meshgen.low_freq_mask(m)
    diag = m.bbox_diag()
    for tool in (capi.TOOL_DRAW, capi.TOOL_INFLATE, capi.TOOL_CLAY_STRIPS, capi.TOOL_GRAB):
        ses = capi.SculptSession(m, leaf_limit=0)
        auto = np.zeros(m.totvert, dtype=np.float32)
        capi.host_lib().DUNE_sculpt_automask_boundary_edges(ses.pbvh, 1, capi.fptr(auto))
        ses.close()
        dabs = stroke.c4_tool_stroke(tool, diag, dabs=12)
        r = run_parity(m, dabs, mask=mask, automask=auto, check_every=3)
        assert r["moved"] > 0, tool


def test_c3_radius_sweep_reduced_grid():
    """config C3 on a 768^2 grid (589,824 verts, multi-tile leaves): draw + normals + bounds per dab,
    radius sweep 1-50 % of the diagonal, 2 dabs per radius"""
    m = meshgen.grid(768)
    dabs = stroke.c3_radius_sweep(m.bbox_diag(), dabs_per_radius=2)
    r = run_parity(m, dabs, check_every=1)
    assert r["moved"] > 0


def test_checkpoint_rollback_repeats_the_stroke_bit_for_bit():
    """dsc_state_save / dsc_state_restore: after a rollback the same stroke gives the same mesh
    (what bench.py relies on to make every step identical), and the host arrays follow"""
    m = meshgen.grid(300)
    dabs = stroke.c3_radius_sweep(m.bbox_diag(), dabs_per_radius=2)
    ses = capi.SculptSession(m, device=0)
    try:
        co0, no0 = ses.co(), ses.no()
        bb0, obb0 = ses.node_bb()
        ses.checkpoint()
        outs = []
        for _ in range(2):
            ses.stroke_begin()
            for d in dabs:
                ses.dab(d)
            ses.stroke_end()
            outs.append((ses.co(), ses.no(), ses.node_bb(), ses.stats()["vertex_dabs"]))
            ses.rollback()
            assert np.array_equal(ses.co(), co0) and np.array_equal(ses.no(), no0)
            bb, obb = ses.node_bb()
            assert np.array_equal(bb, bb0) and np.array_equal(obb, obb0)
        assert not np.array_equal(outs[0][0], co0)
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
        assert np.array_equal(outs[0][2][0], outs[1][2][0]) and outs[0][3] == outs[1][3]
    finally:
        ses.close()


def test_batched_dabs_through_cuda_graphs_match_the_oracle():
    """dsc_dabs: runs of dabs with one launch sequence are replayed as CUDA graphs over the device dab
    ring (batches of 32 / 16 / 4, singles for the rest) -- same bits as the oracle's dab-by-dab stroke"""
    import ctypes as C
    from oracle_py import Oracle
    m = meshgen.grid(400)
    diag = m.bbox_diag()
    dabs = stroke.c3_radius_sweep(diag, dabs_per_radius=9)               # 63 draw dabs: 32 + 16 + 4 x 3 + 3 singles
    dabs += _line_dabs(capi.TOOL_INFLATE, (-0.5, 0.2, 0.0), (0.5, -0.3, 0.0), 0.25, 6)   # another signature: 4 + 2
    orc = Oracle(m)
    ses = capi.SculptSession(m, device=0)
    try:
        orc.stroke_begin()
        for d in dabs:
            orc.dab(d)
        orc.stroke_end()
        arr = (capi.DscDab * len(dabs))(*dabs)
        for _ in range(2):  # the second stroke replays the cached graphs
            ses.checkpoint()
            ses.stroke_begin()
            ses.dabs(arr, len(dabs))
            st = ses.stats()
            ses.stroke_end()
            assert st["vertex_dabs"] == orc.vertex_dabs() and st["dabs"] == len(dabs)
            assert np.array_equal(ses.co(), orc.co()), "positions differ in bits"
            assert np.array_equal(ses.no(), orc.no()), "normals differ in bits"
            bb, obb = ses.node_bb()
            na = orc.node_arrays()
            assert np.array_equal(bb, na["vb"]) and np.array_equal(obb, na["orig_vb"])
            assert np.array_equal(ses.orig_co(), orc.orig_co())
            ses.rollback()
    finally:
        ses.close()
        orc.close()


def test_mixed_polygons_general_and_tile_paths_together():
    """triangles, quads and hexagons in one mesh: leaves with an n-gon take the general normals / bounds
    kernels (and keep their update flags until the dab's clear pass), their neighbours the tile kernel;
    draw, inflate and smooth strokes across both"""
    m = meshgen.mixed_grid(97)
    assert set(np.unique(m.poly_len).tolist()) == {3, 4, 6}
    dabs = _line_dabs(capi.TOOL_DRAW, (-0.7, -0.6, 0.0), (0.7, 0.5, 0.0), 0.22, 6)
    dabs += _line_dabs(capi.TOOL_INFLATE, (0.6, -0.6, 0.0), (-0.6, 0.6, 0.0), 0.3, 4)
    r = run_parity(m, dabs, leaf_limit=300)
    assert r["moved"] > 0
    r = run_parity(m, _line_dabs(capi.TOOL_SMOOTH, (-0.5, 0.0, 0.0), (0.5, 0.1, 0.0), 0.35, 4, alpha=0.8), leaf_limit=300)
    assert r["moved"] > 0
    # batched submission on a mesh with slow leaves goes dab by dab (no graph) and still matches
    arr = (capi.DscDab * len(dabs))(*dabs)
    ses = capi.SculptSession(m, leaf_limit=300, device=0)
    from oracle_py import Oracle
    orc = Oracle(m, leaf_limit=300)
    try:
        orc.stroke_begin()
        for d in dabs:
            orc.dab(d)
        orc.stroke_end()
        ses.stroke_begin()
        ses.dabs(arr, len(dabs))
        ses.stroke_end()
        assert np.array_equal(ses.co(), orc.co()) and np.array_equal(ses.no(), orc.no())
        bb, _ = ses.node_bb()
        assert np.array_equal(bb, orc.node_arrays()["vb"])
    finally:
        ses.close()
        orc.close()


def test_dab_that_gathers_nothing_and_degenerate_inputs():
    """a dab far from the mesh gathers no node (BKE_pbvh_search_gather returns NULL, 0, pbvh.c:2760-2766)
    and changes nothing; a zero-radius dab is refused; a one-quad mesh is one leaf"""
    m = meshgen.grid(33)
    far = capi.make_dab(capi.TOOL_DRAW, (5.0, 5.0, 5.0), 0.3, bstrength=0.3, view_normal=(0, 0, 1))
    r = run_parity(m, [far, capi.make_dab(capi.TOOL_DRAW, (0.1, 0.1, 0.0), 0.3, bstrength=0.3, view_normal=(0, 0, 1)), far], leaf_limit=100)
    assert r["moved"] > 0
    ses = capi.SculptSession(m, leaf_limit=100, device=0)
    try:
        ses.capture(True)
        ses.stroke_begin()
        ses.dab(far)
        assert ses.hits().size == 0 and ses.moved().size == 0
        bad = capi.make_dab(capi.TOOL_DRAW, (0, 0, 0), 0.3, bstrength=0.3)
        bad.radius = 0.0
        with pytest.raises(capi.DeviceError):
            ses.dab(bad)
        ses.stroke_end()
    finally:
        ses.close()
    tiny = meshgen.grid(2)
    r = run_parity(tiny, [capi.make_dab(capi.TOOL_DRAW, (0.0, 0.0, 0.0), 2.0, bstrength=0.2, view_normal=(0, 0, 1))])
    assert r["moved"] == 4


def test_persistent_batch_kernel_is_bit_identical(monkeypatch):
    """DSC_BATCH_KERNEL=1: a run of dabs in one cooperative launch, stages separated by grid barriers (an
    experiment that measured slower than graph replay; kept honest here)"""
    from oracle_py import Oracle
    monkeypatch.setenv("DSC_BATCH_KERNEL", "1")
    m = meshgen.grid(300)
    dabs = stroke.c3_radius_sweep(m.bbox_diag(), dabs_per_radius=3)
    dabs += _line_dabs(capi.TOOL_CLAY_STRIPS, (-0.5, 0.2, 0.0), (0.5, -0.3, 0.0), 0.25, 5)
    orc = Oracle(m)
    ses = capi.SculptSession(m, device=0)
    try:
        orc.stroke_begin()
        for d in dabs:
            orc.dab(d)
        orc.stroke_end()
        arr = (capi.DscDab * len(dabs))(*dabs)
        ses.stroke_begin()
        ses.dabs(arr, len(dabs))
        st = ses.stats()
        ses.stroke_end()
        assert st["vertex_dabs"] == orc.vertex_dabs()
        assert st["kernel_launches"] < 3 * len(dabs), "the batch kernel was not used"
        assert np.array_equal(ses.co(), orc.co()) and np.array_equal(ses.no(), orc.no())
        bb, obb = ses.node_bb()
        na = orc.node_arrays()
        assert np.array_equal(bb, na["vb"]) and np.array_equal(obb, na["orig_vb"])
    finally:
        ses.close()
        orc.close()


@pytest.mark.parametrize("smooth", [True, False])
def test_draw_buffers_filled_on_the_device_match_gpu_pbvh_mesh_buffers_update(smooth):
    """after a stroke the leaves flagged PBVH_UpdateDrawBuffers get their 36-byte-per-corner
    vertex buffers (gpu_buffers.c:84-100, 174-305) packed on the device; bytes equal the oracle's; flags cleared;
    leaves the stroke did not touch are only filled by the first (rebuild) pass"""
    from oracle_py import Oracle
    m = meshgen.mixed_grid(65) if not smooth else meshgen.grid(129)
    mask = meshgen.low_freq_mask(m)
    dabs = _line_dabs(capi.TOOL_DRAW, (-0.6, -0.5, 0.0), (0.5, 0.4, 0.0), 0.2, 5)
    orc = Oracle(m, mask=mask, leaf_limit=200)
    ses = capi.SculptSession(m, mask=mask, leaf_limit=200, device=0, draw_buffers=True)
    try:
        na = orc.node_arrays()
        leaves = np.nonzero(na["flag"] & 1)[0]
        for rnd in range(2):
            if rnd == 1:
                orc.stroke_begin()
                ses.stroke_begin()
                for d in dabs:
                    orc.dab(d)
                    ses.dab(d)
                orc.stroke_end()
                ses.stroke_end()
            flagged = [int(n) for n in leaves if orc.node_arrays()["flag"][n] & (capi.PBVH_UpdateDrawBuffers | capi.PBVH_RebuildDrawBuffers)]
            assert flagged and (rnd == 0 or len(flagged) < leaves.size)
            ses.update_draw_buffers(smooth=smooth, show_mask=True)
            for n in flagged:
                ref = orc.draw_buffer(n, int(na["totprim"][n]), smooth=smooth, show_mask=True)
                got = ses.draw_buffer(n)
                assert got.shape == ref.shape and np.array_equal(got, ref), "leaf %d round %d" % (n, rnd)
            keep = capi.PBVH_UpdateDrawBuffers | capi.PBVH_RebuildDrawBuffers
            assert not np.any(ses.node_flags() & keep) and not np.any(orc.node_arrays()["flag"] & keep)
    finally:
        ses.close()
        orc.close()


def _ray_cases(m, rng, count):
    """rays at a mesh from outside its box: at random points, exactly at vertices and edge midpoints (ties between
    the looptris round them, across leaf borders too), axis-parallel ones, and ones that miss"""
    co = m.co
    lo, hi = co.min(axis=0), co.max(axis=0)
    ctr, ext = 0.5 * (lo + hi), float(np.linalg.norm(hi - lo))
    rays = []
    for i in range(count):
        kind = i % 5
        if kind == 0:
            target = lo + rng.random(3).astype(np.float32) * (hi - lo)
        elif kind == 1:
            target = co[rng.integers(m.totvert)]
        elif kind == 2:
            f = rng.integers(m.totpoly)
            a, b = m.loop_v[m.poly_start[f]], m.loop_v[m.poly_start[f] + 1]
            target = (0.5 * (co[a] + co[b])).astype(np.float32)
        elif kind == 3:
            target = co[rng.integers(m.totvert)]
        else:
            target = ctr + (hi - lo) * 3.0 * np.sign(rng.normal(size=3)).astype(np.float32)  # far off: a miss
        if kind == 3:
            d = np.zeros(3, np.float32)
            d[rng.integers(3)] = 1.0 if rng.random() < 0.5 else -1.0
        else:
            d = rng.normal(size=3).astype(np.float32)
            if abs(d[2]) < 0.3:
                d[2] = 0.7
            d /= np.float32(np.linalg.norm(d))
        start = (np.asarray(target, np.float32) - d * np.float32(2.0 * ext)).astype(np.float32)
        rays.append((start, d.astype(np.float32)))
    return rays


def _same_hit(ref, got, what):
    assert (ref is None) == (got is None), what
    if ref is None:
        return 0
    assert np.float32(ref["depth"]).tobytes() == np.float32(got["depth"]).tobytes(), what
    assert ref["face"] == got["face"] and ref["vertex"] == got["vertex"] and ref["node"] == got["node"], (what, ref, got)
    assert np.array_equal(ref["normal"].view(np.uint32), got["normal"].view(np.uint32)), what
    return 1


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["grid", "mixed", "sphere"])
def test_raycast_matches_bke_pbvh_raycast(name):
    """the nearest hit of a ray -- depth, active vertex, face, triangle normal, leaf -- as
    BKE_pbvh_raycast (pbvh.c:3915-3928) with the stroke operator's per-leaf callback over BKE_pbvh_node_raycast
    (pbvh.c:4041-4100, 4203-4260) returns it; bit-equal, including which of several tied looptris wins.  Mid-stroke:
    original=True sees the stroke-start surface, original=False the deformed one."""
    from oracle_py import Oracle
    m = {"grid": lambda: meshgen.grid(129), "mixed": lambda: meshgen.mixed_grid(65), "sphere": lambda: meshgen.icosphere(24, noise=0.02)}[name]()
    orc = Oracle(m, leaf_limit=300)
    ses = capi.SculptSession(m, leaf_limit=300, device=0, raycast=True)
    try:
        rng = np.random.default_rng(17)
        rays = _ray_cases(m, rng, 150)
        hits = 0
        for i, (s, d) in enumerate(rays):
            hits += _same_hit(orc.raycast(s, d), ses.raycast(s, d), "ray %d" % i)
        assert hits > 60 and hits < len(rays)
        # a depth limit in front of / behind the surface
        s, d = rays[0]
        full = orc.raycast(s, d)
        if full is not None:
            for md in (float(full["depth"]) * 0.5, float(full["depth"]), float(full["depth"]) * 1.5):
                _same_hit(orc.raycast(s, d, max_depth=md), ses.raycast(s, d, max_depth=md), "max_depth %g" % md)
        # mid-stroke: some leaves hold undo coordinates
        top = m.co[np.argmax(m.co[:, 2])]
        dabs = _line_dabs(capi.TOOL_DRAW, tuple(top - np.float32([0.3, 0.2, 0.0])), tuple(top + np.float32([0.2, 0.3, 0.0])), 0.25, 4)
        orc.stroke_begin()
        ses.stroke_begin()
        for dab in dabs:
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
        for i, (s, d) in enumerate(rays[:40]):
            _same_hit(orc.raycast(s, d), ses.raycast(s, d), "after stroke ray %d" % i)
    finally:
        ses.close()
        orc.close()


# ---- rows a10 / a11 / a19: hidden verts, tube falloff, clipping and axis locks, grab normal weight ----------------

def _hidden_flags(m, seed=5, frac=0.2):
    rng = np.random.default_rng(seed)
    vf = np.zeros(m.totvert, np.uint8)
    vf[rng.random(m.totvert) < frac] = capi.ME_HIDE
    # a fully hidden patch too, so whole leaves are flagged at build
    co = m.co
    vf[(co[:, 0] > 0.55) & (co[:, 1] > 0.55)] = capi.ME_HIDE
    return vf


@pytest.mark.parametrize("tool", [capi.TOOL_DRAW, capi.TOOL_INFLATE, capi.TOOL_CLAY_STRIPS, capi.TOOL_SMOOTH, capi.TOOL_GRAB])
def test_hidden_vertices_are_skipped_by_every_tool(tool):
    m = meshgen.grid(97, height=0.05)
    vf = _hidden_flags(m)
    if tool == capi.TOOL_GRAB:
        bs = stroke._strength(tool, 0.8)
        dabs = [capi.make_dab(tool, (0.3, 0.3, 0.0), 0.45, bstrength=bs, grab_delta=np.array([0.1, -0.1, 0.3]) * (i + 1) / 6,
                              flags=capi.DAB_FIRST_STEP if i == 0 else 0) for i in range(6)]
    else:
        dabs = _line_dabs(tool, (-0.6, -0.5, 0), (0.8, 0.7, 0), 0.3, 12)
    before = m.co.copy()
    res = run_parity(m, dabs, leaf_limit=300, vert_flag=vf)
    hid = vf != 0
    assert np.array_equal(res["co"][hid], before[hid]), "a hidden vertex moved"
    assert res["moved"] > 0


@pytest.mark.parametrize("tool", [capi.TOOL_DRAW, capi.TOOL_INFLATE, capi.TOOL_SMOOTH, capi.TOOL_CLAY_STRIPS, capi.TOOL_GRAB])
def test_tube_falloff_tests_the_view_line(tool):
    # an icosphere seen along a tilted view: the tube reaches the far side too, the node test is the line-box distance
    m = meshgen.icosphere(5)
    vn = np.array([0.3, -0.2, 0.93], np.float32)
    vn /= np.linalg.norm(vn)
    kw = dict(view_normal=vn, falloff_shape=capi.FALLOFF_TUBE)
    if tool == capi.TOOL_GRAB:
        bs = stroke._strength(tool, 0.8)
        dabs = [capi.make_dab(tool, (0.2, 0.1, 0.97), 0.3, bstrength=bs, grab_delta=np.array([0.1, 0.05, 0.2]) * (i + 1) / 5,
                              flags=capi.DAB_FIRST_STEP if i == 0 else 0, **kw) for i in range(5)]
    else:
        dabs = _line_dabs(tool, (-0.3, -0.2, 0.93), (0.4, 0.3, 0.86), 0.25, 10, **kw)
    res = run_parity(m, dabs, leaf_limit=400)
    moved_far = (np.abs(res["co"] - m.co).max(axis=1) > 0) & (m.co @ vn < -0.5)
    # (the clay-strips cube test bounds the depth in brush space whatever the falloff shape)
    assert tool == capi.TOOL_CLAY_STRIPS or moved_far.any(), "the tube did not reach the far side of the sphere"


def test_tube_falloff_with_hidden_and_mask_on_grid():
    m = meshgen.grid(129, height=0.08)
    vf = _hidden_flags(m, seed=9)
    rng = np.random.default_rng(3)
    mask = rng.random(m.totvert).astype(np.float32)
    dabs = _line_dabs(capi.TOOL_DRAW, (-0.7, 0.1, 0.4), (0.7, -0.3, -0.4), 0.2, 14, falloff_shape=capi.FALLOFF_TUBE,
                      view_normal=(0.0, 0.6, 0.8))
    run_parity(m, dabs, mask=mask, leaf_limit=500, vert_flag=vf)


@pytest.mark.parametrize("tool", [capi.TOOL_DRAW, capi.TOOL_INFLATE, capi.TOOL_SMOOTH, capi.TOOL_CLAY_STRIPS, capi.TOOL_GRAB])
def test_mirror_clipping_and_axis_locks(tool):
    m = meshgen.grid(97, height=0.05)
    on_plane = np.abs(m.co[:, 0]) <= 0.011
    assert on_plane.any()
    kw = dict(clip_flags=capi.CLIP_X | capi.LOCK_Y, clip_tolerance=(0.011, 0.0, 0.0))
    if tool == capi.TOOL_GRAB:
        bs = stroke._strength(tool, 0.8)
        dabs = [capi.make_dab(tool, (0.05, 0.1, 0.0), 0.4, bstrength=bs, grab_delta=np.array([0.2, 0.2, 0.3]) * (i + 1) / 5,
                              flags=capi.DAB_FIRST_STEP if i == 0 else 0, **kw) for i in range(5)]
    else:
        dabs = _line_dabs(tool, (-0.3, -0.4, 0), (0.3, 0.5, 0), 0.3, 10, sculpt_plane=capi.DIR_X if tool == capi.TOOL_DRAW else capi.DIR_AREA, **kw)
    res = run_parity(m, dabs, leaf_limit=300)
    moved = np.abs(res["co"] - m.co).max(axis=1) > 0
    assert moved.any()
    assert np.array_equal(res["co"][:, 1], m.co[:, 1]), "a locked axis changed"
    touched_on_plane = on_plane & moved
    assert np.all(res["co"][touched_on_plane, 0] == 0.0), "a vertex inside the clip tolerance left the mirror plane"


@pytest.mark.parametrize("weight,plane", [(0.35, capi.DIR_AREA), (1.0, capi.DIR_AREA), (0.5, capi.DIR_VIEW), (0.5, capi.DIR_Z)])
def test_grab_normal_weight(weight, plane):
    m = meshgen.icosphere(5)
    bs = stroke._strength(capi.TOOL_GRAB, 0.9)
    vn = np.array([0.2, 0.1, 0.97], np.float32)
    vn /= np.linalg.norm(vn)
    dabs = [capi.make_dab(capi.TOOL_GRAB, (0.3, 0.2, 0.93), 0.4, bstrength=bs, view_normal=vn, normal_weight=weight, sculpt_plane=plane,
                          grab_delta=np.array([0.15, -0.05, 0.1]) * (i + 1) / 8, flags=capi.DAB_FIRST_STEP if i == 0 else 0)
            for i in range(8)]
    res = run_parity(m, dabs, leaf_limit=400)
    plain = [capi.make_dab(capi.TOOL_GRAB, (0.3, 0.2, 0.93), 0.4, bstrength=bs, view_normal=vn, sculpt_plane=plane,
                           grab_delta=np.array([0.15, -0.05, 0.1]) * (i + 1) / 8, flags=capi.DAB_FIRST_STEP if i == 0 else 0)
             for i in range(8)]
    ref = run_parity(m, plain, leaf_limit=400)
    assert not np.array_equal(res["co"], ref["co"]), "the normal weight changed nothing"


def test_symmetry_passes_through_the_dab_helper():
    # X|Z symmetry: four passes per dab, each a mirror image (location, view normal, drag)
    m = meshgen.icosphere(4)
    base = _line_dabs(capi.TOOL_DRAW, (0.5, 0.1, 0.8), (0.6, 0.4, 0.65), 0.25, 4, view_normal=(0.4, 0.0, 0.9165))
    dabs = []
    for d in base:
        dabs += capi.dab_symmetry(d, 1 | 4)
    assert len(dabs) == 16
    res = run_parity(m, dabs, leaf_limit=300)
    # both halves were sculpted
    moved = np.abs(res["co"] - m.co).max(axis=1) > 0
    assert (moved & (m.co[:, 0] > 0.2) & (m.co[:, 2] > 0.2)).any() and (moved & (m.co[:, 0] < -0.2) & (m.co[:, 2] > 0.2)).any()
    assert (moved & (m.co[:, 0] > 0.2) & (m.co[:, 2] < -0.2)).any() and (moved & (m.co[:, 0] < -0.2) & (m.co[:, 2] < -0.2)).any()


def test_host_marks_made_while_the_device_is_ahead_reach_it():
    """BKE_pbvh_node_fully_hidden_set / _mark_update / BKE_pbvh_vert_mark_update on the host PBVH between device calls, when the
    device already holds newer state than the host: the marks are pushed with the next device call and survive the sync."""
    from oracle_py import Oracle
    m = meshgen.grid(97)
    dabs = _line_dabs(capi.TOOL_DRAW, (-0.6, -0.3, 0.0), (0.6, 0.4, 0.0), 0.3, 8)
    orc = Oracle(m, leaf_limit=200)
    ses = capi.SculptSession(m, leaf_limit=200, device=0)
    try:
        leaves = np.nonzero(orc.node_arrays()["flag"] & 1)[0]
        orc.stroke_begin(None)
        ses.stroke_begin(None)
        ses.capture(True)
        for i, d in enumerate(dabs):
            if i == 3:
                assert ses.pbvh.contents.device_dirty
                for k, n in enumerate(leaves[::3]):
                    orc.set_node_flag(int(n), capi.PBVH_FullyMasked if k % 2 else capi.PBVH_FullyHidden, True)
                    (ses.bke_node_fully_masked_set if k % 2 else ses.bke_node_fully_hidden_set)(int(n), True)
            if i == 6:   # and taken back again
                for k, n in enumerate(leaves[::6]):
                    orc.set_node_flag(int(n), capi.PBVH_FullyHidden, False)
                    ses.bke_node_fully_hidden_set(int(n), False)
            orc.dab(d)
            ses.dab(d)
            assert np.array_equal(orc.hits(), ses.hits()), "dab %d: node-hit list" % i
        orc.stroke_end()
        ses.stroke_end()
        assert np.array_equal(orc.co(), ses.co()) and np.array_equal(orc.no(), ses.no())
        keep = capi.PBVH_FullyHidden | capi.PBVH_FullyMasked
        assert np.array_equal(orc.node_arrays()["flag"] & keep, ses.node_arrays()["flag"] & keep)
        # the advisor's case: a device call leaves the device ahead, then the host marks nodes and verts
        ses.bke_update_bounds(capi.PBVH_UpdateBB)
        assert ses.pbvh.contents.device_dirty
        marked = [int(n) for n in leaves[1::5]]
        orc.L.or_node_mark_update.argtypes = [C.c_void_p, C.c_int]
        orc.L.or_vert_mark_update.argtypes = [C.c_void_p, C.c_int]
        for n in marked:
            ses.bke_node_mark_update(n)
            orc.L.or_node_mark_update(orc.p, n)
            for v in ses.node_vert_indices(n)[:4]:
                ses.bke_vert_mark_update(int(v))
                orc.L.or_vert_mark_update(orc.p, int(v))
        ses.bke_update_normals()
        orc.update_normals()
        ses.sync_to_host()
        fl = ses.node_arrays()["flag"]
        stay = capi.PBVH_UpdateBB | capi.PBVH_UpdateOriginalBB | capi.PBVH_UpdateDrawBuffers | capi.PBVH_UpdateRedraw
        assert all((fl[n] & stay) == stay for n in marked), "marks made on the host were lost in the sync"
        assert all(not (fl[n] & capi.PBVH_UpdateNormals) for n in marked), "the normals update did not see the marked nodes"
        assert np.array_equal(ses.node_flags(), fl)
        ses.bke_update_bounds(capi.PBVH_UpdateBB | capi.PBVH_UpdateOriginalBB)
        ses.sync_to_host()
        fl = ses.node_arrays()["flag"]
        assert all(not (fl[n] & (capi.PBVH_UpdateBB | capi.PBVH_UpdateOriginalBB)) for n in marked)
        # a marked vert gets the faces of the flagged leaves only (pbvh.c:3335-3375): whatever that gives, it is the same on both sides
        assert np.array_equal(orc.no(), ses.no()), "normals of the verts marked on the host"
    finally:
        ses.close()
        orc.close()


def test_draw_buffers_take_their_shading_per_leaf_from_the_material_flags():
    """gpu_buffers.c:221-222: a leaf's buffer is smooth or flat by ME_SMOOTH of the poly of its first looptri; the build splits
    leaves by material / smooth flag (pbvh.c:411-466), so both kinds exist side by side"""
    from oracle_py import Oracle
    m = meshgen.grid(97)
    cx = m.co[m.loop_v.reshape(-1, 4), 0].mean(axis=1)
    poly_flag = (cx > 0.1).astype(np.uint8)          # ME_SMOOTH on one side
    poly_mat = np.zeros(m.totpoly, np.int16)
    mask = meshgen.low_freq_mask(m)
    orc = Oracle(m, mask=mask, leaf_limit=200, poly_mat=poly_mat, poly_flag=poly_flag)
    ses = capi.SculptSession(m, mask=mask, leaf_limit=200, device=0, draw_buffers=True, poly_mat=poly_mat, poly_flag=poly_flag)
    try:
        na = orc.node_arrays()
        prim, tri_poly = orc.prim_indices(), orc.tri_poly()
        leaves = [int(n) for n in np.nonzero(na["flag"] & 1)[0]]
        smooth_of = {n: bool(poly_flag[tri_poly[prim[na["prim_offset"][n]]]] & 1) for n in leaves}
        assert any(smooth_of.values()) and not all(smooth_of.values())
        ses.update_draw_buffers(smooth=-1, show_mask=True)
        for n in leaves:
            ref = orc.draw_buffer(n, int(na["totprim"][n]), smooth=smooth_of[n], show_mask=True)
            assert np.array_equal(ses.draw_buffer(n), ref), "leaf %d (%s)" % (n, "smooth" if smooth_of[n] else "flat")
    finally:
        ses.close()
        orc.close()
