"""Host C side (reference-named PBVH entry points) checked against the oracle on the CPU: build,
tessellation, session tables, traversal with arbitrary callbacks, stroke-start helpers.  No GPU."""
import ctypes as C

import numpy as np
import pytest

from dune_sculpt_b200 import capi, meshgen
from oracle_py import Oracle, lib as oracle_lib, iptr, fptr


CASES = [("grid65", lambda: meshgen.grid(65), 200), ("cube4", lambda: meshgen.cube(4), 100),
         ("ico8", lambda: meshgen.icosphere(8), 77), ("grid300_default_limit", lambda: meshgen.grid(300), 0),
         ("single_leaf", lambda: meshgen.icosphere(6), 0)]


@pytest.mark.parametrize("name,mk,ll", CASES, ids=[c[0] for c in CASES])
def test_build_mesh_matches_oracle(name, mk, ll):
    m = mk()
    o = Oracle(m, leaf_limit=ll)
    s = capi.SculptSession(m, leaf_limit=ll)
    a, b = o.node_arrays(), s.node_arrays()
    assert o.totnode == s.totnode
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(o.prim_indices(), s.prim_indices())
    for i in range(o.totnode):
        if a["flag"][i] & 1:
            cnt = a["uniq_verts"][i] + a["face_verts"][i]
            assert np.array_equal(o.node_vert_indices(i, cnt), s.node_vert_indices(i))
            assert np.array_equal(o.node_face_vert_indices(i, a["totprim"][i]), s.node_face_vert_indices(i))
    s.close()


def test_looptri_rule_and_quad_flip():
    # a concave quad: the 0-2 diagonal is degenerate, the tessellation must flip to 1-3
    co = np.array([[0, 0, 0], [0.3, 0.5, 0], [1, 1, 0], [0, 1, 0]], np.float32)  # v1 is reflex
    m = meshgen.Mesh(co, np.array([0], np.int32), np.array([4], np.int32), np.array([0, 1, 2, 3], np.int32))
    s = capi.SculptSession(m)
    tri = s.looptri["tri"][:2].astype(np.int64)
    L = oracle_lib()
    t = np.zeros((2, 3), np.int32)
    tp = np.zeros(2, np.int32)
    L.or_looptri_calc(1, iptr(m.poly_start), iptr(m.poly_len), iptr(m.loop_v), fptr(co), iptr(t), iptr(tp))
    assert np.array_equal(tri, t) and tri.tolist() == [[0, 1, 3], [1, 2, 3]]
    s.close()
    m2 = meshgen.grid(9)
    s2 = capi.SculptSession(m2)
    assert s2.tottri == 2 * m2.totpoly
    assert np.array_equal(s2.looptri["tri"][0], [0, 1, 2]) and np.array_equal(s2.looptri["tri"][1], [0, 2, 3])
    s2.close()


@pytest.mark.parametrize("mk", [lambda: meshgen.grid(33), lambda: meshgen.icosphere(6), lambda: meshgen.cube(3)])
def test_neighbor_tables_match_oracle(mk):
    m = mk()
    s = capi.SculptSession(m)
    auto = np.zeros(m.totvert, np.float32)
    s.H.DUNE_sculpt_automask_boundary_edges(s.pbvh, 1, capi.fptr(auto))  # builds the tables
    off, idx, bnd = s.neighbor_tables()
    L = oracle_lib()
    o_off = np.zeros(m.totvert + 1, np.int32)
    o_idx = np.zeros(2 * m.totloop + 1, np.int32)
    o_b = np.zeros(m.totvert, np.uint8)
    n = L.or_vert_neighbors(m.totvert, m.totpoly, iptr(m.poly_start), iptr(m.poly_len), iptr(m.loop_v), iptr(o_off),
                            iptr(o_idx), o_b.ctypes.data)
    assert np.array_equal(off, o_off) and np.array_equal(idx, o_idx[:n]) and np.array_equal(bnd, o_b)
    s.close()


def test_grid_neighbors_and_boundary_values():
    m = meshgen.grid(5)
    s = capi.SculptSession(m)
    auto = np.zeros(m.totvert, np.float32)
    s.H.DUNE_sculpt_automask_boundary_edges(s.pbvh, 2, capi.fptr(auto))
    off, idx, bnd = s.neighbor_tables()
    deg = np.diff(off).reshape(5, 5)
    assert deg[0, 0] == 2 and deg[0, 2] == 3 and deg[2, 2] == 4
    assert bnd.reshape(5, 5)[0].all() and not bnd.reshape(5, 5)[1:4, 1:4].any()
    a = auto.reshape(5, 5)
    # distance 0 -> factor 0, distance 1 (2 steps) -> 1 - (1 - 1/2)^2, centre untouched
    assert a[0, 0] == 0.0 and a[1, 1] == pytest.approx(0.75) and a[2, 2] == 1.0
    s.close()


def test_topology_automask_floodfill():
    # two disconnected grids: only the component of the seed gets factor 1
    g = meshgen.grid(9)
    co = np.concatenate([g.co, g.co + np.array([5, 0, 0], np.float32)])
    loops = np.concatenate([g.loop_v, g.loop_v + g.totvert])
    m = meshgen.Mesh(co, np.concatenate([g.poly_start, g.poly_start + g.totloop]), np.concatenate([g.poly_len, g.poly_len]), loops)
    s = capi.SculptSession(m)
    f = np.zeros(m.totvert, np.float32)
    loc = np.zeros(3, np.float32)
    s.H.DUNE_sculpt_automask_topology(s.pbvh, 3, capi.fptr(loc), C.c_float(0.0), capi.fptr(f))
    assert f[:g.totvert].all() and not f[g.totvert:].any()
    s.close()


def test_brush_strength_scalar():
    H = capi.host_lib()
    bs = lambda tool, a, p=1.0, di=False, inv=False: H.DUNE_sculpt_brush_strength(tool, a, p, di, inv, 1.0, 1.0)  # noqa: E731
    assert bs(capi.TOOL_DRAW, 0.5) == pytest.approx(0.25)                     # alpha squared
    assert bs(capi.TOOL_DRAW, 0.5, inv=True) == pytest.approx(-0.25)
    assert bs(capi.TOOL_DRAW, 0.5, di=True, inv=True) == pytest.approx(0.25)
    assert bs(capi.TOOL_CLAY_STRIPS, 1.0, p=0.25) == pytest.approx(0.3 * 0.125)  # 0.3 * pressure^1.5
    assert bs(capi.TOOL_INFLATE, 1.0) == pytest.approx(0.25) and bs(capi.TOOL_INFLATE, 1.0, inv=True) == pytest.approx(-0.125)
    assert bs(capi.TOOL_SMOOTH, 0.75) == pytest.approx(0.5625)
    assert bs(capi.TOOL_GRAB, 0.8) == pytest.approx(0.8)                       # root alpha
    d = capi.make_dab(capi.TOOL_DRAW, (0, 0, 0), 1.0)
    assert d.normal_radius_factor == 0.5 and d.plane_trim == 0.5 and d.sculpt_plane == capi.DIR_AREA and d.hardness == 0.0


def test_search_gather_host_path_with_arbitrary_callback():
    """no device attached: BKE_pbvh_search_gather walks the tree on the host like pbvh.c:2664-2767
    (children left first, NULL/0 when empty, caller frees with MEM_freeN)"""
    m = meshgen.cube(4)
    s = capi.SculptSession(m, leaf_limit=100)
    o = Oracle(m, leaf_limit=100)
    H = s.H
    base = C.addressof(s.pbvh.contents.nodes.contents)

    def run(cb, data=None):
        arr = C.POINTER(C.POINTER(capi.PBVHNode))()
        tot = C.c_int(0)
        H.BKE_pbvh_search_gather(s.pbvh, cb, data, C.byref(arr), C.byref(tot))
        got = [(C.addressof(arr[i].contents) - base) // C.sizeof(capi.PBVHNode) for i in range(tot.value)]
        if tot.value:
            H.MEM_freeN(arr)
        else:
            assert not arr
        return got

    # x > 0 half-space test on the node box, as a Python callback
    cb = capi.SEARCH_CB(lambda node, data: node.contents.vb.bmax[0] > 0.25)
    got = run(C.cast(cb, C.c_void_p))
    na = o.node_arrays()
    leaves = np.nonzero(na["flag"] & 1)[0]
    leaves = leaves[np.argsort(na["prim_offset"][leaves])]
    assert got == [int(n) for n in leaves if na["vb"][n, 3] > 0.25]
    assert run(C.cast(capi.SEARCH_CB(lambda node, data: False), C.c_void_p)) == []
    assert run(None) == [int(n) for n in leaves]  # no callback: every leaf (pbvh.c:2687)
    # the sphere callback on the host path
    c = np.array([0.9, 0.1, 1.0], np.float32)
    data = capi.SculptSearchSphereData(capi.fptr(c), 0.3, False, True)
    got = run(C.cast(H.SCULPT_search_sphere_cb, C.c_void_p), C.byref(data))
    assert got == list(o.gather_sphere(c, 0.3))
    s.close()


def test_node_accessors():
    m = meshgen.grid(33)
    s = capi.SculptSession(m, leaf_limit=100)
    H = s.H
    H.BKE_pbvh_node_num_verts.argtypes = [C.POINTER(capi.PBVH), C.POINTER(capi.PBVHNode), capi.c_int_p, capi.c_int_p]
    H.BKE_pbvh_node_get_BB.argtypes = [C.POINTER(capi.PBVHNode), capi.c_float_p, capi.c_float_p]
    na = s.node_arrays()
    leaf = int(np.nonzero(na["flag"] & 1)[0][0])
    node = C.pointer(s.pbvh.contents.nodes[leaf])
    u, t = C.c_int(), C.c_int()
    H.BKE_pbvh_node_num_verts(s.pbvh, node, C.byref(u), C.byref(t))
    assert u.value == na["uniq_verts"][leaf] and t.value == na["uniq_verts"][leaf] + na["face_verts"][leaf]
    lo, hi = np.zeros(3, np.float32), np.zeros(3, np.float32)
    H.BKE_pbvh_node_get_BB(node, capi.fptr(lo), capi.fptr(hi))
    assert np.array_equal(np.concatenate([lo, hi]), na["vb"][leaf])
    H.BKE_pbvh_node_mark_update(node)
    f = s.pbvh.contents.nodes[leaf].flag
    assert f & capi.PBVH_UpdateNormals and f & capi.PBVH_UpdateBB and f & capi.PBVH_UpdateOriginalBB
    s.close()


def test_host_build_matches_the_oracle_at_sizes_that_run_the_task_parallel_partition():
    """BKE_pbvh_build_mesh splits subtrees of >= 65536 prims into OpenMP tasks; the tree, the prim order and the leaves'
    vertex lists must still be the serial build's (the oracle's), default leaf limit and a small one"""
    import numpy as np
    from dune_sculpt_b200 import capi, meshgen
    from oracle_py import Oracle
    for mesh, ll in ((meshgen.grid(420), 0), (meshgen.icosphere(130, noise=0.002), 2500)):
        assert mesh.totloop - 2 * mesh.totpoly >= 3 * 65536   # looptris: several levels of tasks
        ses = capi.SculptSession(mesh, leaf_limit=ll, device=None)
        orc = Oracle(mesh, leaf_limit=ll)
        na, nb = orc.node_arrays(), ses.node_arrays()
        for k in ("vb", "orig_vb", "children_offset", "totprim", "uniq_verts", "face_verts", "prim_offset"):
            assert np.array_equal(na[k], nb[k]), k
        assert np.array_equal(orc.prim_indices(), ses.prim_indices())
        for n in np.nonzero(na["flag"] & 1)[0]:
            cnt = int(na["uniq_verts"][n] + na["face_verts"][n])
            assert np.array_equal(orc.node_vert_indices(int(n), cnt), ses.node_vert_indices(int(n))), n
        ses.close()
        orc.close()


def _patchy(n, seed, nmat=3):
    rng = np.random.default_rng(seed)
    mat = np.zeros(n, dtype=np.int16)
    flag = np.zeros(n, dtype=np.uint8)
    i = 0
    while i < n:
        run = int(rng.integers(1, max(2, n // 12)))
        mat[i:i + run] = rng.integers(0, nmat)
        flag[i:i + run] = rng.integers(0, 2)
        i += run
    return mat, flag


@pytest.mark.parametrize("mk,ll", [(lambda: meshgen.grid(96), 700), (lambda: meshgen.mixed_grid(40), 300)])
def test_build_mesh_material_split_and_hidden_leaves(mk, ll):
    """leaf_needs_material_split / partition_indices_material (pbvh.c:2091-2132, 2329-2359) and the fully-hidden leaf flag
    (pbvh.c:2188-2208) -- the oracle these are compared with is itself pinned to the reference (tests/test_ref_pin.py)"""
    m = mk()
    mat, pflag = _patchy(m.totpoly, 1)
    co = np.asarray(m.co)
    vflag = np.where((co[:, 0] > 0.2) & (co[:, 1] > -0.1), 16, 0).astype(np.uint8)
    o = Oracle(m, leaf_limit=ll, poly_mat=mat, poly_flag=pflag, vert_flag=vflag)
    s = capi.SculptSession(m, leaf_limit=ll, poly_mat=mat, poly_flag=pflag, vert_flag=vflag)
    a, b = o.node_arrays(), s.node_arrays()
    assert o.totnode == s.totnode
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(o.prim_indices(), s.prim_indices())
    leaves = np.nonzero(a["flag"] & 1)[0]
    assert 0 < ((a["flag"][leaves] & capi.PBVH_FullyHidden) != 0).sum() < leaves.size
    for i in leaves:
        assert np.array_equal(o.node_vert_indices(int(i), a["uniq_verts"][i] + a["face_verts"][i]), s.node_vert_indices(int(i)))
    s.close()
    o.close()


def test_build_grids_material_split_and_hidden_leaves():
    from oracle_py import GridOracle
    mr = meshgen.multires_cube(2, 4)
    mat, gflag = _patchy(mr.totgrid, 2)
    gs2 = mr.grid_size * mr.grid_size
    hidden = np.zeros(mr.totelem, dtype=np.uint8)
    hidden[:gs2 * 24] = 1
    hidden[gs2 * 40 + 5:gs2 * 40 + 9] = 1
    for ll in (0, 2):
        o = GridOracle(mr, leaf_limit=ll, recalc_normals=False, grid_mat=mat, grid_flag=gflag, hidden=hidden)
        s = capi.GridSession(mr, leaf_limit=ll, device=None, grid_mat=mat, grid_flag=gflag, hidden=hidden)
        a, b = o.node_arrays(), s.node_arrays()
        assert o.totnode == s.totnode
        for k in a:
            assert np.array_equal(a[k], b[k]), (ll, k)
        assert np.array_equal(o.prim_indices(), s.prim_indices())
        s.close()
        o.close()
