"""Pins the oracle to the REFERENCE ITSELF: oracle/_ref/libref.so holds the reference's own functions, cut verbatim
out of /root/reference (oracle/ref_extract.py; the manifest with file:line of every function is
oracle/_ref/manifest.txt) and compiled behind a shim for the absent headers.  Same arrays in, bit-equal results out:

  build            BKE_pbvh_build_mesh / _grids, build_sub, partition_indices, build_mesh_leaf_node  pbvh.c:2070-2561
  traversal        BKE_pbvh_search_gather, pbvh_iter_next                                           pbvh.c:2622-2767
  normals          pbvh_faces_update_normals, BKE_mesh_calc_poly_normal, normal_tri/quad_v3        pbvh.c:2912-3036
  bounds           BKE_pbvh_update_bounds, update_node_vb, pbvh_flush_bb                            pbvh.c:2026-2046, 3124-3339
  CCG normals      subdiv_ccg_recalc_inner_face_normals, subdiv_ccg_average_inner_face_normals      subdiv_ccg.c:670-740
  CCG averaging    subdiv_ccg_average_inner_face_grids, _grids_boundary, _grids_corners             subdiv_ccg.c:873-1104

The library is built where /root/reference exists and travels as a prebuilt .so; without either the tests skip."""
import numpy as np
import pytest

from dune_sculpt_b200 import meshgen
from oracle_py import GridOracle, Oracle
import ref_py

pytestmark = pytest.mark.skipif(ref_py.lib() is None, reason="oracle/_ref/libref.so absent and /root/reference not here to build it")

F_Leaf, F_UpdateNormals, F_UpdateBB, F_UpdateOriginalBB = 1, 2, 4, 8
NODE_KEYS = ("children_offset", "flag", "prim_offset", "totprim", "uniq_verts", "face_verts", "vb", "orig_vb")


def _same_tree(orc, ref):
    assert orc.totnode == ref.totnode
    a, b = orc.node_arrays(), ref.node_arrays()
    for k in NODE_KEYS:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(orc.prim_indices(), ref.prim_indices())
    return a


MESHES = {
    "grid": lambda: (meshgen.grid(96), 700),
    "cube": lambda: (meshgen.cube(5), 500),
    "icosphere": lambda: (meshgen.icosphere(24, noise=0.002), 900),
    "mixed": lambda: (meshgen.mixed_grid(40), 300),
}


@pytest.mark.parametrize("name", sorted(MESHES))
def test_build_mesh_is_the_references(name):
    mesh, leaf_limit = MESHES[name]()
    orc, ref = Oracle(mesh, leaf_limit=leaf_limit), ref_py.RefMesh(mesh, leaf_limit=leaf_limit)
    try:
        na = _same_tree(orc, ref)
        leaves = np.nonzero(na["flag"] & F_Leaf)[0]
        assert leaves.size > 8
        for n in leaves:
            cnt = int(na["uniq_verts"][n] + na["face_verts"][n])
            assert np.array_equal(orc.node_vert_indices(int(n), cnt), ref.node_vert_indices(int(n), cnt)), n
            assert np.array_equal(orc.node_face_vert_indices(int(n), int(na["totprim"][n])),
                                  ref.node_face_vert_indices(int(n), int(na["totprim"][n]))), n
    finally:
        orc.close()
        ref.close()


def test_default_leaf_limit_tree():
    """LEAF_LIMIT = 10000 (pbvh.c:1952), the limit every GPU parity test and the bench build with"""
    mesh = meshgen.grid(300)
    orc, ref = Oracle(mesh), ref_py.RefMesh(mesh)
    try:
        assert ref.L.ref_leaf_limit_used(ref.p) == 10000
        _same_tree(orc, ref)
    finally:
        orc.close()
        ref.close()


def test_search_gather_order_and_pruning():
    mesh = meshgen.icosphere(20)
    orc, ref = Oracle(mesh, leaf_limit=400), ref_py.RefMesh(mesh, leaf_limit=400)
    try:
        na = _same_tree(orc, ref)
        leaves = np.nonzero(na["flag"] & F_Leaf)[0]
        for n in leaves[::7]:
            orc.set_node_flag(int(n), 1 << 10)  # FullyHidden
            ref.set_node_flag(int(n), 1 << 10)
        rng = np.random.default_rng(3)
        nonempty = 0
        for i in range(60):
            c = rng.normal(size=3)
            c = (c / np.linalg.norm(c) * rng.uniform(0.6, 1.2)).astype(np.float32)
            r = np.float32(rng.uniform(0.02, 0.9))
            for ignore in (True, False):
                a, b = orc.gather_sphere(c, float(r * r), ignore=ignore), ref.gather_sphere(c, float(r * r), ignore=ignore)
                assert np.array_equal(a, b), (i, ignore)
                nonempty += a.size > 0
        assert nonempty > 60
        # empty result: NULL, 0 (pbvh.c:2760-2766)
        assert ref.gather_sphere(np.array([9, 9, 9], np.float32), 0.01).size == 0
    finally:
        orc.close()
        ref.close()


@pytest.mark.parametrize("name", ["grid", "icosphere", "mixed"])
def test_update_normals_and_bounds_are_the_references(name):
    """marks + moved vertices, then pbvh_faces_update_normals and BKE_pbvh_update_bounds: normals of the marked verts
    (others untouched), vert_bitmap cleared, flags cleared, leaf boxes, inner boxes of the flushed ancestors, orig_vb"""
    mesh, leaf_limit = MESHES[name]()
    orc, ref = Oracle(mesh, leaf_limit=leaf_limit), ref_py.RefMesh(mesh, leaf_limit=leaf_limit)
    try:
        na = _same_tree(orc, ref)
        ref.set_no(orc.no())
        rng = np.random.default_rng(11)
        co = np.array(mesh.co, dtype=np.float32)
        for rnd in range(3):
            centre = co[rng.integers(0, mesh.totvert)]
            r2 = float(np.float32(0.25) ** 2) if rnd else float(np.float32(0.6) ** 2)
            hit = orc.gather_sphere(centre, r2)
            assert np.array_equal(hit, ref.gather_sphere(centre, r2))
            inside = np.nonzero(((co - centre) ** 2).sum(axis=1) <= r2)[0]
            # only verts that are unique in a gathered leaf move (what a brush does)
            uniq = np.concatenate([orc.node_vert_indices(int(n), int(na["uniq_verts"][n])) for n in hit]) if hit.size else np.zeros(0, np.int32)
            moved = np.intersect1d(inside, uniq)
            assert moved.size > 0
            co[moved] += rng.normal(scale=0.01, size=(moved.size, 3)).astype(np.float32)
            orc.set_co(co)
            ref.set_co(co)
            for n in hit:
                orc.L.or_node_mark_update(orc.p, int(n))
                ref.node_mark_update(int(n))
            for v in moved:
                orc.L.or_vert_mark_update(orc.p, int(v))
                ref.vert_mark_update(int(v))
            assert np.array_equal(orc.node_arrays()["flag"], ref.node_arrays()["flag"])
            orc.update_normals()
            ref.update_normals()
            assert np.array_equal(orc.no(), ref.no()), "round %d: vertex normals differ in bits" % rnd
            assert not any(ref.vert_marked(int(v)) for v in moved[:50])
            orc.update_bounds(F_UpdateBB)
            ref.update_bounds(F_UpdateBB)
            a, b = orc.node_arrays(), ref.node_arrays()
            for k in ("flag", "vb", "orig_vb"):
                assert np.array_equal(a[k], b[k]), (rnd, k)
            orc.update_bounds(F_UpdateOriginalBB)
            ref.update_bounds(F_UpdateOriginalBB)
            a, b = orc.node_arrays(), ref.node_arrays()
            for k in ("flag", "vb", "orig_vb"):
                assert np.array_equal(a[k], b[k]), (rnd, k)
    finally:
        orc.close()
        ref.close()


GRIDS = {
    "cube": lambda: meshgen.multires_cube(2, 4, with_mask=True),
    "cube_nomask": lambda: meshgen.multires_cube(1, 5),
    "plane": lambda: meshgen.multires_plane(4, 3, with_mask=True),
}


@pytest.mark.parametrize("name", sorted(GRIDS))
def test_build_grids_is_the_references(name):
    mr = GRIDS[name]()
    for leaf_limit in (0, 3):
        orc, ref = GridOracle(mr, leaf_limit=leaf_limit, recalc_normals=False), ref_py.RefGrids(mr, leaf_limit=leaf_limit)
        try:
            _same_tree(orc, ref)
            c = np.array(mr.co[mr.totelem // 3], dtype=np.float32)
            for r in (0.05, 0.3, 1.0):
                assert np.array_equal(orc.gather_sphere(c, r * r), ref.gather_sphere(c, r * r))
        finally:
            orc.close()
            ref.close()


@pytest.mark.parametrize("name", sorted(GRIDS))
def test_ccg_normals_and_averaging_are_the_references(name):
    mr = GRIDS[name]()
    orc, ref = GridOracle(mr, recalc_normals=False), ref_py.RefGrids(mr)
    try:
        # displaced coordinates so that duplicated elements disagree before the averaging
        rng = np.random.default_rng(5)
        co = (np.array(mr.co, dtype=np.float32) + rng.normal(scale=0.003, size=(mr.totelem, 3)).astype(np.float32))
        orc.set_co(co)
        ref.set_co(co)
        orc.L.or_grids_inner_normals(orc.p)
        ref.inner_normals()
        assert np.array_equal(orc.no(), ref.no()), "inner CCG normals differ in bits"
        assert np.abs(orc.no()).max() > 0.5
        orc.L.or_grids_average_all(orc.p)
        ref.average_all()
        assert np.array_equal(orc.co(), ref.co()), "averaged coordinates"
        assert np.array_equal(orc.no(), ref.no()), "averaged normals"
        if mr.mask is not None:
            assert np.array_equal(orc.mask(), ref.mask()), "averaged mask"
        assert not np.array_equal(orc.co(), co)
    finally:
        orc.close()
        ref.close()


def _materials(n, seed, nmat=3):
    """patchy materials + smooth flags: runs of equal material so that some leaves need a split and some do not"""
    rng = np.random.default_rng(seed)
    mat = np.zeros(n, dtype=np.int16)
    flag = np.zeros(n, dtype=np.uint8)
    i = 0
    while i < n:
        run = int(rng.integers(1, max(2, n // 12)))
        mat[i:i + run] = rng.integers(0, nmat)
        flag[i:i + run] = rng.integers(0, 2)  # ME_SMOOTH
        i += run
    return mat, flag


@pytest.mark.parametrize("name", ["grid", "icosphere", "mixed"])
def test_material_split_and_hidden_leaves_mesh(name):
    """leaf_needs_material_split / partition_indices_material (pbvh.c:2091-2132, 2329-2359) and the fully-hidden
    flag of build_mesh_leaf_node (pbvh.c:2188-2208, 2235)"""
    mesh, leaf_limit = MESHES[name]()
    mat, pflag = _materials(mesh.totpoly, 1)
    co = np.asarray(mesh.co)
    vflag = np.where((co[:, 0] > 0.2) & (co[:, 1] > -0.1), 16, 0).astype(np.uint8)  # ME_HIDE
    orc = Oracle(mesh, leaf_limit=leaf_limit, poly_mat=mat, poly_flag=pflag, vert_flag=vflag)
    ref = ref_py.RefMesh(mesh, leaf_limit=leaf_limit, poly_mat=mat, poly_flag=pflag, vert_flag=vflag)
    plain = Oracle(mesh, leaf_limit=leaf_limit)
    try:
        na = _same_tree(orc, ref)
        assert orc.totnode > plain.totnode, "the materials split no leaf"
        leaves = np.nonzero(na["flag"] & F_Leaf)[0]
        hidden = (na["flag"][leaves] & (1 << 10)) != 0
        assert 0 < hidden.sum() < leaves.size
        for n in leaves:
            cnt = int(na["uniq_verts"][n] + na["face_verts"][n])
            assert np.array_equal(orc.node_vert_indices(int(n), cnt), ref.node_vert_indices(int(n), cnt)), n
        # every leaf holds one material / shading mode
        prim, tp = orc.prim_indices(), orc.tri_poly()
        for n in leaves:
            polys = tp[prim[na["prim_offset"][n]:na["prim_offset"][n] + na["totprim"][n]]]
            assert np.unique(mat[polys]).size == 1 and np.unique(pflag[polys] & 1).size == 1
    finally:
        orc.close()
        ref.close()
        plain.close()


def test_material_split_and_hidden_leaves_grids():
    mr = meshgen.multires_cube(2, 4)
    mat, gflag = _materials(mr.totgrid, 2)
    gs2 = mr.grid_size * mr.grid_size
    hidden = np.zeros(mr.totelem, dtype=np.uint8)
    hidden[:gs2 * 24] = 1                    # the grids of six faces wholly hidden
    hidden[gs2 * 40 + 5:gs2 * 40 + 9] = 1    # one grid partly
    for leaf_limit in (0, 2):
        orc = GridOracle(mr, leaf_limit=leaf_limit, recalc_normals=False, grid_mat=mat, grid_flag=gflag, hidden=hidden)
        ref = ref_py.RefGrids(mr, leaf_limit=leaf_limit, grid_mat=mat, grid_flag=gflag, hidden=hidden)
        try:
            na = _same_tree(orc, ref)
            leaves = np.nonzero(na["flag"] & F_Leaf)[0]
            nhid = int(((na["flag"][leaves] & (1 << 10)) != 0).sum())
            assert nhid < leaves.size and (leaf_limit == 0 or nhid > 0)
        finally:
            orc.close()
            ref.close()
