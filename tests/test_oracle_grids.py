"""CPU: the grids (multires CCG) part of the oracle -- invariants of kernel/intern/subdiv_ccg.c and of
BKE_pbvh_build_grids (pbvh.c:2516-2561) -- and the host library's grids PBVH against it."""
import numpy as np

from dune_sculpt_b200 import capi, meshgen
from oracle_py import GridOracle


def _dups_equal(mr, arr):
    gs = mr.grid_size
    rows = mr.edge_elems.reshape(-1, 2 * gs)
    for e in range(mr.edge_off.shape[0] - 1):
        a, b = mr.edge_off[e], mr.edge_off[e + 1]
        for f in range(a + 1, b):
            if not np.array_equal(arr[rows[a]], arr[rows[f]]):
                return False
    for v in range(mr.cvert_off.shape[0] - 1):
        el = mr.cvert_elems[mr.cvert_off[v]:mr.cvert_off[v + 1]]
        if not np.all(arr[el] == arr[el[0]]):
            return False
    gs2 = gs * gs
    for f in range(mr.face_start.shape[0]):
        n = mr.face_num[f]
        for c in range(n):
            prev = mr.face_start[f] + (c + n - 1) % n
            cur = mr.face_start[f] + c
            i = np.arange(gs)
            if not np.array_equal(arr[prev * gs2 + i], arr[cur * gs2 + i * gs]):
                return False
    return True


def test_generator_tables_are_consistent():
    mr = meshgen.multires_cube(1, 4, with_mask=True)
    assert mr.grid_size == 9 and mr.totgrid == 6 * 4 * 4
    assert np.all(np.diff(mr.edge_off) == 2)            # closed cube: every coarse edge has two faces
    assert set(np.diff(mr.cvert_off).tolist()) == {3, 4}  # cube corners have 3 faces, the other verts 4
    assert _dups_equal(mr, mr.co) and _dups_equal(mr, mr.mask)


def test_grids_pbvh_structure_and_host_library_agree():
    mr = meshgen.multires_cube(1, 4)
    for leaf_limit in (0, 5, 1):
        orc = GridOracle(mr, leaf_limit=leaf_limit)
        ses = capi.GridSession(mr, leaf_limit=leaf_limit, device=None)
        na, nb = orc.node_arrays(), ses.node_arrays()
        for k in ("vb", "orig_vb", "children_offset", "totprim", "uniq_verts", "face_verts"):
            assert np.array_equal(na[k], nb[k]), k
        assert np.array_equal(na["flag"] & 1, nb["flag"] & 1)
        assert np.array_equal(orc.prim_indices(), ses.prim_indices())
        leaf = (na["flag"] & 1) != 0
        # pbvh.c:2533: leaf_limit = max(LEAF_LIMIT / gridsize^2, 1); every grid in exactly one leaf
        lim = leaf_limit if leaf_limit else max(10000 // (mr.grid_size ** 2), 1)
        assert na["totprim"][leaf].max() <= lim
        assert sorted(orc.prim_indices().tolist()) == list(range(mr.totgrid))
        assert np.array_equal(na["uniq_verts"][leaf], na["totprim"][leaf] * mr.grid_size ** 2)
        ses.close()
        orc.close()


def test_ccg_normals_of_a_flat_grid_and_unit_length_on_the_sphere():
    mr = meshgen.multires_cube(1, 3, noise=0.0, spherify=False)   # flat faces
    orc = GridOracle(mr)
    no, co = orc.no(), orc.co()
    gs2 = mr.grid_size ** 2
    # interior elements of a +Z face grid point along +Z exactly
    g_top = [g for g in range(mr.totgrid) if np.all(co[g * gs2:(g + 1) * gs2, 2] == 1.0)]
    assert g_top
    inner = g_top[0] * gs2 + 1 * mr.grid_size + 1
    assert np.array_equal(no[inner], np.array([0, 0, 1], np.float32))
    orc.close()
    mr = meshgen.multires_cube(1, 4)
    orc = GridOracle(mr)
    ln = np.linalg.norm(orc.no(), axis=1)
    assert ln.min() > 0.99 and ln.max() < 1.001    # mean of up to four unit quad normals, not renormalised
    assert _dups_equal(mr, orc.no())
    orc.close()


def test_dab_keeps_duplicates_stitched_and_updates_only_gathered_leaves():
    mr = meshgen.multires_cube(1, 4, with_mask=True)
    orc = GridOracle(mr, leaf_limit=4)
    co0, no0 = orc.co(), orc.no()
    d = capi.make_dab(capi.TOOL_DRAW, (0.0, 0.0, 1.0), 0.6, bstrength=0.3, view_normal=(0, 0, 1))
    orc.stroke_begin()
    nh = orc.dab(d)
    orc.stroke_end()
    assert 0 < nh < (orc.node_arrays()["flag"] & 1).sum()
    co1 = orc.co()
    assert np.abs(co1 - co0).max() > 0.01
    assert _dups_equal(mr, co1) and _dups_equal(mr, orc.no()) and _dups_equal(mr, orc.mask())
    na = orc.node_arrays()
    assert not np.any(na["flag"] & (capi.PBVH_UpdateNormals | capi.PBVH_UpdateBB))
    # leaf boxes contain their elements
    prims = orc.prim_indices()
    gs2 = mr.grid_size ** 2
    for n in np.nonzero(na["flag"] & 1)[0]:
        g = prims[na["prim_offset"][n]:na["prim_offset"][n] + na["totprim"][n]]
        el = (g[:, None] * gs2 + np.arange(gs2)[None, :]).reshape(-1)
        assert np.all(co1[el].min(axis=0) == na["vb"][n, :3]) and np.all(co1[el].max(axis=0) == na["vb"][n, 3:])
    assert orc.vertex_dabs() == int(na["uniq_verts"][orc.hits()].sum())
    orc.close()


def test_openmp_and_single_thread_agree_on_grids():
    mr = meshgen.multires_cube(1, 4)
    outs = []
    for threads in (1, 4):
        orc = GridOracle(mr, leaf_limit=3, threads=threads)
        orc.stroke_begin()
        for k in range(4):
            orc.dab(capi.make_dab(capi.TOOL_DRAW, (0.2 * k - 0.3, 0.1, 1.0), 0.5, bstrength=0.2, view_normal=(0, 0, 1)))
        orc.stroke_end()
        outs.append((orc.co(), orc.no()))
        orc.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def _welded_graph(mr):
    """independent of subdiv_ccg.c: weld the duplicated elements by position (the generator makes them
    bit-identical) and collect the edges of every grid's quad lattice between welded ids"""
    gs, gs2 = mr.grid_size, mr.grid_size ** 2
    _, weld = np.unique(mr.co.view(np.dtype((np.void, 12))).reshape(-1), return_inverse=True)
    el = np.arange(mr.totelem, dtype=np.int64).reshape(mr.totgrid, gs, gs)
    pairs = np.concatenate([np.stack([el[:, :, :-1].reshape(-1), el[:, :, 1:].reshape(-1)], 1),
                            np.stack([el[:, :-1, :].reshape(-1), el[:, 1:, :].reshape(-1)], 1)])
    adj = {}
    for a, b in weld[pairs]:
        adj.setdefault(int(a), set()).add(int(b))
        adj.setdefault(int(b), set()).add(int(a))
    return weld, adj


def test_grid_neighbours_are_the_welded_lattice_neighbours():
    """KERNEL_subdiv_ccg_neighbor_coords_get (subdiv_ccg.c:1882-1909) restated: for EVERY element -- interior,
    inner boundaries, face centres, coarse edges (both edge directions), grid corners on edges, coarse vertices
    of valence 3 and 4, open boundaries -- the neighbours are exactly the lattice neighbours of the welded
    vertex, each once, none of them a duplicate of the element itself"""
    for mr in (meshgen.multires_cube(1, 3), meshgen.multires_plane(3, 3), meshgen.multires_cube_n(3, 2)):
        orc = GridOracle(mr, leaf_limit=4)
        weld, adj = _welded_graph(mr)
        gs, gs1 = mr.grid_size, mr.grid_size - 1
        for e in range(mr.totelem):
            nb = orc.neighbors(e)
            w = weld[nb]
            assert len(set(w.tolist())) == len(w), "element %d: a neighbour listed twice" % e
            assert set(w.tolist()) == adj[int(weld[e])], "element %d" % e
            x, y = (e % (gs * gs)) % gs, (e % (gs * gs)) // gs
            if 0 < x < gs1 and 0 < y < gs1:   # subdiv_ccg.c:1870-1880: prev row, next row, prev col, next col
                assert nb.tolist() == [e - gs, e + gs, e - 1, e + 1]
        orc.close()


def test_grid_boundary_elements_follow_the_coarse_mesh():
    closed = GridOracle(meshgen.multires_cube(1, 3), leaf_limit=4)
    assert not any(closed.is_boundary(e) for e in range(closed.totvert))
    closed.close()
    mr = meshgen.multires_plane(3, 3, noise=0.0)
    orc = GridOracle(mr, leaf_limit=4)
    lim = np.abs(mr.co[:, :2]).max()
    on_rim = (np.abs(np.abs(mr.co[:, 0]) - lim) < 1e-3) | (np.abs(np.abs(mr.co[:, 1]) - lim) < 1e-3)
    b = np.array([orc.is_boundary(e) for e in range(mr.totelem)])
    # every rim element is a boundary element; the converse fails only where the reference's rule says so: a
    # coarse edge whose two ends are boundary vertices counts even if it runs through the interior (none here
    # with 3 x 3 base quads: inner edges touch at most one rim vertex)
    assert np.array_equal(b, on_rim)
    orc.close()


def test_grid_smooth_brush_oracle_runs_and_is_thread_invariant():
    from dune_sculpt_b200 import stroke
    mr = meshgen.multires_cube(1, 4)
    d = capi.make_dab(capi.TOOL_SMOOTH, (0.0, 0.0, 1.0), 0.6, bstrength=0.6)
    out = []
    for threads in (1, 4):
        orc = GridOracle(mr, leaf_limit=4, threads=threads)
        orc.stroke_begin(None)
        for _ in range(3):
            orc.dab(d)
        orc.stroke_end()
        out.append((orc.co(), orc.no()))
        assert _dups_equal(mr, out[-1][0]), "stitch keeps duplicated elements equal"
        orc.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert not np.array_equal(out[0][0], mr.co)


def test_host_library_neighbours_match_the_oracle_and_duplicates_are_the_other_copies():
    """the host library's BKE_subdiv_ccg_neighbor_coords_get against the oracle's restatement (order included), and
    include_duplicates = true appends exactly the other elements welded to the same vertex"""
    for mr in (meshgen.multires_cube(1, 3), meshgen.multires_plane(3, 3)):
        orc = GridOracle(mr, leaf_limit=4)
        ses = capi.GridSession(mr, leaf_limit=4, device=None)
        weld, _ = _welded_graph(mr)
        copies = {}
        for e, w in enumerate(weld):
            copies.setdefault(int(w), set()).add(e)
        for e in range(mr.totelem):
            nb, nd = ses.neighbors(e)
            assert nd == 0 and nb.tolist() == orc.neighbors(e).tolist(), "element %d" % e
            nb2, nd2 = ses.neighbors(e, True)
            assert nb2[:nb2.size - nd2].tolist() == nb.tolist()
            assert set(nb2[nb2.size - nd2:].tolist()) == copies[int(weld[e])] - {e} and nd2 == len(copies[int(weld[e])]) - 1, e
        ses.close()
        orc.close()


def test_multires_write_back_from_the_ccg_without_a_device():
    """multires_reshape_assign_final_coords_from_ccg (multires_reshape_ccg.c:10-70): element (x, y) of grid g lands in
    mdisps[g].disps[y * gs + x], its mask in grid_paint_masks[g].data[...]"""
    mr = meshgen.multires_cube(1, 3, with_mask=True)
    ses = capi.GridSession(mr, leaf_limit=4, device=None)
    disps, masks = ses.multires_write_back()
    assert np.array_equal(disps.reshape(-1, 3), mr.co) and np.array_equal(masks.reshape(-1), mr.mask)
    ses.close()


def test_grids_raycast_oracle_closed_forms():
    """pbvh_grids_node_raycast restated (pbvh.c:4102-4200): a ray down the z axis onto the spherified multires cube
    hits just inside the unit sphere (the quads are chords), at the element nearest to the pole; rays that leave, or
    whose search starts nearer than the surface, hit nothing; mid-stroke the original coordinates answer for the
    leaves that carry an undo node"""
    mr = meshgen.multires_cube(1, 5, noise=0.0)
    orc = GridOracle(mr, leaf_limit=2)
    hit = orc.raycast((0.02, 0.01, 3.0), (0.0, 0.0, -1.0))
    assert hit is not None and 2.0 <= hit["depth"] < 2.01
    assert abs(abs(hit["normal"][2]) - 1.0) < 0.02
    v = mr.co[hit["vertex"]]
    assert v[2] > 0.99 and np.hypot(v[0] - 0.02, v[1] - 0.01) < 0.08
    assert hit["face"] == hit["vertex"] // (mr.grid_size ** 2)          # the active grid holds the active element
    assert orc.raycast((0.02, 0.01, 3.0), (0.0, 0.0, 1.0)) is None
    assert orc.raycast((0.02, 0.01, 3.0), (0.0, 0.0, -1.0), max_depth=1.5) is None
    # from inside the sphere the far side is hit (no back-face culling in the ray test)
    inside = orc.raycast((0.0, 0.0, 0.0), (0.0, 1.0, 0.0))
    assert inside is not None and 0.99 < inside["depth"] <= 1.0
    # a draw dab raises the pole; the original coordinates still answer with the rest surface
    d0 = hit["depth"]
    orc.stroke_begin(None)
    orc.dab(capi.make_dab(capi.TOOL_DRAW, (0.0, 0.0, 1.0), 0.4, bstrength=0.5, view_normal=(0.0, 0.0, 1.0)))
    now = orc.raycast((0.02, 0.01, 3.0), (0.0, 0.0, -1.0))
    orig = orc.raycast((0.02, 0.01, 3.0), (0.0, 0.0, -1.0), original=True)
    assert now["depth"] < d0 - 0.01 and orig["depth"] == d0
    orc.stroke_end()
    orc.close()


def test_grids_draw_buffer_oracle_layout():
    """gpu_pbvh_grid_buffers_update restated (gpu_buffers.c:548-725): smooth = one 36-byte record per element with its
    position, i16 normal and u8 mask; flat = four records per quad sharing the quad normal (which faces outwards on the
    spherified cube) and the mean mask; the draw flags are cleared"""
    mr = meshgen.multires_cube(1, 3, noise=0.0, with_mask=True)
    orc = GridOracle(mr, leaf_limit=2)
    na = orc.node_arrays()
    leaf = int(np.nonzero(na["flag"] & 1)[0][3])
    grids = orc.prim_indices()[na["prim_offset"][leaf]:na["prim_offset"][leaf] + na["totprim"][leaf]]
    gs, gs2 = mr.grid_size, mr.grid_size ** 2
    buf = orc.draw_buffer(leaf, int(na["totprim"][leaf]), smooth=True)
    el = np.concatenate([np.arange(g * gs2, (g + 1) * gs2) for g in grids])
    assert np.array_equal(buf[:, 0:12].copy().view(np.float32).reshape(-1, 3), orc.co()[el])
    assert np.array_equal(buf[:, 16:22].copy().view(np.int16).reshape(-1, 3), (orc.no()[el] * np.float32(32767.0)).astype(np.int16))
    assert np.array_equal(buf[:, 22], (orc.mask()[el] * np.float32(255)).astype(np.uint8))
    assert (buf[:, 32:35] == 255).all() and (buf[:, 24:32] == 0).all()
    assert not (orc.node_arrays()["flag"][leaf] & (capi.PBVH_UpdateDrawBuffers | capi.PBVH_RebuildDrawBuffers))
    flat = orc.draw_buffer(leaf, int(na["totprim"][leaf]), smooth=False)
    assert flat.shape[0] == len(grids) * (gs - 1) ** 2 * 4
    pos = flat[:, 0:12].copy().view(np.float32).reshape(-1, 4, 3)
    nor = flat[:, 16:22].copy().view(np.int16).reshape(-1, 4, 3).astype(np.float32) / 32767.0
    assert np.array_equal(pos[0, 0], orc.co()[grids[0] * gs2]) and np.array_equal(pos[0, 2], orc.co()[grids[0] * gs2 + gs + 1])
    assert (nor[:, 0] == nor[:, 1]).all() and (nor[:, 0] == nor[:, 3]).all()
    centre = pos.mean(axis=1)
    assert (np.einsum("ij,ij->i", nor[:, 0], centre / np.linalg.norm(centre, axis=1, keepdims=True)) > 0.9).all()
    assert (flat[:, 24:32] == 255).all()
    orc.close()
