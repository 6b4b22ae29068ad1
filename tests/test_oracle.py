"""CPU checks of the oracle itself: invariants the reference code guarantees, closed forms, and
single-thread vs OpenMP agreement (SURVEY.md section 8c pins 1-3).  No GPU."""
import numpy as np
import pytest

from dune_sculpt_b200 import capi, meshgen, stroke
from oracle_py import Oracle, lib as oracle_lib


def _meshes():
    return [("grid", meshgen.grid(97), 300), ("cube", meshgen.cube(5), 150), ("ico", meshgen.icosphere(20), 111)]


@pytest.mark.parametrize("name,mesh,ll", _meshes(), ids=[m[0] for m in _meshes()])
def test_pbvh_invariants(name, mesh, ll):
    o = Oracle(mesh, leaf_limit=ll)
    na = o.node_arrays()
    leaf = (na["flag"] & 1) != 0
    owner = np.full(mesh.totvert, -1)
    co = o.co()
    prims = o.prim_indices()
    assert np.array_equal(np.sort(prims), np.arange(o.tottri))          # prim_indices is a permutation
    expect = 0
    for i in np.nonzero(leaf)[0][np.argsort(na["prim_offset"][leaf])]:
        assert na["prim_offset"][i] == expect and 0 < na["totprim"][i] <= ll  # leaves tile prim_indices, leaf limit
        expect += na["totprim"][i]
        vi = o.node_vert_indices(i, na["uniq_verts"][i] + na["face_verts"][i])
        u = vi[:na["uniq_verts"][i]]
        assert (owner[u] == -1).all()                                     # unique in exactly one leaf (pbvh.c:2157-2165)
        owner[u] = i
        assert len(set(vi.tolist())) == vi.size
        assert (co[vi] >= na["vb"][i, :3] - 0).all() and (co[vi] <= na["vb"][i, 3:]).all()  # leaf box holds its verts
        fvi = o.node_face_vert_indices(i, na["totprim"][i])
        tv = o.tri_verts()[prims[na["prim_offset"][i]:na["prim_offset"][i] + na["totprim"][i]]]
        assert np.array_equal(vi[fvi], tv)                                # face_vert_indices index vert_indices
    assert expect == o.tottri and (owner >= 0).all() and na["uniq_verts"].sum() == mesh.totvert
    for i in np.nonzero(~leaf)[0]:
        c = na["children_offset"][i]
        assert np.array_equal(na["vb"][i, :3], np.minimum(na["vb"][c, :3], na["vb"][c + 1, :3]))   # pbvh.c:2040-2043
        assert np.array_equal(na["vb"][i, 3:], np.maximum(na["vb"][c, 3:], na["vb"][c + 1, 3:]))
    ln = np.linalg.norm(o.no().astype(np.float64), axis=1)
    assert np.all((np.abs(ln - 1.0) < 1e-6) | (ln == 0.0))
    assert np.array_equal(na["vb"], na["orig_vb"])


def test_gather_is_the_flat_leaf_test():
    """the DFS of pbvh.c:2664-2705 with the sphere callback returns exactly the leaves that pass the
    callback themselves, in ascending prim offset -- what the CUDA gather relies on"""
    m = meshgen.cube(5)
    o = Oracle(m, leaf_limit=100)
    na = o.node_arrays()
    leaves = np.nonzero(na["flag"] & 1)[0]
    leaves = leaves[np.argsort(na["prim_offset"][leaves])]
    rng = np.random.default_rng(1)
    for _ in range(50):
        c = rng.uniform(-1.3, 1.3, size=3).astype(np.float32)
        rsq = np.float32(rng.uniform(0.0, 2.0))
        bb = na["vb"][leaves]
        nearest = np.clip(c[None, :], bb[:, :3], bb[:, 3:])
        t = (c[None, :] - nearest).astype(np.float32)
        d = (t[:, 0] * t[:, 0] + t[:, 1] * t[:, 1]) + t[:, 2] * t[:, 2]
        assert np.array_equal(o.gather_sphere(c, float(rsq)), leaves[d < rsq])


def test_curve_presets_known_values():
    o = Oracle(meshgen.grid(5))
    f = lambda preset, p: o.curve_strength(preset, np.float32((1.0 - p) * 2.0), 2.0)  # noqa: E731  p = 1 - len/r
    for p in (0.0, 0.25, 0.5, 1.0):
        assert f(capi.CURVE_SMOOTH, p) == pytest.approx(3 * p * p - 2 * p ** 3, abs=1e-6)
        assert f(capi.CURVE_SMOOTHER, p) == pytest.approx(p ** 3 * (p * (6 * p - 15) + 10), abs=1e-6)
        assert f(capi.CURVE_SPHERE, p) == pytest.approx(np.sqrt(2 * p - p * p), abs=1e-6)
        assert f(capi.CURVE_ROOT, p) == pytest.approx(np.sqrt(p), abs=1e-6)
        assert f(capi.CURVE_SHARP, p) == pytest.approx(p * p, abs=1e-6)
        assert f(capi.CURVE_LIN, p) == pytest.approx(p, abs=1e-6)
        assert f(capi.CURVE_POW4, p) == pytest.approx(p ** 4, abs=1e-6)
        assert f(capi.CURVE_INVSQUARE, p) == pytest.approx(p * (2 - p), abs=1e-6)
    assert f(capi.CURVE_CONSTANT, 0.3) == 1.0
    assert o.curve_strength(capi.CURVE_CONSTANT, 2.0, 2.0) == 0.0  # p >= len -> 0
    t = np.linspace(0, 1, 257, dtype=np.float32) ** 2
    o.set_custom_curve(t)
    assert f(capi.CURVE_CUSTOM, 0.5) == pytest.approx(0.25, abs=1e-4)  # LUT is indexed by 1 - p (colortools.c:942-965)


def test_draw_closed_form():
    m = meshgen.grid(65, height=0.0)
    o = Oracle(m, leaf_limit=100)
    o.stroke_begin()
    d = capi.make_dab(capi.TOOL_DRAW, (0, 0, 0), 0.4, curve_preset=capi.CURVE_CONSTANT, sculpt_plane=capi.DIR_Z, bstrength=0.25)
    o.dab(d)
    inside = (m.co[:, 0] ** 2 + m.co[:, 1] ** 2) <= np.float32(0.4) ** 2
    co = o.co()
    assert np.array_equal(co[inside, 2], np.full(inside.sum(), np.float32(0.4) * np.float32(0.25), np.float32))
    assert np.array_equal(co[~inside], m.co[~inside])
    assert np.array_equal(np.sort(o.moved()), np.nonzero(inside)[0])
    # area normal of a flat patch is +Z whatever the summation order
    o2 = Oracle(m, leaf_limit=100)
    o2.stroke_begin()
    o2.dab(capi.make_dab(capi.TOOL_DRAW, (0.1, 0.1, 0), 0.5))
    no, _ = o2.last_area()
    assert no[0] == 0.0 and no[1] == 0.0 and abs(no[2] - 1.0) < 2e-7


def test_inflate_sphere_radius_grows():
    m = meshgen.icosphere(32)
    o = Oracle(m, leaf_limit=500)
    o.stroke_begin()
    o.dab(capi.make_dab(capi.TOOL_INFLATE, (0, 0, 1), 0.4, curve_preset=capi.CURVE_CONSTANT, bstrength=0.2))
    inside = np.linalg.norm(m.co - np.array([0, 0, 1], np.float32), axis=1) <= np.float32(0.4)
    grow = np.linalg.norm(o.co().astype(np.float64), axis=1) - 1.0
    assert np.allclose(grow[inside], 0.2 * 0.4, atol=5e-4) and np.all(grow[~inside] == 0.0 + (np.linalg.norm(m.co[~inside].astype(np.float64), axis=1) - 1.0))


def test_undo_membership_and_original_boxes():
    m = meshgen.grid(97)
    o = Oracle(m, leaf_limit=200)
    o.stroke_begin()
    hit = set()
    for d in stroke.c4_tool_stroke(capi.TOOL_DRAW, m.bbox_diag(), dabs=6):
        o.dab(d)
        hit |= set(o.hits().tolist())
        na = o.node_arrays()
        assert np.array_equal(na["orig_vb"], Oracle(m, leaf_limit=200).node_arrays()["vb"])  # orig_vb frozen during the stroke
    assert set(o.touched().tolist()) == hit
    oc = o.orig_co()
    for n in hit:
        na = o.node_arrays()
        u = o.node_vert_indices(n, na["uniq_verts"][n])
        assert np.array_equal(oc[u], m.co[u])  # snapshot = stroke-start coordinates
    o.stroke_end()
    na = o.node_arrays()
    assert np.array_equal(na["orig_vb"], na["vb"])


@pytest.mark.parametrize("tool", [capi.TOOL_DRAW, capi.TOOL_INFLATE, capi.TOOL_GRAB, capi.TOOL_CLAY_STRIPS, capi.TOOL_SMOOTH])
def test_single_thread_vs_openmp(tool):
    """the timed OpenMP mode may differ from the deterministic mode only through the order of the
    float atomics of the normal accumulation (pbvh.c:2970-2976)"""
    m = meshgen.grid(129)
    dabs = stroke.c4_tool_stroke(tool, m.bbox_diag(), dabs=8) if tool != capi.TOOL_SMOOTH else \
        [capi.make_dab(capi.TOOL_SMOOTH, (0.1 * i - 0.3, 0.05 * i, 0.0), 0.4, bstrength=0.6) for i in range(6)]
    res = []
    for threads in (1, 4):
        o = Oracle(m, leaf_limit=300, threads=threads)
        o.stroke_begin()
        for d in dabs:
            o.dab(d)
        o.stroke_end()
        res.append((o.co(), o.no(), o.node_arrays()["vb"], o.touched()))
        o.set_threads(1)
    tol = 1e-5 * m.bbox_diag()
    assert np.abs(res[0][0] - res[1][0]).max() <= tol
    assert np.abs(res[0][1] - res[1][1]).max() <= 1e-5
    assert np.abs(res[0][2] - res[1][2]).max() <= tol
    assert np.array_equal(res[0][3], res[1][3])


@pytest.mark.parametrize("mk", [lambda: meshgen.grid(129), lambda: meshgen.mixed_grid(60), lambda: meshgen.icosphere(20, noise=0.003)])
def test_threaded_oracle_with_ordered_normals_is_bit_identical_to_the_serial_one(mk):
    """or_set_ordered_normals: what bench.py's full-size parity legs run -- threads for speed, per-vertex sums in the serial
    loop's order (ascending looptri position) for bits"""
    m = mk()
    dabs = stroke.c3_radius_sweep(m.bbox_diag(), dabs_per_radius=2)
    res = []
    for threads, ordered in ((1, 0), (4, 1)):
        oracle_lib().or_set_ordered_normals(ordered)  # before the build: the initial normals are computed with it
        o = Oracle(m, leaf_limit=200, threads=threads)
        o.stroke_begin()
        for d in dabs:
            o.dab(d)
        o.stroke_end()
        na = o.node_arrays()
        res.append((o.co(), o.no(), na["vb"], na["flag"], o.touched()))
        o.L.or_set_ordered_normals(0)
        o.set_threads(1)
        o.close()
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)


def test_golden_fixtures():
    """regression pins generated by tests/golden/make_golden.py FROM THE ORACLE (the reference ships
    no vectors for this path: parity unpinned)"""
    import json
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden
    want = json.load(open(os.path.join(here, "golden", "oracle_strokes.json")))
    got = make_golden.compute()
    assert got == want


def test_raycast_closed_forms():
    """BKE_pbvh_raycast + pbvh_faces_node_raycast (pbvh.c:3896-3928, 4041-4100) on shapes with known answers: a unit
    sphere is hit at |start| - 1 from outside along a radius; the hit vertex is a corner of the hit polygon and the
    nearest of them; a ray pointing away misses; max_depth cuts the search; brute force over all triangles agrees"""
    from oracle_py import Oracle
    m = meshgen.icosphere(16)
    orc = Oracle(m, leaf_limit=200)
    try:
        rng = np.random.default_rng(5)
        co = m.co.astype(np.float64)
        tri_v = orc.tri_verts()
        for i in range(40):
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            start = (-3.0 * d).astype(np.float32)
            nrm = d.astype(np.float32)
            hit = orc.raycast(start, nrm)
            assert hit is not None
            # facets lie just inside the unit sphere
            assert 2.0 - 1e-5 <= hit["depth"] < 2.02
            # brute force (Moeller-Trumbore in f64) over all looptris
            a, b, c = co[tri_v[:, 0]], co[tri_v[:, 1]], co[tri_v[:, 2]]
            e1, e2 = b - a, c - a
            pv = np.cross(d, e2)
            det = np.einsum("ij,ij->i", e1, pv)
            ok = np.abs(det) > 1e-14
            inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
            tv = start.astype(np.float64) - a
            u = np.einsum("ij,ij->i", tv, pv) * inv
            qv = np.cross(tv, e1)
            v = (qv @ d) * inv
            t = np.einsum("ij,ij->i", e2, qv) * inv
            inside = ok & (u >= -1e-9) & (v >= -1e-9) & (u + v <= 1 + 1e-9) & (t > 0)
            assert abs(t[inside].min() - hit["depth"]) < 1e-5
            f = hit["face"]
            poly = [int(x) for x in m.loop_v[m.poly_start[f]:m.poly_start[f] + m.poly_len[f]]]
            assert hit["vertex"] in poly
            loc = start.astype(np.float64) + d * float(hit["depth"])
            near = min(poly, key=lambda vv: np.sum((co[vv] - loc) ** 2))
            assert np.sum((co[hit["vertex"]] - loc) ** 2) <= np.sum((co[near] - loc) ** 2) + 1e-9
            assert np.dot(hit["normal"], d) < 0  # the outward facet faces the ray
            assert orc.raycast(start, (-d).astype(np.float32)) is None
            assert orc.raycast(start, nrm, max_depth=1.5) is None
    finally:
        orc.close()


# ---- rows a10 / a11 / a19 (dagger rows): closed forms of the hidden / tube / clip / normal-weight rules ----------

def test_hidden_vertices_stay_and_do_not_count_in_the_area_normal():
    m = meshgen.grid(65, height=0.0)
    vf = np.zeros(m.totvert, np.uint8)
    vf[::3] = capi.ME_HIDE
    o = Oracle(m, leaf_limit=100, vert_flag=vf)
    o.stroke_begin()
    o.dab(capi.make_dab(capi.TOOL_DRAW, (0, 0, 0), 0.4, curve_preset=capi.CURVE_CONSTANT, sculpt_plane=capi.DIR_Z, bstrength=0.25))
    inside = (m.co[:, 0] ** 2 + m.co[:, 1] ** 2) <= np.float32(0.4) ** 2
    co = o.co()
    vis = vf == 0
    assert np.array_equal(co[~vis], m.co[~vis])
    assert np.array_equal(co[inside & vis, 2], np.full((inside & vis).sum(), np.float32(0.4) * np.float32(0.25), np.float32))
    assert np.array_equal(np.sort(o.moved()), np.nonzero(inside & vis)[0])
    o.close()


def test_tube_falloff_is_the_distance_to_the_view_line():
    # two parallel sheets: the sphere brush moves the near one only, the tube both, by the same falloff of the XY distance
    a = meshgen.grid(33, height=0.0)
    co = np.concatenate([a.co, a.co + np.array([0, 0, -0.9], np.float32)])
    m = meshgen.Mesh(np.ascontiguousarray(co, np.float32), np.concatenate([a.poly_start, a.poly_start + a.loop_v.size]).astype(np.int32),
                     np.concatenate([a.poly_len, a.poly_len]).astype(np.int32), np.concatenate([a.loop_v, a.loop_v + a.totvert]).astype(np.int32))
    kw = dict(curve_preset=capi.CURVE_LIN, sculpt_plane=capi.DIR_Z, bstrength=0.5, view_normal=(0, 0, 1))
    o = Oracle(m, leaf_limit=100)
    o.stroke_begin()
    o.dab(capi.make_dab(capi.TOOL_DRAW, (0, 0, 0), 0.5, **kw))
    dz = o.co()[:, 2] - m.co[:, 2]
    assert dz[: a.totvert].max() > 0 and np.all(dz[a.totvert:] == 0)
    o.close()
    o = Oracle(m, leaf_limit=100)
    o.stroke_begin()
    o.dab(capi.make_dab(capi.TOOL_DRAW, (0, 0, 0), 0.5, falloff_shape=capi.FALLOFF_TUBE, **kw))
    dz = o.co()[:, 2] - m.co[:, 2]
    assert dz[: a.totvert].max() > 0
    # within a few ulps: z of the far sheet is -0.9 + dz
    assert np.allclose(dz[: a.totvert], dz[a.totvert:], atol=2e-7)
    r = np.sqrt(m.co[: a.totvert, 0].astype(np.float64) ** 2 + m.co[: a.totvert, 1].astype(np.float64) ** 2)
    want = np.where(r <= 0.5, 0.5 * 0.5 * (1.0 - r / 0.5), 0.0)
    assert np.allclose(dz[: a.totvert], want, atol=1e-6)
    o.close()


def test_clip_holds_the_mirror_plane_and_lock_keeps_the_axis():
    m = meshgen.grid(65, height=0.0)
    o = Oracle(m, leaf_limit=100)
    o.stroke_begin()
    o.dab(capi.make_dab(capi.TOOL_DRAW, (0, 0, 0), 0.5, curve_preset=capi.CURVE_CONSTANT, sculpt_plane=capi.DIR_X, bstrength=0.25,
                        clip_flags=capi.CLIP_X | capi.LOCK_Z, clip_tolerance=(0.02, 0, 0)))
    co = o.co()
    inside = (m.co[:, 0] ** 2 + m.co[:, 1] ** 2) < np.float32(0.5) ** 2  # on the rim the falloff is 0
    near = np.abs(m.co[:, 0]) <= np.float32(0.02)
    assert np.all(co[inside & near, 0] == 0.0)
    assert np.array_equal(co[inside & ~near, 0], m.co[inside & ~near, 0] + np.float32(0.5) * np.float32(0.25))
    assert np.array_equal(co[:, 2], m.co[:, 2]) and np.array_equal(co[:, 1], m.co[:, 1])
    o.close()


def test_grab_normal_weight_one_moves_along_the_normal_only():
    # flat sheet, drag (dx, 0, dz), weight 1: the drag becomes n * dot(n, drag) * 1/|dot(n - view (n.view), n)| -- with the view
    # along the normal the in-plane part of n vanishes and the scale falls back to 1: pure +Z motion of dz
    m = meshgen.grid(65, height=0.0)
    o = Oracle(m, leaf_limit=100)
    o.stroke_begin()
    o.dab(capi.make_dab(capi.TOOL_GRAB, (0, 0, 0), 0.4, curve_preset=capi.CURVE_CONSTANT, bstrength=1.0, grab_delta=(0.3, 0.0, 0.2),
                        normal_weight=1.0, flags=capi.DAB_FIRST_STEP))
    co = o.co()
    inside = (m.co[:, 0] ** 2 + m.co[:, 1] ** 2) <= np.float32(0.4) ** 2
    assert np.array_equal(co[inside, 0], m.co[inside, 0]) and np.array_equal(co[inside, 1], m.co[inside, 1])
    assert np.allclose(co[inside, 2], 0.2, atol=1e-6)
    assert np.array_equal(co[~inside], m.co[~inside])
    o.close()


def test_symmetry_helper_lists_the_valid_mirror_passes():
    d = capi.make_dab(capi.TOOL_GRAB, (0.5, 0.25, -0.125), 0.3, view_normal=(0.6, 0.0, 0.8), grab_delta=(0.1, 0.2, 0.3))
    for symm, want in [(0, [0]), (1, [0, 1]), (2, [0, 2]), (4, [0, 4]), (3, [0, 1, 2, 3]), (5, [0, 1, 4, 5]), (6, [0, 2, 4, 6]),
                       (7, [0, 1, 2, 3, 4, 5, 6, 7])]:
        out = capi.dab_symmetry(d, symm)
        assert len(out) == len(want)
        for o, i in zip(out, want):
            sgn = np.array([-1.0 if i & (1 << k) else 1.0 for k in range(3)])
            assert np.array_equal(np.array(o.location[:]), np.array(d.location[:]) * sgn)
            assert np.array_equal(np.array(o.view_normal[:]), np.array(d.view_normal[:]) * sgn)
            assert np.array_equal(np.array(o.grab_delta[:]), np.array(d.grab_delta[:]) * sgn)
            assert o.radius == d.radius and o.tool == d.tool
