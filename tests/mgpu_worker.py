"""One rank of the multi-GPU parity check (launched by tests/test_gpu_multi.py, one process per GPU).
Every rank holds the whole mesh, computes the leaves it owns, and after stroke end must hold exactly
the state the single-process CPU oracle produces."""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from dune_sculpt_b200 import capi, meshgen, stroke  # noqa: E402
from oracle_py import GridOracle, Oracle  # noqa: E402


def skipped(ses):
    import ctypes
    n = ctypes.c_int(0)
    ses.D.dsc_dist_exchanges_skipped(ses.ctx, ctypes.byref(n))
    return n.value


def own_part_on_host(orc, ses, rank, what):
    """after stroke end the HOST arrays of a rank hold the oracle's values for the vertices / elements it owns"""
    rng, owner = ses.partition(ses.dist[0])
    na = ses.node_arrays()
    co_o, no_o = orc.co(), orc.no()
    if hasattr(ses, "host_elements"):
        co_h, no_h, _ = ses.host_elements()
        gs2 = ses.mesh.grid_size ** 2
        prim = ses.prim_indices()
        mine = np.concatenate([prim[na["prim_offset"][n]:na["prim_offset"][n] + na["totprim"][n]] for n in np.nonzero(owner == rank)[0]] or
                              [np.zeros(0, np.int32)])
        idx = (mine[:, None].astype(np.int64) * gs2 + np.arange(gs2)[None, :]).reshape(-1)
    else:
        co_h = ses.mvert["co"] if not ses.pbvh.contents.deformed else np.ctypeslib.as_array(
            C.cast(ses.pbvh.contents.verts, C.POINTER(C.c_float)), shape=(ses.mesh.totvert, 4))[:, :3]
        no_h = np.ctypeslib.as_array(ses.pbvh.contents.vert_normals, shape=(ses.mesh.totvert * 3,)).reshape(-1, 3)
        idx = np.concatenate([ses.node_vert_indices(int(n))[:na["uniq_verts"][n]] for n in np.nonzero(owner == rank)[0]] or
                             [np.zeros(0, np.int32)])
    bad = idx[(co_o[idx] != co_h[idx]).any(axis=1)]
    assert bad.size == 0, "rank %d %s: host positions of the owned part differ at %d of %d (first %s, max %g)" % (
        rank, what, bad.size, idx.size, bad[:6], np.abs(co_o[idx] - co_h[idx]).max())
    assert np.array_equal(no_o[idx], no_h[idx]), "rank %d %s: host normals of the owned part differ" % (rank, what)


def batched_stroke(orc, ses, dabs, rank, what):
    """a second stroke submitted as ONE dsc_dabs call: runs of one launch sequence replay as CUDA graphs, the
    peer-memory exchanges inside them (their round numbers live on the device)"""
    dabs = sorted(dabs, key=lambda d: d.tool)       # runs of equal signature
    arr = (capi.DscDab * len(dabs))(*dabs)
    orc.stroke_begin(None)
    for d in dabs:
        orc.dab(d)
    orc.stroke_end()
    ses.stroke_begin(None)
    ses.dabs(arr, len(dabs))
    ses.stroke_end()
    own_part_on_host(orc, ses, rank, what)
    ses.gather()
    assert np.array_equal(orc.co(), ses.co()), "rank %d: batched %s stroke: positions differ" % (rank, what)
    assert np.array_equal(orc.no(), ses.no()), "rank %d: batched %s stroke: normals differ" % (rank, what)
    na = orc.node_arrays()
    bb, obb = ses.node_bb()
    assert np.array_equal(na["vb"], bb) and np.array_equal(na["orig_vb"], obb), "rank %d: batched stroke: boxes differ" % rank


def main():
    world, rank, idfile, scenario = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    if rank == 0:
        nid = capi.nccl_unique_id()
        with open(idfile + ".tmp", "wb") as f:
            f.write(nid)
        os.replace(idfile + ".tmp", idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            time.sleep(0.05)
            assert time.time() - t0 < 120, "no NCCL id from rank 0"
        nid = open(idfile, "rb").read()
    if scenario.startswith("multires"):
        return multires(world, rank, nid, scenario)
    if scenario == "grid":
        mesh, ll = meshgen.grid(257), 900
        diag = mesh.bbox_diag()
        dabs = []
        for tool in (capi.TOOL_DRAW, capi.TOOL_INFLATE, capi.TOOL_CLAY_STRIPS, capi.TOOL_GRAB):
            dabs += stroke.c4_tool_stroke(tool, diag, dabs=6, radius_pct=14.0)
        dabs += [capi.make_dab(capi.TOOL_SMOOTH, (0.12 * i - 0.5, 0.05 * i - 0.1, 0.0), 0.45, bstrength=0.8) for i in range(6)]
        # small dabs well inside one rank's region: their halo exchanges are skipped on every rank
        for c in ((-0.8, -0.8, 0.0), (0.8, 0.8, 0.0), (-0.8, 0.8, 0.0), (0.8, -0.8, 0.0)):
            dabs.append(capi.make_dab(capi.TOOL_DRAW, c, 0.08, bstrength=0.2, view_normal=(0, 0, 1)))
            dabs.append(capi.make_dab(capi.TOOL_SMOOTH, c, 0.1, bstrength=0.6))
        mask = meshgen.low_freq_mask(mesh)
    else:
        mesh, ll = meshgen.icosphere(48, noise=0.003), 700
        dabs = stroke.c2_smooth_stroke(dabs=8, radius=0.6) + \
            [capi.make_dab(capi.TOOL_DRAW, p, 0.5, bstrength=0.2, view_normal=p) for p in ((0, 0, 1), (0.6, 0, 0.8), (1, 0, 0))]
        mask = None
    orc = Oracle(mesh, mask=mask, leaf_limit=ll)
    ses = capi.SculptSession(mesh, mask=mask, leaf_limit=ll, device=rank, dist=(world, rank, nid))
    rng, owner = ses.partition(world)
    orc.stroke_begin()
    ses.stroke_begin()
    vd = 0
    for i, d in enumerate(dabs):
        orc.dab(d)
        ses.dab(d)
        ho = orc.hits()
        mine = ho[owner[ho] == rank]
        hg = ses.hits()
        assert np.array_equal(mine, hg), "rank %d dab %d: own hit list differs (%d vs %d)" % (rank, i, mine.size, hg.size)
    orc.stroke_end()
    ses.stroke_end()  # leaf boxes / flags / stroke state travel; the vertex data stays with its owner
    own_part_on_host(orc, ses, rank, scenario)
    ses.gather()      # every replica whole again
    vd = ses.stats()["vertex_dabs"]
    co_o, co_g = orc.co(), ses.co()
    assert np.array_equal(co_o, co_g), "rank %d: positions differ (max %g)" % (rank, np.abs(co_o - co_g).max())
    assert np.array_equal(orc.no(), ses.no()), "rank %d: normals differ" % rank
    na = orc.node_arrays()
    bb, obb = ses.node_bb()
    assert np.array_equal(na["vb"], bb) and np.array_equal(na["orig_vb"], obb), "rank %d: boxes differ" % rank
    assert np.array_equal(orc.orig_co(), ses.orig_co()), "rank %d: undo snapshot differs" % rank
    assert np.array_equal(orc.touched(), ses.touched()), "rank %d: undo membership differs" % rank
    batched_stroke(orc, ses, dabs, rank, scenario)
    print("MGPU_OK rank %d/%d scenario %s own leaves [%d,%d) vertex_dabs %d of %d peer_memory %d skipped_exchanges %d" %
          (rank, world, scenario, rng[rank], rng[rank + 1], vd, orc.vertex_dabs(), ses.D.dsc_dist_uses_peer_memory(ses.ctx),
           skipped(ses)), ses.dist_dab_counts(), flush=True)
    ses.close()


def multires(world, rank, nid, scenario):
    """partitioned multires grids: smooth, draw, inflate, grab and clay strips dabs that straddle the partition cuts;
    after stroke end every replica must hold the oracle's elements, normals, mask layer and boxes, bit for bit"""
    small_r = 0.04
    if scenario == "multires_open":
        mr, ll = meshgen.multires_plane(6, 4, noise=0.04, freq=9.0, with_mask=True), 3
        centres = [(-0.5 + 0.25 * i, 0.1 * i - 0.2, 0.0) for i in range(5)]
    elif scenario == "multires_wide":
        # many coarse faces, small dabs: most dabs reach one or two ranks only -- the others skip them and run ahead
        mr, ll = meshgen.multires_plane(16, 3, noise=0.04, freq=9.0, with_mask=True), 4
        centres = [(-0.7 + 0.35 * i, 0.3 * i - 0.6, 0.0) for i in range(5)]
        small_r = 0.02
    else:
        mr, ll = meshgen.multires_cube(2, 4, noise=0.03, freq=13.0, with_mask=True), 5
        rng_ = np.random.default_rng(11)
        centres = [p / np.linalg.norm(p) for p in rng_.normal(size=(6, 3))]
    diag = mr.bbox_diag()
    dabs = []
    for i, c in enumerate(centres):
        c = np.asarray(c, dtype=np.float32)
        n = tuple(c / max(np.linalg.norm(c), 1e-9)) if scenario == "multires" else (0.0, 0.0, 1.0)
        dabs.append(capi.make_dab(capi.TOOL_SMOOTH, c, diag * 0.22, bstrength=stroke._strength(capi.TOOL_SMOOTH, 0.6)))
        dabs.append(capi.make_dab(capi.TOOL_DRAW, c, diag * (0.12 + 0.05 * i), bstrength=stroke._strength(capi.TOOL_DRAW, 0.5),
                                  view_normal=n, flags=capi.DAB_FIRST_STEP if i == 0 else 0))
        dabs.append(capi.make_dab(capi.TOOL_INFLATE, c, diag * 0.3, bstrength=stroke._strength(capi.TOOL_INFLATE, 0.5), view_normal=n))
        dabs.append(capi.make_dab(capi.TOOL_CLAY_STRIPS, c, diag * 0.25, bstrength=stroke._strength(capi.TOOL_CLAY_STRIPS, 0.5),
                                  view_normal=n, grab_delta=(0.03, 0.01, 0.02)))
    big = 0.3 if scenario != "multires_wide" else 0.08
    if scenario == "multires_wide":
        for d in dabs:
            d.radius = d.radius * 0.3
    dabs.append(capi.make_dab(capi.TOOL_GRAB, np.asarray(centres[0], dtype=np.float32), diag * big,
                              bstrength=stroke._strength(capi.TOOL_GRAB, 0.5), grab_delta=(0.02, -0.03, 0.04)))
    # small dabs: some of them gather no leaf near a partition cut, and their halo exchanges are skipped on every rank
    rs = np.random.default_rng(23)
    for k in range(48 if scenario == "multires_wide" else 16):
        if scenario != "multires":
            c = np.array([rs.uniform(-0.9, 0.9), rs.uniform(-0.9, 0.9), 0.0], dtype=np.float32)
            n = (0.0, 0.0, 1.0)
        else:
            c = rs.normal(size=3)
            c = (c / np.linalg.norm(c)).astype(np.float32)
            n = tuple(c)
        if k % 2:
            dabs.append(capi.make_dab(capi.TOOL_SMOOTH, c, diag * small_r, bstrength=stroke._strength(capi.TOOL_SMOOTH, 0.6)))
        else:
            dabs.append(capi.make_dab(capi.TOOL_DRAW, c, diag * small_r, bstrength=stroke._strength(capi.TOOL_DRAW, 0.5), view_normal=n))
    orc = GridOracle(mr, leaf_limit=ll)
    ses = capi.GridSession(mr, leaf_limit=ll, device=rank, dist=(world, rank, nid))
    rng, owner = ses.partition(world)
    orc.stroke_begin(None)
    ses.stroke_begin(None)
    for i, d in enumerate(dabs):
        orc.dab(d)
        ses.dab(d)
        ho = orc.hits()
        mine = ho[owner[ho] == rank]
        hg = ses.hits()
        assert np.array_equal(mine, hg), "rank %d dab %d: own hit list differs (%d vs %d)" % (rank, i, mine.size, hg.size)
    orc.stroke_end()
    ses.stroke_end()
    own_part_on_host(orc, ses, rank, scenario)
    ses.gather()
    co_o, co_g = orc.co(), ses.co()
    bad = np.nonzero((co_o != co_g).any(axis=1))[0]
    gs2 = mr.grid_size ** 2
    assert bad.size == 0, "rank %d: %d positions differ (max %g), first elements %s (grid, y, x) %s" % (
        rank, bad.size, np.abs(co_o - co_g).max(), bad[:8], [(int(b) // gs2, (int(b) % gs2) // mr.grid_size, int(b) % mr.grid_size) for b in bad[:8]])
    no_o, no_g = orc.no(), ses.no()
    badn = np.nonzero((no_o != no_g).any(axis=1))[0]
    assert badn.size == 0, "rank %d: %d normals differ, first (grid, y, x) %s" % (
        rank, badn.size, [(int(b) // gs2, (int(b) % gs2) // mr.grid_size, int(b) % mr.grid_size) for b in badn[:8]])
    assert np.array_equal(orc.mask(), ses.mask()), "rank %d: mask layer differs" % rank
    na = orc.node_arrays()
    bb, obb = ses.node_bb()
    assert np.array_equal(na["vb"], bb) and np.array_equal(na["orig_vb"], obb), "rank %d: boxes differ" % rank
    assert np.array_equal(orc.orig_co(), ses.orig_co()), "rank %d: undo snapshot differs" % rank
    assert np.array_equal(orc.touched(), ses.touched()), "rank %d: undo membership differs" % rank
    batched_stroke(orc, ses, dabs, rank, scenario)
    assert np.array_equal(orc.mask(), ses.mask()), "rank %d: mask layer differs after the batched stroke" % rank
    print("MGPU_OK rank %d/%d scenario %s own leaves [%d,%d) vertex_dabs %d of %d peer_memory %d skipped_exchanges %d" %
          (rank, world, scenario, rng[rank], rng[rank + 1], ses.stats()["vertex_dabs"], orc.vertex_dabs(),
           ses.D.dsc_dist_uses_peer_memory(ses.ctx), skipped(ses)), ses.dist_dab_counts(), flush=True)
    ses.close()


if __name__ == "__main__":
    main()
