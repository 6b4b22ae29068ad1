"""Host logic of the multi-GPU path on the CPU, world_size 2 over gloo: the spatial partition, the
symmetry of the halo plan, and a simulated halo exchange (what NCCL send/recv carries on the box)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dune_sculpt_b200 import capi, meshgen
        mesh = meshgen.icosphere(24, noise=0.002)
        ses = capi.SculptSession(mesh, leaf_limit=300)
        rng, owner = ses.partition(world)
        na = ses.node_arrays()
        leaves = np.nonzero(na["flag"] & 1)[0]
        leaves = leaves[np.argsort(na["prim_offset"][leaves])]
        # contiguous runs in traversal order, every leaf owned exactly once
        assert rng[0] == 0 and rng[-1] == leaves.size and np.all(np.diff(rng) > 0)
        assert np.array_equal(owner[leaves], np.repeat(np.arange(world), np.diff(rng)))
        soff, sv, roff, rv = ses.halo_plan(world, rank)
        # vertex ownership: unique verts of owned leaves
        vowner = np.full(mesh.totvert, -1)
        for n in leaves:
            vi = ses.node_vert_indices(int(n))
            vowner[vi[:na["uniq_verts"][n]]] = owner[n]
        assert (vowner >= 0).all()
        assert np.all(vowner[sv] == rank) and np.all(vowner[rv] != rank)
        # symmetry: what I send to p is what p expects from me (sizes first, then the ids)
        peer = 1 - rank
        mine = torch.tensor([int(soff[peer + 1] - soff[peer]), int(roff[peer + 1] - roff[peer])], dtype=torch.long)
        theirs = [torch.zeros(2, dtype=torch.long) for _ in range(world)]
        dist.all_gather(theirs, mine)
        assert theirs[peer][1].item() == mine[0].item() and theirs[peer][0].item() == mine[1].item()
        send_ids = torch.from_numpy(sv[soff[peer]:soff[peer + 1]].astype(np.int64))
        recv_ids = torch.zeros(int(mine[1]), dtype=torch.long)
        if rank == 0:
            dist.send(send_ids, peer); dist.recv(recv_ids, peer)
        else:
            dist.recv(recv_ids, peer); dist.send(send_ids, peer)
        assert np.array_equal(recv_ids.numpy(), rv[roff[peer]:roff[peer + 1]])
        # simulated dab: each rank moves the verts it owns, then the halo exchange must make every
        # vertex its leaves read (shared verts and smooth neighbours) current
        truth = mesh.co.astype(np.float64) * (1.0 + 0.01 * (vowner[:, None] + 1))
        local = mesh.co.astype(np.float64).copy()
        local[vowner == rank] = truth[vowner == rank]
        payload = torch.from_numpy(local[sv[soff[peer]:soff[peer + 1]]])
        got = torch.zeros((int(mine[1]), 3), dtype=torch.float64)
        if rank == 0:
            dist.send(payload, peer); dist.recv(got, peer)
        else:
            dist.recv(got, peer); dist.send(payload, peer)
        local[rv[roff[peer]:roff[peer + 1]]] = got.numpy()
        off, idx, _ = ses.neighbor_tables()
        for n in leaves[owner[leaves] == rank]:
            vi = ses.node_vert_indices(int(n))
            assert np.array_equal(local[vi], truth[vi])           # unique + shared verts of my leaves
            for v in vi[:na["uniq_verts"][n]]:
                nb = idx[off[v]:off[v + 1]]
                assert np.array_equal(local[nb], truth[nb])       # one-ring of my unique verts
        q.put((rank, "ok", int(sv.size), int(rv.size)))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "fail: %s\n%s" % (e, traceback.format_exc()), 0, 0))
    finally:
        dist.destroy_process_group()


def test_partition_and_halo_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status, ns, nr in sorted(res):
        assert status == "ok", "rank %d: %s" % (rank, status)
        assert ns > 0 and nr > 0


def test_partition_shapes():
    sys.path.insert(0, ROOT)
    from dune_sculpt_b200 import capi, meshgen
    ses = capi.SculptSession(meshgen.grid(129), leaf_limit=300)
    for world in (1, 2, 3, 4, 8):
        rng, owner = ses.partition(world)
        assert rng[0] == 0 and rng[-1] == int((ses.node_arrays()["flag"] & 1).sum())
        assert sorted(set(owner[owner >= 0].tolist())) == list(range(world))
    # power-of-two worlds cut the tree at depth log2(world): every rank owns whole subtrees
    rng, owner = ses.partition(4)
    na = ses.node_arrays()
    c = na["children_offset"]
    depth2 = [c[c[0]], c[c[0]] + 1, c[c[0] + 1], c[c[0] + 1] + 1]

    def leaves_under(n):
        if na["flag"][n] & 1:
            return [n]
        return leaves_under(c[n]) + leaves_under(c[n] + 1)
    for n in depth2:
        assert len(set(owner[leaves_under(int(n))].tolist())) == 1
    ses.close()


def _grids_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dune_sculpt_b200 import capi, meshgen
        mr = meshgen.multires_cube(1, 3)      # 24 faces, 96 grids of 5 x 5
        gs, gs2 = mr.grid_size, mr.grid_size ** 2
        ses = capi.GridSession(mr, leaf_limit=3, device=None)
        plan = ses.grids_plan(world, rank)
        rng, node_owner = ses.partition(world)
        na = ses.node_arrays()
        prims = ses.prim_indices()
        # a grid belongs to the rank of its leaf
        gowner = np.full(mr.totgrid, -1)
        for n in np.nonzero(na["flag"] & 1)[0]:
            gowner[prims[na["prim_offset"][n]:na["prim_offset"][n] + na["totprim"][n]]] = node_owner[n]
        assert np.array_equal(gowner, plan["grid_owner"]) and set(gowner.tolist()) == set(range(world))
        eowner = np.repeat(gowner, gs2)
        mine = eowner == rank
        have = mine.copy()
        have[plan["recv_elem"]] = True
        assert not mine[plan["recv_elem"]].any() and mine[plan["send_elem"]].all()
        # every group of duplicated elements with an owned member is computed here, on current inputs
        F = mr.face_start.shape[0]
        fdom = np.array([(gowner[mr.face_start[f]:mr.face_start[f] + mr.face_num[f]] == rank).any() for f in range(F)])
        assert np.array_equal((plan["face_dom"] & 1).astype(bool), fdom)
        for f in np.nonzero(fdom)[0]:
            for c in range(mr.face_num[f]):
                g = mr.face_start[f] + c
                i = np.arange(gs)
                assert have[g * gs2 + i].all() and have[g * gs2 + i * gs].all()
        rows = mr.edge_elems.reshape(-1, 2 * gs)
        for e in range(mr.edge_off.shape[0] - 1):
            r_ = rows[mr.edge_off[e]:mr.edge_off[e + 1]]
            for h in range(2):
                owned_half = mine[r_[:, h * gs:(h + 1) * gs]].any()
                assert bool((plan["edge_mine"][e] >> h) & 1) == bool(owned_half), (e, h)
                if owned_half:
                    assert have[r_[:, h * gs:(h + 1) * gs]].all()
                    assert have[r_[:, gs - 1:gs + 1]].all()           # the middle pair of every face on the edge
                    for k in range(r_.shape[0]):                       # ... which that face's inner pass runs here
                        f = int(np.searchsorted(mr.face_start, r_[k, 0] // gs2, side="right") - 1)
                        assert plan["face_dom"][f] & 3
        for v in range(mr.cvert_off.shape[0] - 1):
            el = mr.cvert_elems[mr.cvert_off[v]:mr.cvert_off[v + 1]]
            assert bool(plan["cvert_mine"][v]) == bool(mine[el].any())
            if mine[el].any():
                assert have[el].all()
                for x in el:                                            # every face around it lists it
                    f = int(np.searchsorted(mr.face_start, x // gs2, side="right") - 1)
                    assert plan["face_dom"][f] != 0
        # smooth brush: the neighbours of every owned element
        for e in np.nonzero(mine)[0]:
            assert have[ses.neighbors(int(e))[0]].all(), "element %d reads a stale neighbour" % e
        # symmetry with the peer over the wire, then a simulated exchange
        peer = 1 - rank
        so, ro = plan["send_off"], plan["recv_off"]
        send_ids = torch.from_numpy(plan["send_elem"][so[peer]:so[peer + 1]].astype(np.int64))
        n_recv = torch.zeros(1, dtype=torch.long)
        n_send = torch.tensor([send_ids.numel()], dtype=torch.long)
        if rank == 0:
            dist.send(n_send, peer); dist.recv(n_recv, peer)
        else:
            dist.recv(n_recv, peer); dist.send(n_send, peer)
        assert int(n_recv) == ro[peer + 1] - ro[peer]
        recv_ids = torch.zeros(int(n_recv), dtype=torch.long)
        truth = mr.co.astype(np.float64) * (1.0 + 0.01 * (eowner[:, None] + 1))
        local = mr.co.astype(np.float64).copy()
        local[mine] = truth[mine]
        payload = torch.from_numpy(local[send_ids.numpy()])
        got = torch.zeros((int(n_recv), 3), dtype=torch.float64)
        if rank == 0:
            dist.send(send_ids, peer); dist.recv(recv_ids, peer); dist.send(payload, peer); dist.recv(got, peer)
        else:
            dist.recv(recv_ids, peer); dist.send(send_ids, peer); dist.recv(got, peer); dist.send(payload, peer)
        assert np.array_equal(recv_ids.numpy(), plan["recv_elem"][ro[peer]:ro[peer + 1]])
        local[recv_ids.numpy()] = got.numpy()
        assert np.array_equal(local[have], truth[have])
        ses.close()
        q.put((rank, "ok", int(plan["send_elem"].size), int(plan["recv_elem"].size)))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "fail: %s\n%s" % (e, traceback.format_exc()), 0, 0))
    finally:
        dist.destroy_process_group()


def test_grids_partition_plan_world2_gloo():
    """partitioned multires grids: ownership, the groups a rank averages, the halo closure (incl. the middle pairs
    of shared edges and the smooth brush's neighbours) and the symmetry of the send / receive lists"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grids_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status, ns, nr in sorted(res):
        assert status == "ok", "rank %d: %s" % (rank, status)
        assert ns > 0 and nr > 0


def test_grids_partition_plans_are_symmetric_and_closed_for_every_world_size():
    """the same plan for 3, 4 and 8 ranks, all ranks evaluated in one process: what rank a sends to rank b is what b
    expects from a (same elements, same order), nobody receives an element it owns, and every averaging group with an
    owned member reads only owned or received elements"""
    sys.path.insert(0, ROOT)
    from dune_sculpt_b200 import capi, meshgen
    mr = meshgen.multires_cube(2, 3)     # 96 faces, 384 grids of 5 x 5
    gs, gs2 = mr.grid_size, mr.grid_size ** 2
    ses = capi.GridSession(mr, leaf_limit=4, device=None)
    rows = mr.edge_elems.reshape(-1, 2 * gs)
    for world in (3, 4, 8):
        plans = [ses.grids_plan(world, r, with_neighbors=(world == 4)) for r in range(world)]
        owner = plans[0]["grid_owner"]
        assert set(owner.tolist()) == set(range(world))
        eowner = np.repeat(owner, gs2)
        for a in range(world):
            pa = plans[a]
            assert np.array_equal(pa["grid_owner"], owner)
            assert (eowner[pa["send_elem"]] == a).all() and (eowner[pa["recv_elem"]] != a).all()
            for b in range(world):
                if a == b:
                    continue
                sent = pa["send_elem"][pa["send_off"][b]:pa["send_off"][b + 1]]
                want = plans[b]["recv_elem"][plans[b]["recv_off"][a]:plans[b]["recv_off"][a + 1]]
                assert np.array_equal(sent, want), (world, a, b)
            have = eowner == a
            have[pa["recv_elem"]] = True
            for e in range(mr.edge_off.shape[0] - 1):
                r_ = rows[mr.edge_off[e]:mr.edge_off[e + 1]]
                for h in range(2):
                    if (eowner[r_[:, h * gs:(h + 1) * gs]] == a).any():
                        assert (pa["edge_mine"][e] >> h) & 1
                        assert have[r_[:, h * gs:(h + 1) * gs]].all() and have[r_[:, gs - 1:gs + 1]].all()
            for v in range(mr.cvert_off.shape[0] - 1):
                el = mr.cvert_elems[mr.cvert_off[v]:mr.cvert_off[v + 1]]
                if (eowner[el] == a).any():
                    assert pa["cvert_mine"][v] and have[el].all()
            for f in range(mr.face_start.shape[0]):
                g0 = mr.face_start[f]
                if (owner[g0:g0 + mr.face_num[f]] == a).any():
                    assert pa["face_dom"][f] & 1
                    for c in range(mr.face_num[f]):
                        i = np.arange(gs)
                        assert have[(g0 + c) * gs2 + i].all() and have[(g0 + c) * gs2 + i * gs].all()
    ses.close()


def test_mesh_halo_plans_are_symmetric_for_every_world_size():
    sys.path.insert(0, ROOT)
    from dune_sculpt_b200 import capi, meshgen
    ses = capi.SculptSession(meshgen.icosphere(20, noise=0.002), leaf_limit=200)
    for world in (3, 4, 8):
        plans = [ses.halo_plan(world, r) for r in range(world)]
        for a in range(world):
            soff, sv, roff, rv = plans[a]
            for b in range(world):
                if a == b:
                    assert soff[b + 1] == soff[b] and roff[b + 1] == roff[b]
                    continue
                boff, bsv, broff, brv = plans[b]
                assert np.array_equal(sv[soff[b]:soff[b + 1]], brv[broff[a]:broff[a + 1]]), (world, a, b)
    ses.close()
