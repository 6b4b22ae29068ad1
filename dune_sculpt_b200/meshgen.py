"""Synthetic meshes of the benchmark configs, seeded and file-free.

A mesh is the reference's polygon soup: MVert.co, MPoly (loopstart, totloop) and MLoop.v
(types/types_meshdata.h:13-17, 50-57, 69-74).
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class Mesh:
    co: np.ndarray          # (V, 3) float32
    poly_start: np.ndarray  # (P,) int32   MPoly.loopstart
    poly_len: np.ndarray    # (P,) int32   MPoly.totloop
    loop_v: np.ndarray      # (L,) int32   MLoop.v

    @property
    def totvert(self):
        return int(self.co.shape[0])

    @property
    def totpoly(self):
        return int(self.poly_start.shape[0])

    @property
    def totloop(self):
        return int(self.loop_v.shape[0])

    def bbox_diag(self):
        return float(np.linalg.norm(self.co.max(axis=0).astype(np.float64) - self.co.min(axis=0).astype(np.float64)))


def _from_faces(co, faces):
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    p, k = faces.shape
    return Mesh(
        co=np.ascontiguousarray(co, dtype=np.float32),
        poly_start=(np.arange(p, dtype=np.int32) * k).astype(np.int32),
        poly_len=np.full(p, k, dtype=np.int32),
        loop_v=faces.reshape(-1).copy(),
    )


def grid(n, height=0.05, freq=8.0):
    """n x n vertices on [-1,1]^2, z = height*sin(freq*x)*cos(freq*y), (n-1)^2 quads facing +Z.
    C3: n=4096 (V=16,777,216); C4: n=2048."""
    t = np.linspace(-1.0, 1.0, n, dtype=np.float64)
    x, y = np.meshgrid(t, t, indexing="xy")  # x varies fastest
    z = height * np.sin(freq * x) * np.cos(freq * y)
    co = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(n - 1, dtype=np.int64), np.arange(n - 1, dtype=np.int64), indexing="xy")
    v00 = (j * n + i).reshape(-1)
    faces = np.stack([v00, v00 + 1, v00 + n + 1, v00 + n], axis=-1)
    return _from_faces(co, faces)


def cube(levels=8):
    """Cube [-1,1]^3 with every face split into 2^levels x 2^levels quads, shared edges welded.
    C1: levels=8 -> V = 6*256^2 + 2 = 393,218."""
    n = 1 << levels
    s = n + 1
    a, b = np.meshgrid(np.arange(s, dtype=np.int64), np.arange(s, dtype=np.int64), indexing="xy")
    a = a.reshape(-1)
    b = b.reshape(-1)
    zero = np.zeros_like(a)
    full = np.full_like(a, n)
    # (x, y, z) integer coordinates per face; (a, b) ordered so that a x b points outward
    face_xyz = [
        (full, a, b),   # +X
        (zero, b, a),   # -X
        (b, full, a),   # +Y
        (a, zero, b),   # -Y
        (a, b, full),   # +Z
        (b, a, zero),   # -Z
    ]
    keys = [x * s * s + y * s + z for (x, y, z) in face_xyz]
    allkeys = np.concatenate(keys)
    uniq, inv = np.unique(allkeys, return_inverse=True)
    ux = uniq // (s * s)
    uy = (uniq // s) % s
    uz = uniq % s
    co = (np.stack([ux, uy, uz], axis=-1).astype(np.float64) * (2.0 / n) - 1.0).astype(np.float32)
    ci, cj = np.meshgrid(np.arange(n, dtype=np.int64), np.arange(n, dtype=np.int64), indexing="xy")
    c00 = (cj * s + ci).reshape(-1)
    faces = []
    for f in range(6):
        g = inv[f * s * s:(f + 1) * s * s]
        faces.append(np.stack([g[c00], g[c00 + 1], g[c00 + s + 1], g[c00 + s]], axis=-1))
    return _from_faces(co, np.concatenate(faces))


_ICO_T = (1.0 + 5.0 ** 0.5) / 2.0
_ICO_V = np.array(
    [[-1, _ICO_T, 0], [1, _ICO_T, 0], [-1, -_ICO_T, 0], [1, -_ICO_T, 0], [0, -1, _ICO_T], [0, 1, _ICO_T],
     [0, -1, -_ICO_T], [0, 1, -_ICO_T], [_ICO_T, 0, -1], [_ICO_T, 0, 1], [-_ICO_T, 0, -1], [-_ICO_T, 0, 1]],
    dtype=np.float64)
_ICO_F = np.array(
    [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
     [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7],
     [9, 8, 1]], dtype=np.int64)


def icosphere(f, radius=1.0, noise=0.0, seed=1234):
    """Geodesic icosphere of frequency f: V = 10 f^2 + 2 verts, 20 f^2 triangles.
    C2: f=316 -> V = 998,562.  noise: per-vertex displacement U(-noise, noise) along the normal."""
    m = f + 1
    keys_all, pts_all, tris_all = [], [], []
    base = 0
    # barycentric lattice of one face
    ii, jj = np.meshgrid(np.arange(m, dtype=np.int64), np.arange(m, dtype=np.int64), indexing="ij")
    sel = (ii + jj) <= f
    wi = ii[sel]
    wj = jj[sel]
    wk = f - wi - wj
    lid = -np.ones((m, m), dtype=np.int64)
    lid[wi, wj] = np.arange(wi.shape[0])
    # triangles of the lattice: (i,j),(i+1,j),(i,j+1) and (i+1,j),(i+1,j+1),(i,j+1)
    up = (ii + jj) <= (f - 1)
    ui, uj = ii[up], jj[up]
    t_up = np.stack([lid[ui, uj], lid[ui + 1, uj], lid[ui, uj + 1]], axis=-1)
    dn = (ii + jj) <= (f - 2)
    di, dj = ii[dn], jj[dn]
    t_dn = np.stack([lid[di + 1, dj], lid[di + 1, dj + 1], lid[di, dj + 1]], axis=-1)
    lat_tris = np.concatenate([t_up, t_dn])
    big = np.int64(f + 1)
    for (a, b, c) in _ICO_F:
        # canonical key: the (vertex id, weight) pairs with non-zero weight, sorted by vertex id
        vid = np.stack([np.full_like(wi, a), np.full_like(wi, b), np.full_like(wi, c)], axis=-1)
        w = np.stack([wi, wj, wk], axis=-1)
        vid = np.where(w > 0, vid, 99)
        order = np.argsort(vid, axis=-1, kind="stable")
        vid = np.take_along_axis(vid, order, axis=-1)
        w = np.take_along_axis(w, order, axis=-1)
        vid = np.where(vid == 99, 12, vid)
        key = np.zeros(wi.shape[0], dtype=np.int64)
        for k in range(3):
            key = (key * 13 + vid[:, k]) * big + w[:, k]
        keys_all.append(key)
        # position from the canonical ordering so welded points are bit-identical
        p = np.zeros((wi.shape[0], 3), dtype=np.float64)
        for k in range(3):
            vk = np.where(vid[:, k] == 12, 0, vid[:, k])
            p += _ICO_V[vk] * w[:, k:k + 1]
        pts_all.append(p)
        tris_all.append(lat_tris + base)
        base += wi.shape[0]
    keys = np.concatenate(keys_all)
    pts = np.concatenate(pts_all)
    tris = np.concatenate(tris_all)
    uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    p = pts[first]
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    if noise > 0.0:
        rng = np.random.default_rng(seed)
        p = p * (1.0 + rng.uniform(-noise, noise, size=(p.shape[0], 1)))
    co = (p * radius).astype(np.float32)
    return _from_faces(co, inv[tris])


def low_freq_mask(mesh, seed=11, waves=4):
    """mask[v] = smoothstep of a seeded low-frequency field, in [0, 1] (config C4)."""
    rng = np.random.default_rng(seed)
    co = mesh.co.astype(np.float64)
    f = np.zeros(co.shape[0])
    for _ in range(waves):
        k = rng.uniform(-3.0, 3.0, size=3)
        ph = rng.uniform(0, 2 * np.pi)
        f += np.sin(co @ k + ph)
    f = (f - f.min()) / max(f.max() - f.min(), 1e-12)
    return (f * f * (3.0 - 2.0 * f)).astype(np.float32)


@dataclass
class Multires:
    """A SubdivCCG as flat tables (kernel/intern/subdiv_ccg.c): grids of one level over a quad base mesh.
    Element index = grid * grid_size^2 + y * grid_size + x; face f owns the grids [face_start[f],
    face_start[f] + face_num[f]), one per corner in MLoop order; (0, 0) is the face centre and
    (grid_size - 1, grid_size - 1) the coarse vertex of the corner (subdiv_ccg.c:397-530)."""
    grid_size: int
    co: np.ndarray           # (E, 3) float32
    no: np.ndarray           # (E, 3) float32
    mask: np.ndarray         # (E,) float32 or None
    face_start: np.ndarray   # (F,) int32
    face_num: np.ndarray     # (F,) int32
    edge_off: np.ndarray     # (NE + 1,) int32, in adjacent faces
    edge_elems: np.ndarray   # (edge_off[-1] * 2 * grid_size,) int32: SubdivCCGAdjacentEdge.boundary_coords
    cvert_off: np.ndarray    # (NV + 1,) int32
    cvert_elems: np.ndarray  # (cvert_off[-1],) int32: SubdivCCGAdjacentVertex.corner_coords
    grid_edge: np.ndarray    # (G,) int32: coarse edge leaving the grid's corner vertex (MLoop order)
    grid_cvert: np.ndarray   # (G,) int32: the grid's corner vertex
    # the topology-refiner queries KERNEL_subdiv_ccg_neighbor_coords_get makes (subdiv_ccg.c:1558-1582, 1649-1651);
    # OpenSubdiv is not in the reference tree, so the order is this generator's: a vertex's edges ascend
    edge_verts: np.ndarray = None      # (NE, 2) int32: getEdgeVertices
    cvert_edge_off: np.ndarray = None  # (NV + 1,) int32
    cvert_edges: np.ndarray = None     # getVertexEdges, ascending edge index

    @property
    def totgrid(self):
        return int(self.grid_edge.shape[0])

    @property
    def totelem(self):
        return int(self.co.shape[0])

    @property
    def totvert(self):
        return self.totelem

    def bbox_diag(self):
        return float(np.linalg.norm(self.co.max(axis=0).astype(np.float64) - self.co.min(axis=0).astype(np.float64)))


def multires_cube(base_levels=2, level=4, noise=0.01, freq=5.0, with_mask=False, spherify=True):
    """Multires cube: base = cube with 2^base_levels x 2^base_levels quads per side, `level` multires
    levels -> grid_size = 2^(level - 1) + 1, 4 grids per base quad.  C5: base 25 x 25 per side
    (use multires_cube_n), level 7.  Positions: the base quad's bilinear patch pushed onto the unit
    sphere plus a smooth deterministic bump field; duplicated elements are made bit-identical by one
    averaging pass, as the reference does when it creates the CCG (subdiv_ccg.c:1170-1189)."""
    return _multires_from_base(cube(base_levels), level, noise, freq, with_mask, spherify)


def multires_cube_n(n_per_side, level, **kw):
    """same with n x n base quads per side (n need not be a power of two)"""
    return _multires_from_base(_cube_n(n_per_side), level, kw.get("noise", 0.01), kw.get("freq", 5.0),
                               kw.get("with_mask", False), kw.get("spherify", True))


def multires_plane(n_per_side, level, noise=0.01, freq=5.0, with_mask=False):
    """open base: n x n quads of a planar height field; coarse boundary edges have one face, so the elements
    along them are boundary elements of the smooth brush"""
    return _multires_from_base(grid(n_per_side + 1, height=0.1, freq=3.0), level, noise, freq, with_mask, False)


def _cube_n(n):
    s = n + 1
    a, b = np.meshgrid(np.arange(s, dtype=np.int64), np.arange(s, dtype=np.int64), indexing="xy")
    a = a.reshape(-1)
    b = b.reshape(-1)
    zero = np.zeros_like(a)
    full = np.full_like(a, n)
    face_xyz = [(full, a, b), (zero, b, a), (b, full, a), (a, zero, b), (a, b, full), (b, a, zero)]
    keys = [x * s * s + y * s + z for (x, y, z) in face_xyz]
    uniq, inv = np.unique(np.concatenate(keys), return_inverse=True)
    co = (np.stack([uniq // (s * s), (uniq // s) % s, uniq % s], axis=-1).astype(np.float64) * (2.0 / n) - 1.0).astype(np.float32)
    ci, cj = np.meshgrid(np.arange(n, dtype=np.int64), np.arange(n, dtype=np.int64), indexing="xy")
    c00 = (cj * s + ci).reshape(-1)
    faces = []
    for f in range(6):
        g = inv[f * s * s:(f + 1) * s * s]
        faces.append(np.stack([g[c00], g[c00 + 1], g[c00 + s + 1], g[c00 + s]], axis=-1))
    return _from_faces(co, np.concatenate(faces))


def _multires_from_base(base, level, noise, freq, with_mask, spherify):
    assert np.all(base.poly_len == 4), "quad base mesh"
    gs = (1 << (level - 1)) + 1
    gs2 = gs * gs
    F = base.totpoly
    G = 4 * F
    quads = base.loop_v.reshape(F, 4).astype(np.int64)
    P = base.co.astype(np.float64)[quads]                    # (F, 4, 3)
    centre = P.mean(axis=1)                                    # (F, 3)
    mid_next = 0.5 * (P + np.roll(P, -1, axis=1))              # midpoint of edge c -> c+1
    mid_prev = np.roll(mid_next, 1, axis=1)                    # midpoint of edge c-1 -> c
    t = np.arange(gs, dtype=np.float64) / (gs - 1)
    u, v = np.meshgrid(t, t, indexing="xy")                    # u along x, v along y; element (x, y) at [y, x]
    u = u.reshape(1, 1, gs2, 1)
    v = v.reshape(1, 1, gs2, 1)
    C0 = centre[:, None, None, :]
    pos = ((1 - u) * (1 - v) * C0 + u * (1 - v) * mid_next[:, :, None, :] + u * v * P[:, :, None, :] +
           (1 - u) * v * mid_prev[:, :, None, :])              # (F, 4, gs2, 3)
    pos = pos.reshape(G * gs2, 3)
    if spherify:
        pos = pos / np.linalg.norm(pos, axis=1, keepdims=True)
    if noise:
        r = 1.0 + noise * np.sin(freq * pos[:, 0]) * np.cos(freq * pos[:, 1]) * np.sin(freq * pos[:, 2] + 0.5)
        pos = pos * r[:, None]
    co = pos.astype(np.float32)
    mask = None
    if with_mask:
        mask = (0.5 + 0.5 * np.sin(3.0 * pos[:, 0] + 1.0) * np.cos(2.0 * pos[:, 2])).astype(np.float32)
    # coarse edges: undirected vertex pairs, orientation of the first face that brings them
    loop_next = np.roll(quads, -1, axis=1)
    ea = quads.reshape(-1)
    eb = loop_next.reshape(-1)
    key = np.minimum(ea, eb) * (base.totvert + 1) + np.maximum(ea, eb)
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    NE = uniq.shape[0]
    edge_v0 = ea[first]                                        # getEdgeVertices()[0]
    grid_edge = inv.astype(np.int32)                           # per (face, corner) = per grid
    grid_cvert = ea.astype(np.int32)
    # boundary coords (subdiv_ccg.c:432-456), faces in ascending order per edge
    grid_ids = np.arange(G, dtype=np.int64)
    cur = grid_ids
    nxt = (grid_ids // 4) * 4 + (grid_ids % 4 + 1) % 4
    flipped = edge_v0[inv] != ea
    i = np.arange(gs, dtype=np.int64)

    def elem(g, x, y):
        return g[:, None] * gs2 + y * gs + x

    not_flipped_a = elem(cur, np.full_like(i, gs - 1)[None, :], (gs - 1 - i)[None, :])
    not_flipped_b = elem(nxt, i[None, :], np.full_like(i, gs - 1)[None, :])
    flipped_a = elem(nxt, (gs - 1 - i)[None, :], np.full_like(i, gs - 1)[None, :])
    flipped_b = elem(cur, np.full_like(i, gs - 1)[None, :], i[None, :])
    rows = np.where(flipped[:, None], np.concatenate([flipped_a, flipped_b], axis=1),
                    np.concatenate([not_flipped_a, not_flipped_b], axis=1))     # (G, 2 gs)
    order = np.argsort(inv, kind="stable")                     # by edge, faces (= grids) ascending
    counts = np.bincount(inv, minlength=NE)
    edge_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    edge_elems = rows[order].reshape(-1).astype(np.int32)
    # corner coords (subdiv_ccg.c:512-527)
    NV = base.totvert
    vorder = np.argsort(ea, kind="stable")
    vcounts = np.bincount(ea, minlength=NV)
    cvert_off = np.concatenate([[0], np.cumsum(vcounts)]).astype(np.int32)
    cvert_elems = (grid_ids[vorder] * gs2 + (gs - 1) * gs + (gs - 1)).astype(np.int32)
    m = Multires(grid_size=gs, co=co, no=np.zeros_like(co), mask=mask,
                 face_start=(np.arange(F, dtype=np.int32) * 4), face_num=np.full(F, 4, dtype=np.int32),
                 edge_off=edge_off, edge_elems=edge_elems, cvert_off=cvert_off, cvert_elems=cvert_elems,
                 grid_edge=grid_edge, grid_cvert=grid_cvert)
    ev = np.stack([edge_v0, eb[first]], axis=1).astype(np.int32)
    inc_v = ev.reshape(-1).astype(np.int64)
    inc_e = np.repeat(np.arange(NE, dtype=np.int64), 2)
    o = np.lexsort((inc_e, inc_v))
    m.edge_verts = ev
    m.cvert_edge_off = np.concatenate([[0], np.cumsum(np.bincount(inc_v, minlength=NV))]).astype(np.int32)
    m.cvert_edges = inc_e[o].astype(np.int32)
    _multires_make_consistent(m)
    return m


def _multires_make_consistent(m):
    """duplicated elements get one value (float32 mean in list order, as element_accumulator does);
    numpy only, so the generator does not depend on the oracle"""
    gs = m.grid_size
    gs2 = gs * gs
    F = m.face_start.shape[0]

    def avg(groups):
        # groups: (N, k) element indices; sequential float32 sum in column order, times 1/k
        k = groups.shape[1]
        for arr in (m.co, m.mask):
            if arr is None:
                continue
            acc = np.zeros((groups.shape[0],) + arr.shape[1:], dtype=np.float32)
            for j in range(k):
                acc = (acc + arr[groups[:, j]]).astype(np.float32)
            acc = (acc * np.float32(1.0 / k)).astype(np.float32)
            for j in range(k):
                arr[groups[:, j]] = acc

    g0 = m.face_start.astype(np.int64)
    i = np.arange(1, gs, dtype=np.int64)
    for c in range(4):
        prev = g0 + (c + 3) % 4
        cur = g0 + c
        a = (prev[:, None] * gs2 + 0 * gs + i[None, :]).reshape(-1)       # prev (i, 0)
        b = (cur[:, None] * gs2 + i[None, :] * gs + 0).reshape(-1)        # cur (0, i)
        avg(np.stack([a, b], axis=1))
    avg(np.stack([(g0 + c) * gs2 for c in range(4)], axis=1))
    nf = np.diff(m.edge_off)
    rows = m.edge_elems.reshape(-1, 2 * gs).astype(np.int64)
    for k in np.unique(nf):
        if k < 2:
            continue
        e = np.nonzero(nf == k)[0]
        idx = (m.edge_off[e][:, None] + np.arange(k)[None, :])           # (ne, k) rows
        grp = rows[idx][:, :, 1:2 * gs - 1]                               # (ne, k, 2gs-2)
        avg(np.transpose(grp, (0, 2, 1)).reshape(-1, k))
    nv = np.diff(m.cvert_off)
    for k in np.unique(nv):
        if k < 2:
            continue
        v = np.nonzero(nv == k)[0]
        idx = m.cvert_off[v][:, None] + np.arange(k)[None, :]
        avg(m.cvert_elems.astype(np.int64)[idx])
    assert F > 0


def mixed_grid(n, height=0.05, freq=8.0):
    """the height-field grid with mixed polygon sizes: every third cell pair of every other row is
    merged into a hexagon, every fifth remaining cell is split into two triangles.  N-gons push their
    leaves onto the general normals / bounds path; the rest stay on the tile path."""
    base = grid(n, height, freq)
    polys = []
    used = np.zeros((n - 1, n - 1), dtype=bool)

    def vid(i, j):
        return j * n + i

    for j in range(n - 1):
        for i in range(n - 1):
            if used[j, i]:
                continue
            if j % 2 == 0 and i % 3 == 0 and i + 1 < n - 1 and not used[j, i + 1]:
                used[j, i] = used[j, i + 1] = True
                polys.append([vid(i, j), vid(i + 1, j), vid(i + 2, j), vid(i + 2, j + 1), vid(i + 1, j + 1), vid(i, j + 1)])
            elif (i + j) % 5 == 1:
                used[j, i] = True
                polys.append([vid(i, j), vid(i + 1, j), vid(i + 1, j + 1)])
                polys.append([vid(i, j), vid(i + 1, j + 1), vid(i, j + 1)])
            else:
                used[j, i] = True
                polys.append([vid(i, j), vid(i + 1, j), vid(i + 1, j + 1), vid(i, j + 1)])
    lens = np.array([len(p) for p in polys], dtype=np.int32)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    loops = np.concatenate([np.asarray(p, dtype=np.int32) for p in polys])
    return Mesh(co=base.co.copy(), poly_start=starts, poly_len=lens, loop_v=loops)
