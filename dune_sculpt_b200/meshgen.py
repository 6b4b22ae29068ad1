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
