"""The tables of the tile kernel (dune_sculpt_b200/csrc/dsc_tile_tables.h: vertex -> looptri CSR, staged verts, local poly
entries, sliced-ELL index words, tile descriptors) are built leaf by leaf in parallel and appended in leaf order.  This
check compiles the header with g++ (no CUDA) next to the serial construction it replaced (tests/native/tile_tables_ref.inc)
and holds every table to it byte for byte, for several meshes, leaf limits, tile sizes, scrambled tiles and thread counts."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from dune_sculpt_b200 import capi, meshgen

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["tri_leaf", "vt_off+vt_idx", "leaf_sslots+leaf_sbeg", "stage", "e_pv", "e_halo_leaf", "tile_meta", "v2_goff", "v2_idx",
         "leaf_fast", "tile_dims", "any_slow_leaf"]


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(HERE, "native", "libtile_tables_check.so")
    src = os.path.join(HERE, "native", "tile_tables_check.cpp")
    deps = [src, os.path.join(HERE, "native", "tile_tables_ref.inc"),
            os.path.join(HERE, "..", "dune_sculpt_b200", "csrc", "dsc_tile_tables.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-o", out, src], check=True)
    L = C.CDLL(out)
    L.tt_hashes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    return L


def _hashes(lib, me, pd, tile, scramble, mode):
    h = (C.c_uint64 * 12)()
    r = lib.tt_hashes(C.byref(me), C.byref(pd), tile, scramble, mode, h)
    assert r == 0, r
    return list(h)


CASES = [
    ("grid", lambda: meshgen.grid(97), 200),
    ("grid-tiny-leaves", lambda: meshgen.grid(33), 7),
    ("ico-large-leaves", lambda: meshgen.icosphere(40, noise=0.02), 300),
    ("grid-default-leaves", lambda: meshgen.grid(257), 0),
    ("mixed", lambda: meshgen.mixed_grid(65), 150),      # tris, quads and n-gons: some leaves take the general path
    ("ico", lambda: meshgen.icosphere(20, noise=0.01), 111),
    ("cube", lambda: meshgen.cube(5), 150),
]


@pytest.mark.parametrize("name,mk,ll", CASES, ids=[c[0] for c in CASES])
def test_parallel_tile_tables_are_the_serial_ones(lib, name, mk, ll):
    ses = capi.SculptSession(mk(), leaf_limit=ll)
    me, pd, keep = ses.descs(with_neighbors=False)
    for tile, scramble in ((1024, 0), (256, 0), (96, 1), (1024, 1), (32, 0)):
        ref = _hashes(lib, me, pd, tile, scramble, 0)
        for threads in (1, 3, 8, 16):
            got = _hashes(lib, me, pd, tile, scramble, threads)
            bad = [NAMES[i] for i in range(12) if ref[i] != got[i]]
            assert not bad, "%s tile %d scramble %d threads %d: %s differ" % (name, tile, scramble, threads, bad)
        if name == "mixed":
            assert ref[11] == 1   # n-gons: some leaves take the general path
        if name == "grid" and not scramble:
            assert ref[11] == 0   # compact quad tiles: every leaf on the tile path
