"""ctypes binding of oracle/_ref/libref.so: the REFERENCE's own hot-path functions, cut verbatim out of
/root/reference by oracle/ref_extract.py and compiled behind oracle/ref_shim.h.  TEST INFRASTRUCTURE: it exists to
pin the oracle (tests/test_ref_pin.py)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

from oracle_py import fptr, iptr, lib as oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    """the library, (re)built when /root/reference is present, else the prebuilt one; None when there is neither"""
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "oracle", "_ref", "libref.so")
        srcs = [os.path.join(ROOT, "oracle", f) for f in ("ref_extract.py", "ref_shim.h", "ref_api.c", "ref_glue.inc")]
        if os.path.isdir("/root/reference") and (not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)):
            subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_extract.py")], check=True, stdout=subprocess.DEVNULL)
        if not os.path.exists(so):
            return None
        L = C.CDLL(so)
        L.ref_build_mesh.restype = C.c_void_p
        L.ref_build_mesh.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_build_grids.restype = C.c_void_p
        L.ref_build_grids.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_free.argtypes = [C.c_void_p]
        for fn in ("ref_totnode", "ref_totprim", "ref_leaf_limit_used", "ref_update_normals", "ref_grids_inner_normals"):
            getattr(L, fn).argtypes = [C.c_void_p]
        L.ref_nodes.argtypes = [C.c_void_p] * 9
        L.ref_prim_indices.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_node_vert_indices.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_node_face_vert_indices.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_gather_sphere.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p]
        L.ref_gather_flag.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_set_co.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_get_no.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_set_no.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_node_mark_update.argtypes = [C.c_void_p, C.c_int]
        L.ref_node_set_flag.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ref_vert_mark_update.argtypes = [C.c_void_p, C.c_int]
        L.ref_vert_marked.argtypes = [C.c_void_p, C.c_int]
        L.ref_update_bounds.argtypes = [C.c_void_p, C.c_int]
        L.ref_grids_set_adjacency.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                              C.c_void_p, C.c_void_p]
        L.ref_grids_average_all.argtypes = [C.c_void_p]
        L.ref_get_co.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_get_mask.argtypes = [C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def looptris(mesh):
    """MLoopTri.tri (loop indices) and .poly by the reference's tessellation rule -- taken from the oracle's
    restatement of mesh_tessellate.c:420-447 (the tessellator itself is not part of the extracted unit)"""
    O = oracle_lib()
    n = O.or_looptri_count(mesh.totpoly, iptr(mesh.poly_len))
    tri = np.zeros((n, 3), dtype=np.int32)
    poly = np.zeros(n, dtype=np.int32)
    co = np.ascontiguousarray(mesh.co, dtype=np.float32)
    O.or_looptri_calc(mesh.totpoly, iptr(mesh.poly_start), iptr(mesh.poly_len), iptr(mesh.loop_v), fptr(co), iptr(tri), iptr(poly))
    return tri, poly


class _Session:
    def node_arrays(self):
        n = self.totnode
        vb = np.zeros((n, 6), dtype=np.float32)
        ovb = np.zeros((n, 6), dtype=np.float32)
        out = {k: np.zeros(n, dtype=np.int32) for k in ("children_offset", "flag", "prim_offset", "totprim", "uniq_verts", "face_verts")}
        self.L.ref_nodes(self.p, vb.ctypes.data, ovb.ctypes.data, *[out[k].ctypes.data for k in
                                                                     ("children_offset", "flag", "prim_offset", "totprim", "uniq_verts", "face_verts")])
        out["vb"], out["orig_vb"] = vb, ovb
        return out

    def prim_indices(self):
        out = np.zeros(self.L.ref_totprim(self.p), dtype=np.int32)
        self.L.ref_prim_indices(self.p, out.ctypes.data)
        return out

    def gather_sphere(self, center, radius_sq, original=False, ignore=True):
        c = np.asarray(center, dtype=np.float32)
        buf = np.zeros(self.totnode + 1, dtype=np.int32)
        n = self.L.ref_gather_sphere(self.p, c.ctypes.data, C.c_float(radius_sq), int(original), int(ignore), buf.ctypes.data)
        return buf[:n].copy()

    def gather_flag(self, flag):
        buf = np.zeros(self.totnode + 1, dtype=np.int32)
        n = self.L.ref_gather_flag(self.p, int(flag), buf.ctypes.data)
        return buf[:n].copy()

    def set_co(self, co):
        a = np.ascontiguousarray(co, dtype=np.float32)
        self.L.ref_set_co(self.p, a.ctypes.data)

    def no(self):
        out = np.zeros((self.totvert, 3), dtype=np.float32)
        self.L.ref_get_no(self.p, out.ctypes.data)
        return out

    def set_node_flag(self, node, flag, on=True):
        self.L.ref_node_set_flag(self.p, int(node), int(flag), int(on))

    def node_mark_update(self, node):
        self.L.ref_node_mark_update(self.p, int(node))

    def update_bounds(self, flag):
        self.L.ref_update_bounds(self.p, int(flag))

    def close(self):
        if self.p is not None:
            self.L.ref_free(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RefMesh(_Session):
    """BKE_pbvh_build_mesh (pbvh.c:2452-2514) of the reference itself over a meshgen.Mesh"""

    def __init__(self, mesh, leaf_limit=0, mask=None, poly_mat=None, poly_flag=None, vert_flag=None):
        self.L = lib()
        tri, poly = looptris(mesh)
        co = np.ascontiguousarray(mesh.co, dtype=np.float32)
        keep = [co, tri, poly]

        def opt(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data
        self.p = C.c_void_p(self.L.ref_build_mesh(
            mesh.totvert, co.ctypes.data, opt(vert_flag, np.uint8), mesh.totpoly, mesh.poly_start.ctypes.data, mesh.poly_len.ctypes.data,
            opt(poly_mat, np.int16), opt(poly_flag, np.uint8), mesh.totloop, mesh.loop_v.ctypes.data, tri.shape[0], tri.ctypes.data,
            poly.ctypes.data, opt(mask, np.float32), int(leaf_limit)))
        self.totnode = self.L.ref_totnode(self.p)
        self.totvert = mesh.totvert

    def node_vert_indices(self, i, count):
        out = np.zeros(count, dtype=np.int32)
        self.L.ref_node_vert_indices(self.p, int(i), out.ctypes.data)
        return out

    def node_face_vert_indices(self, i, totprim):
        out = np.zeros((totprim, 3), dtype=np.int32)
        self.L.ref_node_face_vert_indices(self.p, int(i), out.ctypes.data)
        return out

    def set_no(self, no):
        a = np.ascontiguousarray(no, dtype=np.float32)
        self.L.ref_set_no(self.p, a.ctypes.data)

    def vert_mark_update(self, v):
        self.L.ref_vert_mark_update(self.p, int(v))

    def vert_marked(self, v):
        return bool(self.L.ref_vert_marked(self.p, int(v)))

    def update_normals(self):
        self.L.ref_update_normals(self.p)


class RefGrids(_Session):
    """BKE_pbvh_build_grids (pbvh.c:2516-2561) of the reference itself over a meshgen.Multires"""

    def __init__(self, mr, leaf_limit=0, grid_mat=None, grid_flag=None, hidden=None):
        self.L = lib()
        co = np.ascontiguousarray(mr.co, dtype=np.float32)
        no = np.ascontiguousarray(mr.no, dtype=np.float32)
        mask = None if mr.mask is None else np.ascontiguousarray(mr.mask, dtype=np.float32)
        keep = []

        def opt(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data
        self.p = C.c_void_p(self.L.ref_build_grids(mr.totgrid, mr.grid_size, co.ctypes.data, no.ctypes.data,
                                                   None if mask is None else mask.ctypes.data, opt(grid_mat, np.int16),
                                                   opt(grid_flag, np.uint8), opt(hidden, np.uint8), int(leaf_limit)))
        self.totnode = self.L.ref_totnode(self.p)
        self.totvert = mr.totelem

        self.has_mask = mask is not None
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)  # noqa: E731
        t = [i32(mr.face_start), i32(mr.face_num), i32(mr.edge_off), i32(mr.edge_elems), i32(mr.cvert_off), i32(mr.cvert_elems)]
        self.L.ref_grids_set_adjacency(self.p, t[0].shape[0], t[0].ctypes.data, t[1].ctypes.data, t[2].shape[0] - 1, t[2].ctypes.data,
                                       t[3].ctypes.data, t[4].shape[0] - 1, t[4].ctypes.data, t[5].ctypes.data)

    def inner_normals(self):
        """subdiv_ccg_recalc_inner_face_normals + subdiv_ccg_average_inner_face_normals on every grid"""
        self.L.ref_grids_inner_normals(self.p)

    def average_all(self):
        """subdiv_ccg_average_inner_face_grids, _grids_boundary, _grids_corners over all faces / edges / vertices"""
        self.L.ref_grids_average_all(self.p)

    def co(self):
        out = np.zeros((self.totvert, 3), dtype=np.float32)
        self.L.ref_get_co(self.p, out.ctypes.data)
        return out

    def mask(self):
        if not self.has_mask:
            return None
        out = np.zeros(self.totvert, dtype=np.float32)
        self.L.ref_get_mask(self.p, out.ctypes.data)
        return out
