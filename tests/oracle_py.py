"""ctypes binding of the CPU oracle (oracle/oracle.h).  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)


class OrBB(C.Structure):
    _fields_ = [("bmin", C.c_float * 3), ("bmax", C.c_float * 3)]


class OrDab(C.Structure):
    _fields_ = [
        ("tool", C.c_int), ("curve_preset", C.c_int), ("flags", C.c_int), ("sculpt_plane", C.c_int),
        ("location", C.c_float * 3), ("radius", C.c_float), ("view_normal", C.c_float * 3),
        ("bstrength", C.c_float), ("scale", C.c_float * 3), ("hardness", C.c_float),
        ("normal_radius_factor", C.c_float), ("plane_offset", C.c_float), ("plane_trim", C.c_float),
        ("tip_roundness", C.c_float), ("grab_delta", C.c_float * 3), ("radius_scale", C.c_float),
        ("falloff_shape", C.c_int), ("clip_flags", C.c_int), ("clip_tolerance", C.c_float * 3), ("normal_weight", C.c_float),
    ]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "build", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(path)
        L.or_pbvh_build_mesh.restype = C.c_void_p
        L.or_pbvh_build_mesh.argtypes = [C.c_int, c_float_p, c_float_p, c_float_p, C.c_int, c_int_p, c_int_p, C.c_int,
                                         c_int_p, C.c_int]
        L.or_pbvh_free.argtypes = [C.c_void_p]
        for fn in ("or_pbvh_totnode", "or_pbvh_tottri", "or_pbvh_totvert"):
            getattr(L, fn).argtypes = [C.c_void_p]
        L.or_pbvh_nodes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [c_int_p] * 6
        for fn in ("or_pbvh_prim_indices", "or_pbvh_tri_verts", "or_pbvh_tri_poly"):
            getattr(L, fn).restype = c_int_p
            getattr(L, fn).argtypes = [C.c_void_p]
        for fn in ("or_pbvh_node_vert_indices", "or_pbvh_node_face_vert_indices"):
            getattr(L, fn).restype = c_int_p
            getattr(L, fn).argtypes = [C.c_void_p, C.c_int]
        for fn in ("or_pbvh_co", "or_pbvh_no", "or_pbvh_orig_co", "or_pbvh_orig_no"):
            getattr(L, fn).restype = c_float_p
            getattr(L, fn).argtypes = [C.c_void_p]
        L.or_pbvh_node_set_flag.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.or_gather_sphere.argtypes = [C.c_void_p, c_float_p, C.c_float, C.c_int, C.c_int, c_int_p]
        L.or_gather_flag.argtypes = [C.c_void_p, C.c_int, c_int_p]
        L.or_vert_mark_update.argtypes = [C.c_void_p, C.c_int]
        L.or_node_mark_update.argtypes = [C.c_void_p, C.c_int]
        L.or_update_normals.argtypes = [C.c_void_p]
        L.or_update_bounds.argtypes = [C.c_void_p, C.c_int]
        L.or_recalc_all_normals.argtypes = [C.c_void_p]
        L.or_set_threads.argtypes = [C.c_int]
        L.or_set_ordered_normals.argtypes = [C.c_int]
        L.or_stroke_begin.argtypes = [C.c_void_p, c_float_p]
        L.or_stroke_end.argtypes = [C.c_void_p]
        L.or_set_custom_curve.argtypes = [C.c_void_p, c_float_p]
        L.or_dab.argtypes = [C.c_void_p, C.POINTER(OrDab)]
        L.or_last_hits.argtypes = [C.c_void_p, c_int_p]
        L.or_last_moved.argtypes = [C.c_void_p, c_int_p]
        L.or_last_area.argtypes = [C.c_void_p, c_float_p, c_float_p]
        L.or_touched_nodes.argtypes = [C.c_void_p, c_int_p]
        L.or_stroke_vertex_dabs.argtypes = [C.c_void_p]
        L.or_stroke_vertex_dabs.restype = C.c_int64
        L.or_brush_curve_strength.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.or_brush_curve_strength.restype = C.c_float
        L.or_looptri_count.argtypes = [C.c_int, c_int_p]
        L.or_looptri_calc.argtypes = [C.c_int, c_int_p, c_int_p, c_int_p, c_float_p, c_int_p, c_int_p]
        L.or_vert_neighbors.argtypes = [C.c_int, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, C.c_void_p]
        L.or_raycast.argtypes = [C.c_void_p, c_float_p, c_float_p, C.c_int, C.c_float, c_float_p, c_int_p, c_int_p, c_float_p, c_int_p]
        L.or_draw_buffers_update.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.or_pbvh_build_grids.restype = C.c_void_p
        L.or_pbvh_build_grids.argtypes = [C.c_int, C.c_int, c_float_p, c_float_p, c_float_p, C.c_int, c_int_p, c_int_p, C.c_int,
                                          c_int_p, c_int_p, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p, C.c_int]
        L.or_grids_average_all.argtypes = [C.c_void_p]
        L.or_grids_recalc_normals.argtypes = [C.c_void_p]
        L.or_grids_inner_normals.argtypes = [C.c_void_p]
        L.or_pbvh_next_build_attrs.argtypes = [C.c_void_p] * 6
        L.or_grids_set_topology.argtypes = [C.c_void_p, c_int_p, c_int_p, c_int_p]
        L.or_grids_neighbors.argtypes = [C.c_void_p, C.c_int, c_int_p]
        L.or_grids_max_neighbors.argtypes = [C.c_void_p]
        L.or_grids_is_boundary.argtypes = [C.c_void_p, C.c_int]
        L.or_pbvh_mask.restype = c_float_p
        L.or_pbvh_mask.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def fptr(a):
    return a.ctypes.data_as(c_float_p)


def iptr(a):
    return a.ctypes.data_as(c_int_p)


def dab_from(d):
    """copy a product DscDab (same field layout by design of the dab descriptor) into an OrDab"""
    o = OrDab()
    C.memmove(C.byref(o), C.byref(d), C.sizeof(OrDab))
    return o


def _next_attrs(L, poly_mat=None, poly_flag=None, vert_flag=None, grid_mat=None, grid_flag=None, grid_hidden=None):
    """material / visibility inputs of the next build (the build copies them)"""
    keep = []

    def opt(a, dt):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data
    L.or_pbvh_next_build_attrs(opt(poly_mat, np.int16), opt(poly_flag, np.uint8), opt(vert_flag, np.uint8), opt(grid_mat, np.int16),
                               opt(grid_flag, np.uint8), opt(grid_hidden, np.uint8))
    L._attrs_keep = keep  # alive until the build has copied them


class Oracle:
    def __init__(self, mesh, mask=None, no=None, leaf_limit=0, threads=1, poly_mat=None, poly_flag=None, vert_flag=None):
        L = lib()
        _next_attrs(L, poly_mat=poly_mat, poly_flag=poly_flag, vert_flag=vert_flag)
        self.L = L
        self.mesh = mesh
        co = np.ascontiguousarray(mesh.co, dtype=np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.float32)
        n = None if no is None else np.ascontiguousarray(no, dtype=np.float32)
        L.or_set_threads(int(threads))
        self.p = C.c_void_p(L.or_pbvh_build_mesh(mesh.totvert, fptr(co), None if n is None else fptr(n),
                                                 None if m is None else fptr(m), mesh.totpoly, iptr(mesh.poly_start),
                                                 iptr(mesh.poly_len), mesh.totloop, iptr(mesh.loop_v), int(leaf_limit)))
        self.totnode = L.or_pbvh_totnode(self.p)
        self.tottri = L.or_pbvh_tottri(self.p)
        self.totvert = mesh.totvert
        if no is None:
            L.or_recalc_all_normals(self.p)

    def set_threads(self, n):
        self.L.or_set_threads(int(n))

    def node_arrays(self):
        n = self.totnode
        vb = np.zeros((n, 6), dtype=np.float32)
        ovb = np.zeros((n, 6), dtype=np.float32)
        out = {k: np.zeros(n, dtype=np.int32) for k in ("children_offset", "flag", "prim_offset", "totprim", "uniq_verts", "face_verts")}
        self.L.or_pbvh_nodes(self.p, vb.ctypes.data, ovb.ctypes.data, iptr(out["children_offset"]), iptr(out["flag"]),
                             iptr(out["prim_offset"]), iptr(out["totprim"]), iptr(out["uniq_verts"]), iptr(out["face_verts"]))
        out["vb"] = vb
        out["orig_vb"] = ovb
        return out

    def prim_indices(self):
        return np.ctypeslib.as_array(self.L.or_pbvh_prim_indices(self.p), shape=(self.tottri,)).copy()

    def tri_verts(self):
        return np.ctypeslib.as_array(self.L.or_pbvh_tri_verts(self.p), shape=(self.tottri * 3,)).copy().reshape(-1, 3)

    def tri_poly(self):
        return np.ctypeslib.as_array(self.L.or_pbvh_tri_poly(self.p), shape=(self.tottri,)).copy()

    def node_vert_indices(self, i, count):
        return np.ctypeslib.as_array(self.L.or_pbvh_node_vert_indices(self.p, i), shape=(count,)).copy()

    def node_face_vert_indices(self, i, totprim):
        return np.ctypeslib.as_array(self.L.or_pbvh_node_face_vert_indices(self.p, i), shape=(totprim * 3,)).copy().reshape(-1, 3)

    def _v3(self, fn):
        return np.ctypeslib.as_array(getattr(self.L, fn)(self.p), shape=(self.totvert * 3,)).copy().reshape(-1, 3)

    def co(self):
        return self._v3("or_pbvh_co")

    def no(self):
        return self._v3("or_pbvh_no")

    def orig_co(self):
        return self._v3("or_pbvh_orig_co")

    def orig_no(self):
        return self._v3("or_pbvh_orig_no")

    def set_co(self, co):
        a = np.ctypeslib.as_array(self.L.or_pbvh_co(self.p), shape=(self.totvert * 3,))
        a[:] = np.ascontiguousarray(co, dtype=np.float32).reshape(-1)

    def set_node_flag(self, node, flag, on=True):
        self.L.or_pbvh_node_set_flag(self.p, int(node), int(flag), int(on))

    def gather_sphere(self, center, radius_sq, original=False, ignore=True):
        c = np.asarray(center, dtype=np.float32)
        buf = np.zeros(self.totnode + 1, dtype=np.int32)
        n = self.L.or_gather_sphere(self.p, fptr(c), C.c_float(radius_sq), int(original), int(ignore), iptr(buf))
        return buf[:n].copy()

    def stroke_begin(self, automask=None):
        a = None if automask is None else np.ascontiguousarray(automask, dtype=np.float32)
        self.L.or_stroke_begin(self.p, None if a is None else fptr(a))

    def stroke_end(self):
        self.L.or_stroke_end(self.p)

    def set_custom_curve(self, table):
        t = np.ascontiguousarray(table, dtype=np.float32)
        self.L.or_set_custom_curve(self.p, fptr(t))

    def dab(self, d):
        o = d if isinstance(d, OrDab) else dab_from(d)
        r = self.L.or_dab(self.p, C.byref(o))
        assert r >= 0, "oracle does not implement tool %d" % o.tool
        return r

    def hits(self):
        buf = np.zeros(self.totnode + 1, dtype=np.int32)
        n = self.L.or_last_hits(self.p, iptr(buf))
        return buf[:n].copy()

    def moved(self):
        buf = np.zeros(self.totvert + 1, dtype=np.int32)
        n = self.L.or_last_moved(self.p, iptr(buf))
        return buf[:n].copy()

    def last_area(self):
        no = np.zeros(3, dtype=np.float32)
        co = np.zeros(3, dtype=np.float32)
        self.L.or_last_area(self.p, fptr(no), fptr(co))
        return no, co

    def touched(self):
        buf = np.zeros(self.totnode + 1, dtype=np.int32)
        n = self.L.or_touched_nodes(self.p, iptr(buf))
        return buf[:n].copy()

    def vertex_dabs(self):
        return int(self.L.or_stroke_vertex_dabs(self.p))

    def update_normals(self):
        self.L.or_update_normals(self.p)

    def update_bounds(self, flag):
        self.L.or_update_bounds(self.p, int(flag))

    def raycast(self, start, normal, original=False, max_depth=3.4028234663852886e38):
        """BKE_pbvh_raycast + the stroke operator's hit callback: None or dict(depth, vertex, face, normal, node)"""
        s_ = np.asarray(start, dtype=np.float32)
        n_ = np.asarray(normal, dtype=np.float32)
        depth = np.zeros(1, np.float32)
        vert = np.zeros(1, np.int32)
        face = np.zeros(1, np.int32)
        node = np.zeros(1, np.int32)
        fno = np.zeros(3, np.float32)
        if not self.L.or_raycast(self.p, fptr(s_), fptr(n_), int(original), C.c_float(max_depth), fptr(depth), iptr(vert), iptr(face), fptr(fno), iptr(node)):
            return None
        return {"depth": depth[0], "vertex": int(vert[0]), "face": int(face[0]), "normal": fno, "node": int(node[0])}

    def draw_buffer(self, node, totprim, smooth=True, show_mask=True):
        """the leaf's packed VBO (gpu_buffers.c:174-305), (totprim * 3, 36) bytes"""
        out = np.zeros((totprim * 3, 36), dtype=np.uint8)
        n = self.L.or_draw_buffers_update(self.p, int(node), int(smooth), int(show_mask), out.ctypes.data)
        assert n == totprim * 3
        return out

    def curve_strength(self, preset, p, length):
        return float(self.L.or_brush_curve_strength(self.p, int(preset), C.c_float(p), C.c_float(length)))

    def close(self):
        if self.p is not None:
            self.L.or_pbvh_free(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GridOracle(Oracle):
    """the oracle over a multires CCG (meshgen.Multires): PBVH_GRIDS, prims are grids, "vertices" are
    grid elements"""

    def __init__(self, mr, leaf_limit=0, threads=1, recalc_normals=True, grid_mat=None, grid_flag=None, hidden=None):
        L = lib()
        _next_attrs(L, grid_mat=grid_mat, grid_flag=grid_flag, grid_hidden=hidden)
        self.L = L
        self.mesh = mr
        L.or_set_threads(int(threads))
        co = np.ascontiguousarray(mr.co, dtype=np.float32)
        no = np.ascontiguousarray(mr.no, dtype=np.float32)
        mask = None if mr.mask is None else np.ascontiguousarray(mr.mask, dtype=np.float32)
        self.p = C.c_void_p(L.or_pbvh_build_grids(
            mr.totgrid, mr.grid_size, fptr(co), fptr(no), None if mask is None else fptr(mask), int(mr.face_start.shape[0]),
            iptr(mr.face_start), iptr(mr.face_num), int(mr.edge_off.shape[0] - 1), iptr(mr.edge_off), iptr(mr.edge_elems),
            int(mr.cvert_off.shape[0] - 1), iptr(mr.cvert_off), iptr(mr.cvert_elems), iptr(mr.grid_edge), iptr(mr.grid_cvert),
            int(leaf_limit)))
        self.totnode = L.or_pbvh_totnode(self.p)
        self.tottri = L.or_pbvh_tottri(self.p)  # prims = grids
        self.totvert = mr.totelem
        if getattr(mr, "edge_verts", None) is not None:
            L.or_grids_set_topology(self.p, iptr(np.ascontiguousarray(mr.edge_verts, dtype=np.int32)),
                                    iptr(np.ascontiguousarray(mr.cvert_edge_off, dtype=np.int32)),
                                    iptr(np.ascontiguousarray(mr.cvert_edges, dtype=np.int32)))
        if recalc_normals:
            L.or_grids_recalc_normals(self.p)

    def draw_buffer(self, node, totprim, smooth=True, show_mask=True):
        """the grid leaf's packed VBO (gpu_buffers.c:548-725): (records, 36) bytes"""
        gs = self.mesh.grid_size
        per = gs * gs if smooth else (gs - 1) * (gs - 1) * 4
        out = np.zeros((totprim * per, 36), dtype=np.uint8)
        n = self.L.or_draw_buffers_update(self.p, int(node), int(smooth), int(show_mask), out.ctypes.data)
        assert n == totprim * per
        return out

    def neighbors(self, elem):
        """KERNEL_subdiv_ccg_neighbor_coords_get without duplicates: element indices, reference order"""
        buf = np.zeros(self.L.or_grids_max_neighbors(self.p), dtype=np.int32)
        n = self.L.or_grids_neighbors(self.p, int(elem), iptr(buf))
        return buf[:n].copy()

    def is_boundary(self, elem):
        return bool(self.L.or_grids_is_boundary(self.p, int(elem)))

    def mask(self):
        ptr = self.L.or_pbvh_mask(self.p)
        if not ptr:
            return None
        return np.ctypeslib.as_array(ptr, shape=(self.totvert,)).copy()
