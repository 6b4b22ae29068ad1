"""ctypes bindings of the two native libraries (no torch types cross this boundary).

libdune_sculpt_cuda.so  -- the C ABI of include/dune_sculpt_cuda.h (CUDA kernels, sm_100a)
libdune_sculpt_host.so  -- the reference-named host entry points of include/dune_pbvh.h

Loading never needs a GPU; creating a device context does and fails loudly without one.
"""
import ctypes as C
import os

import numpy as np

from .meshgen import Mesh

_LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_ubyte_p = C.POINTER(C.c_ubyte)

DSC_NUM_STAGES = 8

# tools / presets / flags (include/dune_sculpt_cuda.h)
TOOL_DRAW, TOOL_SMOOTH, TOOL_INFLATE, TOOL_GRAB, TOOL_CLAY_STRIPS = 1, 2, 4, 5, 18
CURVE_CUSTOM, CURVE_SMOOTH, CURVE_SPHERE, CURVE_ROOT, CURVE_SHARP, CURVE_LIN = 0, 1, 2, 3, 4, 5
CURVE_POW4, CURVE_INVSQUARE, CURVE_CONSTANT, CURVE_SMOOTHER = 6, 7, 8, 9
DIR_AREA, DIR_VIEW, DIR_X, DIR_Y, DIR_Z = 0, 1, 2, 3, 4
DAB_FRONTFACE, DAB_PLANE_TRIM, DAB_FIRST_STEP, DAB_NO_NORMALS, DAB_NO_BOUNDS = 1, 2, 4, 8, 16
FALLOFF_SPHERE, FALLOFF_TUBE = 0, 1
CLIP_X, CLIP_Y, CLIP_Z, LOCK_X, LOCK_Y, LOCK_Z = 1, 2, 4, 8, 16, 32
ME_HIDE = 16
PBVH_Leaf, PBVH_UpdateNormals, PBVH_UpdateBB, PBVH_UpdateOriginalBB = 1, 2, 4, 8
PBVH_FullyHidden, PBVH_FullyMasked = 1 << 10, 1 << 11
PBVH_UpdateDrawBuffers, PBVH_UpdateRedraw, PBVH_RebuildDrawBuffers = 1 << 4, 1 << 5, 1 << 9


class DscDab(C.Structure):
    _fields_ = [
        ("tool", C.c_int), ("curve_preset", C.c_int), ("flags", C.c_int), ("sculpt_plane", C.c_int),
        ("location", C.c_float * 3), ("radius", C.c_float), ("view_normal", C.c_float * 3),
        ("bstrength", C.c_float), ("scale", C.c_float * 3), ("hardness", C.c_float),
        ("normal_radius_factor", C.c_float), ("plane_offset", C.c_float), ("plane_trim", C.c_float),
        ("tip_roundness", C.c_float), ("grab_delta", C.c_float * 3), ("radius_scale", C.c_float),
        ("falloff_shape", C.c_int), ("clip_flags", C.c_int), ("clip_tolerance", C.c_float * 3), ("normal_weight", C.c_float),
    ]


class DscStrokeStats(C.Structure):
    _fields_ = [("vertex_dabs", C.c_int64), ("node_hits", C.c_int64), ("moved_verts", C.c_int64),
                ("dabs", C.c_int64), ("kernel_launches", C.c_int64), ("area_verts", C.c_int64), ("area_inside", C.c_int64),
                ("all_verts", C.c_int64), ("prims", C.c_int64), ("first_touch_verts", C.c_int64), ("refit_nodes", C.c_int64)]


class BB(C.Structure):
    _fields_ = [("bmin", C.c_float * 3), ("bmax", C.c_float * 3)]


class PBVHNode(C.Structure):
    _fields_ = [
        ("vb", BB), ("orig_vb", BB), ("children_offset", C.c_int), ("prim_indices", c_int_p),
        ("totprim", C.c_uint), ("vert_indices", c_int_p), ("uniq_verts", C.c_uint), ("face_verts", C.c_uint),
        ("face_vert_indices", c_int_p), ("flag", C.c_uint),
    ]


class PBVH(C.Structure):
    _fields_ = [
        ("nodes", C.POINTER(PBVHNode)), ("node_mem_count", C.c_int), ("totnode", C.c_int),
        ("prim_indices", c_int_p), ("totprim", C.c_int), ("totvert", C.c_int), ("leaf_limit", C.c_int),
        ("vert_normals", c_float_p), ("verts", C.c_void_p), ("mpoly", C.c_void_p), ("mloop", C.c_void_p),
        ("looptri", C.c_void_p), ("totpoly", C.c_int), ("totloop", C.c_int), ("vmask", c_float_p),
        ("vert_bitmap", C.POINTER(C.c_uint)), ("deformed", C.c_bool), ("owns_normals", C.c_bool),
        ("is_grids", C.c_int), ("grids", C.c_void_p), ("gridfaces", C.c_void_p), ("grid_flag_mats", C.c_void_p),
        ("totgrid", C.c_int), ("gridkey", C.c_int * 9), ("grid_hidden", C.c_void_p), ("subdiv_ccg", C.c_void_p),
        ("want_draw_buffers", C.c_int),
        ("device", C.c_void_p), ("device_dirty", C.c_bool), ("in_stroke", C.c_bool),
        ("normals_pinned", C.c_bool), ("verts_pinned", C.c_bool), ("grids_pinned", C.c_bool),
        ("nb_offsets", c_int_p), ("nb_indices", c_int_p), ("boundary", c_ubyte_p),
        ("dist_world", C.c_int), ("gather_whole", C.c_bool), ("synced_flag", C.c_void_p), ("host_vert_marks", C.c_bool),
    ]


MVERT = np.dtype([("co", np.float32, 3), ("flag", np.int8), ("bweight", np.int8), ("_pad", np.int8, 2)])
MPOLY = np.dtype([("loopstart", np.int32), ("totloop", np.int32), ("mat_nr", np.int16), ("flag", np.int8), ("_pad", np.int8)])
MLOOP = np.dtype([("v", np.uint32), ("e", np.uint32)])
MLOOPTRI = np.dtype([("tri", np.uint32, 3), ("poly", np.uint32)])

SEARCH_CB = C.CFUNCTYPE(C.c_bool, C.POINTER(PBVHNode), C.c_void_p)


class SculptSearchSphereData(C.Structure):
    _fields_ = [("center", c_float_p), ("radius_squared", C.c_float), ("original", C.c_bool),
                ("ignore_fully_ineffective", C.c_bool)]


# every symbol include/dune_sculpt_cuda.h declares (checked by tests/test_abi.py)
CUDA_SYMBOLS = [
    "dsc_ctx_create", "dsc_ctx_destroy", "dsc_last_error", "dsc_abi_version", "dsc_mesh_upload", "dsc_pbvh_upload",
    "dsc_recalc_normals", "dsc_set_custom_curve", "dsc_set_mask", "dsc_node_flag_set", "dsc_node_flags_apply", "dsc_vert_marks_or", "dsc_stroke_begin", "dsc_dab",
    "dsc_dabs", "dsc_state_save", "dsc_state_restore", "dsc_grids_upload", "dsc_download_mask",
    "dsc_raycast_enable", "dsc_raycast", "dsc_draw_enable", "dsc_draw_leaf_shading", "dsc_draw_update", "dsc_draw_node_buffer", "dsc_draw_download",
    "dsc_gather_readback", "dsc_search_sphere", "dsc_last_area", "dsc_debug_capture", "dsc_last_moved",
    "dsc_stroke_stats", "dsc_stroke_end", "dsc_update_normals", "dsc_update_bounds", "dsc_node_mark_update",
    "dsc_download_co", "dsc_download_mvert", "dsc_download_ccg", "dsc_host_register", "dsc_host_unregister", "dsc_download_no", "dsc_download_orig_co", "dsc_download_orig_no", "dsc_download_node_bb",
    "dsc_download_node_flags", "dsc_download_touched", "dsc_upload_co", "dsc_synchronize", "dsc_timer_start",
    "dsc_timer_stop", "dsc_stream", "dsc_stage_timing", "dsc_stage_times", "dsc_stage_name",
    "dsc_dist_unique_id", "dsc_dist_init", "dsc_dist_partition", "dsc_dist_halo_plan", "dsc_dist_free",
    "dsc_dist_owned_range", "dsc_dist_grids_plan", "dsc_dist_uses_peer_memory", "dsc_dist_exchanges_skipped",
    "dsc_dist_gather", "dsc_dist_dab_counts", "dsc_download_owned_mvert", "dsc_download_owned_ccg",
]
HOST_SYMBOLS = [
    "BKE_mesh_poly_to_tri_count", "BKE_mesh_recalc_looptri", "BKE_pbvh_new", "BKE_pbvh_build_mesh", "BKE_pbvh_free",
    "DUNE_pbvh_mesh_sizes_set", "DUNE_pbvh_mask_layer_set", "DUNE_pbvh_vert_normals_set", "DUNE_pbvh_leaf_limit_set",
    "DUNE_pbvh_device_attach", "DUNE_pbvh_device_attach_dist", "DUNE_pbvh_device_gather", "DUNE_pbvh_device_detach", "DUNE_pbvh_device_sync_to_host", "DUNE_pbvh_device_error",
    "DUNE_pbvh_device_checkpoint", "DUNE_pbvh_device_rollback", "BKE_pbvh_build_grids", "BKE_pbvh_count_grid_quads", "BKE_pbvh_node_get_grids",
    "BKE_subdiv_ccg_key_top_level", "DUNE_subdiv_ccg_from_tables", "DUNE_subdiv_ccg_free", "DUNE_pbvh_device_attach_grids",
    "DUNE_pbvh_device_attach_grids_dist", "DUNE_multires_reshape_assign_final_coords",
    "DUNE_subdiv_ccg_topology_set", "BKE_subdiv_ccg_neighbor_coords_get", "BKE_subdiv_ccg_coarse_mesh_adjacency_info_get",
    "DUNE_pbvh_draw_buffers_enable", "DUNE_pbvh_update_draw_buffers", "DUNE_pbvh_node_draw_buffer",
    "DUNE_pbvh_raycast_enable", "DUNE_pbvh_raycast_nearest",
    "BKE_pbvh_search_gather", "SCULPT_search_sphere_cb", "BKE_pbvh_node_mark_update", "BKE_pbvh_vert_mark_update",
    "BKE_pbvh_node_fully_hidden_set", "BKE_pbvh_node_fully_hidden_get", "BKE_pbvh_node_fully_masked_set",
    "BKE_pbvh_node_fully_masked_get", "BKE_pbvh_node_get_verts", "BKE_pbvh_node_num_verts", "BKE_pbvh_node_get_BB",
    "BKE_pbvh_node_get_original_BB", "BKE_pbvh_update_normals", "BKE_pbvh_update_bounds", "BKE_pbvh_vert_coords_alloc",
    "BKE_pbvh_vert_coords_apply", "BKE_pbvh_get_verts", "BKE_pbvh_get_vert_normals", "DUNE_sculpt_brush_strength",
    "DUNE_sculpt_dab_defaults", "DUNE_sculpt_dab_symmetry", "DUNE_sculpt_stroke_begin", "DUNE_sculpt_dab", "DUNE_sculpt_stroke_end",
    "DUNE_sculpt_automask_boundary_edges", "DUNE_sculpt_automask_topology", "MEM_freeN",
]

_cuda = None
_host = None


class NativeLibraryMissing(RuntimeError):
    pass


def _load(name):
    path = os.path.join(_LIB_DIR, name)
    if not os.path.exists(path):
        raise NativeLibraryMissing(
            "%s is not built: run `python -m dune_sculpt_b200.build` (nvcc, sm_100a). There is no fallback path." % path)
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


def cuda_lib():
    global _cuda
    if _cuda is None:
        L = _load("libdune_sculpt_cuda.so")
        L.dsc_last_error.restype = C.c_char_p
        L.dsc_last_error.argtypes = [C.c_void_p]
        L.dsc_stage_name.restype = C.c_char_p
        L.dsc_stream.restype = C.c_void_p
        L.dsc_stream.argtypes = [C.c_void_p]
        L.dsc_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.dsc_ctx_destroy.argtypes = [C.c_void_p]
        L.dsc_ctx_destroy.restype = None
        for fn in ("dsc_recalc_normals", "dsc_stroke_end", "dsc_update_normals", "dsc_synchronize", "dsc_timer_start",
                   "dsc_state_save", "dsc_state_restore"):
            getattr(L, fn).argtypes = [C.c_void_p]
        L.dsc_stroke_begin.argtypes = [C.c_void_p, c_float_p]
        L.dsc_dab.argtypes = [C.c_void_p, C.POINTER(DscDab)]
        L.dsc_dabs.argtypes = [C.c_void_p, C.POINTER(DscDab), C.c_int]
        L.dsc_gather_readback.argtypes = [C.c_void_p, c_int_p, C.c_int, c_int_p]
        L.dsc_search_sphere.argtypes = [C.c_void_p, c_float_p, C.c_float, C.c_int, C.c_int, c_int_p, C.c_int, c_int_p]
        L.dsc_last_area.argtypes = [C.c_void_p, c_float_p, c_float_p]
        L.dsc_debug_capture.argtypes = [C.c_void_p, C.c_int]
        L.dsc_last_moved.argtypes = [C.c_void_p, c_int_p, C.c_int, c_int_p]
        L.dsc_stroke_stats.argtypes = [C.c_void_p, C.POINTER(DscStrokeStats)]
        L.dsc_update_bounds.argtypes = [C.c_void_p, C.c_int]
        L.dsc_node_mark_update.argtypes = [C.c_void_p, C.c_int]
        L.dsc_node_flag_set.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.dsc_download_mask.argtypes = [C.c_void_p, c_float_p]
        L.dsc_draw_enable.argtypes = [C.c_void_p]
        L.dsc_draw_update.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.dsc_draw_node_buffer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), c_int_p]
        L.dsc_draw_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, c_int_p]
        for fn in ("dsc_download_co", "dsc_download_mvert", "dsc_host_register", "dsc_host_unregister", "dsc_download_no", "dsc_download_orig_co", "dsc_download_orig_no", "dsc_upload_co",
                   "dsc_set_custom_curve", "dsc_set_mask"):
            getattr(L, fn).argtypes = [C.c_void_p, c_float_p]
        L.dsc_host_register.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.dsc_host_unregister.argtypes = [C.c_void_p, C.c_void_p]
        L.dsc_download_node_bb.argtypes = [C.c_void_p, c_float_p, c_float_p]
        L.dsc_download_node_flags.argtypes = [C.c_void_p, c_int_p]
        L.dsc_download_touched.argtypes = [C.c_void_p, c_ubyte_p]
        L.dsc_timer_stop.argtypes = [C.c_void_p, c_float_p]
        L.dsc_stage_timing.argtypes = [C.c_void_p, C.c_int]
        L.dsc_stage_times.argtypes = [C.c_void_p, c_float_p, c_int_p]
        L.dsc_dist_unique_id.argtypes = [C.c_char_p]
        L.dsc_dist_owned_range.argtypes = [C.c_void_p, c_int_p]
        L.dsc_dist_grids_plan.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, c_int_p, c_ubyte_p, c_ubyte_p, c_ubyte_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.dsc_dist_uses_peer_memory.argtypes = [C.c_void_p]
        L.dsc_dist_exchanges_skipped.argtypes = [C.c_void_p, c_int_p]
        L.dsc_dist_gather.argtypes = [C.c_void_p]
        L.dsc_dist_dab_counts.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        L.dsc_download_owned_mvert.argtypes = [C.c_void_p, C.c_void_p, c_float_p]
        L.dsc_download_owned_ccg.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.dsc_dist_free.argtypes = [C.c_void_p]
        L.dsc_dist_free.restype = None
        _cuda = L
    return _cuda


def host_lib():
    global _host
    if _host is None:
        cuda_lib()
        L = _load("libdune_sculpt_host.so")
        L.BKE_pbvh_new.restype = C.POINTER(PBVH)
        L.BKE_pbvh_free.argtypes = [C.POINTER(PBVH)]
        L.BKE_pbvh_free.restype = None
        L.BKE_mesh_recalc_looptri.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.BKE_mesh_recalc_looptri.restype = None
        L.BKE_pbvh_build_mesh.argtypes = [C.POINTER(PBVH), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.BKE_pbvh_build_mesh.restype = None
        L.DUNE_pbvh_mesh_sizes_set.argtypes = [C.POINTER(PBVH), C.c_int, C.c_int]
        L.DUNE_pbvh_mask_layer_set.argtypes = [C.POINTER(PBVH), c_float_p]
        L.DUNE_pbvh_vert_normals_set.argtypes = [C.POINTER(PBVH), c_float_p]
        L.DUNE_pbvh_leaf_limit_set.argtypes = [C.POINTER(PBVH), C.c_int]
        L.DUNE_pbvh_device_attach.argtypes = [C.POINTER(PBVH), C.c_int]
        L.DUNE_pbvh_device_detach.argtypes = [C.POINTER(PBVH)]
        L.DUNE_pbvh_device_attach_dist.argtypes = [C.POINTER(PBVH), C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.DUNE_pbvh_device_gather.argtypes = [C.POINTER(PBVH)]
        L.DUNE_pbvh_device_sync_to_host.argtypes = [C.POINTER(PBVH)]
        L.BKE_pbvh_build_grids.argtypes = [C.POINTER(PBVH), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.BKE_pbvh_build_grids.restype = None
        L.BKE_subdiv_ccg_key_top_level.argtypes = [C.c_void_p, C.c_void_p]
        L.BKE_subdiv_ccg_key_top_level.restype = None
        L.DUNE_subdiv_ccg_from_tables.restype = C.c_void_p
        L.DUNE_subdiv_ccg_from_tables.argtypes = [C.c_int, C.c_int, c_float_p, c_float_p, c_float_p, C.c_int, c_int_p, c_int_p,
                                                  C.c_int, c_int_p, c_int_p, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p]
        L.DUNE_subdiv_ccg_free.argtypes = [C.c_void_p]
        L.DUNE_subdiv_ccg_free.restype = None
        L.DUNE_multires_reshape_assign_final_coords.argtypes = [C.POINTER(PBVH), C.c_void_p, C.c_void_p, C.c_void_p]
        L.DUNE_multires_reshape_assign_final_coords.restype = C.c_bool
        L.DUNE_subdiv_ccg_topology_set.argtypes = [C.c_void_p, c_int_p, c_int_p, c_int_p]
        L.DUNE_subdiv_ccg_topology_set.restype = None
        L.BKE_subdiv_ccg_neighbor_coords_get.argtypes = [C.c_void_p, C.c_void_p, C.c_bool, C.c_void_p]
        L.BKE_subdiv_ccg_neighbor_coords_get.restype = None
        L.DUNE_pbvh_device_attach_grids.argtypes = [C.POINTER(PBVH), C.c_void_p, C.c_int]
        L.DUNE_pbvh_device_attach_grids_dist.argtypes = [C.POINTER(PBVH), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.DUNE_pbvh_draw_buffers_enable.argtypes = [C.POINTER(PBVH)]
        L.DUNE_pbvh_draw_buffers_enable.restype = None
        L.DUNE_pbvh_update_draw_buffers.argtypes = [C.POINTER(PBVH), C.c_int, C.c_bool]
        L.DUNE_pbvh_raycast_enable.argtypes = [C.POINTER(PBVH)]
        L.DUNE_pbvh_raycast_enable.restype = None
        L.DUNE_pbvh_raycast_nearest.argtypes = [C.POINTER(PBVH), c_float_p, c_float_p, C.c_bool, C.c_float, c_float_p, c_int_p, c_int_p,
                                                c_float_p, C.POINTER(C.c_void_p)]
        L.DUNE_pbvh_raycast_nearest.restype = C.c_bool
        L.DUNE_pbvh_node_draw_buffer.argtypes = [C.POINTER(PBVH), C.c_void_p, C.POINTER(C.c_void_p), c_int_p]
        L.DUNE_pbvh_device_checkpoint.argtypes = [C.POINTER(PBVH)]
        L.DUNE_pbvh_device_rollback.argtypes = [C.POINTER(PBVH)]
        L.DUNE_pbvh_device_error.argtypes = [C.POINTER(PBVH)]
        L.DUNE_pbvh_device_error.restype = C.c_char_p
        L.BKE_pbvh_search_gather.argtypes = [C.POINTER(PBVH), C.c_void_p, C.c_void_p,
                                             C.POINTER(C.POINTER(C.POINTER(PBVHNode))), c_int_p]
        L.BKE_pbvh_search_gather.restype = None
        L.BKE_pbvh_update_normals.argtypes = [C.POINTER(PBVH), C.c_void_p]
        L.BKE_pbvh_update_normals.restype = None
        L.BKE_pbvh_update_bounds.argtypes = [C.POINTER(PBVH), C.c_int]
        L.BKE_pbvh_update_bounds.restype = None
        L.BKE_pbvh_node_mark_update.argtypes = [C.POINTER(PBVHNode)]
        L.BKE_pbvh_node_mark_update.restype = None
        L.BKE_pbvh_vert_mark_update.argtypes = [C.POINTER(PBVH), C.c_int]
        L.BKE_pbvh_vert_mark_update.restype = None
        for fn in ("BKE_pbvh_node_fully_hidden_set", "BKE_pbvh_node_fully_masked_set"):
            getattr(L, fn).argtypes = [C.POINTER(PBVHNode), C.c_int]
            getattr(L, fn).restype = None
        L.BKE_pbvh_vert_coords_alloc.argtypes = [C.POINTER(PBVH)]
        L.BKE_pbvh_vert_coords_alloc.restype = c_float_p
        L.BKE_pbvh_vert_coords_apply.argtypes = [C.POINTER(PBVH), c_float_p, C.c_int]
        L.BKE_pbvh_vert_coords_apply.restype = None
        L.BKE_pbvh_get_verts.argtypes = [C.POINTER(PBVH)]
        L.BKE_pbvh_get_verts.restype = C.c_void_p
        L.BKE_pbvh_get_vert_normals.argtypes = [C.POINTER(PBVH)]
        L.BKE_pbvh_get_vert_normals.restype = c_float_p
        L.MEM_freeN.argtypes = [C.c_void_p]
        L.MEM_freeN.restype = None
        L.DUNE_sculpt_brush_strength.argtypes = [C.c_int, C.c_float, C.c_float, C.c_bool, C.c_bool, C.c_float, C.c_float]
        L.DUNE_sculpt_brush_strength.restype = C.c_float
        L.DUNE_sculpt_dab_defaults.argtypes = [C.POINTER(DscDab), C.c_int]
        L.DUNE_sculpt_dab_defaults.restype = None
        L.DUNE_sculpt_dab_symmetry.argtypes = [C.POINTER(DscDab), C.c_int, C.POINTER(DscDab)]
        L.DUNE_sculpt_dab_symmetry.restype = C.c_int
        L.DUNE_sculpt_stroke_begin.argtypes = [C.POINTER(PBVH), c_float_p]
        L.DUNE_sculpt_dab.argtypes = [C.POINTER(PBVH), C.POINTER(DscDab)]
        L.DUNE_sculpt_stroke_end.argtypes = [C.POINTER(PBVH)]
        L.DUNE_sculpt_automask_boundary_edges.argtypes = [C.POINTER(PBVH), C.c_int, c_float_p]
        L.DUNE_sculpt_automask_boundary_edges.restype = None
        L.DUNE_sculpt_automask_topology.argtypes = [C.POINTER(PBVH), C.c_int, c_float_p, C.c_float, c_float_p]
        L.DUNE_sculpt_automask_topology.restype = None
        _host = L
    return _host


def nccl_unique_id():
    """rank 0 makes the NCCL id; the host broadcasts the 128 bytes to the other ranks"""
    buf = C.create_string_buffer(128)
    r = cuda_lib().dsc_dist_unique_id(buf)
    if r != 0:
        raise DeviceError("dsc_dist_unique_id: %s" % (cuda_lib().dsc_last_error(None) or b"").decode())
    return buf.raw


class DscMeshDesc(C.Structure):
    _fields_ = [("totvert", C.c_int), ("co", c_float_p), ("no", c_float_p), ("mask", c_float_p), ("totpoly", C.c_int),
                ("totloop", C.c_int), ("poly_loopstart", c_int_p), ("poly_totloop", c_int_p), ("loop_vert", c_int_p),
                ("tottri", C.c_int), ("tri_vert", c_int_p), ("tri_poly", c_int_p), ("nb_offsets", c_int_p),
                ("nb_indices", c_int_p), ("boundary", c_ubyte_p), ("vert_tail", C.POINTER(C.c_uint))]


class DscGridsDesc(C.Structure):
    _fields_ = [("totgrid", C.c_int), ("grid_size", C.c_int), ("co", c_float_p), ("no", c_float_p), ("mask", c_float_p),
                ("totface", C.c_int), ("face_start_grid", c_int_p), ("face_num_grids", c_int_p), ("totedge", C.c_int),
                ("edge_offsets", c_int_p), ("edge_elems", c_int_p), ("totcvert", C.c_int), ("cvert_offsets", c_int_p),
                ("cvert_elems", c_int_p), ("grid_edge", c_int_p), ("grid_cvert", c_int_p), ("rim_width", C.c_int),
                ("rim_neighbors", c_int_p), ("rim_boundary", c_ubyte_p), ("hidden", c_ubyte_p)]


class DscPbvhDesc(C.Structure):
    _fields_ = [("totnode", C.c_int), ("node_bb", c_float_p), ("node_orig_bb", c_float_p), ("children_offset", c_int_p),
                ("flag", c_int_p), ("prim_offset", c_int_p), ("totprim", c_int_p), ("prim_indices", c_int_p),
                ("uniq_verts", c_int_p), ("face_verts", c_int_p), ("vert_offset", c_int_p), ("vert_indices", c_int_p)]


def fptr(a):
    return a.ctypes.data_as(c_float_p)


def iptr(a):
    return a.ctypes.data_as(c_int_p)


class DeviceError(RuntimeError):
    pass


def make_dab(tool, location, radius, **kw):
    """Dab descriptor with the Brush defaults (types/types_brush_defaults.h) for `tool`."""
    d = DscDab()
    host_lib().DUNE_sculpt_dab_defaults(C.byref(d), int(tool))
    d.location[:] = [float(x) for x in location]
    d.radius = float(radius)
    for k, v in kw.items():
        if k in ("view_normal", "scale", "grab_delta", "clip_tolerance"):
            getattr(d, k)[:] = [float(x) for x in v]
        elif k in ("flags", "curve_preset", "sculpt_plane", "falloff_shape", "clip_flags"):
            setattr(d, k, int(v))
        else:
            setattr(d, k, float(v))
    return d


def dab_symmetry(d, symm):
    """the symmetry passes of one dab (DUNE_sculpt_dab_symmetry): list of DscDab, the dab itself first"""
    out = (DscDab * 8)()
    n = host_lib().DUNE_sculpt_dab_symmetry(C.byref(d), int(symm), out)
    res = []
    for i in range(n):
        c = DscDab()
        C.memmove(C.byref(c), C.byref(out[i]), C.sizeof(DscDab))
        res.append(c)
    return res


class SculptSession:
    """A mesh + its PBVH through the reference-named host API, optionally attached to a device."""

    def __init__(self, mesh: Mesh, mask=None, no=None, leaf_limit=0, device=None, dist=None, draw_buffers=False, raycast=False,
                 poly_mat=None, poly_flag=None, vert_flag=None):
        """dist = (world, rank, nccl_id_bytes) attaches this process as one rank of a partitioned PBVH;
        poly_mat / poly_flag = MPoly.mat_nr / .flag (ME_SMOOTH), vert_flag = MVert.flag (ME_HIDE)"""
        H = host_lib()
        self.H = H
        self.mesh = mesh
        self.mvert = np.zeros(mesh.totvert, dtype=MVERT)
        self.mvert["co"] = mesh.co
        if vert_flag is not None:
            self.mvert["flag"] = np.asarray(vert_flag).astype(np.int8)
        self.mpoly = np.zeros(mesh.totpoly, dtype=MPOLY)
        self.mpoly["loopstart"] = mesh.poly_start
        self.mpoly["totloop"] = mesh.poly_len
        if poly_mat is not None:
            self.mpoly["mat_nr"] = poly_mat
        if poly_flag is not None:
            self.mpoly["flag"] = np.asarray(poly_flag).astype(np.int8)
        self.mloop = np.zeros(mesh.totloop, dtype=MLOOP)
        self.mloop["v"] = mesh.loop_v.astype(np.uint32)
        self.tottri = int(H.BKE_mesh_poly_to_tri_count(mesh.totpoly, mesh.totloop))
        self.looptri = np.zeros(max(self.tottri, 1), dtype=MLOOPTRI)
        H.BKE_mesh_recalc_looptri(self.mloop.ctypes.data, self.mpoly.ctypes.data, self.mvert.ctypes.data, mesh.totloop,
                                  mesh.totpoly, self.looptri.ctypes.data)
        self.mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.float32)
        self.vnors = None if no is None else np.ascontiguousarray(no, dtype=np.float32)
        self.pbvh = H.BKE_pbvh_new()
        H.DUNE_pbvh_mesh_sizes_set(self.pbvh, mesh.totpoly, mesh.totloop)
        if leaf_limit:
            H.DUNE_pbvh_leaf_limit_set(self.pbvh, int(leaf_limit))
        if self.mask is not None:
            H.DUNE_pbvh_mask_layer_set(self.pbvh, fptr(self.mask))
        if self.vnors is not None:
            H.DUNE_pbvh_vert_normals_set(self.pbvh, fptr(self.vnors))
        H.BKE_pbvh_build_mesh(self.pbvh, None, self.mpoly.ctypes.data, self.mloop.ctypes.data, self.mvert.ctypes.data,
                              mesh.totvert, None, None, None, self.looptri.ctypes.data, self.tottri)
        self.ctx = None
        self.dist = dist
        if draw_buffers:
            H.DUNE_pbvh_draw_buffers_enable(self.pbvh)
        self.raycast_enabled = bool(raycast)
        if raycast:
            H.DUNE_pbvh_raycast_enable(self.pbvh)
        if device is not None:
            self.attach(device)

    # ---- host-side structure -------------------------------------------------------------
    @property
    def totnode(self):
        return int(self.pbvh.contents.totnode)

    def node_arrays(self):
        p = self.pbvh.contents
        n = p.totnode
        out = {k: np.zeros(n, dtype=np.int32) for k in ("children_offset", "flag", "prim_offset", "totprim", "uniq_verts", "face_verts")}
        out["vb"] = np.zeros((n, 6), dtype=np.float32)
        out["orig_vb"] = np.zeros((n, 6), dtype=np.float32)
        base = C.addressof(p.prim_indices.contents) if p.totprim else 0
        for i in range(n):
            nd = p.nodes[i]
            out["vb"][i, :3] = nd.vb.bmin[:]
            out["vb"][i, 3:] = nd.vb.bmax[:]
            out["orig_vb"][i, :3] = nd.orig_vb.bmin[:]
            out["orig_vb"][i, 3:] = nd.orig_vb.bmax[:]
            out["children_offset"][i] = nd.children_offset
            out["flag"][i] = nd.flag
            if nd.flag & PBVH_Leaf:
                out["prim_offset"][i] = (C.addressof(nd.prim_indices.contents) - base) // 4
                out["totprim"][i] = nd.totprim
                out["uniq_verts"][i] = nd.uniq_verts
                out["face_verts"][i] = nd.face_verts
        return out

    def prim_indices(self):
        p = self.pbvh.contents
        return np.ctypeslib.as_array(p.prim_indices, shape=(p.totprim,)).copy()

    def node_vert_indices(self, i):
        nd = self.pbvh.contents.nodes[i]
        return np.ctypeslib.as_array(nd.vert_indices, shape=(nd.uniq_verts + nd.face_verts,)).copy()

    def node_face_vert_indices(self, i):
        nd = self.pbvh.contents.nodes[i]
        return np.ctypeslib.as_array(nd.face_vert_indices, shape=(nd.totprim * 3,)).copy().reshape(-1, 3)

    def neighbor_tables(self):
        p = self.pbvh.contents
        v = p.totvert
        off = np.ctypeslib.as_array(p.nb_offsets, shape=(v + 1,)).copy()
        idx = np.ctypeslib.as_array(p.nb_indices, shape=(int(off[-1]),)).copy()
        bnd = np.ctypeslib.as_array(p.boundary, shape=(v,)).copy()
        return off, idx, bnd

    def descs(self, with_neighbors=True):
        """the flattened mesh / PBVH descriptors of the C ABI, from the host-side PBVH (keeps the arrays alive)"""
        m = self.mesh
        na = self.node_arrays()
        keep = {}
        keep["tri_vert"] = np.ascontiguousarray(self.mloop["v"][self.looptri["tri"][:self.tottri]].astype(np.int32))
        keep["tri_poly"] = np.ascontiguousarray(self.looptri["poly"][:self.tottri].astype(np.int32))
        keep["co"] = np.ascontiguousarray(m.co)
        me = DscMeshDesc()
        me.totvert, me.totpoly, me.totloop, me.tottri = m.totvert, m.totpoly, m.totloop, self.tottri
        me.co = fptr(keep["co"])
        me.poly_loopstart, me.poly_totloop, me.loop_vert = iptr(m.poly_start), iptr(m.poly_len), iptr(m.loop_v)
        me.tri_vert, me.tri_poly = iptr(keep["tri_vert"]), iptr(keep["tri_poly"])
        if with_neighbors:
            auto = np.zeros(m.totvert, np.float32)
            self.H.DUNE_sculpt_automask_boundary_edges(self.pbvh, 1, fptr(auto))  # builds the session tables
            off, idx, bnd = self.neighbor_tables()
            keep["off"], keep["idx"], keep["bnd"] = off.astype(np.int32), idx.astype(np.int32), bnd
            me.nb_offsets, me.nb_indices = iptr(keep["off"]), iptr(keep["idx"])
            me.boundary = keep["bnd"].ctypes.data_as(c_ubyte_p)
        leaf = (na["flag"] & PBVH_Leaf) != 0
        cnt = np.where(leaf, na["uniq_verts"] + na["face_verts"], 0).astype(np.int64)
        voff = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int32)
        vi = np.zeros(max(int(cnt.sum()), 1), np.int32)
        for i in np.nonzero(leaf)[0]:
            vi[voff[i]:voff[i] + cnt[i]] = self.node_vert_indices(int(i))
        keep.update(vb=np.ascontiguousarray(na["vb"]), ovb=np.ascontiguousarray(na["orig_vb"]), voff=voff, vi=vi,
                    prims=self.prim_indices().astype(np.int32), na=na)
        pd = DscPbvhDesc()
        pd.totnode = self.totnode
        pd.node_bb, pd.node_orig_bb = fptr(keep["vb"]), fptr(keep["ovb"])
        pd.children_offset, pd.flag = iptr(na["children_offset"]), iptr(na["flag"])
        pd.prim_offset, pd.totprim, pd.prim_indices = iptr(na["prim_offset"]), iptr(na["totprim"]), iptr(keep["prims"])
        pd.uniq_verts, pd.face_verts = iptr(na["uniq_verts"]), iptr(na["face_verts"])
        pd.vert_offset, pd.vert_indices = iptr(voff), iptr(vi)
        return me, pd, keep

    def partition(self, world):
        """leaf ranges and per-node owner of the spatial partition (host only)"""
        me, pd, keep = self.descs(with_neighbors=False)
        rng = np.zeros(world + 1, np.int32)
        owner = np.zeros(self.totnode, np.int32)
        r = cuda_lib().dsc_dist_partition(C.byref(pd), int(world), iptr(rng), iptr(owner))
        assert r == 0
        return rng, owner

    def halo_plan(self, world, rank):
        """(send_off, send_vert, recv_off, recv_vert) of one rank, in vertex ids (host only)"""
        me, pd, keep = self.descs(with_neighbors=True)
        L = cuda_lib()
        ptrs = [c_int_p() for _ in range(4)]
        r = L.dsc_dist_halo_plan(C.byref(me), C.byref(pd), int(world), int(rank), *[C.byref(p) for p in ptrs])
        assert r == 0
        soff = np.ctypeslib.as_array(ptrs[0], shape=(world + 1,)).copy()
        roff = np.ctypeslib.as_array(ptrs[2], shape=(world + 1,)).copy()
        sv = np.ctypeslib.as_array(ptrs[1], shape=(max(int(soff[-1]), 1),)).copy()[:soff[-1]]
        rv = np.ctypeslib.as_array(ptrs[3], shape=(max(int(roff[-1]), 1),)).copy()[:roff[-1]]
        for p in ptrs:
            L.dsc_dist_free(p)
        return soff, sv, roff, rv

    # ---- device ----------------------------------------------------------------------------
    def _chk(self, r):
        if r != 0:
            msg = self.H.DUNE_pbvh_device_error(self.pbvh)
            raise DeviceError("dune_sculpt_cuda error %d: %s" % (r, (msg or b"").decode()))

    def attach(self, device=0):
        if self.dist is not None and self.dist[0] > 1:
            world, rank, nid = self.dist
            self._chk(self.H.DUNE_pbvh_device_attach_dist(self.pbvh, int(device), int(world), int(rank), nid))
        else:
            self._chk(self.H.DUNE_pbvh_device_attach(self.pbvh, int(device)))
        self.ctx = C.c_void_p(self.pbvh.contents.device)
        self.D = cuda_lib()

    def stroke_begin(self, automask=None):
        a = None if automask is None else np.ascontiguousarray(automask, dtype=np.float32)
        self._chk(self.H.DUNE_sculpt_stroke_begin(self.pbvh, None if a is None else fptr(a)))

    def dab(self, d):
        self._chk(self.H.DUNE_sculpt_dab(self.pbvh, C.byref(d)))

    def dabs(self, dab_array, count):
        """a run of dabs in one C call (dab_array: ctypes array of DscDab)"""
        self.pbvh.contents.device_dirty = True
        self._chk(self.D.dsc_dabs(self.ctx, dab_array, int(count)))

    def stroke_end(self):
        self._chk(self.H.DUNE_sculpt_stroke_end(self.pbvh))

    def raycast(self, start, normal, original=False, max_depth=3.4028234663852886e38):
        """BKE_pbvh_raycast + the stroke operator's hit callback: None or dict(depth, vertex, face, normal, node)"""
        if not self.raycast_enabled:
            raise DeviceError("SculptSession(raycast=True) keeps the tables the device ray-cast needs")
        s_ = np.ascontiguousarray(start, dtype=np.float32)
        n_ = np.ascontiguousarray(normal, dtype=np.float32)
        depth = C.c_float(0)
        vert = C.c_int(0)
        face = C.c_int(0)
        fno = np.zeros(3, np.float32)
        node = C.c_void_p(0)
        if not self.H.DUNE_pbvh_raycast_nearest(self.pbvh, fptr(s_), fptr(n_), bool(original), C.c_float(max_depth), C.byref(depth), C.byref(vert),
                                                C.byref(face), fptr(fno), C.byref(node)):
            return None
        base = C.addressof(self.pbvh.contents.nodes.contents)
        return {"depth": np.float32(depth.value), "vertex": vert.value, "face": face.value, "normal": fno,
                "node": (node.value - base) // C.sizeof(PBVHNode)}

    def update_draw_buffers(self, smooth=True, show_mask=True):
        """pack the vertex buffers of the leaves flagged for a draw update, on the device"""
        self._chk(self.H.DUNE_pbvh_update_draw_buffers(self.pbvh, int(smooth), bool(show_mask)))

    def draw_buffer(self, node):
        """the packed vertex buffer of a leaf, (verts, 36) bytes"""
        n = C.c_int(0)
        self._chk(self.D.dsc_draw_download(self.ctx, int(node), None, 0, C.byref(n)))
        out = np.zeros((n.value, 36), dtype=np.uint8)
        self._chk(self.D.dsc_draw_download(self.ctx, int(node), out.ctypes.data, out.nbytes, C.byref(n)))
        return out

    def gather(self):
        """partitioned: every replica (device and host side) whole again -- collective, every rank calls it"""
        self._chk(self.H.DUNE_pbvh_device_gather(self.pbvh))

    def dist_dab_counts(self):
        """partitioned: dabs this rank skipped / ran alone / exchanged"""
        c = (C.c_longlong * 3)()
        self._chk(self.D.dsc_dist_dab_counts(self.ctx, c))
        return {"skipped": int(c[0]), "local": int(c[1]), "exchanged": int(c[2])}

    def checkpoint(self):
        """remember the resident mesh state (device-to-device)"""
        self._chk(self.H.DUNE_pbvh_device_checkpoint(self.pbvh))

    def rollback(self):
        """back to the checkpoint (device-to-device); the host arrays follow at the next sync"""
        self._chk(self.H.DUNE_pbvh_device_rollback(self.pbvh))

    def hits(self):
        n = C.c_int(0)
        buf = np.zeros(max(self.totnode, 1), dtype=np.int32)
        self._chk(self.D.dsc_gather_readback(self.ctx, iptr(buf), buf.size, C.byref(n)))
        return buf[:n.value].copy()

    def search_sphere(self, center, radius_sq, original=False, ignore=True):
        c = np.asarray(center, dtype=np.float32)
        n = C.c_int(0)
        buf = np.zeros(max(self.totnode, 1), dtype=np.int32)
        self._chk(self.D.dsc_search_sphere(self.ctx, fptr(c), C.c_float(radius_sq), int(original), int(ignore), iptr(buf),
                                           buf.size, C.byref(n)))
        return buf[:n.value].copy()

    def capture(self, on=True):
        self._chk(self.D.dsc_debug_capture(self.ctx, int(on)))

    def moved(self):
        n = C.c_int(0)
        buf = np.zeros(self.mesh.totvert, dtype=np.int32)
        self._chk(self.D.dsc_last_moved(self.ctx, iptr(buf), buf.size, C.byref(n)))
        return buf[:n.value].copy()

    def last_area(self):
        no = np.zeros(3, dtype=np.float32)
        co = np.zeros(3, dtype=np.float32)
        self._chk(self.D.dsc_last_area(self.ctx, fptr(no), fptr(co)))
        return no, co

    def _dl3(self, fn):
        out = np.zeros((self.mesh.totvert, 3), dtype=np.float32)
        self._chk(getattr(self.D, fn)(self.ctx, fptr(out)))
        return out

    def co(self):
        return self._dl3("dsc_download_co")

    def no(self):
        return self._dl3("dsc_download_no")

    def orig_co(self):
        return self._dl3("dsc_download_orig_co")

    def orig_no(self):
        return self._dl3("dsc_download_orig_no")

    def node_bb(self):
        bb = np.zeros((self.totnode, 6), dtype=np.float32)
        obb = np.zeros((self.totnode, 6), dtype=np.float32)
        self._chk(self.D.dsc_download_node_bb(self.ctx, fptr(bb), fptr(obb)))
        return bb, obb

    def node_flags(self):
        f = np.zeros(self.totnode, dtype=np.int32)
        self._chk(self.D.dsc_download_node_flags(self.ctx, iptr(f)))
        return f

    def touched(self):
        t = np.zeros(self.totnode, dtype=np.uint8)
        self._chk(self.D.dsc_download_touched(self.ctx, t.ctypes.data_as(c_ubyte_p)))
        return np.nonzero(t)[0].astype(np.int32)

    def stats(self):
        s = DscStrokeStats()
        self._chk(self.D.dsc_stroke_stats(self.ctx, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in DscStrokeStats._fields_}

    def set_node_flag(self, node, flag, on=True):
        self._chk(self.D.dsc_node_flag_set(self.ctx, int(node), int(flag), int(on)))

    # the reference's own node setters on the host PBVH: the change reaches the device with the next device call
    def node_ptr(self, node):
        return C.pointer(self.pbvh.contents.nodes[int(node)])

    def bke_node_mark_update(self, node):
        self.H.BKE_pbvh_node_mark_update(self.node_ptr(node))

    def bke_vert_mark_update(self, vert):
        self.H.BKE_pbvh_vert_mark_update(self.pbvh, int(vert))

    def bke_node_fully_hidden_set(self, node, on=True):
        self.H.BKE_pbvh_node_fully_hidden_set(self.node_ptr(node), int(on))

    def bke_node_fully_masked_set(self, node, on=True):
        self.H.BKE_pbvh_node_fully_masked_set(self.node_ptr(node), int(on))

    def bke_update_normals(self):
        self.H.BKE_pbvh_update_normals(self.pbvh, None)

    def bke_update_bounds(self, flag):
        self.H.BKE_pbvh_update_bounds(self.pbvh, int(flag))

    def sync_to_host(self):
        self._chk(self.H.DUNE_pbvh_device_sync_to_host(self.pbvh))

    def set_custom_curve(self, table):
        t = np.ascontiguousarray(table, dtype=np.float32)
        assert t.size == 257
        self._chk(self.D.dsc_set_custom_curve(self.ctx, fptr(t)))

    def synchronize(self):
        self._chk(self.D.dsc_synchronize(self.ctx))

    def timer_start(self):
        self._chk(self.D.dsc_timer_start(self.ctx))

    def timer_stop(self):
        ms = C.c_float(0)
        self._chk(self.D.dsc_timer_stop(self.ctx, C.byref(ms)))
        return float(ms.value)

    def stage_timing(self, on):
        self._chk(self.D.dsc_stage_timing(self.ctx, int(on)))

    def stage_times(self):
        ms = (C.c_float * DSC_NUM_STAGES)()
        ln = (C.c_int * DSC_NUM_STAGES)()
        self._chk(self.D.dsc_stage_times(self.ctx, ms, ln))
        return {self.D.dsc_stage_name(i).decode(): (float(ms[i]), int(ln[i])) for i in range(DSC_NUM_STAGES)}

    def close(self):
        if self.pbvh is not None:
            self.H.BKE_pbvh_free(self.pbvh)
            self.pbvh = None
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SubdivCCGStruct(C.Structure):
    _fields_ = [
        ("level", C.c_int), ("grid_size", C.c_int), ("grid_element_size", C.c_int), ("num_grids", C.c_int),
        ("grids", C.c_void_p), ("grids_storage", C.c_void_p), ("has_normal", C.c_bool), ("has_mask", C.c_bool),
        ("normal_offset", C.c_int), ("mask_offset", C.c_int), ("num_faces", C.c_int), ("faces", C.c_void_p),
        ("grid_faces", C.c_void_p), ("num_adjacent_edges", C.c_int), ("adjacent_edges", C.c_void_p),
        ("num_adjacent_vertices", C.c_int), ("adjacent_vertices", C.c_void_p), ("grid_edge", c_int_p), ("grid_vertex", c_int_p),
        ("edge_vertices", c_int_p), ("vertex_edge_offsets", c_int_p), ("vertex_edges", c_int_p),
    ]


class MDisps(C.Structure):
    _fields_ = [("totdisp", C.c_int), ("level", C.c_int), ("disps", c_float_p), ("hidden", C.c_void_p)]


class GridPaintMask(C.Structure):
    _fields_ = [("data", c_float_p), ("level", C.c_uint), ("_pad", C.c_char * 4)]


class SubdivCCGCoord(C.Structure):
    _fields_ = [("grid_index", C.c_int), ("x", C.c_short), ("y", C.c_short)]


class SubdivCCGNeighbors(C.Structure):
    _fields_ = [("coords", C.POINTER(SubdivCCGCoord)), ("size", C.c_int), ("num_duplicates", C.c_int),
                ("coords_fixed", SubdivCCGCoord * 256)]


class GridSession(SculptSession):
    """a multires CCG (meshgen.Multires) behind the reference's grids entry points: SubdivCCG ->
    BKE_pbvh_build_grids -> DUNE_pbvh_device_attach_grids; the dab / download methods are the mesh ones
    (a grid element is a vertex to them)"""

    def __init__(self, mr, leaf_limit=0, device=0, dist=None, draw_buffers=False, grid_mat=None, grid_flag=None, hidden=None, raycast=False):
        """grid_mat / grid_flag = DMFlagMat per grid; hidden = [totelem] grid_hidden bit per element"""
        H = host_lib()
        self.H = H
        self.mesh = mr
        self.flagmats = None
        if grid_mat is not None or grid_flag is not None:
            self.flagmats = np.zeros(mr.totgrid, dtype=np.dtype([("mat_nr", np.int16), ("flag", np.int8), ("_pad", np.int8)]))
            if grid_mat is not None:
                self.flagmats["mat_nr"] = grid_mat
            if grid_flag is not None:
                self.flagmats["flag"] = np.asarray(grid_flag).astype(np.int8)
        self.grid_hidden = None
        if hidden is not None:
            # BLI_bitmap ** : one bitmap per grid that hides something, NULL otherwise
            area = mr.grid_size * mr.grid_size
            h = np.asarray(hidden, dtype=np.uint8).reshape(mr.totgrid, area)
            self._gh_maps = []
            self.grid_hidden = (C.c_void_p * mr.totgrid)()
            for g in range(mr.totgrid):
                if not h[g].any():
                    continue
                words = np.zeros((area >> 5) + 1, dtype=np.uint32)
                idx = np.nonzero(h[g])[0]
                np.bitwise_or.at(words, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
                self._gh_maps.append(words)
                self.grid_hidden[g] = words.ctypes.data
        self.dist = dist
        level = int(np.log2(mr.grid_size - 1)) + 1
        assert (1 << (level - 1)) + 1 == mr.grid_size
        co = np.ascontiguousarray(mr.co, dtype=np.float32)
        no = np.ascontiguousarray(mr.no, dtype=np.float32)
        mask = None if mr.mask is None else np.ascontiguousarray(mr.mask, dtype=np.float32)
        have_no = bool(np.any(no != 0.0))
        self.ccg = C.c_void_p(H.DUNE_subdiv_ccg_from_tables(
            level, mr.totgrid, fptr(co), fptr(no) if have_no else None, None if mask is None else fptr(mask),
            int(mr.face_start.shape[0]), iptr(mr.face_start), iptr(mr.face_num), int(mr.edge_off.shape[0] - 1), iptr(mr.edge_off),
            iptr(mr.edge_elems), int(mr.cvert_off.shape[0] - 1), iptr(mr.cvert_off), iptr(mr.cvert_elems), iptr(mr.grid_edge),
            iptr(mr.grid_cvert)))
        if getattr(mr, "edge_verts", None) is not None:
            # the coarse topology the element-neighbour lookup asks the refiner for (smooth brush)
            H.DUNE_subdiv_ccg_topology_set(self.ccg, iptr(np.ascontiguousarray(mr.edge_verts, dtype=np.int32)),
                                           iptr(np.ascontiguousarray(mr.cvert_edge_off, dtype=np.int32)),
                                           iptr(np.ascontiguousarray(mr.cvert_edges, dtype=np.int32)))
        self.key = (C.c_int * 9)()
        H.BKE_subdiv_ccg_key_top_level(self.key, self.ccg)
        ccg = C.cast(self.ccg, C.POINTER(SubdivCCGStruct)).contents
        self.pbvh = H.BKE_pbvh_new()
        if leaf_limit:
            H.DUNE_pbvh_leaf_limit_set(self.pbvh, int(leaf_limit))
        H.BKE_pbvh_build_grids(self.pbvh, ccg.grids, mr.totgrid, self.key, ccg.grid_faces,
                               None if self.flagmats is None else self.flagmats.ctypes.data, self.grid_hidden)
        self.ctx = None
        if draw_buffers:
            H.DUNE_pbvh_draw_buffers_enable(self.pbvh)
        self.raycast_enabled = bool(raycast)
        if raycast:
            H.DUNE_pbvh_raycast_enable(self.pbvh)
        if device is not None:
            if dist is None:
                self._chk(H.DUNE_pbvh_device_attach_grids(self.pbvh, self.ccg, int(device)))
            else:
                world, rank, nid = dist
                self._chk(H.DUNE_pbvh_device_attach_grids_dist(self.pbvh, self.ccg, int(device), int(world), int(rank), nid))
            self.ctx = C.c_void_p(self.pbvh.contents.device)
            self.D = cuda_lib()

    def mask(self):
        out = np.zeros(self.mesh.totelem, dtype=np.float32)
        self._chk(self.D.dsc_download_mask(self.ctx, fptr(out)))
        return out

    def descs(self, with_neighbors=True):
        """(DscGridsDesc, DscPbvhDesc, keep-alive) of the C ABI from the generator's tables and the host-side grids PBVH;
        with_neighbors builds the rim neighbour table through the host's BKE_subdiv_ccg_neighbor_coords_get (small meshes)"""
        mr = self.mesh
        na = self.node_arrays()
        keep = {"na": na, "co": np.ascontiguousarray(mr.co, dtype=np.float32), "prims": self.prim_indices().astype(np.int32),
                "vb": np.ascontiguousarray(na["vb"]), "ovb": np.ascontiguousarray(na["orig_vb"])}
        gd = DscGridsDesc()
        gd.totgrid, gd.grid_size, gd.co = mr.totgrid, mr.grid_size, fptr(keep["co"])
        gd.totface, gd.face_start_grid, gd.face_num_grids = int(mr.face_start.shape[0]), iptr(mr.face_start), iptr(mr.face_num)
        gd.totedge, gd.edge_offsets, gd.edge_elems = int(mr.edge_off.shape[0] - 1), iptr(mr.edge_off), iptr(mr.edge_elems)
        gd.totcvert, gd.cvert_offsets, gd.cvert_elems = int(mr.cvert_off.shape[0] - 1), iptr(mr.cvert_off), iptr(mr.cvert_elems)
        gd.grid_edge, gd.grid_cvert = iptr(mr.grid_edge), iptr(mr.grid_cvert)
        if with_neighbors and getattr(mr, "edge_verts", None) is not None:
            gs = mr.grid_size
            rim = [(b, 0) for b in range(gs)] + [(b, gs - 1) for b in range(gs)] + [(0, y) for y in range(1, gs - 1)] + \
                  [(gs - 1, y) for y in range(1, gs - 1)]
            rows = [[self.neighbors(g * gs * gs + y * gs + x)[0] for (x, y) in rim] for g in range(mr.totgrid)]
            width = max(4, max(len(r) for gr in rows for r in gr))
            tab = np.full((mr.totgrid, len(rim), width), -1, dtype=np.int32)
            for g, gr in enumerate(rows):
                for b, r in enumerate(gr):
                    tab[g, b, :len(r)] = r
            keep["rim"] = tab
            gd.rim_width, gd.rim_neighbors = width, iptr(tab)
        pd = DscPbvhDesc()
        pd.totnode = self.totnode
        pd.node_bb, pd.node_orig_bb = fptr(keep["vb"]), fptr(keep["ovb"])
        pd.children_offset, pd.flag = iptr(na["children_offset"]), iptr(na["flag"])
        pd.prim_offset, pd.totprim, pd.prim_indices = iptr(na["prim_offset"]), iptr(na["totprim"]), iptr(keep["prims"])
        pd.uniq_verts, pd.face_verts = iptr(na["uniq_verts"]), iptr(na["face_verts"])
        return gd, pd, keep

    def grids_plan(self, world, rank, with_neighbors=True):
        """dsc_dist_grids_plan of one rank (host only): dict of grid_owner, face_dom, edge_mine, cvert_mine and the
        per-peer send / receive element lists"""
        gd, pd, keep = self.descs(with_neighbors=with_neighbors)
        L = cuda_lib()
        owner = np.zeros(gd.totgrid, np.int32)
        face_dom = np.zeros(max(gd.totface, 1), np.uint8)
        edge_mine = np.zeros(max(gd.totedge, 1), np.uint8)
        cvert_mine = np.zeros(max(gd.totcvert, 1), np.uint8)
        ptrs = [c_int_p() for _ in range(4)]
        r = L.dsc_dist_grids_plan(C.byref(gd), C.byref(pd), int(world), int(rank), iptr(owner), face_dom.ctypes.data_as(c_ubyte_p),
                                  edge_mine.ctypes.data_as(c_ubyte_p), cvert_mine.ctypes.data_as(c_ubyte_p),
                                  *[C.byref(p) for p in ptrs])
        assert r == 0
        soff = np.ctypeslib.as_array(ptrs[0], shape=(world + 1,)).copy()
        roff = np.ctypeslib.as_array(ptrs[2], shape=(world + 1,)).copy()
        se = np.ctypeslib.as_array(ptrs[1], shape=(max(int(soff[-1]), 1),)).copy()[:soff[-1]]
        re_ = np.ctypeslib.as_array(ptrs[3], shape=(max(int(roff[-1]), 1),)).copy()[:roff[-1]]
        for p in ptrs:
            L.dsc_dist_free(p)
        return dict(grid_owner=owner, face_dom=face_dom[:gd.totface], edge_mine=edge_mine[:gd.totedge],
                    cvert_mine=cvert_mine[:gd.totcvert], send_off=soff, send_elem=se, recv_off=roff, recv_elem=re_)

    def multires_write_back(self, with_mask=True):
        """multires_reshape_assign_final_coords_from_ccg through the host library: (disps [G, gs^2, 3], masks [G, gs^2] or
        None) as the per-loop MDisps / GridPaintMask arrays receive them"""
        mr = self.mesh
        G, area = mr.totgrid, mr.grid_size ** 2
        level = int(np.log2(mr.grid_size - 1)) + 1
        disps = np.full((G, area, 3), np.nan, dtype=np.float32)
        masks = np.full((G, area), np.nan, dtype=np.float32) if (with_mask and mr.mask is not None) else None
        md = (MDisps * G)()
        gm = (GridPaintMask * G)() if masks is not None else None
        for g in range(G):
            md[g].totdisp, md[g].level = area, level
            md[g].disps = disps[g].ctypes.data_as(c_float_p)
            if gm is not None:
                gm[g].data, gm[g].level = masks[g].ctypes.data_as(c_float_p), level
        ok = self.H.DUNE_multires_reshape_assign_final_coords(self.pbvh, self.ccg, md, gm)
        assert ok
        return disps, masks

    def neighbors(self, elem, include_duplicates=False):
        """BKE_subdiv_ccg_neighbor_coords_get of the host library -> (element indices, num_duplicates)"""
        gs = self.mesh.grid_size
        c = SubdivCCGCoord(int(elem) // (gs * gs), (int(elem) % (gs * gs)) % gs, (int(elem) % (gs * gs)) // gs)
        nb = SubdivCCGNeighbors()
        self.H.BKE_subdiv_ccg_neighbor_coords_get(self.ccg, C.byref(c), bool(include_duplicates), C.byref(nb))
        out = np.array([nb.coords[i].grid_index * gs * gs + nb.coords[i].y * gs + nb.coords[i].x for i in range(nb.size)],
                       dtype=np.int64)
        if C.addressof(nb.coords.contents) != C.addressof(nb.coords_fixed):
            self.H.MEM_freeN(nb.coords)
        return out, nb.num_duplicates

    def host_elements(self):
        """(co, no, mask) as the host's CCGElem storage holds them after a sync"""
        ccg = C.cast(self.ccg, C.POINTER(SubdivCCGStruct)).contents
        nfl = ccg.grid_element_size // 4
        raw = np.ctypeslib.as_array(C.cast(ccg.grids_storage, c_float_p), shape=(self.mesh.totelem * nfl,)).reshape(-1, nfl)
        co = raw[:, 0:3].copy()
        no = raw[:, ccg.normal_offset // 4:ccg.normal_offset // 4 + 3].copy()
        mask = raw[:, ccg.mask_offset // 4].copy() if ccg.has_mask else None
        return co, no, mask

    def close(self):
        super().close()
        if getattr(self, "ccg", None):
            self.H.DUNE_subdiv_ccg_free(self.ccg)
            self.ccg = None
