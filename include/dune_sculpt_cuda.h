/* libdune_sculpt_cuda -- C ABI of the B200 sculpt-stroke path.
 *
 * Plain C, POD only, caller-allocated outputs.  These are the entry points the reference's host C
 * code (kernel/intern/pbvh.c, kernel/intern/paint.c and the stroke operator) binds to; each one
 * names the reference interface it sits behind.  Citations are relative to
 * /root/reference/source/dune/ ; "pbvh.c" = kernel/intern/pbvh.c second copy (lines 1913-4986).
 * See INTEGRATION.md for the call sites a maintainer adds.
 *
 * Threading: one DscContext per PBVH, driven from one host thread (the reference calls the PBVH
 * API from the main thread only, pbvh.c:4953-4959).  All device work is queued on the context's
 * stream; dsc_dab() returns without waiting.  Functions that hand data back synchronise.
 *
 * Errors: every function returns DSC_OK (0) or a negative DscStatus; dsc_last_error() gives text.
 * There is no CPU fallback: without a usable device dsc_ctx_create() fails.
 */
#ifndef DUNE_SCULPT_CUDA_H
#define DUNE_SCULPT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct DscContext DscContext;

typedef enum DscStatus {
  DSC_OK = 0,
  DSC_ERR_NO_DEVICE = -1,
  DSC_ERR_CUDA = -2,
  DSC_ERR_INVALID = -3,
  DSC_ERR_STATE = -4,
  DSC_ERR_UNSUPPORTED = -5,
  DSC_ERR_NCCL = -6,
} DscStatus;

/* PBVHNodeFlags (names at pbvh.c:3643-3726; values of the absent public header) */
enum {
  DSC_PBVH_Leaf = 1 << 0,
  DSC_PBVH_UpdateNormals = 1 << 1,
  DSC_PBVH_UpdateBB = 1 << 2,
  DSC_PBVH_UpdateOriginalBB = 1 << 3,
  DSC_PBVH_UpdateDrawBuffers = 1 << 4,
  DSC_PBVH_UpdateRedraw = 1 << 5,
  DSC_PBVH_UpdateMask = 1 << 6,
  DSC_PBVH_UpdateVisibility = 1 << 8,
  DSC_PBVH_RebuildDrawBuffers = 1 << 9,
  DSC_PBVH_FullyHidden = 1 << 10,
  DSC_PBVH_FullyMasked = 1 << 11,
  DSC_PBVH_FullyUnmasked = 1 << 12,
  DSC_PBVH_UpdateColor = 1 << 14,
};

/* Brush.sculpt_tool, types/types_brush_enums.h:410-443 */
enum { DSC_TOOL_DRAW = 1, DSC_TOOL_SMOOTH = 2, DSC_TOOL_INFLATE = 4, DSC_TOOL_GRAB = 5, DSC_TOOL_CLAY_STRIPS = 18 };
/* Brush.curve_preset, types/types_brush_enums.h:176-187 */
enum {
  DSC_CURVE_CUSTOM = 0, DSC_CURVE_SMOOTH = 1, DSC_CURVE_SPHERE = 2, DSC_CURVE_ROOT = 3, DSC_CURVE_SHARP = 4,
  DSC_CURVE_LIN = 5, DSC_CURVE_POW4 = 6, DSC_CURVE_INVSQUARE = 7, DSC_CURVE_CONSTANT = 8, DSC_CURVE_SMOOTHER = 9,
};
/* Brush.sculpt_plane, types/types_brush_enums.h:580-586 */
enum { DSC_DIR_AREA = 0, DSC_DIR_VIEW = 1, DSC_DIR_X = 2, DSC_DIR_Y = 3, DSC_DIR_Z = 4 };

enum {
  DSC_DAB_FRONTFACE = 1 << 0,  /* BRUSH_FRONTFACE, types_brush_enums.h:365 */
  DSC_DAB_PLANE_TRIM = 1 << 1, /* BRUSH_PLANE_TRIM, types_brush_enums.h:364 */
  DSC_DAB_FIRST_STEP = 1 << 2, /* first dab of the stroke */
  DSC_DAB_NO_NORMALS = 1 << 3, /* do not run the BKE_pbvh_update_normals stage for this dab */
  DSC_DAB_NO_BOUNDS = 1 << 4,  /* do not run the BKE_pbvh_update_bounds stage for this dab */
};

/* Mesh arrays, original vertex order.  Replaces the borrowed pointers BKE_pbvh_build_mesh stores
 * (pbvh.c:2467-2480: mpoly, mloop, verts, vert_normals, looptri, CD_PAINT_MASK pbvh.c:4894) and
 * the session tables of sculpt_update_object (kernel/intern/paint.c:1685-1688, pmap). */
typedef struct DscMeshDesc {
  int totvert;
  const float *co;   /* [totvert][3], MVert.co (types/types_meshdata.h:13-17) */
  const float *no;   /* [totvert][3] vert_normals, or NULL: computed on the device */
  const float *mask; /* [totvert] CD_PAINT_MASK, or NULL */
  int totpoly, totloop;
  const int *poly_loopstart; /* MPoly.loopstart */
  const int *poly_totloop;   /* MPoly.totloop */
  const int *loop_vert;      /* MLoop.v */
  int tottri;
  const int *tri_vert; /* [tottri][3] mloop[MLoopTri.tri[j]].v (pbvh.c:2951-2956) */
  const int *tri_poly; /* [tottri] MLoopTri.poly */
  /* vertex -> edge-neighbour CSR in the order the neighbour iterator lists them (smooth brush);
   * NULL if the smooth brush is not used */
  const int *nb_offsets; /* [totvert + 1] */
  const int *nb_indices;
  const unsigned char *boundary; /* [totvert] boundary-vertex flags, or NULL */
  /* [totvert] the 4 bytes that follow MVert.co (flag, bweight, pad: types_meshdata.h:13-17), kept so
   * dsc_download_mvert can hand back whole MVert records; NULL = zeros */
  const unsigned int *vert_tail;
} DscMeshDesc;

/* Multires grids instead of a mesh: a SubdivCCG (kernel/intern/subdiv_ccg.c) as flat tables.  Replaces
 * the arrays BKE_pbvh_build_grids borrows (pbvh.c:2516-2561: grids, gridkey, gridfaces) and the
 * adjacency SubdivCCG keeps (faces, adjacent_edges, adjacent_vertices, subdiv_ccg.c:397-530).  Element
 * index = grid * grid_size^2 + y * grid_size + x; CCGElem fields de-interleaved. */
typedef struct DscGridsDesc {
  int totgrid, grid_size;
  const float *co;   /* [totgrid * grid_size^2][3] CCG_elem_co */
  const float *no;   /* same shape, CCG_elem_no, or NULL: computed on the device */
  const float *mask; /* [totgrid * grid_size^2] CCG_elem_mask, or NULL */
  int totface;
  const int *face_start_grid; /* SubdivCCGFace.start_grid_index */
  const int *face_num_grids;  /* SubdivCCGFace.num_grids */
  int totedge;
  const int *edge_offsets;    /* [totedge + 1] in adjacent faces (SubdivCCGAdjacentEdge.num_adjacent_faces) */
  const int *edge_elems;      /* [edge_offsets[totedge]][2 * grid_size] boundary_coords as element indices */
  int totcvert;
  const int *cvert_offsets;   /* [totcvert + 1] (SubdivCCGAdjacentVertex.num_adjacent_faces) */
  const int *cvert_elems;     /* corner_coords as element indices */
  const int *grid_edge;       /* [totgrid] coarse edge of the grid's face corner (getFaceEdges order) */
  const int *grid_cvert;      /* [totgrid] coarse vertex of the grid's face corner (getFaceVertices order) */
  /* smooth brush on grids: the neighbours of the elements on the rim of every grid, as
   * KERNEL_subdiv_ccg_neighbor_coords_get (subdiv_ccg.c:1882-1909, no duplicates) lists them -- the neighbours of
   * interior elements are (x, y - 1), (x, y + 1), (x - 1, y), (x + 1, y) (subdiv_ccg.c:1870-1880) and need no table.
   * Rim index b of element (x, y), gs = grid_size: y == 0: x;  y == gs - 1: gs + x;  x == 0: 2 gs + y - 1;
   * x == gs - 1: 3 gs - 2 + y - 1  (4 gs - 4 per grid).  NULL: the smooth brush is refused. */
  int rim_width;               /* row length: the most neighbours any rim element has */
  const int *rim_neighbors;    /* [totgrid][4 * grid_size - 4][rim_width] element indices, -1 = none */
  const unsigned char *rim_boundary; /* [totgrid][4 * grid_size - 4] element is a boundary element of the coarse mesh
                                        (subdiv_ccg.c:1972-2008), or NULL: none is */
  /* [totgrid * grid_size^2] non-zero = the element is hidden (grid_hidden, pbvh.c:2527: a BLI_bitmap per grid) and
   * the vertex iterator skips it; NULL: none is */
  const unsigned char *hidden;
} DscGridsDesc;

/* The built PBVH, flattened.  Replaces PBVH.nodes / PBVHNode (kernel/intern/pbvh_intern.h:15-162).
 * For grids the prims are grids, uniq_verts = totprim * grid_size^2, face_verts = 0 and vert_offset /
 * vert_indices are not read (a grid leaf's elements are its grids' elements in order). */
typedef struct DscPbvhDesc {
  int totnode;
  const float *node_bb;      /* [totnode][6] PBVHNode.vb (bmin, bmax) */
  const float *node_orig_bb; /* [totnode][6] PBVHNode.orig_vb */
  const int *children_offset; /* [totnode], inner nodes */
  const int *flag;            /* [totnode] PBVHNodeFlags */
  const int *prim_offset;     /* [totnode] leaf: PBVHNode.prim_indices - PBVH.prim_indices */
  const int *totprim;         /* [totnode] */
  const int *prim_indices;    /* [tottri] PBVH.prim_indices */
  const int *uniq_verts;      /* [totnode] */
  const int *face_verts;      /* [totnode] */
  const int *vert_offset;     /* [totnode] leaf: start of its vert_indices in the array below */
  const int *vert_indices;    /* concatenated PBVHNode.vert_indices, unique verts first */
} DscPbvhDesc;

/* One dab.  Host fills it from Brush / StrokeCache; bstrength is the brush_strength() scalar. */
typedef struct DscDab {
  int tool;
  int curve_preset;
  int flags;
  int sculpt_plane;
  float location[3];
  float radius;
  float view_normal[3];
  float bstrength;
  float scale[3];
  float hardness;
  float normal_radius_factor;
  float plane_offset;
  float plane_trim;
  float tip_roundness;
  float grab_delta[3];
  float radius_scale;
  /* Brush.falloff_shape (types_brush_enums.h PAINT_FALLOFF_SHAPE_SPHERE / _TUBE): the tube tests and fades by the distance to
   * the view line through the location (SCULPT_brush_test_circle_sq, sculpt.c:1560-1574) */
  int falloff_shape;
  /* StrokeCache.flag CLIP_X/Y/Z (1, 2, 4) and Sculpt.flags SCULPT_LOCK_X/Y/Z (8, 16, 32), already shifted down: bits 0..2
   * clip the axis to 0 when the result is within clip_tolerance of it, bits 3..5 lock it (SCULPT_clip, sculpt.c:1484-1510) */
  int clip_flags;
  float clip_tolerance[3];
  /* grab: Brush.normal_weight (sculpt_project_v3_normal_align on the drag, sculpt.c:3015-3018); 0 = plain drag */
  float normal_weight;
} DscDab;

enum { DSC_FALLOFF_SPHERE = 0, DSC_FALLOFF_TUBE = 1 };
enum { DSC_CLIP_X = 1, DSC_CLIP_Y = 2, DSC_CLIP_Z = 4, DSC_LOCK_X = 8, DSC_LOCK_Y = 16, DSC_LOCK_Z = 32 };

/* Counters of the running stroke (device-side, read back on demand). */
typedef struct DscStrokeStats {
  int64_t vertex_dabs;  /* sum over dabs of uniq_verts of the gathered leaves */
  int64_t node_hits;    /* sum over dabs of the number of gathered leaves */
  int64_t moved_verts;  /* sum over dabs of vertices the brush displaced */
  int64_t dabs;
  int64_t kernel_launches; /* kernels launched by this context since stroke begin */
  int64_t area_verts;   /* sum over dabs of uniq_verts of the gathered leaves that reach the normal-sampling sphere */
  int64_t area_inside;  /* sum over dabs of the vertices inside it (what the area normal / centre averages) */
  /* the other terms of the algorithmic byte count, summed over dabs */
  int64_t all_verts;         /* unique + shared verts of the gathered leaves (A) */
  int64_t prims;             /* their looptris, or grids (T) */
  int64_t first_touch_verts; /* A of the leaves first touched in the stroke (undo snapshot) */
  int64_t refit_nodes;       /* inner nodes whose box the bottom-up refit rewrote */
} DscStrokeStats;

/* --- context ---------------------------------------------------------------------------- */
int dsc_ctx_create(int device, DscContext **r_ctx);
void dsc_ctx_destroy(DscContext *ctx);
const char *dsc_last_error(const DscContext *ctx); /* ctx may be NULL: last create error */
int dsc_abi_version(void);

/* --- session start: behind BKE_pbvh_build_mesh (pbvh.c:2452-2514) ------------------------- */
int dsc_mesh_upload(DscContext *ctx, const DscMeshDesc *mesh);
/* behind BKE_pbvh_build_grids (pbvh.c:2516-2561): instead of dsc_mesh_upload.  On grids the dab runs
 * gather -> brush -> stitch of duplicated boundary elements (multires.c:1171-1196) -> CCG normal update
 * of the gathered leaves' faces (subdiv_ccg.c:847-866) -> bounds; draw / inflate / grab / clay strips, and smooth
 * when the rim neighbour table is given. */
int dsc_grids_upload(DscContext *ctx, const DscGridsDesc *grids);
int dsc_pbvh_upload(DscContext *ctx, const DscPbvhDesc *pbvh);
/* vertex normals of the whole mesh from the current positions (all vertices dirty, all leaves
 * flagged): what BKE_pbvh_vert_coords_apply triggers (pbvh.c:4739-4747) */
int dsc_recalc_normals(DscContext *ctx);
int dsc_set_custom_curve(DscContext *ctx, const float *table257); /* colortools.c:942-965 LUT */
int dsc_set_mask(DscContext *ctx, const float *mask /* [totvert] or NULL */);
/* PBVHNode.flag bits a host pass owns (FullyHidden / FullyMasked, pbvh.c:3678-3710) */
int dsc_node_flag_set(DscContext *ctx, int node, int flag, int on);
/* the host-side marks made since the last push, in one batch and in stream order (no host synchronisation): for each of the
 * `count` listed nodes, set_bits are OR-ed into PBVHNode.flag and clear_bits removed.  What BKE_pbvh_node_mark_update,
 * BKE_pbvh_node_mark_redraw, BKE_pbvh_node_fully_hidden_set / _masked_set (pbvh.c:3620-3710) change on a host PBVH whose
 * truth lives on the device. */
int dsc_node_flags_apply(DscContext *ctx, int count, const int *nodes, const int *set_bits, const int *clear_bits);
/* BKE_pbvh_vert_mark_update (pbvh.c:3609-3613) marks made on the host: bitmap[totvert / 32 + 1], bit per vertex, OR-ed into the
 * device's vertex bitmap (the verts whose normals the next normals update recomputes) */
int dsc_vert_marks_or(DscContext *ctx, const unsigned int *bitmap);

/* --- stroke: behind the stroke operator's per-dab sequence --------------- */
int dsc_stroke_begin(DscContext *ctx, const float *automask /* [totvert] or NULL */);
/* gather -> undo snapshot + mark -> brush -> normals -> bounds, queued on the stream */
int dsc_dab(DscContext *ctx, const DscDab *dab);
/* the same for a run of dabs (a whole stroke segment); nothing is waited for */
int dsc_dabs(DscContext *ctx, const DscDab *dabs, int count);
/* BKE_pbvh_search_gather result of the last dab (pbvh.c:2736-2767): node indices in traversal
 * order.  r_nodes may be NULL to get the count only.  Synchronises. */
int dsc_gather_readback(DscContext *ctx, int *r_nodes, int capacity, int *r_tot);
/* stand-alone BKE_pbvh_search_gather with the sphere callback (no marking) */
int dsc_search_sphere(DscContext *ctx, const float center[3], float radius_sq, int original,
                      int ignore_fully_ineffective, int *r_nodes, int capacity, int *r_tot);
/* area normal / centre the last dab used (zero normal if the tool did not need one) */
int dsc_last_area(DscContext *ctx, float r_no[3], float r_co[3]);
/* vertices the last dab marked (BKE_pbvh_vert_mark_update, pbvh.c:3729), ascending vertex index;
 * only recorded while dsc_debug_capture(ctx, 1) is on */
int dsc_debug_capture(DscContext *ctx, int on);
int dsc_last_moved(DscContext *ctx, int *r_verts, int capacity, int *r_tot);
int dsc_stroke_stats(DscContext *ctx, DscStrokeStats *r_stats);
/* flushes PBVH_UpdateOriginalBB (pbvh.c:3298-3314) and closes the stroke */
int dsc_stroke_end(DscContext *ctx);

/* --- the PBVH update entry points on their own ------------------------------------------- */
int dsc_update_normals(DscContext *ctx);         /* BKE_pbvh_update_normals, pbvh.c:4559 */
int dsc_update_bounds(DscContext *ctx, int flag); /* BKE_pbvh_update_bounds, pbvh.c:3319 */
int dsc_node_mark_update(DscContext *ctx, int node); /* BKE_pbvh_node_mark_update, pbvh.c:3641 */

/* --- sync: behind BKE_pbvh_vert_coords_alloc/apply, get_verts, get_vert_normals
 *     (pbvh.c:4689-4749, 4961-4971) and the undo push (paint_hide.c:78) --------------------- */
int dsc_download_co(DscContext *ctx, float *r_co /* [totvert][3] */);
/* positions as complete 16-byte MVert records (BKE_pbvh_get_verts): one DMA, no host-side scatter */
int dsc_download_mvert(DscContext *ctx, void *r_mvert /* [totvert] MVert */);
/* page-lock / release a host array the download calls write into (cudaHostRegister); optional, it
 * lets the stroke-end downloads run at PCIe speed */
int dsc_host_register(DscContext *ctx, void *ptr, size_t bytes);
int dsc_host_unregister(DscContext *ctx, void *ptr);
int dsc_download_no(DscContext *ctx, float *r_no /* [totvert][3] */);
int dsc_download_mask(DscContext *ctx, float *r_mask /* [totvert]: the mask layer (grids average it when stitching) */);
/* grids: whole CCGElem records (co, then mask, then no -- subdiv_ccg.c:62-90) in element order, interleaved on the
 * device and copied by DMA straight into the CCG's storage; offsets in floats, -1 = the layer is absent */
int dsc_download_ccg(DscContext *ctx, void *r_elems, int elem_floats, int mask_offset_floats, int normal_offset_floats);
int dsc_download_orig_co(DscContext *ctx, float *r_co /* [totvert][3] */);
int dsc_download_orig_no(DscContext *ctx, float *r_no /* [totvert][3] */);
int dsc_download_node_bb(DscContext *ctx, float *r_bb /* [totnode][6] */, float *r_orig_bb /* or NULL */);
int dsc_download_node_flags(DscContext *ctx, int *r_flags /* [totnode] */);
/* undo-node membership of the running / last stroke: r_touched[node] = 1 */
int dsc_download_touched(DscContext *ctx, unsigned char *r_touched /* [totnode] */);
int dsc_upload_co(DscContext *ctx, const float *co /* [totvert][3] */); /* vert_coords_apply */
/* --- draw-buffer fill from the device: behind pbvh_update_draw_buffers (pbvh.c:3169-3285) ->
 *     GPU_pbvh_mesh_buffers_update (gpu/intern/gpu_buffers.c:174-305).  Every leaf flagged
 *     PBVH_UpdateDrawBuffers / PBVH_RebuildDrawBuffers gets its vertex buffer packed on the device in the
 *     format of gpu_pbvh_init (gpu_buffers.c:84-100): 36 bytes per looptri corner -- pos f32 x 3 @0,
 *     nor i16 x 3 @16, msk u8 @22, col u16 x 4 @24 (zero), fset u8 x 3 @32 (white); the flags are cleared
 *     (pbvh.c:3276).  The reference's caller copies the node's run into its GL buffer (CUDA-GL interop
 *     on the pointer dsc_draw_node_buffer returns) instead of re-reading the node's triangles on the CPU.
 *     dsc_draw_enable comes between dsc_mesh_upload and dsc_pbvh_upload.  Grids (gpu_pbvh_grid_buffers_update,
 *     gpu_buffers.c:548-725): grid_size^2 records per grid when smooth, 4 (grid_size - 1)^2 when flat -- the shading
 *     mode (one for all leaves, or per leaf after dsc_draw_leaf_shading) of a grids context is fixed by its first dsc_draw_update.  Only flagged leaves are refilled, like the
 *     reference (an unflagged leaf keeps its last records even when a neighbour's stitch moved its rim). ---------- */
int dsc_draw_enable(DscContext *ctx);

/* --- ray-cast: behind BKE_pbvh_raycast (pbvh.c:3896-3928) + BKE_pbvh_node_raycast (pbvh.c:4041-4100), the step
 *     that finds the dab location under the cursor.  Nearest intersection of the ray with
 *     the mesh: depth, the looptri's MLoopTri.poly, the corner nearest to the hit point, the triangle normal and
 *     the leaf.  `original`: stroke-start boxes and, for leaves with an undo node, stroke-start coordinates.
 *     `max_depth`: the depth the caller's search starts from (the ray's length to the far clip); only nearer hits count.
 *     dsc_raycast_enable comes between dsc_mesh_upload / dsc_grids_upload and dsc_pbvh_upload.  Synchronises.
 *     Grids (pbvh_grids_node_raycast, pbvh.c:4102-4200): face = the active grid, vertex = the nearest corner of the hit
 *     quad as an element index, the normal is the quad's; quads with a hidden corner are skipped. ---- */
typedef struct DscRayHit {
  int hit;
  float depth;
  int vertex; /* active vertex (original index) */
  int face;   /* MLoopTri.poly */
  int node;
  float face_normal[3];
} DscRayHit;
int dsc_raycast_enable(DscContext *ctx);
int dsc_raycast(DscContext *ctx, const float ray_start[3], const float ray_normal[3], int original, float max_depth,
                DscRayHit *r_hit);
/* The reference picks the shading per leaf: ME_SMOOTH of the poly of the leaf's first looptri (gpu_buffers.c:221-222) or of the
 * leaf's first grid (grid_flag_mats, gpu_buffers.c:574).  node_smooth[totnode] (leaves are read) hands those flags down, after
 * dsc_pbvh_upload and before the first dsc_draw_update; on grids it also fixes the per-leaf record layout. */
int dsc_draw_leaf_shading(DscContext *ctx, const unsigned char *node_smooth);
enum { DSC_DRAW_SHADING_PER_LEAF = -1, DSC_DRAW_SHADING_FLAT = 0, DSC_DRAW_SHADING_SMOOTH = 1 };
int dsc_draw_update(DscContext *ctx, int smooth /* DSC_DRAW_SHADING_*: every leaf flat / smooth, or per leaf */, int show_mask);
int dsc_draw_node_buffer(DscContext *ctx, int node, void **r_device_ptr, int *r_vert_len);
int dsc_draw_download(DscContext *ctx, int node, void *r_host, size_t capacity_bytes, int *r_vert_len);

/* Checkpoint / rollback of the resident mesh state (positions, normals, node boxes and flags) by
 * device-to-device copies: what operator cancel / the undo restore do on the host side of the
 * reference (undo nodes written back, paint_hide.c:78, then BKE_pbvh_update_bounds), without the
 * mesh crossing PCIe.  Not inside a stroke.  Queued on the stream like a dab. */
int dsc_state_save(DscContext *ctx);
int dsc_state_restore(DscContext *ctx);
int dsc_synchronize(DscContext *ctx);

/* --- multi-GPU: one process per GPU, the PBVH partitioned spatially (contiguous runs of leaves in
 *     traversal order = subtrees) across the ranks of one box.  Every rank holds the whole mesh and
 *     computes the leaves it owns; NCCL carries, per dab, one small all-reduce (area-normal sums +
 *     the bitmask of gathered leaves) and one exchange of the one-ring halo positions; at stroke
 *     end the owned vertex runs and leaf boxes are all-gathered (the BB-root reduction) so every
 *     replica is whole again.  Call order: dsc_ctx_create, dsc_dist_init, dsc_mesh_upload,
 *     dsc_pbvh_upload.  Not in the reference.
 *     Grids (dsc_grids_upload instead of dsc_mesh_upload): per dab two more exchanges -- the positions of the halo
 *     elements after the brush (and after every smoothing iteration), their normals after the CCG normal pass -- and
 *     every rank averages the groups of duplicated elements that hold one of its own (see dsc_dist_grids_plan). --- */
#define DSC_NCCL_ID_BYTES 128
int dsc_dist_unique_id(char id[DSC_NCCL_ID_BYTES]); /* rank 0 makes it, the host broadcasts it */
int dsc_dist_init(DscContext *ctx, int world, int rank, const char id[DSC_NCCL_ID_BYTES]);
/* the partition alone, no device needed: r_leaf_range[world + 1] bounds in traversal-order leaf
 * ranks, r_owner[totnode] owning rank of each leaf node (-1 for inner nodes) */
int dsc_dist_partition(const DscPbvhDesc *pbvh, int world, int *r_leaf_range, int *r_owner);
/* halo plan of one rank, no device needed: for every peer the vertices this rank sends (it owns
 * them, the peer's leaves reference them) and receives.  Arrays are malloc'd; free with
 * dsc_dist_free.  r_send_off / r_recv_off have world + 1 entries. */
int dsc_dist_halo_plan(const DscMeshDesc *mesh, const DscPbvhDesc *pbvh, int world, int rank, int **r_send_off,
                       int **r_send_vert, int **r_recv_off, int **r_recv_vert);
/* the same for partitioned grids, no device needed: r_grid_owner[totgrid]; what the rank computes after the brush --
 * r_face_dom[totface] (1: a grid of the face is owned, all its inner boundaries and its centre are averaged here; 2: only
 * the middle pairs on edges with an owned half), r_edge_mine[totedge] (bit h: half h of the coarse edge's points),
 * r_cvert_mine[totcvert] -- and per peer the elements it sends (owns) and receives (other ranks' elements those groups and
 * the smooth brush's neighbour lookups read), ascending element indices.  Arrays malloc'd; dsc_dist_free. */
int dsc_dist_grids_plan(const DscGridsDesc *grids, const DscPbvhDesc *pbvh, int world, int rank, int *r_grid_owner,
                        unsigned char *r_face_dom, unsigned char *r_edge_mine, unsigned char *r_cvert_mine, int **r_send_off,
                        int **r_send_elem, int **r_recv_off, int **r_recv_elem);
void dsc_dist_free(void *p);
/* How a partitioned stroke runs (peer-memory transport).  A dab is executed by, and exchanged among, only the ranks whose
 * region it can reach -- decided identically on every rank from the region boxes all ranks hold at stroke begin and the dabs
 * so far; the other ranks skip it and run ahead (on grids they replay the averaging of all coarse vertices every dab does).
 * Stroke end moves leaf boxes, flags and stroke state only; the vertex data stays with its owner:
 *   dsc_download_owned_mvert / _ccg  the runs this rank owns, scattered into the host arrays (the other entries untouched);
 *   dsc_dist_gather                  every replica whole again (what the whole-array download calls then return);
 *   dsc_dist_dab_counts              dabs this rank skipped / ran alone / exchanged since the context was made. */
int dsc_dist_gather(DscContext *ctx);
int dsc_dist_dab_counts(DscContext *ctx, long long r_counts[3]);
int dsc_download_owned_mvert(DscContext *ctx, void *r_mvert /* [totvert] MVert */, float *r_no /* [totvert][3] or NULL */);
int dsc_download_owned_ccg(DscContext *ctx, void *r_elems, int elem_floats, int mask_offset_floats, int normal_offset_floats);
/* leaf nodes this rank owns: r_range[2] = first and one-past-last leaf in traversal order */
int dsc_dist_owned_range(DscContext *ctx, int r_range[2]);
/* 1 when the per-dab exchanges run as stores into the peers' HBM over NVLink (cudaIpc-mapped inboxes, flag
 * handshakes), 0 when NCCL send / recv / all-reduce carries them (mapping refused, or DSC_NO_P2P set) */
int dsc_dist_uses_peer_memory(DscContext *ctx);
/* halo exchanges that returned at once because the dab gathered no leaf near a partition cut (peer-memory transport:
 * every rank reads that off the all-reduced bitmask of gathered leaves).  Synchronises. */
int dsc_dist_exchanges_skipped(DscContext *ctx, int *r_skipped);

/* --- timing helpers (CUDA events on the context's stream) --------------------------------- */
int dsc_timer_start(DscContext *ctx);
int dsc_timer_stop(DscContext *ctx, float *r_ms); /* synchronises */
void *dsc_stream(DscContext *ctx);                /* cudaStream_t */
/* per-stage device time of the dabs since the last reset; stage names via dsc_stage_name() */
#define DSC_NUM_STAGES 8
/* enable = 1: dabs are launched kernel by kernel with an event pair around each (a diagnostic path);
 * enable = 2: the dabs run as they do untimed -- replayed CUDA graphs -- and the event pairs are nodes of those graphs,
 * read after every replay (each pair also measures the launch latency of its kernel node, about 8 us);
 * enable = 3: the dabs run exactly as they do untimed and the durations are the hardware timestamps CUPTI records for
 * every kernel (libcupti is looked up at run time; DSC_ERR_UNSUPPORTED when it cannot trace) */
int dsc_stage_timing(DscContext *ctx, int enable);
int dsc_stage_times(DscContext *ctx, float r_ms[DSC_NUM_STAGES], int r_launches[DSC_NUM_STAGES]);
const char *dsc_stage_name(int stage);

#ifdef __cplusplus
}
#endif
#endif
