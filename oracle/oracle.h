/* ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
 *
 * CPU restatement (plain C11) of the reference's sculpt-stroke hot path, used only as the checker
 * by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing
 * in the product path (dune_sculpt_b200/, include/) may include, link or call this.
 *
 * "Parity unpinned": the reference (/root/reference) ships no test, golden vector or fixture for
 * PBVH / sculpt / normals, cannot be compiled here (pbvh.c is two concatenated
 * copies, its public header is absent) and has no source for the brush layer.
 * The PBVH parts below follow source/dune/kernel/intern/pbvh.c line by line in behaviour (each
 * function cites the lines); the brush parts (marked DAGGER) restate the upstream project's
 * published behaviour and are the specification by decision.
 *
 * All paths cited are relative to /root/reference/source/dune/ .  "pbvh.c" is
 * kernel/intern/pbvh.c, second (complete) copy, lines 1913-4986.
 */
#ifndef DUNE_ORACLE_H
#define DUNE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* PBVHNodeFlags: names used at pbvh.c:3643-3726; numeric values from the (absent) public header,
 * upstream values, SURVEY.md section 8a row a8. */
enum {
  OR_PBVH_Leaf = 1 << 0,
  OR_PBVH_UpdateNormals = 1 << 1,
  OR_PBVH_UpdateBB = 1 << 2,
  OR_PBVH_UpdateOriginalBB = 1 << 3,
  OR_PBVH_UpdateDrawBuffers = 1 << 4,
  OR_PBVH_UpdateRedraw = 1 << 5,
  OR_PBVH_UpdateMask = 1 << 6,
  OR_PBVH_UpdateVisibility = 1 << 8,
  OR_PBVH_RebuildDrawBuffers = 1 << 9,
  OR_PBVH_FullyHidden = 1 << 10,
  OR_PBVH_FullyMasked = 1 << 11,
  OR_PBVH_FullyUnmasked = 1 << 12,
  OR_PBVH_UpdateColor = 1 << 14,
};

/* eBrushSculptTool, types/types_brush_enums.h:410-443 */
enum { OR_TOOL_DRAW = 1, OR_TOOL_SMOOTH = 2, OR_TOOL_INFLATE = 4, OR_TOOL_GRAB = 5, OR_TOOL_CLAY_STRIPS = 18 };
/* eBrushCurvePreset, types/types_brush_enums.h:176-187 */
enum {
  OR_CURVE_CUSTOM = 0, OR_CURVE_SMOOTH = 1, OR_CURVE_SPHERE = 2, OR_CURVE_ROOT = 3, OR_CURVE_SHARP = 4,
  OR_CURVE_LIN = 5, OR_CURVE_POW4 = 6, OR_CURVE_INVSQUARE = 7, OR_CURVE_CONSTANT = 8, OR_CURVE_SMOOTHER = 9,
};
/* sculpt_plane, types/types_brush_enums.h:580-586 */
enum { OR_DIR_AREA = 0, OR_DIR_VIEW = 1, OR_DIR_X = 2, OR_DIR_Y = 3, OR_DIR_Z = 4 };

enum {
  OR_DAB_FRONTFACE = 1 << 0,  /* BRUSH_FRONTFACE, types_brush_enums.h:365 */
  OR_DAB_PLANE_TRIM = 1 << 1, /* BRUSH_PLANE_TRIM, types_brush_enums.h:364 */
  OR_DAB_FIRST_STEP = 1 << 2, /* first dab of the stroke (clay strips skips it) */
  OR_DAB_NO_NORMALS = 1 << 3, /* skip BKE_pbvh_update_normals after the dab */
  OR_DAB_NO_BOUNDS = 1 << 4,  /* skip BKE_pbvh_update_bounds after the dab */
};

typedef struct OrBB {
  float bmin[3], bmax[3];
} OrBB;

/* One dab.  Field meaning follows SURVEY.md section 8a rows a11-a20. */
typedef struct OrDab {
  int tool;
  int curve_preset;
  int flags;
  int sculpt_plane;
  float location[3];
  float radius;
  float view_normal[3];
  float bstrength; /* signed, brush_strength() result, row a14 */
  float scale[3];
  float hardness;
  float normal_radius_factor;
  float plane_offset;
  float plane_trim;
  float tip_roundness;
  float grab_delta[3];
  float radius_scale; /* gather radius multiplier, 1.0 normally */
  int falloff_shape;  /* 0 sphere, 1 tube: distance to the view line through the location (types_brush_enums.h:600-603) */
  int clip_flags;     /* bits 0-2: mirror clipping on the axis (CLIP_X << i); bits 3-5: axis locked (SCULPT_LOCK_X << i) */
  float clip_tolerance[3];
  float normal_weight; /* grab: blend of the drag towards the sculpt normal (types_brush.h:158) */
} OrDab;
/* DAGGER distance from the line loc + t n to a box, squared: the node test of the tube falloff (the view-line counterpart
 * of the sphere callback, row a7).  0 when the line crosses the box, else the least distance to one of its 12 edges. */
float or_line_aabb_distsq(const float loc[3], const float n[3], const float bmin[3], const float bmax[3]);
struct OrPbvh;
int or_gather_tube(struct OrPbvh *p, const float center[3], const float normal[3], float radius_sq, int original,
                   int ignore_fully_ineffective, int *r_nodes);

typedef struct OrPbvh OrPbvh;

/* ---- mesh helpers (oracle_mesh.c) ---- */
int or_looptri_count(int totpoly, const int *poly_len);
void or_looptri_calc(int totpoly, const int *poly_start, const int *poly_len, const int *loop_v,
                     const float (*co)[3], int (*r_tri_loop)[3], int *r_tri_poly);
void or_vert_poly_map(int totvert, int totpoly, const int *poly_start, const int *poly_len,
                      const int *loop_v, int *r_off /* V+1 */, int *r_idx /* totloop */);
/* returns number of neighbour entries written; r_idx must hold 2*totloop ints */
int or_vert_neighbors(int totvert, int totpoly, const int *poly_start, const int *poly_len,
                      const int *loop_v, int *r_off /* V+1 */, int *r_idx, unsigned char *r_boundary /* V */);

/* ---- PBVH (oracle_pbvh.c) ---- */
OrPbvh *or_pbvh_build_mesh(int totvert, const float (*co)[3], const float (*no)[3], const float *mask,
                           int totpoly, const int *poly_start, const int *poly_len, int totloop,
                           const int *loop_v, int leaf_limit /* 0 = LEAF_LIMIT */);
/* optional inputs of the NEXT or_pbvh_build_mesh / _grids call (copied by it): MPoly.mat_nr / .flag, MVert.flag,
 * DMFlagMat.mat_nr / .flag, grid_hidden per element -- material split, fully hidden leaves, hidden vertices */
void or_pbvh_next_build_attrs(const short *poly_mat, const unsigned char *poly_flag, const unsigned char *vert_flag,
                              const short *grid_mat, const unsigned char *grid_flag, const unsigned char *grid_hidden);
/* BKE_pbvh_build_grids (pbvh.c:2516-2561) over a SubdivCCG given as flat tables: elements
 * [totgrid * grid_size^2] (index = grid * gs^2 + y * gs + x), faces (start grid, grid count), adjacent
 * edges (per edge and adjacent face 2 * gs element indices, subdiv_ccg.c:397-463), adjacent vertices
 * (corner elements, subdiv_ccg.c:483-530), and per grid the coarse edge / vertex of its face corner
 * (the order subdiv_ccg.c:1198-1223 walks them) */
OrPbvh *or_pbvh_build_grids(int totgrid, int grid_size, const float (*co)[3], const float (*no)[3], const float *mask,
                            int totface, const int *face_start, const int *face_num, int totedge, const int *edge_off,
                            const int *edge_elems, int totcvert, const int *cvert_off, const int *cvert_elems,
                            const int *grid_edge, const int *grid_cvert, int leaf_limit /* 0 = LEAF_LIMIT / gs^2 */);
/* the two topology-refiner queries the element-neighbour lookup makes beyond the tables above:
 * edge_verts[totedge][2] (getEdgeVertices), cvert_edges CSR (getVertexEdges); needed by the smooth brush */
void or_grids_set_topology(OrPbvh *p, const int *edge_verts, const int *cvert_edge_off, const int *cvert_edges);
/* KERNEL_subdiv_ccg_neighbor_coords_get (subdiv_ccg.c:1882-1909), include_duplicates = false: neighbours of an
 * element in the reference's order; r holds or_grids_max_neighbors() ints; returns the count */
int or_grids_neighbors(const OrPbvh *p, int elem, int *r);
int or_grids_max_neighbors(const OrPbvh *p);
int or_grids_is_boundary(const OrPbvh *p, int elem); /* DAGGER SCULPT_vertex_is_boundary, subdiv_ccg.c:1949-2008 */
void or_grids_average_all(OrPbvh *p);  /* KERNEL_subdiv_ccg_average_grids, subdiv_ccg.c:1170-1189 */
/* threads > 1 only: accumulate vertex normals per vertex in the serial loop's order instead of with float atomics, so
 * that the threaded run is bit-identical to the single-threaded restatement (bench.py's full-size parity legs) */
void or_set_ordered_normals(int on);
void or_grids_recalc_normals(OrPbvh *p); /* KERNEL_subdiv_ccg_recalc_normals, subdiv_ccg.c:782-790 */
void or_grids_inner_normals(OrPbvh *p);  /* its first half alone: subdiv_ccg.c:670-740 on every grid (pin tests) */
float *or_pbvh_mask(OrPbvh *p);
void or_pbvh_free(OrPbvh *p);
int or_pbvh_totnode(const OrPbvh *p);
int or_pbvh_tottri(const OrPbvh *p);
int or_pbvh_totvert(const OrPbvh *p);
/* flat export: arrays sized totnode */
void or_pbvh_nodes(const OrPbvh *p, OrBB *vb, OrBB *orig_vb, int *children_offset, int *flag,
                   int *prim_offset, int *totprim, int *uniq_verts, int *face_verts);
const int *or_pbvh_prim_indices(const OrPbvh *p);
const int *or_pbvh_node_vert_indices(const OrPbvh *p, int node);
const int *or_pbvh_node_face_vert_indices(const OrPbvh *p, int node);
const int *or_pbvh_tri_verts(const OrPbvh *p); /* [T][3] vertex indices */
const int *or_pbvh_tri_poly(const OrPbvh *p);
float *or_pbvh_co(OrPbvh *p);
float *or_pbvh_no(OrPbvh *p);
float *or_pbvh_orig_co(OrPbvh *p);
float *or_pbvh_orig_no(OrPbvh *p);
void or_pbvh_node_set_flag(OrPbvh *p, int node, int flag, int on);

/* BKE_pbvh_search_gather with the sphere callback; returns count, node indices in r_nodes */
int or_gather_sphere(OrPbvh *p, const float center[3], float radius_sq, int original,
                     int ignore_fully_ineffective, int *r_nodes);
/* BKE_pbvh_search_gather(update_search_cb, flag) */
int or_gather_flag(OrPbvh *p, int flag, int *r_nodes);
void or_vert_mark_update(OrPbvh *p, int v);
void or_node_mark_update(OrPbvh *p, int node);
void or_update_normals(OrPbvh *p);
void or_update_bounds(OrPbvh *p, int flag);
/* BKE_pbvh_raycast (pbvh.c:3896-3928) with the stroke operator's hit callback (DAGGER sculpt_raycast_cb ->
 * BKE_pbvh_node_raycast -> pbvh_faces_node_raycast, pbvh.c:4041-4100): nearest hit of the ray with the mesh.
 * Returns 1 on a hit; r_face = MLoopTri.poly of the hit, r_vertex = its corner nearest to the hit point.  Grids
 * (pbvh_grids_node_raycast, pbvh.c:4102-4200): r_face = the grid of the hit quad, r_vertex = its nearest element. */
int or_raycast(OrPbvh *p, const float ray_start[3], const float ray_normal[3], int original, float max_depth, float *r_depth,
               int *r_vertex, int *r_face, float r_face_normal[3], int *r_node);
/* GPU_pbvh_mesh_buffers_update (gpu/intern/gpu_buffers.c:174-305) of one leaf into `out` (totprim * 3 records of
 * 36 bytes); returns the vertex count.  Grid leaves (gpu_pbvh_grid_buffers_update, gpu_buffers.c:548-725): smooth -- one
 * record per element, totprim * grid_size^2; flat -- four per quad, totprim * (grid_size - 1)^2 * 4 */
int or_draw_buffers_update(OrPbvh *p, int node, int smooth, int show_mask, unsigned char *out);
/* full-mesh vertex normals the way the accumulate pass would produce them with every vertex dirty */
void or_recalc_all_normals(OrPbvh *p);
void or_set_threads(int n);

/* ---- sculpt session (oracle_sculpt.c) ---- */
void or_stroke_begin(OrPbvh *p, const float *automask /* V or NULL */);
void or_stroke_end(OrPbvh *p);
void or_set_custom_curve(OrPbvh *p, const float *table257);
int or_dab(OrPbvh *p, const OrDab *d);
/* results of the last dab */
int or_last_hits(const OrPbvh *p, int *r_nodes);         /* gathered leaves, gather order */
int or_last_moved(const OrPbvh *p, int *r_verts);        /* vertices marked update, iteration order */
void or_last_area(const OrPbvh *p, float r_no[3], float r_co[3]);
int or_touched_nodes(const OrPbvh *p, int *r_nodes);     /* undo-node membership, ascending node index */
int64_t or_stroke_vertex_dabs(const OrPbvh *p);          /* sum of uniq_verts over hit leaves */
float or_brush_curve_strength(const OrPbvh *p, int preset, float final_len, float radius);

#ifdef __cplusplus
}
#endif
#endif
