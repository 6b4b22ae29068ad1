/* Host side (C) of the B200 sculpt-stroke path: the reference's own PBVH entry points, kept by name
 * and argument meaning, backed by the device-resident mesh of libdune_sculpt_cuda.
 *
 * In the reference these functions live in source/dune/kernel/intern/pbvh.c (second copy, lines
 * 1913-4986) and their declarations in the absent BKE_pbvh.h; a maintainer keeps their pbvh.c and
 * adds the DUNE_pbvh_device_* hooks shown in INTEGRATION.md.  This stand-alone build of the same
 * interface is what the parity tests and bench.py drive.  Citations are relative to
 * /root/reference/source/dune/ .
 *
 * Ownership follows the reference: arrays handed back through out-params are MEM_mallocN'd here
 * and freed by the caller with MEM_freeN (pbvh.c:2750, paint_hide.c:371-373); a gather that finds
 * nothing returns NULL, 0 (pbvh.c:2760-2766).  Outside the reference tree MEM_* map to malloc/free.
 */
#ifndef DUNE_PBVH_H
#define DUNE_PBVH_H

#include <stdbool.h>
#include <stddef.h>
#include "dune_sculpt_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef MEM_mallocN
#  define DUNE_PBVH_OWN_MEM 1
void *MEM_mallocN(size_t len, const char *str);
void *MEM_callocN(size_t len, const char *str);
void MEM_freeN(void *vmemh);
#  define MEM_SAFE_FREE(v) \
    do { \
      if (v) { \
        MEM_freeN((void *)(v)); \
        (v) = NULL; \
      } \
    } while (0)
#endif

/* types/types_meshdata.h:13-17, 50-57, 69-74, 178-181 */
typedef struct MVert {
  float co[3];
  char flag, bweight;
  char _pad[2];
} MVert;
enum { ME_HIDE = (1 << 4) };   /* MVert.flag, types/types_meshdata.h:20-24 */
enum { ME_SMOOTH = (1 << 0) }; /* MPoly.flag / DMFlagMat.flag */
typedef struct MPoly {
  int loopstart;
  int totloop;
  short mat_nr;
  char flag, _pad;
} MPoly;
typedef struct MLoop {
  unsigned int v;
  unsigned int e;
} MLoop;
typedef struct MLoopTri {
  unsigned int tri[3];
  unsigned int poly;
} MLoopTri;

struct Mesh;       /* opaque here */
struct CustomData; /* opaque here */

/* ---- multires grids: the CCG types behind BKE_pbvh_build_grids (pbvh.c:2516-2561).  Their headers are
 * absent from the reference; field names are the ones kernel/intern/subdiv_ccg.c uses. ---- */
typedef struct CCGElem CCGElem; /* co[3], then mask, then no[3]: subdiv_ccg.c:62-90 */
typedef struct CCGKey {          /* subdiv_ccg.c:633-646 */
  int level;
  int elem_size;
  int grid_size, grid_area, grid_bytes;
  int normal_offset, mask_offset;
  int has_normals, has_mask;
} CCGKey;
typedef struct DMFlagMat {
  short mat_nr;
  char flag;
} DMFlagMat;
typedef unsigned int BLI_bitmap;
typedef struct SubdivCCGCoord { /* subdiv_ccg.c:368-372 */
  int grid_index;
  short x, y;
} SubdivCCGCoord;
typedef struct SubdivCCGFace { /* subdiv_ccg.c:134-140 */
  int num_grids;
  int start_grid_index;
} SubdivCCGFace;
typedef struct SubdivCCGAdjacentEdge { /* subdiv_ccg.c:383-397 */
  int num_adjacent_faces;
  SubdivCCGCoord **boundary_coords; /* [num_adjacent_faces][2 * grid_size] */
} SubdivCCGAdjacentEdge;
typedef struct SubdivCCGAdjacentVertex { /* subdiv_ccg.c:472-483 */
  int num_adjacent_faces;
  SubdivCCGCoord *corner_coords;
} SubdivCCGAdjacentVertex;
typedef struct SubdivCCG { /* subdiv_ccg.c:104-140 */
  int level;
  int grid_size;
  int grid_element_size;
  int num_grids;
  CCGElem **grids;
  unsigned char *grids_storage;
  bool has_normal, has_mask;
  int normal_offset, mask_offset;
  int num_faces;
  SubdivCCGFace *faces;
  SubdivCCGFace **grid_faces;
  int num_adjacent_edges;
  SubdivCCGAdjacentEdge *adjacent_edges;
  int num_adjacent_vertices;
  SubdivCCGAdjacentVertex *adjacent_vertices;
  /* in place of the OpenSubdiv topology refiner the reference asks for a face's edges and vertices
   * (subdiv_ccg.c:1198-1223): per grid, the coarse edge / vertex of its face corner */
  int *grid_edge, *grid_vertex;
  /* the refiner's getEdgeVertices / getNumVertexEdges / getVertexEdges, which the element-neighbour lookup asks
   * (subdiv_ccg.c:1558-1582, 1649-1651): NULL until DUNE_subdiv_ccg_topology_set */
  int (*edge_vertices)[2];
  int *vertex_edge_offsets, *vertex_edges;
} SubdivCCG;
typedef struct SubdivCCGNeighbors { /* subdiv_ccg.c:1365-1380: the header is absent; fields as the code uses them */
  SubdivCCGCoord *coords;
  int size;
  int num_duplicates;
  SubdivCCGCoord coords_fixed[256];
} SubdivCCGNeighbors;
typedef enum SubdivCCGAdjacencyType { /* subdiv_ccg.c:1972-2008 */
  SUBDIV_CCG_ADJACENT_NONE,
  SUBDIV_CCG_ADJACENT_VERTEX,
  SUBDIV_CCG_ADJACENT_EDGE,
} SubdivCCGAdjacencyType;

/* kernel/intern/pbvh_intern.h:4-7 */
typedef struct BB {
  float bmin[3], bmax[3];
} BB;

/* PBVHNodeFlags: values of the absent public header (names used at pbvh.c:3643-3726) */
typedef enum {
  PBVH_Leaf = 1 << 0,
  PBVH_UpdateNormals = 1 << 1,
  PBVH_UpdateBB = 1 << 2,
  PBVH_UpdateOriginalBB = 1 << 3,
  PBVH_UpdateDrawBuffers = 1 << 4,
  PBVH_UpdateRedraw = 1 << 5,
  PBVH_UpdateMask = 1 << 6,
  PBVH_UpdateVisibility = 1 << 8,
  PBVH_RebuildDrawBuffers = 1 << 9,
  PBVH_FullyHidden = 1 << 10,
  PBVH_FullyMasked = 1 << 11,
  PBVH_FullyUnmasked = 1 << 12,
  PBVH_UpdateColor = 1 << 14,
} PBVHNodeFlags;

/* kernel/intern/pbvh_intern.h:15-90, the fields on this path */
typedef struct PBVHNode {
  BB vb;
  BB orig_vb;
  int children_offset;
  int *prim_indices;
  unsigned int totprim;
  const int *vert_indices;
  unsigned int uniq_verts, face_verts;
  const int (*face_vert_indices)[3];
  unsigned int flag;
} PBVHNode;

/* kernel/intern/pbvh_intern.h:98-162, the fields on this path */
typedef struct PBVH {
  PBVHNode *nodes;
  int node_mem_count, totnode;
  int *prim_indices;
  int totprim;
  int totvert;
  int leaf_limit;
  float (*vert_normals)[3];
  MVert *verts;
  const MPoly *mpoly;
  const MLoop *mloop;
  const MLoopTri *looptri;
  int totpoly, totloop;
  float *vmask; /* CD_PAINT_MASK layer (pbvh.c:4894) or NULL */
  unsigned int *vert_bitmap;
  bool deformed;
  bool owns_normals;

  /* PBVH_GRIDS (pbvh.c:2527-2533) */
  int is_grids;
  CCGElem **grids;
  void **gridfaces;
  const DMFlagMat *grid_flag_mats;
  int totgrid;
  CCGKey gridkey;
  BLI_bitmap **grid_hidden;
  struct SubdivCCG *subdiv_ccg; /* set by DUNE_pbvh_device_attach_grids */
  int want_draw_buffers;

  /* device side */
  DscContext *device;
  bool device_dirty; /* the device holds newer positions / normals / boxes than the host arrays */
  bool in_stroke;
  bool normals_pinned, verts_pinned, grids_pinned; /* host arrays page-locked for the stroke-end DMA */
  /* session tables (kernel/intern/paint.c:1685-1688 pmap, boundary info) */
  int *nb_offsets, *nb_indices;
  unsigned char *boundary;
  /* partitioned over `dist_world` GPUs (one process each): stroke end brings back the runs this rank owns; the entries of
   * the host arrays that other ranks own stay as they were until DUNE_pbvh_device_gather makes the replica whole */
  int dist_world;
  bool gather_whole; /* inside DUNE_pbvh_device_gather */
  /* PBVHNode.flag as the device last saw it: what the BKE_pbvh_node_* setters changed since is pushed down in one batch before
   * the next device call (they take a node, not the PBVH, so the change is found by comparing) */
  unsigned int *synced_flag;
  bool host_vert_marks; /* BKE_pbvh_vert_mark_update was called since the last push */
} PBVH;

typedef bool (*BKE_pbvh_SearchCallback)(PBVHNode *node, void *data);

/* ---- mesh helpers ---- */
/* kernel/intern/mesh_tessellate.c:665 BKE_mesh_recalc_looptri (tri / quad rule :420-447) */
int BKE_mesh_poly_to_tri_count(int totpoly, int totloop);
void BKE_mesh_recalc_looptri(const MLoop *mloop, const MPoly *mpoly, const MVert *mvert, int totloop, int totpoly,
                             MLoopTri *mlooptri);

/* ---- build / free: pbvh.c:2563-2620, 2452-2514 ---- */
PBVH *BKE_pbvh_new(void);
void BKE_pbvh_build_mesh(PBVH *pbvh, struct Mesh *mesh, const MPoly *mpoly, const MLoop *mloop, MVert *verts,
                         int totvert, struct CustomData *vdata, struct CustomData *ldata, struct CustomData *pdata,
                         const MLoopTri *looptri, int looptri_num);
void BKE_pbvh_build_grids(PBVH *pbvh, CCGElem **grids, int totgrid, CCGKey *key, void **gridfaces, DMFlagMat *flagmats,
                          BLI_bitmap **grid_hidden);
void BKE_pbvh_free(PBVH *pbvh);
/* visible quads of the listed grids (pbvh.c:2249-2279); grid_hidden may be NULL, and so may any of its entries */
int BKE_pbvh_count_grid_quads(BLI_bitmap **grid_hidden, const int *grid_indices, int totgrid, int gridsize);
/* pbvh.c:3770-3800 */
void BKE_pbvh_node_get_grids(PBVH *pbvh, PBVHNode *node, int **r_grid_indices, int *r_totgrid, int *r_maxgrid,
                             int *r_gridsize, CCGElem ***r_griddata);
/* subdiv_ccg.c:633-651 */
void BKE_subdiv_ccg_key_top_level(CCGKey *key, const SubdivCCG *subdiv_ccg);
/* not in the reference: a SubdivCCG from flat tables (what BKE_subdiv_to_ccg builds through OpenSubdiv,
 * subdiv_ccg.c:104-140, 397-530); element index = grid * grid_size^2 + y * grid_size + x.  Copies everything. */
SubdivCCG *DUNE_subdiv_ccg_from_tables(int level, int num_grids, const float *co, const float *no, const float *mask,
                                       int num_faces, const int *face_start_grid, const int *face_num_grids, int num_edges,
                                       const int *edge_offsets, const int *edge_elems, int num_vertices,
                                       const int *vert_offsets, const int *vert_elems, const int *grid_edge,
                                       const int *grid_vertex);
void DUNE_subdiv_ccg_free(SubdivCCG *subdiv_ccg);
/* not in the reference: the coarse topology the neighbour lookup needs (edge -> its two vertices, vertex -> its
 * edges in the refiner's order); copies.  Without it the smooth brush is refused on grids. */
void DUNE_subdiv_ccg_topology_set(SubdivCCG *subdiv_ccg, const int *edge_vertices, const int *vertex_edge_offsets,
                                  const int *vertex_edges);
/* subdiv_ccg.c:1882-1909: the neighbours of a grid element; coords is coords_fixed or MEM-allocated when larger
 * (free with MEM_freeN when coords != coords_fixed, as the reference's callers do) */
void BKE_subdiv_ccg_neighbor_coords_get(const SubdivCCG *subdiv_ccg, const SubdivCCGCoord *coord, const bool include_duplicates,
                                        SubdivCCGNeighbors *r_neighbors);
/* subdiv_ccg.c:1972-2008, with the base mesh's loop vertices taken from grid_vertex (mloop[grid_index].v) */
SubdivCCGAdjacencyType BKE_subdiv_ccg_coarse_mesh_adjacency_info_get(const SubdivCCG *subdiv_ccg, const SubdivCCGCoord *coord,
                                                                     int *r_v1, int *r_v2);
/* not in the reference: sizes and layers the reference pulls out of Mesh / CustomData */
void DUNE_pbvh_mesh_sizes_set(PBVH *pbvh, int totpoly, int totloop);
void DUNE_pbvh_mask_layer_set(PBVH *pbvh, float *vmask);
void DUNE_pbvh_vert_normals_set(PBVH *pbvh, float (*vert_normals)[3]);
void DUNE_pbvh_leaf_limit_set(PBVH *pbvh, int leaf_limit);

/* types/types_meshdata.h:274-297: the per-loop displacement grid and paint-mask grid of a multires mesh */
typedef struct MDisps {
  int totdisp;
  int level;
  float (*disps)[3];
  unsigned int *hidden;
} MDisps;
typedef struct GridPaintMask {
  float *data;
  unsigned int level;
  char _pad[4];
} GridPaintMask;
/* multires_reshape_assign_final_coords_from_ccg (kernel/intern/multires_reshape_ccg.c:10-70), the first step of
 * multires_flush_sculpt_updates -> multiresModifier_reshapeFromCCG (kernel/intern/multires.c:401-428): every grid's
 * element coordinates into mdisps[grid].disps[y * grid_size + x] and, when both sides have one, its mask into
 * grid_paint_masks[grid].data[...].  With a device attached the elements come straight from the device (one download,
 * one memcpy per grid -- the CCG storage is not touched); without one, from the CCG as in the reference.  `mdisps` /
 * `grid_paint_masks` may be NULL (multires_reshape_util.c:427-435).  Top level only (reshape level == CCG level). */
bool DUNE_multires_reshape_assign_final_coords(PBVH *pbvh, struct SubdivCCG *subdiv_ccg, MDisps *mdisps,
                                               GridPaintMask *grid_paint_masks);

/* ---- device hooks (new) ---- */
int DUNE_pbvh_device_attach(PBVH *pbvh, int device);
/* before the attach: also keep the tables the device-side draw-buffer fill needs (dsc_draw_enable) */
void DUNE_pbvh_draw_buffers_enable(PBVH *pbvh);
/* pbvh_update_draw_buffers (pbvh.c:3169-3285) on the device: pack the vertex buffers of the leaves flagged
 * PBVH_UpdateDrawBuffers / PBVH_RebuildDrawBuffers (gpu_buffers.c:174-305) and clear the flags; then the
 * device pointer and vertex count of a node's buffer for the GL copy */
/* shading: DSC_DRAW_SHADING_PER_LEAF (-1) takes it per leaf from the material flags the PBVH was built with (ME_SMOOTH of the
 * leaf's first poly / grid, what the reference does); 0 / 1 force flat / smooth on every leaf */
int DUNE_pbvh_update_draw_buffers(PBVH *pbvh, int shading, bool show_mask);
int DUNE_pbvh_node_draw_buffer(PBVH *pbvh, PBVHNode *node, void **r_device_ptr, int *r_vert_len);
/* before the attach: keep the tables the device ray-cast needs (dsc_raycast_enable) */
void DUNE_pbvh_raycast_enable(PBVH *pbvh);
/* BKE_pbvh_raycast (pbvh.c:3915-3928) + BKE_pbvh_node_raycast (pbvh.c:4203-4260) in one call: the nearest hit of
 * the ray with the mesh -- its depth, the active vertex and face (MLoopTri.poly), the triangle normal and the
 * leaf -- as the stroke operator's hit callback accumulates them, starting from `max_depth` (the ray's length).
 * False when nothing nearer than that is hit. */
bool DUNE_pbvh_raycast_nearest(PBVH *pbvh, const float ray_start[3], const float ray_normal[3], bool original,
                               float max_depth, float *r_depth, int *r_active_vertex_index, int *r_active_face_index,
                               float r_face_normal[3], PBVHNode **r_node);
/* the same for a grids PBVH: the CCG's elements and adjacency go to the device */
int DUNE_pbvh_device_attach_grids(PBVH *pbvh, SubdivCCG *subdiv_ccg, int device);
/* one rank of a grids PBVH partitioned across the GPUs of one box (multires meshes too large or too slow for one) */
int DUNE_pbvh_device_attach_grids_dist(PBVH *pbvh, SubdivCCG *subdiv_ccg, int device, int world, int rank, const char *nccl_id);
/* one rank of a PBVH partitioned across the GPUs of one box (see dsc_dist_init) */
int DUNE_pbvh_device_attach_dist(PBVH *pbvh, int device, int world, int rank, const char *nccl_id);
/* partitioned: the owners send their vertex data to every rank and the whole mesh comes back to this rank's host arrays
 * (collective: every rank calls it) */
int DUNE_pbvh_device_gather(PBVH *pbvh);
void DUNE_pbvh_device_detach(PBVH *pbvh);
/* bring host arrays (verts, normals, node boxes, flags) up to date with the device */
int DUNE_pbvh_device_sync_to_host(PBVH *pbvh);
/* checkpoint / rollback of the device-resident mesh (operator cancel, undo restore): device-to-device */
int DUNE_pbvh_device_checkpoint(PBVH *pbvh);
int DUNE_pbvh_device_rollback(PBVH *pbvh);
const char *DUNE_pbvh_device_error(const PBVH *pbvh);

/* ---- traversal: pbvh.c:2736-2767 ---- */
void BKE_pbvh_search_gather(PBVH *pbvh, BKE_pbvh_SearchCallback scb, void *search_data, PBVHNode ***r_array,
                            int *r_tot);

/* The sphere search callback of the stroke operator (absent from the reference).
 * When passed to BKE_pbvh_search_gather on a device-attached PBVH the test runs on the device. */
typedef struct SculptSearchSphereData {
  const float *center;
  float radius_squared;
  bool original;
  bool ignore_fully_ineffective;
} SculptSearchSphereData;
bool SCULPT_search_sphere_cb(PBVHNode *node, void *data_v);

/* ---- node / vertex access: pbvh.c:3641-3840 ---- */
void BKE_pbvh_node_mark_update(PBVHNode *node);
void BKE_pbvh_vert_mark_update(PBVH *pbvh, int index);
void BKE_pbvh_node_fully_hidden_set(PBVHNode *node, int fully_hidden);
bool BKE_pbvh_node_fully_hidden_get(PBVHNode *node);
void BKE_pbvh_node_fully_masked_set(PBVHNode *node, int fully_masked);
bool BKE_pbvh_node_fully_masked_get(PBVHNode *node);
void BKE_pbvh_node_get_verts(PBVH *pbvh, PBVHNode *node, const int **r_vert_indices, MVert **r_verts);
void BKE_pbvh_node_num_verts(PBVH *pbvh, PBVHNode *node, int *r_uniquevert, int *r_totvert);
void BKE_pbvh_node_get_BB(PBVHNode *node, float bb_min[3], float bb_max[3]);
void BKE_pbvh_node_get_original_BB(PBVHNode *node, float bb_min[3], float bb_max[3]);

/* ---- updates: pbvh.c:4559-4587, 3319-3339 ---- */
void BKE_pbvh_update_normals(PBVH *pbvh, struct SubdivCCG *subdiv_ccg);
void BKE_pbvh_update_bounds(PBVH *pbvh, int flag);

/* ---- sync: pbvh.c:4689-4749, 4961-4971 ---- */
float (*BKE_pbvh_vert_coords_alloc(PBVH *pbvh))[3];
void BKE_pbvh_vert_coords_apply(PBVH *pbvh, const float (*vertCos)[3], int totvert);
MVert *BKE_pbvh_get_verts(const PBVH *pbvh);
const float (*BKE_pbvh_get_vert_normals(const PBVH *pbvh))[3];

/* ---- stroke driver (the missing sculpt.c side) ---- */
/* brush_strength(): SURVEY.md row a14.  alpha is Brush.alpha (types/types_brush.h:194), dir_in is
 * BRUSH_DIR_IN (types_brush_enums.h:347), invert the Ctrl modifier */
float DUNE_sculpt_brush_strength(int sculpt_tool, float alpha, float pressure, bool dir_in, bool invert,
                                 float overlap, float feather);
/* Brush defaults (types/types_brush_defaults.h:9-91) into a dab descriptor */
void DUNE_sculpt_dab_defaults(DscDab *dab, int sculpt_tool);
/* the symmetry passes of one dab (Paint.symmetry_flags & PAINT_SYMM_AXIS_ALL, types_scene.h): r_dabs[0] is the dab
 * itself, the others its mirror images in the valid axis combinations; returns how many */
int DUNE_sculpt_dab_symmetry(const DscDab *dab, int symm, DscDab r_dabs[8]);
int DUNE_sculpt_stroke_begin(PBVH *pbvh, const float *automask);
int DUNE_sculpt_dab(PBVH *pbvh, const DscDab *dab);
int DUNE_sculpt_stroke_end(PBVH *pbvh);

/* automasking factors computed at stroke start:
 * boundary-edge propagation (automasking_boundary_edges_propagation_steps, types_brush.h:288) and
 * topology flood fill from a seed vertex inside `radius` of `location` */
void DUNE_sculpt_automask_boundary_edges(const PBVH *pbvh, int propagation_steps, float *r_factor);
void DUNE_sculpt_automask_topology(const PBVH *pbvh, int seed_vert, const float location[3], float radius,
                                   float *r_factor);

#ifdef __cplusplus
}
#endif
#endif
