/* oracle/_ref shim -- TEST INFRASTRUCTURE ONLY.
 *
 * Stands in for the headers the reference's hot-path sources include but /root/reference does not ship
 * (MEM_guardedalloc.h, BLI_bitmap.h, BLI_ghash.h, BLI_task.h, atomic_ops.h, DNA_mesh*_types.h, BKE_ccg.h, BKE_pbvh.h,
 * BKE_subdiv_ccg.h), so that the functions oracle/ref_extract.py cuts verbatim out of
 * pbvh.c / subdiv_ccg.c / mesh_evaluate.c / math_*.c compile unchanged.  Nothing here computes: types, accessor
 * macros, a malloc wrapper, a serial task loop.  Where a definition is restated from the upstream project the tree
 * was forked from (the reference only holds its uses) it is marked DAGGER, like SURVEY.md's dagger rows.
 */
#ifndef DUNE_ORACLE_REF_SHIM_H
#define DUNE_ORACLE_REF_SHIM_H

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned int uint;
#define MINLINE static inline
#define UNUSED(x) UNUSED_##x __attribute__((__unused__))
#define UNLIKELY(x) __builtin_expect(!!(x), 0)
#define LIKELY(x) __builtin_expect(!!(x), 1)
#define SWAP(type, a, b) \
  { \
    type sw_ap; \
    sw_ap = (a); \
    (a) = (b); \
    (b) = sw_ap; \
  } \
  (void)0
#define POINTER_FROM_INT(i) ((void *)(intptr_t)(i))
#define POINTER_AS_INT(i) ((void)0, ((int)(intptr_t)(i)))
#define BLI_assert(a) ((void)0)
#define lib_assert(a) ((void)0)
#define LI_assert(a) ((void)0)

/* ---- MEM_guardedalloc.h: size-prefixed malloc so that MEM_recallocN can zero the tail ---- */
void *ref_mem_alloc(size_t size, int zero);
void *ref_mem_realloc(void *p, size_t size, int zero);
void ref_mem_free(void *p);
#define MEM_mallocN(size, str) ref_mem_alloc((size), 0)
#define MEM_callocN(size, str) ref_mem_alloc((size), 1)
#define MEM_malloc_arrayN(n, size, str) ref_mem_alloc((size_t)(n) * (size_t)(size), 0)
#define MEM_calloc_arrayN(n, size, str) ref_mem_alloc((size_t)(n) * (size_t)(size), 1)
#define MEM_freeN(p) ref_mem_free((void *)(p))
#define MEM_SAFE_FREE(v) \
  do { \
    if (v) { \
      ref_mem_free((void *)(v)); \
      (v) = NULL; \
    } \
  } while (0)
#define MEM_recallocN(p, size) ref_mem_realloc((p), (size), 1)
#define MEM_reallocN(p, size) ref_mem_realloc((p), (size), 0)
#define MEM_recallocN_id(p, size, id) ref_mem_realloc((p), (size), 1)

/* ---- BLI_bitmap.h DAGGER (32-bit blocks; uses: pbvh.c:2157-2165, 2925-2927, paint.c:1237-1240) ---- */
typedef unsigned int BLI_bitmap;
typedef BLI_bitmap LIB_bitmap; /* pbvh_intern.h spells it this way */
#define _BITMAP_NUM_BLOCKS(_num) (((_num) >> 5) + 1)
#define BLI_BITMAP_SIZE(_num) ((size_t)(_BITMAP_NUM_BLOCKS(_num)) * sizeof(BLI_bitmap))
#define BLI_BITMAP_NEW(_num, _alloc_string) ((BLI_bitmap *)MEM_callocN(BLI_BITMAP_SIZE(_num), _alloc_string))
#define BLI_BITMAP_TEST(_bitmap, _index) ((_bitmap)[(_index) >> 5] & (1u << ((_index)&31)))
#define BLI_BITMAP_ENABLE(_bitmap, _index) ((_bitmap)[(_index) >> 5] |= (1u << ((_index)&31)))
#define BLI_BITMAP_DISABLE(_bitmap, _index) ((_bitmap)[(_index) >> 5] &= ~(1u << ((_index)&31)))
#define BLI_BITMAP_SET(_bitmap, _index, _set) \
  { \
    if (_set) { \
      BLI_BITMAP_ENABLE(_bitmap, _index); \
    } \
    else { \
      BLI_BITMAP_DISABLE(_bitmap, _index); \
    } \
  } \
  (void)0
static inline void BLI_bitmap_set_all(BLI_bitmap *bitmap, bool set, size_t bits)
{
  memset(bitmap, set ? 0xff : 0, BLI_BITMAP_SIZE(bits));
}

/* ---- BLI_ghash.h: an int -> pointer map with insertion-ordered iteration (ref_api.c).  build_mesh_leaf_node writes
 * vert_indices[value] = key for every entry, so the iteration order does not reach the result. ---- */
typedef struct GHash GHash;
typedef struct GSet GSet;
typedef struct GHashIterator {
  GHash *gh;
  int i;
} GHashIterator;
typedef struct GSetIterator {
  int unused;
} GSetIterator;
GHash *BLI_ghash_int_new_ex(const char *info, unsigned int reserve);
bool BLI_ghash_ensure_p(GHash *gh, void *key, void ***r_val);
void BLI_ghash_free(GHash *gh, void *keyfree, void *valfree);
void *BLI_ghashIterator_getKey(GHashIterator *ghi);
void *BLI_ghashIterator_getValue(GHashIterator *ghi);
bool ref_ghash_iter_done(GHashIterator *ghi);
#define GHASH_ITER(gh_iter_, ghash_) for ((gh_iter_).gh = (ghash_), (gh_iter_).i = 0; !ref_ghash_iter_done(&(gh_iter_)); (gh_iter_).i++)
#define BLI_gset_len(gs) 0
#define BLI_gsetIterator_init(gsi, gs) ((void)0)

/* ---- BLI_task.h: the callbacks run serially, in index order ---- */
typedef struct TaskParallelTLS {
  void *userdata_chunk;
} TaskParallelTLS;
typedef void (*TaskParallelRangeFunc)(void *__restrict userdata, const int iter, const TaskParallelTLS *__restrict tls);
typedef struct TaskParallelSettings {
  bool use_threading;
  void *userdata_chunk;
  size_t userdata_chunk_size;
  void (*func_free)(const void *__restrict userdata, void *__restrict chunk);
  int min_iter_per_thread;
} TaskParallelSettings;
static inline void BLI_task_parallel_range(const int start, const int stop, void *userdata, TaskParallelRangeFunc func,
                                           const TaskParallelSettings *settings)
{
  TaskParallelTLS tls = {settings ? settings->userdata_chunk : NULL};
  for (int i = start; i < stop; i++) {
    func(userdata, i, &tls);
  }
  if (settings && settings->func_free) {
    settings->func_free(userdata, settings->userdata_chunk);
  }
}
/* atomic_ops.h: atomic_add_and_fetch_fl is a compare-and-swap loop around `old + x` */
static inline float atomic_add_and_fetch_fl(float *p, const float x)
{
  *p = *p + x;
  return *p;
}

/* ---- DNA: types/types_meshdata.h:13-17 (MVert), 47-56 (MPoly), MLoop, MLoopTri; ME_SMOOTH; DMFlagMat ---- */
typedef struct MVert {
  float co[3];
  char flag, bweight;
  char _pad[2];
} MVert;
enum { ME_HIDE = (1 << 4) };
typedef struct MPoly {
  int loopstart;
  int totloop;
  short mat_nr;
  char flag, _pad;
} MPoly;
enum { ME_SMOOTH = (1 << 0) };
typedef struct MLoop {
  unsigned int v;
  unsigned int e;
} MLoop;
typedef struct MLoopTri {
  unsigned int tri[3];
  unsigned int poly;
} MLoopTri;
typedef struct DMFlagMat {
  short mat_nr;
  char flag;
} DMFlagMat;
/* the two layers the path asks CustomData for (pbvh.c:4894-4895) */
typedef struct CustomData {
  float *paint_mask;
  void *prop_color;
} CustomData;
enum { CD_PAINT_MASK = 34, CD_PROP_COLOR = 47 };
static inline void *CustomData_get_layer(const CustomData *data, int type)
{
  if (!data) return NULL;
  return type == CD_PAINT_MASK ? (void *)data->paint_mask : (type == CD_PROP_COLOR ? data->prop_color : NULL);
}
#define CustomData_get_offset(data, type) (-1)
typedef struct Mesh {
  int face_sets_color_seed, face_sets_color_default;
  float (*vert_normals)[3];
} Mesh;
#define BKE_mesh_vertex_normals_ensure(mesh) ((void)0)
#define BKE_mesh_vertex_normals_for_write(mesh) ((mesh)->vert_normals)
typedef struct BMesh {
  CustomData vdata;
} BMesh;
struct BMLog;
struct BMVert;
struct IsectRayPrecalc;
struct GPU_PBVH_Buffers;

/* ---- BKE_ccg.h DAGGER (uses: pbvh.c:2541-2554, subdiv_ccg.c:62-122, 684-738) ---- */
typedef struct CCGElem CCGElem;
typedef struct CCGKey {
  int level;
  int elem_size; /* bytes per element: co, then mask, then no (subdiv_ccg.c:62-90) */
  int grid_size;
  int grid_area;
  int grid_bytes;
  int normal_offset;
  int mask_offset;
  int has_normals;
  int has_mask;
} CCGKey;
static inline float *CCG_elem_co(const CCGKey *key, CCGElem *elem)
{
  (void)key;
  return (float *)elem;
}
static inline float *CCG_elem_no(const CCGKey *key, CCGElem *elem) { return (float *)((char *)elem + key->normal_offset); }
static inline float *CCG_elem_mask(const CCGKey *key, CCGElem *elem) { return (float *)((char *)elem + key->mask_offset); }
static inline CCGElem *CCG_elem_offset(const CCGKey *key, CCGElem *elem, int offset)
{
  return (CCGElem *)(((char *)elem) + key->elem_size * offset);
}
static inline CCGElem *CCG_grid_elem(const CCGKey *key, CCGElem *elem, int x, int y)
{
  return CCG_elem_offset(key, elem, (y * key->grid_size + x));
}
static inline float *CCG_grid_elem_co(const CCGKey *key, CCGElem *elem, int x, int y) { return CCG_elem_co(key, CCG_grid_elem(key, elem, x, y)); }
static inline float *CCG_grid_elem_no(const CCGKey *key, CCGElem *elem, int x, int y) { return CCG_elem_no(key, CCG_grid_elem(key, elem, x, y)); }
static inline float *CCG_elem_offset_co(const CCGKey *key, CCGElem *elem, int offset) { return CCG_elem_co(key, CCG_elem_offset(key, elem, offset)); }
static inline CCGElem *CCG_elem_next(const CCGKey *key, CCGElem *elem) { return CCG_elem_offset(key, elem, 1); }
/* BKE_subdiv_ccg.h DAGGER: the fields the extracted normal / averaging functions read (subdiv_ccg.c:374-379, 676-678,
 * 890-946, 951-1104); boundary_coords[face][2 * grid_size], corner_coords[face] */
typedef struct SubdivCCGCoord {
  int grid_index;
  short x, y;
} SubdivCCGCoord;
typedef struct SubdivCCGFace {
  int num_grids;
  int start_grid_index;
} SubdivCCGFace;
typedef struct SubdivCCGAdjacentEdge {
  int num_adjacent_faces;
  SubdivCCGCoord **boundary_coords;
} SubdivCCGAdjacentEdge;
typedef struct SubdivCCGAdjacentVertex {
  int num_adjacent_faces;
  SubdivCCGCoord *corner_coords;
} SubdivCCGAdjacentVertex;
typedef struct SubdivCCG {
  int grid_size;
  int num_grids;
  CCGElem **grids;
  bool has_normal, has_mask;
  int num_faces;
  SubdivCCGFace *faces;
  int num_adjacent_edges;
  SubdivCCGAdjacentEdge *adjacent_edges;
  int num_adjacent_vertices;
  SubdivCCGAdjacentVertex *adjacent_vertices;
} SubdivCCG;

/* ---- BKE_pbvh.h DAGGER: enum values as SURVEY.md section 8a row a8 lists them ---- */
typedef struct PBVH PBVH;
typedef struct PBVHNode PBVHNode;
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
  PBVH_UpdateTopology = 1 << 13,
  PBVH_UpdateColor = 1 << 14,
} PBVHNodeFlags;
typedef enum { PBVH_FACES, PBVH_GRIDS, PBVH_BMESH } PBVHType;
typedef struct PBVHProxyNode {
  float (*co)[3];
} PBVHProxyNode;
typedef struct PBVHColorBufferNode {
  float (*color)[4];
} PBVHColorBufferNode;
typedef bool (*BKE_pbvh_SearchCallback)(PBVHNode *node, void *data);
#define PBVH_ITER_ALL 0
#define PBVH_ITER_UNIQUE 1
/* pbvh.c's LEAF_LIMIT is the literal 10000; a variable here so that the pin tests can build trees of small meshes */
extern int ref_leaf_limit;
#define LEAF_LIMIT ref_leaf_limit

#include "pbvh_intern.h" /* the reference's own, copied next to the generated unit by ref_extract.py */

void BKE_pbvh_node_mark_rebuild_draw(PBVHNode *node);
void BKE_pbvh_node_fully_hidden_set(PBVHNode *node, int fully_hidden);
void BKE_pbvh_node_get_verts(PBVH *pbvh, PBVHNode *node, const int **r_vert_indices, MVert **r_verts);
void BKE_pbvh_node_num_verts(PBVH *pbvh, PBVHNode *node, int *r_uniquevert, int *r_totvert);
void BKE_pbvh_node_get_grids(PBVH *pbvh, PBVHNode *node, int **r_grid_indices, int *r_totgrid, int *r_maxgrid, int *r_gridsize,
                             CCGElem ***r_griddata);
void BKE_pbvh_parallel_range_settings(TaskParallelSettings *settings, bool use_threading, int totnode);
void BKE_pbvh_search_gather(PBVH *pbvh, BKE_pbvh_SearchCallback scb, void *search_data, PBVHNode ***r_array, int *r_tot);
void BKE_pbvh_build_mesh(PBVH *pbvh, Mesh *mesh, const MPoly *mpoly, const MLoop *mloop, MVert *verts, int totvert,
                         struct CustomData *vdata, struct CustomData *ldata, struct CustomData *pdata, const MLoopTri *looptri,
                         int looptri_num);
void BKE_pbvh_build_grids(PBVH *pbvh, CCGElem **grids, int totgrid, CCGKey *key, void **gridfaces, DMFlagMat *flagmats,
                          BLI_bitmap **grid_hidden);
PBVH *BKE_pbvh_new(void);
void BKE_pbvh_update_bounds(PBVH *pbvh, int flag);
void BKE_pbvh_node_mark_update(PBVHNode *node);
void BKE_pbvh_vert_mark_update(PBVH *pbvh, int index);
void BKE_pbvh_node_get_BB(PBVHNode *node, float bb_min[3], float bb_max[3]);
void BKE_pbvh_node_get_original_BB(PBVHNode *node, float bb_min[3], float bb_max[3]);
void BKE_mesh_calc_poly_normal(const MPoly *mpoly, const MLoop *loopstart, const MVert *mvarray, float r_no[3]);
float normal_tri_v3(float n[3], const float v1[3], const float v2[3], const float v3[3]);
float normal_quad_v3(float n[3], const float v1[3], const float v2[3], const float v3[3], const float v4[3]);
bool paint_is_face_hidden(const MLoopTri *lt, const MVert *mvert, const MLoop *mloop);
bool paint_is_grid_face_hidden(const uint *grid_hidden, int gridsize, int x, int y);

/* PBVHVertexIter and its macro DAGGER (the public header is absent; pbvh_vertex_iter_init, pbvh.c:4840-4897, fills these
 * fields and update_node_vb, pbvh.c:2026-2046, iterates with the macro).  PBVH_ITER_UNIQUE skips hidden vertices / grid
 * elements. */
typedef struct PBVHVertexIter {
  int g, width, height, gx, gy, i, index;
  bool respect_hide;
  CCGKey key;
  CCGElem **grids;
  CCGElem *grid;
  BLI_bitmap **grid_hidden, *gh;
  int *grid_indices;
  int totgrid, gridsize;
  MVert *mverts;
  float (*vert_normals)[3];
  int totvert;
  const int *vert_indices;
  float *vmask;
  void *vcol;
  GSetIterator bm_unique_verts, bm_other_verts;
  CustomData *bm_vdata;
  int cd_vert_mask_offset;
  MVert *mvert;
  struct BMVert *bm_vert;
  float *co, *no, *fno, *mask, *col;
  bool visible;
} PBVHVertexIter;
void pbvh_vertex_iter_init(PBVH *pbvh, PBVHNode *node, PBVHVertexIter *vi, int mode);
#define BKE_pbvh_vertex_iter_begin(pbvh, node, vi, mode) \
  pbvh_vertex_iter_init(pbvh, node, &vi, mode); \
  for (vi.i = 0, vi.g = 0; vi.g < vi.totgrid; vi.g++) { \
    if (vi.grids) { \
      vi.width = vi.gridsize; \
      vi.height = vi.gridsize; \
      vi.index = vi.grid_indices[vi.g] * vi.key.grid_area - 1; \
      vi.grid = vi.grids[vi.grid_indices[vi.g]]; \
      if (mode == PBVH_ITER_UNIQUE) { \
        vi.gh = vi.grid_hidden[vi.grid_indices[vi.g]]; \
      } \
    } \
    else { \
      vi.width = vi.totvert; \
      vi.height = 1; \
    } \
    for (vi.gy = 0; vi.gy < vi.height; vi.gy++) { \
      for (vi.gx = 0; vi.gx < vi.width; vi.gx++, vi.i++) { \
        if (vi.grid) { \
          vi.co = CCG_elem_co(&vi.key, vi.grid); \
          vi.fno = CCG_elem_no(&vi.key, vi.grid); \
          vi.mask = vi.key.has_mask ? CCG_elem_mask(&vi.key, vi.grid) : NULL; \
          vi.grid = CCG_elem_next(&vi.key, vi.grid); \
          vi.index++; \
          vi.visible = true; \
          if (vi.gh) { \
            if (BLI_BITMAP_TEST(vi.gh, vi.gy * vi.gridsize + vi.gx)) { \
              continue; \
            } \
          } \
        } \
        else if (vi.mverts) { \
          vi.mvert = &vi.mverts[vi.vert_indices[vi.gx]]; \
          if (vi.respect_hide) { \
            vi.visible = !(vi.mvert->flag & ME_HIDE); \
            if (mode == PBVH_ITER_UNIQUE && !vi.visible) { \
              continue; \
            } \
          } \
          else { \
            BLI_assert(vi.visible); \
          } \
          vi.co = vi.mvert->co; \
          vi.no = vi.vert_normals[vi.vert_indices[vi.gx]]; \
          vi.index = vi.vert_indices[vi.i]; \
          if (vi.vmask) { \
            vi.mask = &vi.vmask[vi.index]; \
          } \
        }
#define BKE_pbvh_vertex_iter_end \
  } \
  } \
  } \
  ((void)0)

#endif
