/* ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * PBVH build / traversal / normals / bounds restated from kernel/intern/pbvh.c. */
#include "oracle_intern.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#  include <omp.h>
#endif

#define LEAF_LIMIT 10000 /* pbvh.c:1950 */

int or_threads = 1;
void or_set_threads(int n)
{
  or_threads = n < 1 ? 1 : n;
#ifdef _OPENMP
  omp_set_num_threads(or_threads);
#endif
}

static inline float min_ff(float a, float b) { return (a < b) ? a : b; }
static inline float max_ff(float a, float b) { return (a > b) ? a : b; }

/* pbvh.c:1973-1993 */
static void BB_reset(OrBB *bb)
{
  bb->bmin[0] = bb->bmin[1] = bb->bmin[2] = FLT_MAX;
  bb->bmax[0] = bb->bmax[1] = bb->bmax[2] = -FLT_MAX;
}
static void BB_expand(OrBB *bb, const float co[3])
{
  for (int i = 0; i < 3; i++) {
    bb->bmin[i] = min_ff(bb->bmin[i], co[i]);
    bb->bmax[i] = max_ff(bb->bmax[i], co[i]);
  }
}
static void BB_expand_with_bb(OrBB *bb, const OrBB *bb2)
{
  for (int i = 0; i < 3; i++) {
    bb->bmin[i] = min_ff(bb->bmin[i], bb2->bmin[i]);
    bb->bmax[i] = max_ff(bb->bmax[i], bb2->bmax[i]);
  }
}
/* pbvh.c:1995-2016 */
static int BB_widest_axis(const OrBB *bb)
{
  float dim[3];
  for (int i = 0; i < 3; i++) {
    dim[i] = bb->bmax[i] - bb->bmin[i];
  }
  if (dim[0] > dim[1]) {
    return (dim[0] > dim[2]) ? 0 : 2;
  }
  return (dim[1] > dim[2]) ? 1 : 2;
}

/* pbvh.c:2026-2046 update_node_vb (leaf: PBVH_ITER_ALL over unique + shared verts) */
static void update_node_vb(OrPbvh *p, OrNode *node)
{
  OrBB vb;
  BB_reset(&vb);
  if (node->flag & OR_PBVH_Leaf) {
    const int tot = node->uniq_verts + node->face_verts;
    for (int i = 0; i < tot; i++) {
      BB_expand(&vb, p->co[node->vert_indices[i]]);
    }
  }
  else {
    BB_expand_with_bb(&vb, &p->nodes[node->children_offset].vb);
    BB_expand_with_bb(&vb, &p->nodes[node->children_offset + 1].vb);
  }
  node->vb = vb;
}

/* Inputs BKE_pbvh_build_mesh / _grids read besides the geometry: MPoly.mat_nr / .flag & ME_SMOOTH (material split,
 * pbvh.c:2058-2066, 2091-2132, 2329-2359), MVert.flag & ME_HIDE (fully hidden leaves, pbvh.c:2192-2208;
 * paint.c:1227-1232), DMFlagMat and grid_hidden for grids (pbvh.c:2249-2307; paint.c:1234-1241).  Given for the NEXT
 * build through or_pbvh_next_build_attrs (all optional); the build copies them. */
static struct {
  const short *poly_mat, *grid_mat;
  const unsigned char *poly_flag, *vert_flag, *grid_flag, *grid_hidden;
} g_next_attrs;
void or_pbvh_next_build_attrs(const short *poly_mat, const unsigned char *poly_flag, const unsigned char *vert_flag,
                              const short *grid_mat, const unsigned char *grid_flag, const unsigned char *grid_hidden)
{
  g_next_attrs.poly_mat = poly_mat; g_next_attrs.poly_flag = poly_flag; g_next_attrs.vert_flag = vert_flag;
  g_next_attrs.grid_mat = grid_mat; g_next_attrs.grid_flag = grid_flag; g_next_attrs.grid_hidden = grid_hidden;
}
static void *dup_bytes(const void *src, size_t n)
{
  if (!src) return NULL;
  void *d = malloc(n ? n : 1);
  memcpy(d, src, n);
  return d;
}
#define OR_ME_SMOOTH 1
#define OR_ME_HIDE 16

/* pbvh.c:2058-2066 face_materials_match / grid_materials_match over the prim's material record */
static int prim_materials_match(const OrPbvh *p, int prim_a, int prim_b)
{
  int a = prim_a, b = prim_b;
  const short *mat = p->grid_mat;
  const unsigned char *flag = p->grid_flag;
  if (!p->is_grids) {
    a = p->tri_poly[prim_a];
    b = p->tri_poly[prim_b];
    mat = p->poly_mat;
    flag = p->poly_flag;
  }
  const int sa = flag ? (flag[a] & OR_ME_SMOOTH) : 0, sb = flag ? (flag[b] & OR_ME_SMOOTH) : 0;
  const int ma = mat ? mat[a] : 0, mb = mat ? mat[b] : 0;
  return sa == sb && ma == mb;
}

/* pbvh.c:2329-2359 */
static int leaf_needs_material_split(const OrPbvh *p, int offset, int count)
{
  if (count <= 1) return 0;
  if (p->is_grids ? !(p->grid_mat || p->grid_flag) : !(p->poly_mat || p->poly_flag)) return 0;
  const int first = p->prim_indices[offset];
  for (int i = offset + count - 1; i > offset; i--) {
    if (!prim_materials_match(p, first, p->prim_indices[i])) return 1;
  }
  return 0;
}

/* pbvh.c:2091-2132 */
static int partition_indices_material(OrPbvh *p, int lo, int hi)
{
  const int *indices = p->prim_indices;
  const int first = p->prim_indices[lo];
  int i = lo, j = hi;
  for (;;) {
    for (; prim_materials_match(p, first, indices[i]); i++) {
    }
    for (; !prim_materials_match(p, first, indices[j]); j--) {
    }
    if (!(i < j)) {
      return i;
    }
    int t = p->prim_indices[i];
    p->prim_indices[i] = p->prim_indices[j];
    p->prim_indices[j] = t;
    i++;
  }
}

/* pbvh.c:2070-2088 */
static int partition_indices(int *prim_indices, int lo, int hi, int axis, float mid, const OrBBC *prim_bbc)
{
  int i = lo, j = hi;
  for (;;) {
    for (; prim_bbc[prim_indices[i]].bcentroid[axis] < mid; i++) {
    }
    for (; mid < prim_bbc[prim_indices[j]].bcentroid[axis]; j--) {
    }
    if (!(i < j)) {
      return i;
    }
    int t = prim_indices[i];
    prim_indices[i] = prim_indices[j];
    prim_indices[j] = t;
    i++;
  }
}

/* pbvh.c:2134-2145 */
static void pbvh_grow_nodes(OrPbvh *p, int totnode)
{
  if (totnode > p->node_mem_count) {
    int old = p->node_mem_count;
    p->node_mem_count = p->node_mem_count + (p->node_mem_count / 3);
    if (p->node_mem_count < totnode) {
      p->node_mem_count = totnode;
    }
    p->nodes = realloc(p->nodes, sizeof(OrNode) * (size_t)p->node_mem_count);
    memset(p->nodes + old, 0, sizeof(OrNode) * (size_t)(p->node_mem_count - old));
  }
  p->totnode = totnode;
}

/* pbvh.c:2149-2171 map_insert_vert.  The GHash is replaced by a stamp array with the same
 * key -> value semantics (value >= 0: unique slot, value < 0: ~shared slot). */
typedef struct LeafMap {
  int *stamp, *value;
} LeafMap;

static int map_insert_vert(OrPbvh *p, LeafMap *map, int node_index, int *face_verts, int *uniq_verts, int vertex)
{
  if (map->stamp[vertex] != node_index) {
    int value_i;
    map->stamp[vertex] = node_index;
    if (p->vert_bitmap[vertex] == 0) {
      p->vert_bitmap[vertex] = 1;
      value_i = *uniq_verts;
      (*uniq_verts)++;
    }
    else {
      value_i = ~(*face_verts);
      (*face_verts)++;
    }
    map->value[vertex] = value_i;
    return value_i;
  }
  return map->value[vertex];
}

/* pbvh.c:2174-2238 build_mesh_leaf_node */
static void build_mesh_leaf_node(OrPbvh *p, LeafMap *map, int node_index)
{
  OrNode *node = &p->nodes[node_index];
  node->uniq_verts = node->face_verts = 0;
  const int totface = node->totprim;
  const int *prims = p->prim_indices + node->prim_offset;

  int(*fvi)[3] = malloc(sizeof(int[3]) * (size_t)totface);
  node->face_vert_indices = fvi;
  for (int i = 0; i < totface; i++) {
    const int *vt = p->tri_v[prims[i]];
    for (int j = 0; j < 3; j++) {
      fvi[i][j] = map_insert_vert(p, map, node_index, &node->face_verts, &node->uniq_verts, vt[j]);
    }
  }
  int *vert_indices = calloc((size_t)(node->uniq_verts + node->face_verts), sizeof(int));
  node->vert_indices = vert_indices;
  /* Build the vertex list, unique verts first (pbvh.c:2210-2231): every key lands at the slot its
   * value names, so the iteration order of the hash does not matter. */
  for (int i = 0; i < totface; i++) {
    const int *vt = p->tri_v[prims[i]];
    for (int j = 0; j < 3; j++) {
      int ndx = fvi[i][j];
      if (ndx < 0) {
        ndx = -ndx + node->uniq_verts - 1;
        fvi[i][j] = ndx;
      }
      vert_indices[ndx] = vt[j];
    }
  }
  node->flag |= OR_PBVH_RebuildDrawBuffers | OR_PBVH_UpdateDrawBuffers | OR_PBVH_UpdateRedraw; /* pbvh.c:3663 */
  /* pbvh.c:2188-2208, 2235: respect_hide is set by BKE_pbvh_new (pbvh.c:2566); a looptri is hidden when any of its
   * corners is (paint.c:1227-1232) */
  int has_visible = 0;
  for (int i = 0; i < totface && !has_visible; i++) {
    const int *vt = p->tri_v[prims[i]];
    has_visible = !(p->vert_flag && ((p->vert_flag[vt[0]] | p->vert_flag[vt[1]] | p->vert_flag[vt[2]]) & OR_ME_HIDE));
  }
  if (!has_visible) node->flag |= OR_PBVH_FullyHidden;
}

/* pbvh.c:2249-2279 BKE_pbvh_count_grid_quads; paint.c:1234-1241 */
static int count_grid_quads(const OrPbvh *p, const int *grid_indices, int totgrid)
{
  const int gs = p->grid_size, gridarea = (gs - 1) * (gs - 1);
  int totquad = 0;
  for (int i = 0; i < totgrid; i++) {
    const unsigned char *gh = p->grid_hidden ? p->grid_hidden + (size_t)grid_indices[i] * (size_t)(gs * gs) : NULL;
    int any = 0;
    if (gh) {
      for (int e = 0; e < gs * gs && !any; e++) any = gh[e];
    }
    if (!any) { /* no bitmap for this grid */
      totquad += gridarea;
      continue;
    }
    for (int y = 0; y < gs - 1; y++) {
      for (int x = 0; x < gs - 1; x++) {
        if (!(gh[y * gs + x] || gh[y * gs + x + 1] || gh[(y + 1) * gs + x + 1] || gh[(y + 1) * gs + x])) totquad++;
      }
    }
  }
  return totquad;
}

/* pbvh.c:2240-2247 */
static void update_vb(OrPbvh *p, OrNode *node, const OrBBC *prim_bbc, int offset, int count)
{
  BB_reset(&node->vb);
  for (int i = offset + count - 1; i >= offset; i--) {
    BB_expand_with_bb(&node->vb, (const OrBB *)(&prim_bbc[p->prim_indices[i]]));
  }
  node->orig_vb = node->vb;
}

/* pbvh.c:2309-2325 */
static void build_leaf(OrPbvh *p, LeafMap *map, int node_index, const OrBBC *prim_bbc, int offset, int count)
{
  p->nodes[node_index].flag |= OR_PBVH_Leaf;
  p->nodes[node_index].prim_offset = offset;
  p->nodes[node_index].totprim = count;
  update_vb(p, &p->nodes[node_index], prim_bbc, offset, count);
  if (p->is_grids) {
    /* build_grid_leaf_node (pbvh.c:2240-2247 neighbourhood): the node's "verts" are the elements of its
     * grids in the order the vertex iterator walks them (pbvh.c:4840-4897): grid by grid, y, x */
    OrNode *node = &p->nodes[node_index];
    const int gs2 = p->grid_size * p->grid_size;
    node->uniq_verts = count * gs2;
    node->face_verts = 0;
    node->vert_indices = malloc(sizeof(int) * (size_t)(node->uniq_verts > 0 ? node->uniq_verts : 1));
    for (int i = 0; i < count; i++) {
      const int g = p->prim_indices[offset + i];
      for (int j = 0; j < gs2; j++) node->vert_indices[i * gs2 + j] = g * gs2 + j;
    }
    /* build_grid_leaf_node, pbvh.c:2301-2307 */
    if (count_grid_quads(p, p->prim_indices + offset, count) == 0) node->flag |= OR_PBVH_FullyHidden;
    node->flag |= OR_PBVH_RebuildDrawBuffers | OR_PBVH_UpdateDrawBuffers | OR_PBVH_UpdateRedraw; /* pbvh.c:3663 */
    return;
  }
  build_mesh_leaf_node(p, map, node_index);
}

/* pbvh.c:2372-2425 build_sub */
static void build_sub(OrPbvh *p, LeafMap *map, int node_index, OrBB *cb, const OrBBC *prim_bbc, int offset, int count)
{
  int end;
  OrBB cb_backing;

  const int below_leaf_limit = count <= p->leaf_limit;
  if (below_leaf_limit) {
    if (!leaf_needs_material_split(p, offset, count)) {
      build_leaf(p, map, node_index, prim_bbc, offset, count);
      return;
    }
  }

  p->nodes[node_index].children_offset = p->totnode;
  pbvh_grow_nodes(p, p->totnode + 2);

  update_vb(p, &p->nodes[node_index], prim_bbc, offset, count);

  if (!below_leaf_limit) {
    if (!cb) {
      cb = &cb_backing;
      BB_reset(cb);
      for (int i = offset + count - 1; i >= offset; i--) {
        BB_expand(cb, prim_bbc[p->prim_indices[i]].bcentroid);
      }
    }
    const int axis = BB_widest_axis(cb);
    end = partition_indices(p->prim_indices, offset, offset + count - 1, axis,
                            (cb->bmax[axis] + cb->bmin[axis]) * 0.5f, prim_bbc);
  }
  else {
    end = partition_indices_material(p, offset, offset + count - 1); /* pbvh.c:2411-2414 */
  }

  build_sub(p, map, p->nodes[node_index].children_offset, NULL, prim_bbc, offset, end - offset);
  build_sub(p, map, p->nodes[node_index].children_offset + 1, NULL, prim_bbc, end, offset + count - end);
}

/* pbvh.c:2452-2514 BKE_pbvh_build_mesh (+ pbvh_build 2427-2450) */
OrPbvh *or_pbvh_build_mesh(int totvert, const float (*co)[3], const float (*no)[3], const float *mask,
                           int totpoly, const int *poly_start, const int *poly_len, int totloop,
                           const int *loop_v, int leaf_limit)
{
  OrPbvh *p = calloc(1, sizeof(OrPbvh));
  p->totvert = totvert;
  p->totpoly = totpoly;
  p->totloop = totloop;
  p->leaf_limit = leaf_limit > 0 ? leaf_limit : LEAF_LIMIT;
  p->co = malloc(sizeof(float[3]) * (size_t)totvert);
  memcpy(p->co, co, sizeof(float[3]) * (size_t)totvert);
  p->no = calloc((size_t)totvert, sizeof(float[3]));
  if (no) {
    memcpy(p->no, no, sizeof(float[3]) * (size_t)totvert);
  }
  if (mask) {
    p->mask = malloc(sizeof(float) * (size_t)totvert);
    memcpy(p->mask, mask, sizeof(float) * (size_t)totvert);
  }
  p->poly_start = malloc(sizeof(int) * (size_t)totpoly);
  p->poly_len = malloc(sizeof(int) * (size_t)totpoly);
  p->loop_v = malloc(sizeof(int) * (size_t)totloop);
  memcpy(p->poly_start, poly_start, sizeof(int) * (size_t)totpoly);
  memcpy(p->poly_len, poly_len, sizeof(int) * (size_t)totpoly);
  memcpy(p->loop_v, loop_v, sizeof(int) * (size_t)totloop);

  p->poly_mat = dup_bytes(g_next_attrs.poly_mat, sizeof(short) * (size_t)totpoly);
  p->poly_flag = dup_bytes(g_next_attrs.poly_flag, (size_t)totpoly);
  p->vert_flag = dup_bytes(g_next_attrs.vert_flag, (size_t)totvert);
  memset(&g_next_attrs, 0, sizeof(g_next_attrs));

  const int looptri_num = or_looptri_count(totpoly, poly_len);
  p->tri_loop = malloc(sizeof(int[3]) * (size_t)(looptri_num > 0 ? looptri_num : 1));
  p->tri_v = malloc(sizeof(int[3]) * (size_t)(looptri_num > 0 ? looptri_num : 1));
  p->tri_poly = malloc(sizeof(int) * (size_t)(looptri_num > 0 ? looptri_num : 1));
  or_looptri_calc(totpoly, poly_start, poly_len, loop_v, (const float(*)[3])p->co, p->tri_loop, p->tri_poly);
  for (int i = 0; i < looptri_num; i++) {
    for (int j = 0; j < 3; j++) {
      p->tri_v[i][j] = loop_v[p->tri_loop[i][j]];
    }
  }
  p->vert_bitmap = calloc((size_t)totvert + 1, 1);

  OrBB cb;
  BB_reset(&cb);
  OrBBC *prim_bbc = malloc(sizeof(OrBBC) * (size_t)(looptri_num > 0 ? looptri_num : 1));
  for (int i = 0; i < looptri_num; i++) {
    OrBBC *bbc = prim_bbc + i;
    BB_reset((OrBB *)bbc);
    for (int j = 0; j < 3; j++) {
      BB_expand((OrBB *)bbc, p->co[p->tri_v[i][j]]);
    }
    for (int k = 0; k < 3; k++) { /* pbvh.c:2018-2023 */
      bbc->bcentroid[k] = (bbc->bmin[k] + bbc->bmax[k]) * 0.5f;
    }
    BB_expand(&cb, bbc->bcentroid);
  }

  if (looptri_num) {
    p->totprim = looptri_num;
    p->prim_indices = malloc(sizeof(int) * (size_t)looptri_num);
    for (int i = 0; i < looptri_num; i++) {
      p->prim_indices[i] = i;
    }
    p->node_mem_count = 100;
    p->nodes = calloc((size_t)p->node_mem_count, sizeof(OrNode));
    p->totnode = 1;
    LeafMap map;
    map.stamp = malloc(sizeof(int) * (size_t)totvert);
    map.value = malloc(sizeof(int) * (size_t)totvert);
    for (int i = 0; i < totvert; i++) {
      map.stamp[i] = -1;
    }
    build_sub(p, &map, 0, &cb, prim_bbc, 0, looptri_num);
    free(map.stamp);
    free(map.value);
  }
  free(prim_bbc);
  memset(p->vert_bitmap, 0, (size_t)totvert); /* pbvh.c:2512-2513 */

  /* sculpt-session side tables (kernel/intern/paint.c:1685-1688 builds pmap at session start) */
  p->nb_off = malloc(sizeof(int) * ((size_t)totvert + 1));
  p->nb_idx = malloc(sizeof(int) * (size_t)(2 * totloop + 1));
  p->boundary = malloc((size_t)totvert + 1);
  or_vert_neighbors(totvert, totpoly, poly_start, poly_len, loop_v, p->nb_off, p->nb_idx, p->boundary);
  p->orig_co = calloc((size_t)totvert, sizeof(float[3]));
  p->orig_no = calloc((size_t)totvert, sizeof(float[3]));
  p->touched = calloc((size_t)p->totnode + 1, 1);
  p->last_hits = malloc(sizeof(int) * (size_t)(p->totnode + 1));
  p->last_moved = malloc(sizeof(int) * ((size_t)totvert + 1));
  p->scratch = malloc(sizeof(float[3]) * ((size_t)totvert + 1));
  p->iter_flag = calloc((size_t)totvert + 1, 1);
  p->moved_stamp = calloc((size_t)totvert + 1, sizeof(int));
  return p;
}

static int *dup_ints(const int *src, size_t n)
{
  int *d = malloc(sizeof(int) * (n ? n : 1));
  if (n) memcpy(d, src, sizeof(int) * n);
  return d;
}

/* pbvh.c:2516-2561 BKE_pbvh_build_grids (+ pbvh_build 2427-2450) */
OrPbvh *or_pbvh_build_grids(int totgrid, int grid_size, const float (*co)[3], const float (*no)[3], const float *mask,
                            int totface, const int *face_start, const int *face_num, int totedge, const int *edge_off,
                            const int *edge_elems, int totcvert, const int *cvert_off, const int *cvert_elems,
                            const int *grid_edge, const int *grid_cvert, int leaf_limit)
{
  OrPbvh *p = calloc(1, sizeof(OrPbvh));
  const int gs2 = grid_size * grid_size;
  const size_t totelem = (size_t)totgrid * (size_t)gs2;
  p->is_grids = 1;
  p->totgrid = totgrid;
  p->grid_size = grid_size;
  p->totvert = (int)totelem;
  {
    const int lim = LEAF_LIMIT / gs2; /* pbvh.c:2533 */
    p->leaf_limit = leaf_limit > 0 ? leaf_limit : (lim > 1 ? lim : 1);
  }
  p->co = malloc(sizeof(float[3]) * totelem);
  memcpy(p->co, co, sizeof(float[3]) * totelem);
  p->no = calloc(totelem, sizeof(float[3]));
  if (no) memcpy(p->no, no, sizeof(float[3]) * totelem);
  if (mask) {
    p->mask = malloc(sizeof(float) * totelem);
    memcpy(p->mask, mask, sizeof(float) * totelem);
  }
  p->totface = totface;
  p->face_start = dup_ints(face_start, (size_t)totface);
  p->face_num = dup_ints(face_num, (size_t)totface);
  p->grid_face = malloc(sizeof(int) * (size_t)(totgrid ? totgrid : 1));
  for (int f = 0; f < totface; f++) {
    for (int c = 0; c < face_num[f]; c++) p->grid_face[face_start[f] + c] = f;
  }
  p->totedge = totedge;
  p->edge_off = dup_ints(edge_off, (size_t)totedge + 1);
  p->edge_elems = dup_ints(edge_elems, (size_t)edge_off[totedge] * 2 * (size_t)grid_size);
  p->totcvert = totcvert;
  p->cvert_off = dup_ints(cvert_off, (size_t)totcvert + 1);
  p->cvert_elems = dup_ints(cvert_elems, (size_t)cvert_off[totcvert]);
  p->grid_edge = dup_ints(grid_edge, (size_t)totgrid);
  p->grid_cvert = dup_ints(grid_cvert, (size_t)totgrid);
  p->grid_mat = dup_bytes(g_next_attrs.grid_mat, sizeof(short) * (size_t)totgrid);
  p->grid_flag = dup_bytes(g_next_attrs.grid_flag, (size_t)totgrid);
  p->grid_hidden = dup_bytes(g_next_attrs.grid_hidden, totelem);
  memset(&g_next_attrs, 0, sizeof(g_next_attrs));
  p->face_stamp = calloc((size_t)totface + 1, sizeof(int));
  p->edge_stamp = calloc((size_t)totedge + 1, sizeof(int));
  p->cvert_stamp = calloc((size_t)totcvert + 1, sizeof(int));
  p->vert_bitmap = calloc(totelem + 1, 1);

  OrBB cb;
  BB_reset(&cb);
  OrBBC *prim_bbc = malloc(sizeof(OrBBC) * (size_t)(totgrid > 0 ? totgrid : 1));
  for (int i = 0; i < totgrid; i++) {
    OrBBC *bbc = prim_bbc + i;
    BB_reset((OrBB *)bbc);
    for (int j = 0; j < gs2; j++) BB_expand((OrBB *)bbc, p->co[(size_t)i * gs2 + j]);
    for (int k = 0; k < 3; k++) bbc->bcentroid[k] = (bbc->bmin[k] + bbc->bmax[k]) * 0.5f;
    BB_expand(&cb, bbc->bcentroid);
  }
  if (totgrid) {
    p->totprim = totgrid;
    p->prim_indices = malloc(sizeof(int) * (size_t)totgrid);
    for (int i = 0; i < totgrid; i++) p->prim_indices[i] = i;
    p->node_mem_count = 100;
    p->nodes = calloc((size_t)p->node_mem_count, sizeof(OrNode));
    p->totnode = 1;
    build_sub(p, NULL, 0, &cb, prim_bbc, 0, totgrid);
  }
  free(prim_bbc);

  p->orig_co = calloc(totelem, sizeof(float[3]));
  p->orig_no = calloc(totelem, sizeof(float[3]));
  p->touched = calloc((size_t)p->totnode + 1, 1);
  p->last_hits = malloc(sizeof(int) * (size_t)(p->totnode + 1));
  p->last_moved = malloc(sizeof(int) * (totelem + 1));
  p->moved_stamp = calloc(totelem + 1, sizeof(int));
  return p;
}

float *or_pbvh_mask(OrPbvh *p) { return p->mask; }

/* pbvh.c:2570-2620 */
void or_pbvh_free(OrPbvh *p)
{
  if (!p) {
    return;
  }
  for (int i = 0; i < p->totnode; i++) {
    if (p->nodes[i].flag & OR_PBVH_Leaf) {
      free(p->nodes[i].vert_indices);
      free(p->nodes[i].face_vert_indices);
    }
  }
  free(p->nodes); free(p->prim_indices); free(p->co); free(p->no); free(p->mask);
  free(p->poly_start); free(p->poly_len); free(p->loop_v); free(p->tri_loop); free(p->tri_v);
  free(p->vt_off); free(p->vt_pos); free(p->pos_node);
  free(p->poly_mat); free(p->grid_mat); free(p->poly_flag); free(p->grid_flag); free(p->vert_flag); free(p->grid_hidden);
  free(p->tri_poly); free(p->vert_bitmap); free(p->nb_off); free(p->nb_idx); free(p->boundary);
  free(p->face_start); free(p->face_num); free(p->grid_face); free(p->edge_off); free(p->edge_elems);
  free(p->cvert_off); free(p->cvert_elems); free(p->grid_edge); free(p->grid_cvert);
  free(p->edge_verts); free(p->cvert_edge_off); free(p->cvert_edges); free(p->cvert_boundary);
  free(p->face_stamp); free(p->edge_stamp); free(p->cvert_stamp);
  free(p->automask); free(p->orig_co); free(p->orig_no); free(p->touched); free(p->last_hits);
  free(p->last_moved); free(p->scratch); free(p->iter_flag); free(p->moved_stamp);
  free(p);
}

int or_pbvh_totnode(const OrPbvh *p) { return p->totnode; }
int or_pbvh_tottri(const OrPbvh *p) { return p->totprim; }
int or_pbvh_totvert(const OrPbvh *p) { return p->totvert; }
const int *or_pbvh_prim_indices(const OrPbvh *p) { return p->prim_indices; }
const int *or_pbvh_node_vert_indices(const OrPbvh *p, int n) { return p->nodes[n].vert_indices; }
const int *or_pbvh_node_face_vert_indices(const OrPbvh *p, int n) { return (const int *)p->nodes[n].face_vert_indices; }
const int *or_pbvh_tri_verts(const OrPbvh *p) { return (const int *)p->tri_v; }
const int *or_pbvh_tri_poly(const OrPbvh *p) { return p->tri_poly; }
float *or_pbvh_co(OrPbvh *p) { return (float *)p->co; }
float *or_pbvh_no(OrPbvh *p) { return (float *)p->no; }
float *or_pbvh_orig_co(OrPbvh *p) { return (float *)p->orig_co; }
float *or_pbvh_orig_no(OrPbvh *p) { return (float *)p->orig_no; }

void or_pbvh_nodes(const OrPbvh *p, OrBB *vb, OrBB *orig_vb, int *children_offset, int *flag,
                   int *prim_offset, int *totprim, int *uniq_verts, int *face_verts)
{
  for (int i = 0; i < p->totnode; i++) {
    const OrNode *n = &p->nodes[i];
    if (vb) vb[i] = n->vb;
    if (orig_vb) orig_vb[i] = n->orig_vb;
    if (children_offset) children_offset[i] = n->children_offset;
    if (flag) flag[i] = (int)n->flag;
    if (prim_offset) prim_offset[i] = n->prim_offset;
    if (totprim) totprim[i] = n->totprim;
    if (uniq_verts) uniq_verts[i] = n->uniq_verts;
    if (face_verts) face_verts[i] = n->face_verts;
  }
}

void or_pbvh_node_set_flag(OrPbvh *p, int node, int flag, int on)
{
  if (on) {
    p->nodes[node].flag |= (unsigned)flag;
  }
  else {
    p->nodes[node].flag &= ~(unsigned)flag;
  }
}

/* ---- traversal: pbvh.c:2622-2705 (pbvh_iter_begin / pbvh_stack_push / pbvh_iter_next) ---- */
typedef int (*OrSearchCb)(OrPbvh *p, OrNode *node, void *data);

typedef struct OrStackItem {
  int node;
  int revisiting;
} OrStackItem;

static int search_gather(OrPbvh *p, OrSearchCb scb, void *data, int *r_nodes)
{
  /* pbvh.c:2736-2767 BKE_pbvh_search_gather: leaves in pbvh_iter_next order */
  int tot = 0;
  if (p->totnode == 0) {
    return 0;
  }
  int space = 100, size = 0;
  OrStackItem *stack = malloc(sizeof(OrStackItem) * (size_t)space);
  stack[size].node = 0;
  stack[size].revisiting = 0;
  size++;
  while (size) {
    size--;
    const int ni = stack[size].node;
    const int revisiting = stack[size].revisiting;
    OrNode *node = &p->nodes[ni];
    if (revisiting) {
      continue; /* inner node handed back to the caller, which only keeps leaves (pbvh.c:2746) */
    }
    if (scb && !scb(p, node, data)) {
      continue;
    }
    if (node->flag & OR_PBVH_Leaf) {
      r_nodes[tot++] = ni;
      continue;
    }
    if (size + 3 > space) {
      space *= 2;
      stack = realloc(stack, sizeof(OrStackItem) * (size_t)space);
    }
    stack[size].node = ni; stack[size].revisiting = 1; size++;
    stack[size].node = node->children_offset + 1; stack[size].revisiting = 0; size++;
    stack[size].node = node->children_offset; stack[size].revisiting = 0; size++;
  }
  free(stack);
  return tot;
}

/* DAGGER sphere search callback; AABB accessors pbvh.c:3830-3840,
 * fully hidden / masked pbvh.c:3690-3710 */
typedef struct SphereData {
  const float *center;
  float radius_squared;
  int original, ignore_fully_ineffective;
} SphereData;

static int search_sphere_cb(OrPbvh *p, OrNode *node, void *data_v)
{
  (void)p;
  const SphereData *data = data_v;
  const float *center = data->center;
  float nearest[3], t[3];
  if (data->ignore_fully_ineffective) {
    if ((node->flag & OR_PBVH_Leaf) && (node->flag & OR_PBVH_FullyHidden)) {
      return 0;
    }
    if ((node->flag & OR_PBVH_Leaf) && (node->flag & OR_PBVH_FullyMasked)) {
      return 0;
    }
  }
  const OrBB *bb = data->original ? &node->orig_vb : &node->vb;
  for (int i = 0; i < 3; i++) {
    if (bb->bmin[i] > center[i]) {
      nearest[i] = bb->bmin[i];
    }
    else if (bb->bmax[i] < center[i]) {
      nearest[i] = bb->bmax[i];
    }
    else {
      nearest[i] = center[i];
    }
  }
  t[0] = center[0] - nearest[0];
  t[1] = center[1] - nearest[1];
  t[2] = center[2] - nearest[2];
  return (t[0] * t[0] + t[1] * t[1] + t[2] * t[2]) < data->radius_squared;
}

/* DAGGER tube falloff: the node passes when the view line through the brush location comes closer to its box than the
 * radius (upstream: SCULPT_search_circle_cb over dist_squared_ray_to_aabb_v3).  Fixed here as: zero when the line
 * crosses the box (slab test), else the least line-to-edge distance over the box's 12 edges. */
float or_line_aabb_distsq(const float loc[3], const float n[3], const float bmin[3], const float bmax[3])
{
  float tmin = -FLT_MAX, tmax = FLT_MAX;
  int inside = 1;
  for (int k = 0; k < 3; k++) {
    if (n[k] != 0.0f) {
      const float t1 = (bmin[k] - loc[k]) / n[k], t2 = (bmax[k] - loc[k]) / n[k];
      const float lo = t1 < t2 ? t1 : t2, hi = t1 < t2 ? t2 : t1;
      if (lo > tmin) tmin = lo;
      if (hi < tmax) tmax = hi;
    }
    else if (loc[k] < bmin[k] || loc[k] > bmax[k]) {
      inside = 0;
    }
  }
  if (inside && tmin <= tmax) return 0.0f;
  const float a = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
  float best = FLT_MAX;
  for (int axis = 0; axis < 3; axis++) {
    const int u = (axis + 1) % 3, v = (axis + 2) % 3;
    for (int c = 0; c < 4; c++) {
      float p0[3], e[3] = {0.0f, 0.0f, 0.0f};
      p0[axis] = bmin[axis];
      p0[u] = (c & 1) ? bmax[u] : bmin[u];
      p0[v] = (c & 2) ? bmax[v] : bmin[v];
      e[axis] = bmax[axis] - bmin[axis];
      const float w[3] = {p0[0] - loc[0], p0[1] - loc[1], p0[2] - loc[2]};
      const float b = n[0] * e[0] + n[1] * e[1] + n[2] * e[2];
      const float cc = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
      const float dd = n[0] * w[0] + n[1] * w[1] + n[2] * w[2];
      const float ee = e[0] * w[0] + e[1] * w[1] + e[2] * w[2];
      const float denom = a * cc - b * b;
      float sp = 0.0f;
      if (denom > 1.0e-30f) {
        sp = (b * dd - a * ee) / denom;
        sp = sp < 0.0f ? 0.0f : (sp > 1.0f ? 1.0f : sp);
      }
      const float tp = (a > 0.0f) ? (dd + b * sp) / a : 0.0f;
      float dist = 0.0f;
      for (int k = 0; k < 3; k++) {
        const float df = (w[k] + e[k] * sp) - n[k] * tp;
        dist += df * df;
      }
      if (dist < best) best = dist;
    }
  }
  return best;
}
typedef struct TubeData {
  const float *center, *normal;
  float radius_squared;
  int original, ignore_fully_ineffective;
} TubeData;
static int search_tube_cb(OrPbvh *p, OrNode *node, void *data_v)
{
  (void)p;
  const TubeData *data = data_v;
  /* inner nodes never prune: the line distance is not exactly monotone in the box in floating point (the sphere test is),
   * so a leaf passes on its own box alone */
  if (!(node->flag & OR_PBVH_Leaf)) return 1;
  if (data->ignore_fully_ineffective && (node->flag & (OR_PBVH_FullyHidden | OR_PBVH_FullyMasked))) return 0;
  const OrBB *bb = data->original ? &node->orig_vb : &node->vb;
  return or_line_aabb_distsq(data->center, data->normal, bb->bmin, bb->bmax) < data->radius_squared;
}
int or_gather_tube(OrPbvh *p, const float center[3], const float normal[3], float radius_sq, int original,
                   int ignore_fully_ineffective, int *r_nodes)
{
  TubeData d = {center, normal, radius_sq, original, ignore_fully_ineffective};
  return search_gather(p, search_tube_cb, &d, r_nodes);
}

int or_gather_sphere(OrPbvh *p, const float center[3], float radius_sq, int original,
                     int ignore_fully_ineffective, int *r_nodes)
{
  SphereData d = {center, radius_sq, original, ignore_fully_ineffective};
  return search_gather(p, search_sphere_cb, &d, r_nodes);
}

/* pbvh.c:2891-2900 update_search_cb */
static int update_search_cb(OrPbvh *p, OrNode *node, void *data_v)
{
  (void)p;
  const int flag = *(int *)data_v;
  if (node->flag & OR_PBVH_Leaf) {
    return (node->flag & (unsigned)flag) != 0;
  }
  return 1;
}

int or_gather_flag(OrPbvh *p, int flag, int *r_nodes)
{
  return search_gather(p, update_search_cb, &flag, r_nodes);
}

/* pbvh.c:3641-3645, 3729-3733 */
void or_node_mark_update(OrPbvh *p, int node)
{
  p->nodes[node].flag |= OR_PBVH_UpdateNormals | OR_PBVH_UpdateBB | OR_PBVH_UpdateOriginalBB |
                         OR_PBVH_UpdateDrawBuffers | OR_PBVH_UpdateRedraw;
}
void or_vert_mark_update(OrPbvh *p, int v) { p->vert_bitmap[v] = 1; }

/* lib/intern/math_vector_inline.c:1165-1181 normalize_v3 */
static float normalize_v3(float n[3])
{
  float d = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
  if (d > 1.0e-35f) {
    d = sqrtf(d);
    const float f = 1.0f / d;
    n[0] = n[0] * f;
    n[1] = n[1] * f;
    n[2] = n[2] * f;
  }
  else {
    n[0] = n[1] = n[2] = 0.0f;
    d = 0.0f;
  }
  return d;
}

/* kernel/intern/mesh_evaluate.c:39-86 BKE_mesh_calc_poly_normal;
 * lib/intern/math_geom.cc:31-69 normal_tri_v3 / normal_quad_v3;
 * lib/intern/math_vector_inline.c:961-966 add_newell_cross_v3_v3v3 */
static void calc_poly_normal(const OrPbvh *p, int poly, float r_no[3])
{
  const int ls = p->poly_start[poly], n = p->poly_len[poly];
  const int *lv = p->loop_v + ls;
  if (n > 4) {
    const float *v_prev = p->co[lv[n - 1]];
    r_no[0] = r_no[1] = r_no[2] = 0.0f;
    for (int i = 0; i < n; i++) {
      const float *v_curr = p->co[lv[i]];
      r_no[0] += (v_prev[1] - v_curr[1]) * (v_prev[2] + v_curr[2]);
      r_no[1] += (v_prev[2] - v_curr[2]) * (v_prev[0] + v_curr[0]);
      r_no[2] += (v_prev[0] - v_curr[0]) * (v_prev[1] + v_curr[1]);
      v_prev = v_curr;
    }
    if (normalize_v3(r_no) == 0.0f) {
      r_no[2] = 1.0f;
    }
  }
  else if (n == 3) {
    const float *v1 = p->co[lv[0]], *v2 = p->co[lv[1]], *v3 = p->co[lv[2]];
    float n1[3], n2[3];
    n1[0] = v1[0] - v2[0]; n2[0] = v2[0] - v3[0];
    n1[1] = v1[1] - v2[1]; n2[1] = v2[1] - v3[1];
    n1[2] = v1[2] - v2[2]; n2[2] = v2[2] - v3[2];
    r_no[0] = n1[1] * n2[2] - n1[2] * n2[1];
    r_no[1] = n1[2] * n2[0] - n1[0] * n2[2];
    r_no[2] = n1[0] * n2[1] - n1[1] * n2[0];
    normalize_v3(r_no);
  }
  else if (n == 4) {
    const float *v1 = p->co[lv[0]], *v2 = p->co[lv[1]], *v3 = p->co[lv[2]], *v4 = p->co[lv[3]];
    float n1[3], n2[3];
    n1[0] = v1[0] - v3[0]; n1[1] = v1[1] - v3[1]; n1[2] = v1[2] - v3[2];
    n2[0] = v2[0] - v4[0]; n2[1] = v2[1] - v4[1]; n2[2] = v2[2] - v4[2];
    r_no[0] = n1[1] * n2[2] - n1[2] * n2[1];
    r_no[1] = n1[2] * n2[0] - n1[0] * n2[2];
    r_no[2] = n1[0] * n2[1] - n1[1] * n2[0];
    normalize_v3(r_no);
  }
  else {
    r_no[0] = 0.0f; r_no[1] = 0.0f; r_no[2] = 1.0f;
  }
}

/* pbvh.c:2912-3036 pbvh_faces_update_normals: clear / accumulate / store over flagged nodes.
 * One thread: the accumulation order is nodes in gather order, looptris in node order -- for one
 * vertex that is ascending position in prim_indices.  Threads > 1: float atomics as the
 * reference (pbvh.c:2970-2976), order not defined. */
/* Threads > 1 with or_set_ordered_normals(1) (the full-size parity legs of bench.py): the accumulation runs per vertex
 * instead of per looptri -- every dirty unique vert of a flagged node sums the poly normals of its incident looptris in
 * ascending position in prim_indices, looptris of unflagged nodes left out (pbvh.c:2943) -- which is the order the
 * single-threaded loop above produces, so the threaded oracle is bit-identical to the serial one. */
int or_ordered_normals = 0;
void or_set_ordered_normals(int on) { or_ordered_normals = on; }

static void ensure_vert_tri_positions(OrPbvh *p)
{
  if (p->vt_off) return;
  const int T = p->totprim, V = p->totvert;
  p->vt_off = calloc((size_t)V + 2, sizeof(int64_t));
  p->pos_node = malloc(sizeof(int) * (size_t)(T > 0 ? T : 1));
  for (int n = 0; n < p->totnode; n++) {
    const OrNode *node = &p->nodes[n];
    if (!(node->flag & OR_PBVH_Leaf)) continue;
    for (int i = 0; i < node->totprim; i++) p->pos_node[node->prim_offset + i] = n;
  }
  for (int pos = 0; pos < T; pos++) {
    const int *vt = p->tri_v[p->prim_indices[pos]];
    for (int j = 0; j < 3; j++) p->vt_off[vt[j] + 2]++;
  }
  for (int v = 0; v < V; v++) p->vt_off[v + 2] += p->vt_off[v + 1];
  p->vt_pos = malloc(sizeof(int) * (size_t)(3 * (int64_t)T > 0 ? 3 * (int64_t)T : 1));
  /* ascending position by construction; the corners of one looptri in the order the serial loop adds them (j = 2, 1, 0) */
  for (int pos = 0; pos < T; pos++) {
    const int *vt = p->tri_v[p->prim_indices[pos]];
    for (int j = 3; j--;) p->vt_pos[p->vt_off[vt[j] + 1]++] = pos;
  }
}

static void faces_update_normals(OrPbvh *p, const int *nodes, int totnode)
{
  const int par = (or_threads > 1 && totnode > 1); /* pbvh.c:4953-4959 */
  if (par && or_ordered_normals) {
    ensure_vert_tri_positions(p);
#pragma omp parallel for schedule(dynamic)
    for (int n = 0; n < totnode; n++) {
      OrNode *node = &p->nodes[nodes[n]];
      if (!(node->flag & OR_PBVH_UpdateNormals)) continue;
      for (int i = 0; i < node->uniq_verts; i++) {
        const int v = node->vert_indices[i];
        if (!p->vert_bitmap[v]) continue;
        float acc[3] = {0.0f, 0.0f, 0.0f};
        for (int64_t q = p->vt_off[v]; q < p->vt_off[v + 1]; q++) {
          const int pos = p->vt_pos[q];
          if (!(p->nodes[p->pos_node[pos]].flag & OR_PBVH_UpdateNormals)) continue;
          float fn[3];
          calc_poly_normal(p, p->tri_poly[p->prim_indices[pos]], fn);
          acc[2] += fn[2]; acc[1] += fn[1]; acc[0] += fn[0];
        }
        normalize_v3(acc);
        memcpy(p->no[v], acc, sizeof(acc));
      }
    }
    /* flags and dirty bits drop only after every node has read its neighbours' flags */
#pragma omp parallel for schedule(dynamic)
    for (int n = 0; n < totnode; n++) {
      OrNode *node = &p->nodes[nodes[n]];
      if (!(node->flag & OR_PBVH_UpdateNormals)) continue;
      for (int i = 0; i < node->uniq_verts; i++) p->vert_bitmap[node->vert_indices[i]] = 0;
      node->flag &= ~(unsigned)OR_PBVH_UpdateNormals;
    }
    return;
  }
#pragma omp parallel for schedule(dynamic) if (par)
  for (int n = 0; n < totnode; n++) {
    OrNode *node = &p->nodes[nodes[n]];
    if (node->flag & OR_PBVH_UpdateNormals) {
      for (int i = 0; i < node->uniq_verts; i++) {
        const int v = node->vert_indices[i];
        if (p->vert_bitmap[v]) {
          p->no[v][0] = p->no[v][1] = p->no[v][2] = 0.0f;
        }
      }
    }
  }
#pragma omp parallel for schedule(dynamic) if (par)
  for (int n = 0; n < totnode; n++) {
    OrNode *node = &p->nodes[nodes[n]];
    if (node->flag & OR_PBVH_UpdateNormals) {
      int mpoly_prev = -1;
      float fn[3] = {0, 0, 0};
      const int *faces = p->prim_indices + node->prim_offset;
      for (int i = 0; i < node->totprim; i++) {
        const int t = faces[i];
        const int *vtri = p->tri_v[t];
        if (p->tri_poly[t] != mpoly_prev) {
          calc_poly_normal(p, p->tri_poly[t], fn);
          mpoly_prev = p->tri_poly[t];
        }
        for (int j = 3; j--;) {
          const int v = vtri[j];
          if (p->vert_bitmap[v]) {
            for (int k = 3; k--;) {
              if (par) {
#pragma omp atomic
                p->no[v][k] += fn[k];
              }
              else {
                p->no[v][k] += fn[k];
              }
            }
          }
        }
      }
    }
  }
#pragma omp parallel for schedule(dynamic) if (par)
  for (int n = 0; n < totnode; n++) {
    OrNode *node = &p->nodes[nodes[n]];
    if (node->flag & OR_PBVH_UpdateNormals) {
      for (int i = 0; i < node->uniq_verts; i++) {
        const int v = node->vert_indices[i];
        if (p->vert_bitmap[v]) {
          normalize_v3(p->no[v]);
          p->vert_bitmap[v] = 0;
        }
      }
      node->flag &= ~(unsigned)OR_PBVH_UpdateNormals;
    }
  }
}

/* pbvh.c:4559-4587 BKE_pbvh_update_normals (PBVH_FACES branch) */
void or_update_normals(OrPbvh *p)
{
  int *nodes = malloc(sizeof(int) * (size_t)(p->totnode + 1));
  const int totnode = or_gather_flag(p, OR_PBVH_UpdateNormals, nodes);
  if (totnode > 0) {
    if (p->is_grids) {
      /* PBVH_GRIDS branch, pbvh.c:4575-4583 */
      int *faces = malloc(sizeof(int) * (size_t)(p->totface + 1));
      const int num_faces = or_grids_get_updates(p, 1, faces);
      if (num_faces > 0) or_grids_update_normals(p, faces, num_faces);
      free(faces);
    }
    else {
      faces_update_normals(p, nodes, totnode);
    }
  }
  free(nodes);
}

void or_recalc_all_normals(OrPbvh *p)
{
  for (int i = 0; i < p->totnode; i++) {
    if (p->nodes[i].flag & OR_PBVH_Leaf) {
      p->nodes[i].flag |= OR_PBVH_UpdateNormals;
    }
  }
  memset(p->vert_bitmap, 1, (size_t)p->totvert);
  or_update_normals(p);
}

/* ------------------------------------------------------------------------------ ray-cast */
/* lib/intern/math_geom.cc:3017-3030 */
typedef struct RayAABB {
  float ray_origin[3], ray_inv_dir[3];
  int sign[3];
  int original;
} RayAABB;

/* lib/intern/math_geom.cc:3032-3080 isect_ray_aabb_v3 */
static int isect_ray_aabb(const RayAABB *d, const float bb_min[3], const float bb_max[3], float *tmin_out)
{
  float bbox[2][3];
  memcpy(bbox[0], bb_min, sizeof(float[3]));
  memcpy(bbox[1], bb_max, sizeof(float[3]));
  float tmin = (bbox[d->sign[0]][0] - d->ray_origin[0]) * d->ray_inv_dir[0];
  float tmax = (bbox[1 - d->sign[0]][0] - d->ray_origin[0]) * d->ray_inv_dir[0];
  const float tymin = (bbox[d->sign[1]][1] - d->ray_origin[1]) * d->ray_inv_dir[1];
  const float tymax = (bbox[1 - d->sign[1]][1] - d->ray_origin[1]) * d->ray_inv_dir[1];
  if ((tmin > tymax) || (tymin > tmax)) return 0;
  if (tymin > tmin) tmin = tymin;
  if (tymax < tmax) tmax = tymax;
  const float tzmin = (bbox[d->sign[2]][2] - d->ray_origin[2]) * d->ray_inv_dir[2];
  const float tzmax = (bbox[1 - d->sign[2]][2] - d->ray_origin[2]) * d->ray_inv_dir[2];
  if ((tmin > tzmax) || (tzmin > tmax)) return 0;
  if (tzmin > tmin) tmin = tzmin;
  *tmin_out = tmin;
  return 1;
}

/* pbvh.c:3897-3915 ray_aabb_intersect */
static int ray_aabb_cb(OrPbvh *p, OrNode *node, void *data_v)
{
  (void)p;
  const RayAABB *d = data_v;
  const OrBB *bb = d->original ? &node->orig_vb : &node->vb;
  return isect_ray_aabb(d, bb->bmin, bb->bmax, &node->tmin);
}

/* lib/intern/math_geom.cc:1755-1780, 1782-1857 isect_ray_tri_watertight_v3 */
typedef struct RayTri {
  int kx, ky, kz;
  float sx, sy, sz;
} RayTri;

static void ray_tri_precalc(RayTri *pc, const float dir[3])
{
  const float x = fabsf(dir[0]), y = fabsf(dir[1]), z = fabsf(dir[2]);
  int kz = ((x > y) ? ((x > z) ? 0 : 2) : ((y > z) ? 1 : 2));
  int kx = (kz != 2) ? (kz + 1) : 0;
  int ky = (kx != 2) ? (kx + 1) : 0;
  if (dir[kz] < 0.0f) {
    const int t = kx;
    kx = ky;
    ky = t;
  }
  const float inv_dir_z = 1.0f / dir[kz];
  pc->sx = dir[kx] * inv_dir_z;
  pc->sy = dir[ky] * inv_dir_z;
  pc->sz = inv_dir_z;
  pc->kx = kx; pc->ky = ky; pc->kz = kz;
}

static int ray_tri_watertight(const float o[3], const RayTri *pc, const float v0[3], const float v1[3], const float v2[3],
                              float *r_lambda)
{
  const int kx = pc->kx, ky = pc->ky, kz = pc->kz;
  const float sx = pc->sx, sy = pc->sy, sz = pc->sz;
  const float a[3] = {v0[0] - o[0], v0[1] - o[1], v0[2] - o[2]};
  const float b[3] = {v1[0] - o[0], v1[1] - o[1], v1[2] - o[2]};
  const float c[3] = {v2[0] - o[0], v2[1] - o[1], v2[2] - o[2]};
  const float a_kx = a[kx], a_ky = a[ky], a_kz = a[kz];
  const float b_kx = b[kx], b_ky = b[ky], b_kz = b[kz];
  const float c_kx = c[kx], c_ky = c[ky], c_kz = c[kz];
  const float ax = a_kx - sx * a_kz, ay = a_ky - sy * a_kz;
  const float bx = b_kx - sx * b_kz, by = b_ky - sy * b_kz;
  const float cx = c_kx - sx * c_kz, cy = c_ky - sy * c_kz;
  const float u = cx * by - cy * bx;
  const float v = ax * cy - ay * cx;
  const float w = bx * ay - by * ax;
  if ((u < 0.0f || v < 0.0f || w < 0.0f) && (u > 0.0f || v > 0.0f || w > 0.0f)) return 0;
  const float det = u + v + w;
  if (det == 0.0f || !isfinite(det)) return 0;
  union { float f; unsigned i; } ud, ut;
  ud.f = det;
  const unsigned sign_det = ud.i & 0x80000000u;
  const float t = (u * a_kz + v * b_kz + w * c_kz) * sz;
  ut.f = t;
  ut.i ^= sign_det; /* xor_fl */
  if (ut.f < 0.0f) return 0;
  const float inv_det = 1.0f / det;
  *r_lambda = t * inv_det;
  return 1;
}

typedef struct RayNodeRef {
  int node;
  float tmin;
} RayNodeRef;

int or_raycast(OrPbvh *p, const float ray_start[3], const float ray_normal[3], int original, float max_depth, float *r_depth,
               int *r_vertex, int *r_face, float r_face_normal[3], int *r_node)
{
  if (p->totnode == 0) return 0;
  RayAABB d;
  memcpy(d.ray_origin, ray_start, sizeof(float[3]));
  for (int k = 0; k < 3; k++) {
    d.ray_inv_dir[k] = 1.0f / ray_normal[k];
    d.sign[k] = d.ray_inv_dir[k] < 0.0f;
  }
  d.original = original;
  int *nodes = malloc(sizeof(int) * (size_t)(p->totnode + 1));
  const int tot = search_gather(p, ray_aabb_cb, &d, nodes);
  /* BKE_pbvh_search_callback_occluded (pbvh.c:2852-2889): an unbalanced tree keyed by tmin, ties to the right,
   * walked in order = a stable sort of the leaves by tmin */
  RayNodeRef *ord = malloc(sizeof(RayNodeRef) * (size_t)(tot + 1));
  for (int i = 0; i < tot; i++) {
    RayNodeRef r = {nodes[i], p->nodes[nodes[i]].tmin};
    int k = i;
    while (k > 0 && r.tmin < ord[k - 1].tmin) {
      ord[k] = ord[k - 1];
      k--;
    }
    ord[k] = r;
  }
  /* an undo node holds the coordinates of ALL verts of its leaf as they were when it was pushed (row a9); for a
   * vert the leaf shares, that is the owner's snapshot if the owner was pushed too, else the (unmoved) current one */
  int *owner = NULL;
  if (original && !p->is_grids) {
    owner = malloc(sizeof(int) * (size_t)(p->totvert + 1));
    for (int n = 0; n < p->totnode; n++) {
      if (p->nodes[n].flag & OR_PBVH_Leaf) {
        for (int i = 0; i < p->nodes[n].uniq_verts; i++) owner[p->nodes[n].vert_indices[i]] = n;
      }
    }
  }
  RayTri pc;
  ray_tri_precalc(&pc, ray_normal);
  float depth = max_depth; /* the caller starts from the ray's length to the far clip (SCULPT_raycast_init) */
  int hit = 0;
  float tmin = 3.402823466e+38f;
  for (int i = 0; i < tot; i++) {
    const OrNode *node = &p->nodes[ord[i].node];
    if (!(node->tmin < tmin)) continue; /* DAGGER sculpt_raycast_cb: only nodes the ray enters before the best hit */
    const int use_orig = original && p->touched[ord[i].node];
    const int *faces = p->prim_indices + node->prim_offset;
    int node_hit = 0;
    if (p->is_grids) {
      /* pbvh_grids_node_raycast (pbvh.c:4102-4200): the quads of the node's grids in grid, y, x order; a quad is its
       * two triangles (0, 1, 2) then (0, 2, 3), the second only looked at when the first is not a nearer hit
       * (ray_face_intersection_quad, pbvh.c:3930-3949).  A quad with a hidden corner is skipped (paint_is_grid_face_hidden,
       * paint.c:1234-1241).  r_face = the active grid. */
      const int gs = p->grid_size, gs2 = gs * gs;
      float (*src)[3] = use_orig ? p->orig_co : p->co;
      for (int gi = 0; gi < node->totprim; gi++) {
        const int g = faces[gi];
        for (int y = 0; y < gs - 1; y++) {
          for (int x = 0; x < gs - 1; x++) {
            const int e[4] = {g * gs2 + y * gs + x, g * gs2 + y * gs + x + 1, g * gs2 + (y + 1) * gs + x + 1, g * gs2 + (y + 1) * gs + x};
            if (p->grid_hidden && (p->grid_hidden[e[0]] || p->grid_hidden[e[1]] || p->grid_hidden[e[2]] || p->grid_hidden[e[3]])) continue;
            const float *co[4] = {src[e[0]], src[e[1]], src[e[2]], src[e[3]]};
            float depth_test;
            if (!((ray_tri_watertight(ray_start, &pc, co[0], co[1], co[2], &depth_test) && depth_test < depth) ||
                  (ray_tri_watertight(ray_start, &pc, co[0], co[2], co[3], &depth_test) && depth_test < depth)))
              continue;
            depth = depth_test;
            node_hit = 1;
            if (r_face_normal) {
              /* normal_quad_v3, lib/intern/math_geom.cc:51-69 */
              float n1[3], n2[3];
              for (int k = 0; k < 3; k++) {
                n1[k] = co[0][k] - co[2][k];
                n2[k] = co[1][k] - co[3][k];
              }
              r_face_normal[0] = n1[1] * n2[2] - n1[2] * n2[1];
              r_face_normal[1] = n1[2] * n2[0] - n1[0] * n2[2];
              r_face_normal[2] = n1[0] * n2[1] - n1[1] * n2[0];
              normalize_v3(r_face_normal);
            }
            float location[3], nearest[3] = {0.0f, 0.0f, 0.0f};
            for (int k = 0; k < 3; k++) location[k] = ray_start[k] + ray_normal[k] * depth;
            for (int j = 0; j < 4; j++) {
              float da = 0.0f, db = 0.0f;
              for (int k = 0; k < 3; k++) {
                da += (location[k] - co[j][k]) * (location[k] - co[j][k]);
                db += (location[k] - nearest[k]) * (location[k] - nearest[k]);
              }
              if (j == 0 || da < db) {
                memcpy(nearest, co[j], sizeof(float[3]));
                if (r_vertex) *r_vertex = e[j];
              }
            }
            if (r_face) *r_face = g;
            if (r_node) *r_node = ord[i].node;
          }
        }
      }
    }
    else
    for (int f = 0; f < node->totprim; f++) {
      const int *vt = p->tri_v[faces[f]];
      const float *co[3];
      for (int j = 0; j < 3; j++) co[j] = (use_orig && p->touched[owner[vt[j]]]) ? p->orig_co[vt[j]] : p->co[vt[j]];
      float depth_test;
      if (ray_tri_watertight(ray_start, &pc, co[0], co[1], co[2], &depth_test) && depth_test < depth) {
        depth = depth_test;
        node_hit = 1;
        if (r_face_normal) {
          /* normal_tri_v3 */
          float n1[3], n2[3];
          for (int k = 0; k < 3; k++) {
            n1[k] = co[0][k] - co[1][k];
            n2[k] = co[1][k] - co[2][k];
          }
          r_face_normal[0] = n1[1] * n2[2] - n1[2] * n2[1];
          r_face_normal[1] = n1[2] * n2[0] - n1[0] * n2[2];
          r_face_normal[2] = n1[0] * n2[1] - n1[1] * n2[0];
          normalize_v3(r_face_normal);
        }
        float location[3], nearest[3] = {0.0f, 0.0f, 0.0f};
        for (int k = 0; k < 3; k++) location[k] = ray_start[k] + ray_normal[k] * depth;
        for (int j = 0; j < 3; j++) {
          float da = 0.0f, db = 0.0f;
          for (int k = 0; k < 3; k++) {
            da += (location[k] - co[j][k]) * (location[k] - co[j][k]);
            db += (location[k] - nearest[k]) * (location[k] - nearest[k]);
          }
          if (j == 0 || da < db) {
            memcpy(nearest, co[j], sizeof(float[3]));
            if (r_vertex) *r_vertex = vt[j];
            if (r_face) *r_face = p->tri_poly[faces[f]];
          }
        }
        if (r_node) *r_node = ord[i].node;
      }
    }
    if (node_hit) {
      hit = 1;
      tmin = depth;
    }
  }
  free(ord);
  free(nodes);
  free(owner);
  if (hit && r_depth) *r_depth = depth;
  return hit;
}

/* gpu/intern/gpu_buffers.c:174-305 GPU_pbvh_mesh_buffers_update for one leaf, into the packed vertex
 * format of gpu_pbvh_init (gpu_buffers.c:84-100; offsets by VertexFormat_pack, gpu_vertex_format.cc:300-325):
 * 36 bytes per looptri corner -- pos f32 x 3 @0, nor i16 x 3 @16, msk u8 @22, col u16 x 4 @24, fset u8 x 3 @32.
 * No hidden faces, no face sets, no vertex colours on this path: every looptri is visible, col stays zero,
 * fset is white.  Clears PBVH_RebuildDrawBuffers | PBVH_UpdateDrawBuffers like pbvh.c:3276. */
/* gpu/intern/gpu_buffers.c:548-725 gpu_pbvh_grid_buffers_update for one grid leaf, same record format: smooth -- one
 * record per element in (grid, y, x) order with its own normal and mask; flat -- four records per quad (x, y), (x+1, y),
 * (x+1, y+1), (x, y+1), the quad normal taken with the corners reversed (gpu_buffers.c:664-666) and the mean of the four
 * masks.  On this branch col is written too (white, gpu_buffers.c:640-643 with show_vcol, 686-690 always when flat);
 * here: flat writes it, smooth leaves it zero (no vertex colours on this path).  Returns the record count. */
static int grid_draw_buffers_update(OrPbvh *p, OrNode *node, int smooth, int show_mask, unsigned char *out)
{
  const int stride = 36, gs = p->grid_size, gs2 = gs * gs;
  const int *grids = p->prim_indices + node->prim_offset;
  const int use_mask = show_mask && p->mask;
  const int per_grid = smooth ? gs2 : (gs - 1) * (gs - 1) * 4;
  memset(out, 0, (size_t)node->totprim * (size_t)per_grid * stride);
  unsigned char *rec = out;
  for (int i = 0; i < node->totprim; i++) {
    const int g = grids[i];
    if (smooth) {
      for (int e = g * gs2; e < (g + 1) * gs2; e++, rec += stride) {
        short no[3];
        for (int k = 0; k < 3; k++) no[k] = (short)(p->no[e][k] * 32767.0f);
        memcpy(rec, p->co[e], sizeof(float[3]));
        memcpy(rec + 16, no, sizeof(no));
        if (use_mask) rec[22] = (unsigned char)(p->mask[e] * 255);
        rec[32] = rec[33] = rec[34] = 255;
      }
      continue;
    }
    for (int y = 0; y < gs - 1; y++) {
      for (int x = 0; x < gs - 1; x++) {
        const int e[4] = {g * gs2 + y * gs + x, g * gs2 + y * gs + x + 1, g * gs2 + (y + 1) * gs + x + 1, g * gs2 + (y + 1) * gs + x};
        /* normal_quad_v3(fno, co[3], co[2], co[1], co[0]) */
        float n1[3], n2[3], fno[3];
        for (int k = 0; k < 3; k++) {
          n1[k] = p->co[e[3]][k] - p->co[e[1]][k];
          n2[k] = p->co[e[2]][k] - p->co[e[0]][k];
        }
        fno[0] = n1[1] * n2[2] - n1[2] * n2[1];
        fno[1] = n1[2] * n2[0] - n1[0] * n2[2];
        fno[2] = n1[0] * n2[1] - n1[1] * n2[0];
        normalize_v3(fno);
        short no[3];
        for (int k = 0; k < 3; k++) no[k] = (short)(fno[k] * 32767.0f);
        unsigned char cmask = 0;
        if (use_mask) {
          const float fmask = (p->mask[e[0]] + p->mask[e[1]] + p->mask[e[2]] + p->mask[e[3]]) * 0.25f;
          cmask = (unsigned char)(fmask * 255);
        }
        for (int j = 0; j < 4; j++, rec += stride) {
          memcpy(rec, p->co[e[j]], sizeof(float[3]));
          memcpy(rec + 16, no, sizeof(no));
          rec[22] = cmask;
          memset(rec + 24, 0xff, 8); /* col u16 x 4 = USHRT_MAX */
          rec[32] = rec[33] = rec[34] = 255;
        }
      }
    }
  }
  node->flag &= ~(unsigned)(OR_PBVH_RebuildDrawBuffers | OR_PBVH_UpdateDrawBuffers);
  return node->totprim * per_grid;
}

int or_draw_buffers_update(OrPbvh *p, int ni, int smooth, int show_mask, unsigned char *out)
{
  OrNode *node = &p->nodes[ni];
  if (!(node->flag & OR_PBVH_Leaf)) return 0;
  if (p->is_grids) return grid_draw_buffers_update(p, node, smooth, show_mask, out);
  const int stride = 36;
  const int *faces = p->prim_indices + node->prim_offset;
  const int use_mask = show_mask && p->mask;
  int mpoly_prev = -1;
  short no[3] = {0, 0, 0};
  memset(out, 0, (size_t)node->totprim * 3 * stride);
  for (int i = 0; i < node->totprim; i++) {
    const int t = faces[i];
    const int *vtri = p->tri_v[t];
    if (p->tri_poly[t] != mpoly_prev && !smooth) {
      float fno[3];
      calc_poly_normal(p, p->tri_poly[t], fno);
      for (int k = 0; k < 3; k++) no[k] = (short)(fno[k] * 32767.0f); /* normal_float_to_short_v3 */
      mpoly_prev = p->tri_poly[t];
    }
    unsigned char cmask = 0;
    if (use_mask && !smooth) {
      const float fmask = (p->mask[vtri[0]] + p->mask[vtri[1]] + p->mask[vtri[2]]) / 3.0f;
      cmask = (unsigned char)(fmask * 255);
    }
    for (int j = 0; j < 3; j++) {
      unsigned char *rec = out + ((size_t)i * 3 + j) * stride;
      memcpy(rec, p->co[vtri[j]], sizeof(float[3]));
      if (smooth) {
        for (int k = 0; k < 3; k++) no[k] = (short)(p->no[vtri[j]][k] * 32767.0f);
      }
      memcpy(rec + 16, no, sizeof(short[3]));
      if (use_mask && smooth) cmask = (unsigned char)(p->mask[vtri[j]] * 255);
      rec[22] = cmask;
      rec[32] = rec[33] = rec[34] = 255;
    }
  }
  node->flag &= ~(unsigned)(OR_PBVH_RebuildDrawBuffers | OR_PBVH_UpdateDrawBuffers);
  return node->totprim * 3;
}

/* pbvh.c:3287-3317 pbvh_flush_bb */
static int pbvh_flush_bb(OrPbvh *p, OrNode *node, int flag)
{
  int update = 0;
  if (node->flag & OR_PBVH_Leaf) {
    if (flag & OR_PBVH_UpdateBB) {
      update |= (int)(node->flag & OR_PBVH_UpdateBB);
      node->flag &= ~(unsigned)OR_PBVH_UpdateBB;
    }
    if (flag & OR_PBVH_UpdateOriginalBB) {
      update |= (int)(node->flag & OR_PBVH_UpdateOriginalBB);
      node->flag &= ~(unsigned)OR_PBVH_UpdateOriginalBB;
    }
    return update;
  }
  update |= pbvh_flush_bb(p, p->nodes + node->children_offset, flag);
  update |= pbvh_flush_bb(p, p->nodes + node->children_offset + 1, flag);
  if (update & OR_PBVH_UpdateBB) {
    update_node_vb(p, node);
  }
  if (update & OR_PBVH_UpdateOriginalBB) {
    node->orig_vb = node->vb;
  }
  return update;
}

/* pbvh.c:3319-3339 BKE_pbvh_update_bounds (+ 3124-3160 pbvh_update_BB_redraw) */
void or_update_bounds(OrPbvh *p, int flag)
{
  if (!p->nodes) {
    return;
  }
  int *nodes = malloc(sizeof(int) * (size_t)(p->totnode + 1));
  const int totnode = or_gather_flag(p, flag, nodes);
  if (flag & (OR_PBVH_UpdateBB | OR_PBVH_UpdateOriginalBB | OR_PBVH_UpdateRedraw)) {
    const int par = (or_threads > 1 && totnode > 1);
#pragma omp parallel for schedule(dynamic) if (par)
    for (int n = 0; n < totnode; n++) {
      OrNode *node = &p->nodes[nodes[n]];
      if ((flag & OR_PBVH_UpdateBB) && (node->flag & OR_PBVH_UpdateBB)) {
        update_node_vb(p, node);
      }
      if ((flag & OR_PBVH_UpdateOriginalBB) && (node->flag & OR_PBVH_UpdateOriginalBB)) {
        node->orig_vb = node->vb;
      }
      if ((flag & OR_PBVH_UpdateRedraw) && (node->flag & OR_PBVH_UpdateRedraw)) {
        node->flag &= ~(unsigned)OR_PBVH_UpdateRedraw;
      }
    }
  }
  if (flag & (OR_PBVH_UpdateBB | OR_PBVH_UpdateOriginalBB)) {
    pbvh_flush_bb(p, p->nodes, flag);
  }
  free(nodes);
}
