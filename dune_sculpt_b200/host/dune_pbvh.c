/* Host C side of the B200 sculpt-stroke path: the reference's PBVH entry points (see
 * include/dune_pbvh.h) over libdune_sculpt_cuda.  The PBVH topology is built here on the host --
 * it is a once-per-session step and fixes the parity-critical ownership and
 * ordering -- with the split rule, leaf limit and first-touch vertex ownership of
 * kernel/intern/pbvh.c:2070-2514; everything per-dab runs on the device.
 */
#ifndef _POSIX_C_SOURCE
#define _POSIX_C_SOURCE 199309L /* clock_gettime */
#endif
#include "../../include/dune_pbvh.h"

#include <time.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define LEAF_LIMIT 10000 /* pbvh.c:1950 */

#ifdef DUNE_PBVH_OWN_MEM
void *MEM_mallocN(size_t len, const char *str)
{
  (void)str;
  return malloc(len ? len : 1);
}
void *MEM_callocN(size_t len, const char *str)
{
  (void)str;
  return calloc(len ? len : 1, 1);
}
void MEM_freeN(void *vmemh) { free(vmemh); }
#endif

/* ------------------------------------------------------------------------------------ mesh */

int BKE_mesh_poly_to_tri_count(int totpoly, int totloop) { return totloop - 2 * totpoly; }

static bool quad_needs_flip(const float a[3], const float b[3], const float c[3], const float d[3])
{
  /* lib/intern/math_geom.cc:5362-5378: diagonal a-c is degenerate when b and d lie on its same side */
  float ab[3], ac[3], ad[3];
  for (int i = 0; i < 3; i++) {
    ab[i] = b[i] - a[i];
    ac[i] = c[i] - a[i];
    ad[i] = d[i] - a[i];
  }
  const float n1[3] = {ab[1] * ac[2] - ab[2] * ac[1], ab[2] * ac[0] - ab[0] * ac[2], ab[0] * ac[1] - ab[1] * ac[0]};
  const float n2[3] = {ad[1] * ac[2] - ad[2] * ac[1], ad[2] * ac[0] - ad[0] * ac[2], ad[0] * ac[1] - ad[1] * ac[0]};
  return (n1[0] * n2[0] + n1[1] * n2[1] + n1[2] * n2[2]) > 0.0f;
}

void BKE_mesh_recalc_looptri(const MLoop *mloop, const MPoly *mpoly, const MVert *mvert, int totloop, int totpoly,
                             MLoopTri *mlooptri)
{
  (void)totloop;
  MLoopTri *out = mlooptri;
  for (int p = 0; p < totpoly; p++) {
    const unsigned ls = (unsigned)mpoly[p].loopstart;
    const int n = mpoly[p].totloop;
    if (n == 4) {
      /* mesh_tessellate.c:429-447 */
      out[0].tri[0] = ls; out[0].tri[1] = ls + 1; out[0].tri[2] = ls + 2; out[0].poly = (unsigned)p;
      out[1].tri[0] = ls; out[1].tri[1] = ls + 2; out[1].tri[2] = ls + 3; out[1].poly = (unsigned)p;
      if (quad_needs_flip(mvert[mloop[ls].v].co, mvert[mloop[ls + 1].v].co, mvert[mloop[ls + 2].v].co,
                          mvert[mloop[ls + 3].v].co)) {
        out[0].tri[2] = out[1].tri[2];
        out[1].tri[0] = out[0].tri[1];
      }
      out += 2;
    }
    else {
      /* triangle, or fan for n-gons (the reference runs polyfill there; no config uses n-gons) */
      for (int k = 1; k + 1 < n; k++, out++) {
        out->tri[0] = ls; out->tri[1] = ls + (unsigned)k; out->tri[2] = ls + (unsigned)k + 1; out->poly = (unsigned)p;
      }
    }
  }
}

/* ----------------------------------------------------------------------------------- build */

typedef struct PrimBox {
  float lo[3], hi[3], mid[3];
} PrimBox;

static char g_attach_error[512];

static inline float minf(float a, float b) { return (a < b) ? a : b; }
static inline float maxf(float a, float b) { return (a > b) ? a : b; }

static void bb_clear(BB *bb)
{
  for (int i = 0; i < 3; i++) {
    bb->bmin[i] = FLT_MAX;
    bb->bmax[i] = -FLT_MAX;
  }
}

static void ensure_nodes(PBVH *pbvh, int totnode)
{
  /* growth policy of pbvh.c:2134-2145 (capacity only; indices are what matter) */
  if (totnode > pbvh->node_mem_count) {
    int cap = pbvh->node_mem_count + pbvh->node_mem_count / 3;
    if (cap < totnode) cap = totnode;
    PBVHNode *nn = calloc((size_t)cap, sizeof(PBVHNode));
    if (pbvh->nodes) {
      memcpy(nn, pbvh->nodes, sizeof(PBVHNode) * (size_t)pbvh->totnode);
      free(pbvh->nodes);
    }
    pbvh->nodes = nn;
    pbvh->node_mem_count = cap;
  }
  pbvh->totnode = totnode;
}

/* `rank` = the leaf's place in build order (from 1), owner[v] = the lowest rank among the leaves that use v; stamp / local
 * are the calling thread's own scratch (zeroed: rank 0 means "not seen") */
static void leaf_collect_verts(PBVH *pbvh, PBVHNode *node, int rank, const int *owner, int *stamp, int *local)
{
  /* pbvh.c:2149-2238: a vertex belongs ("unique") to the first leaf, in build order, that uses
   * it; later leaves list it after their unique verts.  Within a leaf, order = first use. */
  const int totface = (int)node->totprim;
  int(*fvi)[3] = malloc(sizeof(int[3]) * (size_t)(totface ? totface : 1));
  unsigned uniq = 0, shared = 0;
  for (int i = 0; i < totface; i++) {
    const MLoopTri *lt = &pbvh->looptri[node->prim_indices[i]];
    for (int j = 0; j < 3; j++) {
      const int v = (int)pbvh->mloop[lt->tri[j]].v;
      if (stamp[v] != rank) {
        stamp[v] = rank;
        if (owner[v] == rank) {
          local[v] = (int)uniq++;
        }
        else {
          local[v] = ~(int)(shared++);
        }
      }
      fvi[i][j] = local[v];
    }
  }
  int *vi = calloc((size_t)(uniq + shared ? uniq + shared : 1), sizeof(int));
  for (int i = 0; i < totface; i++) {
    const MLoopTri *lt = &pbvh->looptri[node->prim_indices[i]];
    for (int j = 0; j < 3; j++) {
      int ndx = fvi[i][j];
      if (ndx < 0) {
        ndx = -ndx + (int)uniq - 1;
        fvi[i][j] = ndx;
      }
      vi[ndx] = (int)pbvh->mloop[lt->tri[j]].v;
    }
  }
  node->uniq_verts = uniq;
  node->face_verts = shared;
  node->vert_indices = vi;
  node->face_vert_indices = (const int(*)[3])fvi;
  node->flag |= PBVH_RebuildDrawBuffers | PBVH_UpdateDrawBuffers | PBVH_UpdateRedraw;
  /* pbvh.c:2188-2208, 2235: a leaf none of whose looptris is visible is fully hidden (respect_hide, pbvh.c:2566);
   * a looptri is hidden when one of its corners is (paint.c:1227-1232) */
  bool has_visible = false;
  for (int i = 0; i < totface && !has_visible; i++) {
    const MLoopTri *lt = &pbvh->looptri[node->prim_indices[i]];
    has_visible = !((pbvh->verts[pbvh->mloop[lt->tri[0]].v].flag | pbvh->verts[pbvh->mloop[lt->tri[1]].v].flag |
                     pbvh->verts[pbvh->mloop[lt->tri[2]].v].flag) & ME_HIDE);
  }
  if (!has_visible) node->flag |= PBVH_FullyHidden;
}

/* paint.c:1234-1241 */
static bool grid_face_hidden(const BLI_bitmap *gh, int gridsize, int x, int y)
{
#define GH_TEST(i) ((gh[(i) >> 5] >> ((i)&31)) & 1u)
  return GH_TEST(y * gridsize + x) || GH_TEST(y * gridsize + x + 1) || GH_TEST((y + 1) * gridsize + x + 1) || GH_TEST((y + 1) * gridsize + x);
#undef GH_TEST
}

/* pbvh.c:2249-2279 */
int BKE_pbvh_count_grid_quads(BLI_bitmap **grid_hidden, const int *grid_indices, int totgrid, int gridsize)
{
  const int gridarea = (gridsize - 1) * (gridsize - 1);
  int totquad = 0;
  for (int i = 0; i < totgrid; i++) {
    const BLI_bitmap *gh = grid_hidden ? grid_hidden[grid_indices[i]] : NULL;
    if (gh) {
      for (int y = 0; y < gridsize - 1; y++) {
        for (int x = 0; x < gridsize - 1; x++) {
          if (!grid_face_hidden(gh, gridsize, x, y)) totquad++;
        }
      }
    }
    else {
      totquad += gridarea;
    }
  }
  return totquad;
}

/* pbvh.c:2058-2066 face_materials_match / grid_materials_match on the prims' material records */
static bool prim_materials_match(const PBVH *pbvh, int prim_a, int prim_b)
{
  if (pbvh->is_grids) {
    const DMFlagMat *a = &pbvh->grid_flag_mats[prim_a], *b = &pbvh->grid_flag_mats[prim_b];
    return (a->flag & ME_SMOOTH) == (b->flag & ME_SMOOTH) && a->mat_nr == b->mat_nr;
  }
  const MPoly *a = &pbvh->mpoly[pbvh->looptri[prim_a].poly], *b = &pbvh->mpoly[pbvh->looptri[prim_b].poly];
  return (a->flag & ME_SMOOTH) == (b->flag & ME_SMOOTH) && a->mat_nr == b->mat_nr;
}

/* pbvh.c:2329-2359 */
static bool leaf_needs_material_split(const PBVH *pbvh, int offset, int count)
{
  if (count <= 1) return false;
  if (pbvh->is_grids ? !pbvh->grid_flag_mats : !pbvh->mpoly) return false;
  const int first = pbvh->prim_indices[offset];
  for (int i = offset + count - 1; i > offset; i--) {
    if (!prim_materials_match(pbvh, first, pbvh->prim_indices[i])) return true;
  }
  return false;
}

/* pbvh.c:2091-2132: Hoare partition, prims of the first prim's material to the left */
static int split_by_material(PBVH *pbvh, int lo, int hi)
{
  int *prims = pbvh->prim_indices;
  const int first = prims[lo];
  int i = lo, j = hi;
  for (;;) {
    while (prim_materials_match(pbvh, first, prims[i])) i++;
    while (!prim_materials_match(pbvh, first, prims[j])) j--;
    if (!(i < j)) return i;
    const int t = prims[i];
    prims[i] = prims[j];
    prims[j] = t;
    i++;
  }
}

static int split_by_centroid(int *prims, int lo, int hi, int axis, float mid, const PrimBox *pb)
{
  /* Hoare partition on centroid < mid (pbvh.c:2070-2088); returns first index of the right part */
  int i = lo, j = hi;
  for (;;) {
    while (pb[prims[i]].mid[axis] < mid) i++;
    while (mid < pb[prims[j]].mid[axis]) j--;
    if (!(i < j)) return i;
    const int t = prims[i];
    prims[i] = prims[j];
    prims[j] = t;
    i++;
  }
}

PBVH *BKE_pbvh_new(void)
{
  PBVH *pbvh = calloc(1, sizeof(PBVH));
  pbvh->leaf_limit = 0; /* the build picks LEAF_LIMIT (pbvh.c:2482) or LEAF_LIMIT / gridsize^2 (pbvh.c:2533) */
  return pbvh;
}

void DUNE_pbvh_mesh_sizes_set(PBVH *pbvh, int totpoly, int totloop)
{
  pbvh->totpoly = totpoly;
  pbvh->totloop = totloop;
}
void DUNE_pbvh_mask_layer_set(PBVH *pbvh, float *vmask) { pbvh->vmask = vmask; }
void DUNE_pbvh_vert_normals_set(PBVH *pbvh, float (*vert_normals)[3])
{
  if (pbvh->owns_normals) free(pbvh->vert_normals);
  pbvh->vert_normals = vert_normals;
  pbvh->owns_normals = false;
}
void DUNE_pbvh_leaf_limit_set(PBVH *pbvh, int leaf_limit) { pbvh->leaf_limit = leaf_limit > 0 ? leaf_limit : 0; }

/* pbvh_build / build_sub (pbvh.c:2372-2450) over the prim boxes `pb`, in two passes.
 *
 * Pass 1 partitions.  What a node does to prim_indices -- box, centroid box, widest axis, Hoare split at the midpoint --
 * touches only its own run of the array, so the two children of a split are independent and run as OpenMP tasks: the
 * root's pass over all prims is serial, the two halves below it run side by side, and so on (about 2 N element visits
 * on the critical path instead of depth x N).  The permutation each node leaves behind is the serial one.
 * Pass 2 is serial and cheap: it numbers the nodes in the order the recursive build_sub creates them (a node's two
 * children take the next two indices when the node is visited, the left subtree is finished before the right one,
 * pbvh.c:2387-2425) and collects the leaves' vertices in that order -- vertex ownership is "first leaf in build
 * order", which only a serial walk defines. */
typedef struct TmpNode {
  struct TmpNode *child[2]; /* NULL, NULL: leaf */
  int offset, count;
  BB vb;
} TmpNode;

#define BUILD_TASK_MIN 32768 /* prims below which a subtree is built by the thread that reached it */

static void box_of_run(const int *prims, const PrimBox *pb, int offset, int count, BB *vb, BB *cb)
{
  bb_clear(vb);
  if (cb) bb_clear(cb);
  for (int i = offset + count - 1; i >= offset; i--) {
    const PrimBox *b = &pb[prims[i]];
    for (int k = 0; k < 3; k++) {
      vb->bmin[k] = minf(vb->bmin[k], b->lo[k]);
      vb->bmax[k] = maxf(vb->bmax[k], b->hi[k]);
      if (cb) {
        cb->bmin[k] = minf(cb->bmin[k], b->mid[k]);
        cb->bmax[k] = maxf(cb->bmax[k], b->mid[k]);
      }
    }
  }
}

static TmpNode *partition_rec(PBVH *pbvh, const PrimBox *pb, int offset, int count, const BB *root_cb)
{
  TmpNode *t = calloc(1, sizeof(TmpNode));
  t->offset = offset;
  t->count = count;
  if (count <= pbvh->leaf_limit) {
    box_of_run(pbvh->prim_indices, pb, offset, count, &t->vb, NULL); /* pbvh.c:2240-2247 */
    if (!leaf_needs_material_split(pbvh, offset, count)) return t;
    /* one draw batch per material and shading mode: split by material instead of by position (pbvh.c:2411-2414) */
    const int end = split_by_material(pbvh, offset, offset + count - 1);
    t->child[0] = partition_rec(pbvh, pb, offset, end - offset, NULL);
    t->child[1] = partition_rec(pbvh, pb, end, offset + count - end, NULL);
    return t;
  }
  BB c;
  box_of_run(pbvh->prim_indices, pb, offset, count, &t->vb, &c);
  if (root_cb) c = *root_cb; /* the root splits on the centroid box BKE_pbvh_build_* accumulated (pbvh.c:2440) */
  /* widest centroid extent (pbvh.c:1995-2016), split at the midpoint (pbvh.c:2402-2410) */
  const float dx = c.bmax[0] - c.bmin[0], dy = c.bmax[1] - c.bmin[1], dz = c.bmax[2] - c.bmin[2];
  int axis;
  if (dx > dy) axis = (dx > dz) ? 0 : 2;
  else axis = (dy > dz) ? 1 : 2;
  const int end = split_by_centroid(pbvh->prim_indices, offset, offset + count - 1, axis, (c.bmax[axis] + c.bmin[axis]) * 0.5f, pb);
  const int nl = end - offset, nr = offset + count - end;
  if (count >= 2 * BUILD_TASK_MIN) {
#pragma omp task shared(t) firstprivate(pbvh, pb, offset, nl)
    t->child[0] = partition_rec(pbvh, pb, offset, nl, NULL);
#pragma omp task shared(t) firstprivate(pbvh, pb, end, nr)
    t->child[1] = partition_rec(pbvh, pb, end, nr, NULL);
#pragma omp taskwait
  }
  else {
    t->child[0] = partition_rec(pbvh, pb, offset, nl, NULL);
    t->child[1] = partition_rec(pbvh, pb, end, nr, NULL);
  }
  return t;
}

/* the mesh leaves in build order (number_rec fills it; their vertex lists follow in a parallel pass) */
typedef struct LeafOrder {
  int *node;
  int count, cap;
} LeafOrder;

static void number_rec(PBVH *pbvh, TmpNode *t, int index, LeafOrder *order)
{
  if (!t->child[0]) {
    PBVHNode *node = &pbvh->nodes[index];
    node->flag |= PBVH_Leaf;
    node->prim_indices = pbvh->prim_indices + t->offset;
    node->totprim = (unsigned)t->count;
    node->vb = t->vb;
    node->orig_vb = t->vb;
    if (pbvh->is_grids) {
      /* build_grid_leaf_node: no vertex list, the node's elements are its grids' elements */
      node->uniq_verts = (unsigned)(t->count * pbvh->gridkey.grid_area);
      node->face_verts = 0;
      /* pbvh.c:2301-2307: BKE_pbvh_node_fully_hidden_set(totquads == 0), BKE_pbvh_node_mark_rebuild_draw */
      if (BKE_pbvh_count_grid_quads(pbvh->grid_hidden, node->prim_indices, (int)node->totprim, pbvh->gridkey.grid_size) == 0)
        node->flag |= PBVH_FullyHidden;
      node->flag |= PBVH_RebuildDrawBuffers | PBVH_UpdateDrawBuffers | PBVH_UpdateRedraw;
    }
    else {
      if (order->count == order->cap) {
        order->cap = order->cap ? 2 * order->cap : 1024;
        order->node = realloc(order->node, sizeof(int) * (size_t)order->cap);
      }
      order->node[order->count++] = index;
    }
    free(t);
    return;
  }
  const int child = pbvh->totnode;
  ensure_nodes(pbvh, pbvh->totnode + 2);
  PBVHNode *node = &pbvh->nodes[index]; /* after ensure_nodes: the array may have moved */
  node->children_offset = child;
  node->vb = t->vb;
  node->orig_vb = t->vb;
  number_rec(pbvh, t->child[0], child, order);
  number_rec(pbvh, t->child[1], child + 1, order);
  free(t);
}

static void build_tree(PBVH *pbvh, const PrimBox *pb, int nprims, const BB *cb)
{
  TmpNode *root = NULL;
  struct timespec t0, t1, t2;
  clock_gettime(CLOCK_MONOTONIC, &t0);
#pragma omp parallel
#pragma omp single
  root = partition_rec(pbvh, pb, 0, nprims, cb);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  /* pass 2a, serial and cheap: the nodes numbered in build_sub's order, the mesh leaves listed in that order */
  LeafOrder order = {NULL, 0, 0};
  number_rec(pbvh, root, 0, &order);
  if (order.count) {
    /* pass 2b: a vertex is unique in the FIRST leaf, in build order, that uses it (map_insert_vert + vert_bitmap,
     * pbvh.c:2149-2171) = the lowest rank among its leaves: an order-free minimum, then every leaf collects its verts on
     * its own (a thread's scratch pages are only touched where its leaves -- a contiguous, spatially coherent run -- reach) */
    const int V = pbvh->totvert;
    struct timespec ta, tb, tc; clock_gettime(CLOCK_MONOTONIC, &ta);
    int *owner = malloc(sizeof(int) * (size_t)(V ? V : 1));
#pragma omp parallel for schedule(static)
    for (int v = 0; v < V; v++) owner[v] = 0x7fffffff;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < order.count; k++) {
      const PBVHNode *node = &pbvh->nodes[order.node[k]];
      const int rank = k + 1;
      for (unsigned i = 0; i < node->totprim; i++) {
        const MLoopTri *lt = &pbvh->looptri[node->prim_indices[i]];
        for (int j = 0; j < 3; j++) {
          int *o = &owner[pbvh->mloop[lt->tri[j]].v];
          int cur = __atomic_load_n(o, __ATOMIC_RELAXED);
          while (rank < cur && !__atomic_compare_exchange_n(o, &cur, rank, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
          }
        }
      }
    }
    clock_gettime(CLOCK_MONOTONIC, &tb);
#pragma omp parallel
    {
      int *stamp = calloc((size_t)(V ? V : 1), sizeof(int)), *local = calloc((size_t)(V ? V : 1), sizeof(int));
#pragma omp for schedule(static)
      for (int k = 0; k < order.count; k++) leaf_collect_verts(pbvh, &pbvh->nodes[order.node[k]], k + 1, owner, stamp, local);
      free(stamp);
      free(local);
    }
    free(owner);
    clock_gettime(CLOCK_MONOTONIC, &tc);
    if (getenv("DUNE_PBVH_TIMING")) fprintf(stderr, "owner %.3f collect %.3f\n", (tb.tv_sec - ta.tv_sec) + 1e-9 * (tb.tv_nsec - ta.tv_nsec), (tc.tv_sec - tb.tv_sec) + 1e-9 * (tc.tv_nsec - tb.tv_nsec));
  }
  free(order.node);
  clock_gettime(CLOCK_MONOTONIC, &t2);
  if (getenv("DUNE_PBVH_TIMING")) fprintf(stderr, "partition %.3f number %.3f\n", (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec), (t2.tv_sec - t1.tv_sec) + 1e-9 * (t2.tv_nsec - t1.tv_nsec));
}

void BKE_pbvh_build_mesh(PBVH *pbvh, struct Mesh *mesh, const MPoly *mpoly, const MLoop *mloop, MVert *verts,
                         int totvert, struct CustomData *vdata, struct CustomData *ldata, struct CustomData *pdata,
                         const MLoopTri *looptri, int looptri_num)
{
  (void)mesh; (void)vdata; (void)ldata; (void)pdata;
  pbvh->mpoly = mpoly;
  pbvh->mloop = mloop;
  pbvh->looptri = looptri;
  pbvh->verts = verts;
  pbvh->totvert = totvert;
  if (!pbvh->vert_normals) {
    pbvh->vert_normals = calloc((size_t)(totvert ? totvert : 1), sizeof(float[3]));
    pbvh->owns_normals = true;
  }
  pbvh->vert_bitmap = calloc((size_t)totvert / 32 + 1, sizeof(unsigned));
  if (pbvh->leaf_limit <= 0) pbvh->leaf_limit = LEAF_LIMIT;
  if (!looptri_num) return;

  /* per-looptri box and centroid (pbvh.c:2490-2504): independent per looptri; the centroid box is a min / max, exact
   * in any order */
  PrimBox *pb = malloc(sizeof(PrimBox) * (size_t)looptri_num);
  BB cb;
  bb_clear(&cb);
#pragma omp parallel
  {
    BB mine;
    bb_clear(&mine);
#pragma omp for schedule(static) nowait
    for (int i = 0; i < looptri_num; i++) {
      PrimBox *b = &pb[i];
      for (int k = 0; k < 3; k++) {
        b->lo[k] = FLT_MAX;
        b->hi[k] = -FLT_MAX;
      }
      for (int j = 0; j < 3; j++) {
        const float *co = verts[mloop[looptri[i].tri[j]].v].co;
        for (int k = 0; k < 3; k++) {
          b->lo[k] = minf(b->lo[k], co[k]);
          b->hi[k] = maxf(b->hi[k], co[k]);
        }
      }
      for (int k = 0; k < 3; k++) {
        b->mid[k] = (b->lo[k] + b->hi[k]) * 0.5f;
        mine.bmin[k] = minf(mine.bmin[k], b->mid[k]);
        mine.bmax[k] = maxf(mine.bmax[k], b->mid[k]);
      }
    }
#pragma omp critical
    for (int k = 0; k < 3; k++) {
      cb.bmin[k] = minf(cb.bmin[k], mine.bmin[k]);
      cb.bmax[k] = maxf(cb.bmax[k], mine.bmax[k]);
    }
  }

  pbvh->totprim = looptri_num;
  pbvh->prim_indices = malloc(sizeof(int) * (size_t)looptri_num);
  for (int i = 0; i < looptri_num; i++) pbvh->prim_indices[i] = i;
  pbvh->node_mem_count = 0;
  pbvh->totnode = 0;
  ensure_nodes(pbvh, 100);
  pbvh->totnode = 1;

  build_tree(pbvh, pb, looptri_num, &cb);
  free(pb);
  memset(pbvh->vert_bitmap, 0, sizeof(unsigned) * ((size_t)totvert / 32 + 1)); /* pbvh.c:2512-2513 */
}


/* ------------------------------------------------------------------------------ multires grids */

static inline float *ccg_elem_co(const CCGKey *key, CCGElem *grid, int j)
{
  return (float *)((unsigned char *)grid + (size_t)key->elem_size * (size_t)j);
}

/* pbvh.c:2516-2561 BKE_pbvh_build_grids: per-grid box over all its elements, then the same tree build
 * as for looptris; a grid leaf keeps no vertex list (its elements are its grids' elements) */
void BKE_pbvh_build_grids(PBVH *pbvh, CCGElem **grids, int totgrid, CCGKey *key, void **gridfaces, DMFlagMat *flagmats,
                          BLI_bitmap **grid_hidden)
{
  const int gridsize = key->grid_size;
  pbvh->is_grids = 1;
  pbvh->grids = grids;
  pbvh->gridfaces = gridfaces;
  pbvh->grid_flag_mats = flagmats;
  pbvh->totgrid = totgrid;
  pbvh->gridkey = *key;
  pbvh->grid_hidden = grid_hidden;
  if (pbvh->leaf_limit <= 0) {
    pbvh->leaf_limit = LEAF_LIMIT / (gridsize * gridsize);
    if (pbvh->leaf_limit < 1) pbvh->leaf_limit = 1;
  }
  pbvh->totvert = totgrid * gridsize * gridsize;
  if (!totgrid) return;

  PrimBox *pb = malloc(sizeof(PrimBox) * (size_t)totgrid);
  BB cb;
  bb_clear(&cb);
#pragma omp parallel
  {
    BB mine;
    bb_clear(&mine);
#pragma omp for schedule(static) nowait
    for (int i = 0; i < totgrid; i++) {
      PrimBox *b = &pb[i];
      for (int k = 0; k < 3; k++) {
        b->lo[k] = FLT_MAX;
        b->hi[k] = -FLT_MAX;
      }
      for (int j = 0; j < gridsize * gridsize; j++) {
        const float *co = ccg_elem_co(key, grids[i], j);
        for (int k = 0; k < 3; k++) {
          b->lo[k] = minf(b->lo[k], co[k]);
          b->hi[k] = maxf(b->hi[k], co[k]);
        }
      }
      for (int k = 0; k < 3; k++) {
        b->mid[k] = (b->lo[k] + b->hi[k]) * 0.5f;
        mine.bmin[k] = minf(mine.bmin[k], b->mid[k]);
        mine.bmax[k] = maxf(mine.bmax[k], b->mid[k]);
      }
    }
#pragma omp critical
    for (int k = 0; k < 3; k++) {
      cb.bmin[k] = minf(cb.bmin[k], mine.bmin[k]);
      cb.bmax[k] = maxf(cb.bmax[k], mine.bmax[k]);
    }
  }
  pbvh->totprim = totgrid;
  pbvh->prim_indices = malloc(sizeof(int) * (size_t)totgrid);
  for (int i = 0; i < totgrid; i++) pbvh->prim_indices[i] = i;
  pbvh->node_mem_count = 0;
  pbvh->totnode = 0;
  ensure_nodes(pbvh, 100);
  pbvh->totnode = 1;
  build_tree(pbvh, pb, totgrid, &cb);
  free(pb);
}

void BKE_pbvh_node_get_grids(PBVH *pbvh, PBVHNode *node, int **r_grid_indices, int *r_totgrid, int *r_maxgrid, int *r_gridsize,
                             CCGElem ***r_griddata)
{
  /* pbvh.c:3770-3800, PBVH_GRIDS case */
  if (r_grid_indices) *r_grid_indices = node->prim_indices;
  if (r_totgrid) *r_totgrid = (int)node->totprim;
  if (r_maxgrid) *r_maxgrid = pbvh->totgrid;
  if (r_gridsize) *r_gridsize = pbvh->gridkey.grid_size;
  if (r_griddata) *r_griddata = pbvh->grids;
}

void BKE_subdiv_ccg_key_top_level(CCGKey *key, const SubdivCCG *subdiv_ccg)
{
  /* subdiv_ccg.c:633-651 */
  key->level = subdiv_ccg->level;
  key->elem_size = subdiv_ccg->grid_element_size;
  key->grid_size = subdiv_ccg->grid_size;
  key->grid_area = key->grid_size * key->grid_size;
  key->grid_bytes = key->elem_size * key->grid_area;
  key->normal_offset = subdiv_ccg->normal_offset;
  key->mask_offset = subdiv_ccg->mask_offset;
  key->has_normals = subdiv_ccg->has_normal;
  key->has_mask = subdiv_ccg->has_mask;
}

SubdivCCG *DUNE_subdiv_ccg_from_tables(int level, int num_grids, const float *co, const float *no, const float *mask,
                                       int num_faces, const int *face_start_grid, const int *face_num_grids, int num_edges,
                                       const int *edge_offsets, const int *edge_elems, int num_vertices,
                                       const int *vert_offsets, const int *vert_elems, const int *grid_edge,
                                       const int *grid_vertex)
{
  SubdivCCG *ccg = calloc(1, sizeof(SubdivCCG));
  const int gs = (1 << (level - 1)) + 1, area = gs * gs; /* BKE_subdiv_grid_size_from_level */
  ccg->level = level;
  ccg->grid_size = gs;
  ccg->num_grids = num_grids;
  /* layers: co, then mask, then normals (subdiv_ccg.c:62-90); normals are always kept here */
  int off = (int)sizeof(float[3]);
  ccg->has_mask = mask != NULL;
  ccg->mask_offset = ccg->has_mask ? off : -1;
  if (ccg->has_mask) off += (int)sizeof(float);
  ccg->has_normal = true;
  ccg->normal_offset = off;
  off += (int)sizeof(float[3]);
  ccg->grid_element_size = off;
  ccg->grids = calloc((size_t)(num_grids ? num_grids : 1), sizeof(CCGElem *));
  ccg->grids_storage = calloc((size_t)num_grids * (size_t)area, (size_t)off);
  for (int g = 0; g < num_grids; g++) {
    ccg->grids[g] = (CCGElem *)(ccg->grids_storage + (size_t)g * (size_t)area * (size_t)off);
    for (int j = 0; j < area; j++) {
      unsigned char *e = (unsigned char *)ccg->grids[g] + (size_t)j * (size_t)off;
      const size_t idx = (size_t)g * (size_t)area + (size_t)j;
      memcpy(e, co + 3 * idx, sizeof(float[3]));
      if (mask) memcpy(e + ccg->mask_offset, mask + idx, sizeof(float));
      if (no) memcpy(e + ccg->normal_offset, no + 3 * idx, sizeof(float[3]));
    }
  }
  ccg->num_faces = num_faces;
  ccg->faces = calloc((size_t)(num_faces ? num_faces : 1), sizeof(SubdivCCGFace));
  ccg->grid_faces = calloc((size_t)(num_grids ? num_grids : 1), sizeof(SubdivCCGFace *));
  for (int f = 0; f < num_faces; f++) {
    ccg->faces[f].start_grid_index = face_start_grid[f];
    ccg->faces[f].num_grids = face_num_grids[f];
    for (int c = 0; c < face_num_grids[f]; c++) ccg->grid_faces[face_start_grid[f] + c] = &ccg->faces[f];
  }
  ccg->num_adjacent_edges = num_edges;
  ccg->adjacent_edges = calloc((size_t)(num_edges ? num_edges : 1), sizeof(SubdivCCGAdjacentEdge));
  for (int e = 0; e < num_edges; e++) {
    SubdivCCGAdjacentEdge *ae = &ccg->adjacent_edges[e];
    ae->num_adjacent_faces = edge_offsets[e + 1] - edge_offsets[e];
    ae->boundary_coords = calloc((size_t)(ae->num_adjacent_faces ? ae->num_adjacent_faces : 1), sizeof(SubdivCCGCoord *));
    for (int f = 0; f < ae->num_adjacent_faces; f++) {
      ae->boundary_coords[f] = malloc(sizeof(SubdivCCGCoord) * (size_t)(2 * gs));
      const int *row = edge_elems + (size_t)(edge_offsets[e] + f) * (size_t)(2 * gs);
      for (int i = 0; i < 2 * gs; i++) {
        SubdivCCGCoord *c = &ae->boundary_coords[f][i];
        c->grid_index = row[i] / area;
        c->y = (short)((row[i] % area) / gs);
        c->x = (short)((row[i] % area) % gs);
      }
    }
  }
  ccg->num_adjacent_vertices = num_vertices;
  ccg->adjacent_vertices = calloc((size_t)(num_vertices ? num_vertices : 1), sizeof(SubdivCCGAdjacentVertex));
  for (int v = 0; v < num_vertices; v++) {
    SubdivCCGAdjacentVertex *av = &ccg->adjacent_vertices[v];
    av->num_adjacent_faces = vert_offsets[v + 1] - vert_offsets[v];
    av->corner_coords = malloc(sizeof(SubdivCCGCoord) * (size_t)(av->num_adjacent_faces ? av->num_adjacent_faces : 1));
    for (int f = 0; f < av->num_adjacent_faces; f++) {
      const int el = vert_elems[vert_offsets[v] + f];
      av->corner_coords[f].grid_index = el / area;
      av->corner_coords[f].y = (short)((el % area) / gs);
      av->corner_coords[f].x = (short)((el % area) % gs);
    }
  }
  ccg->grid_edge = malloc(sizeof(int) * (size_t)(num_grids ? num_grids : 1));
  ccg->grid_vertex = malloc(sizeof(int) * (size_t)(num_grids ? num_grids : 1));
  memcpy(ccg->grid_edge, grid_edge, sizeof(int) * (size_t)num_grids);
  memcpy(ccg->grid_vertex, grid_vertex, sizeof(int) * (size_t)num_grids);
  return ccg;
}

void DUNE_subdiv_ccg_topology_set(SubdivCCG *ccg, const int *edge_vertices, const int *vertex_edge_offsets, const int *vertex_edges)
{
  free(ccg->edge_vertices); free(ccg->vertex_edge_offsets); free(ccg->vertex_edges);
  const int ne = ccg->num_adjacent_edges, nv = ccg->num_adjacent_vertices;
  ccg->edge_vertices = malloc(sizeof(int[2]) * (size_t)(ne ? ne : 1));
  memcpy(ccg->edge_vertices, edge_vertices, sizeof(int[2]) * (size_t)ne);
  ccg->vertex_edge_offsets = malloc(sizeof(int) * (size_t)(nv + 1));
  memcpy(ccg->vertex_edge_offsets, vertex_edge_offsets, sizeof(int) * (size_t)(nv + 1));
  ccg->vertex_edges = malloc(sizeof(int) * (size_t)(vertex_edge_offsets[nv] + 1));
  memcpy(ccg->vertex_edges, vertex_edges, sizeof(int) * (size_t)vertex_edge_offsets[nv]);
}

/* ---- element neighbours: subdiv_ccg.c:1365-1909.  An element is classified by where it sits in its grid --
 * interior, on a boundary shared with the next / previous grid of the same face (x == 0 or y == 0), on a coarse edge
 * (x or y == grid_size - 1), at the face centre (0, 0) or at a coarse vertex (both maximal) -- and each class has its own
 * fixed neighbour order; `include_duplicates` appends the other copies of the element itself. ---- */
static SubdivCCGCoord ccg_coord(int grid, int x, int y)
{
  SubdivCCGCoord c;
  c.grid_index = grid;
  c.x = (short)x;
  c.y = (short)y;
  return c;
}

static void neighbors_reserve(SubdivCCGNeighbors *nb, int num_unique, int num_duplicates)
{
  nb->size = num_unique + num_duplicates;
  nb->num_duplicates = num_duplicates;
  nb->coords = nb->size < (int)(sizeof(nb->coords_fixed) / sizeof(nb->coords_fixed[0])) ?
                   nb->coords_fixed :
                   MEM_mallocN(sizeof(SubdivCCGCoord) * (size_t)nb->size, "SubdivCCGNeighbors.coords");
}

/* one step off the grid's rim (subdiv_ccg.c:1448-1471) */
static SubdivCCGCoord step_off_rim(const SubdivCCG *ccg, SubdivCCGCoord c)
{
  const int last = ccg->grid_size - 1;
  if (c.x == last) c.x--;
  else if (c.y == last) c.y--;
  else if (c.x == 0) c.x++;
  else c.y++;
  return c;
}

/* the element lies on a coarse edge and is not a coarse vertex (subdiv_ccg.c:1612-1772) */
static void neighbors_coarse_edge(const SubdivCCG *ccg, const SubdivCCGCoord *co, bool dups, SubdivCCGNeighbors *nb)
{
  const int gs = ccg->grid_size, last = gs - 1;
  const SubdivCCGFace *face = ccg->grid_faces[co->grid_index];
  const int corner = co->grid_index - face->start_grid_index;
  const bool on_x = co->x == last;
  const bool at_grid_corner = (co->x == 0 || co->x == last) && (co->y == 0 || co->y == last);
  /* the face's edge leaving this corner, or the one arriving at it */
  const int edge = on_x ? ccg->grid_edge[co->grid_index] :
                          ccg->grid_edge[face->start_grid_index + (corner == 0 ? face->num_grids - 1 : corner - 1)];
  const SubdivCCGAdjacentEdge *ae = &ccg->adjacent_edges[edge];
  const int nf = ae->num_adjacent_faces;
  /* position along the 2 * grid_size boundary points, flipped when the edge runs against this face's winding */
  int pt = on_x ? gs - co->y - 1 : gs + co->x;
  if (ccg->grid_vertex[co->grid_index] != ccg->edge_vertices[edge][on_x ? 0 : 1]) pt = 2 * gs - pt - 1;
  /* the two middle points of the edge are one vertex (two grid corners): step over the twin */
  const int pt_next = pt == gs - 1 ? pt + 2 : pt + 1;
  const int pt_prev = pt == gs ? pt - 2 : pt - 1;
  const int pt_twin = pt == gs ? pt - 1 : pt + 1;
  neighbors_reserve(nb, nf + 2, dups ? (nf - 1) + (at_grid_corner ? nf : 0) : 0);
  int dup_at = nf + 2;
  for (int i = 0; i < nf; i++) {
    const SubdivCCGCoord *row = ae->boundary_coords[i];
    nb->coords[2 + i] = step_off_rim(ccg, row[pt]);
    if (row[pt].grid_index == co->grid_index) {
      nb->coords[0] = row[pt_prev];
      nb->coords[1] = row[pt_next];
    }
    else if (dups) {
      nb->coords[dup_at++] = row[pt];
    }
    if (dups && at_grid_corner) nb->coords[dup_at++] = row[pt_twin];
  }
}

void BKE_subdiv_ccg_neighbor_coords_get(const SubdivCCG *ccg, const SubdivCCGCoord *co, const bool dups, SubdivCCGNeighbors *nb)
{
  const int gs = ccg->grid_size, last = gs - 1;
  const int g = co->grid_index, x = co->x, y = co->y;
  const SubdivCCGFace *face = ccg->grid_faces[g];
  const int n = face->num_grids, first = face->start_grid_index, corner = g - first;
  if (x > 0 && y > 0 && x < last && y < last) { /* interior: previous / next row, previous / next column */
    neighbors_reserve(nb, 4, 0);
    nb->coords[0] = ccg_coord(g, x, y - 1);
    nb->coords[1] = ccg_coord(g, x, y + 1);
    nb->coords[2] = ccg_coord(g, x - 1, y);
    nb->coords[3] = ccg_coord(g, x + 1, y);
    return;
  }
  if (x == 0 && y == 0) { /* the face centre: one step along every grid of the face */
    neighbors_reserve(nb, n, dups ? n - 1 : 0);
    int dup_at = n;
    for (int c = 0; c < n; c++) {
      nb->coords[c] = ccg_coord(first + c, 1, 0);
      if (dups && first + c != g) nb->coords[dup_at++] = ccg_coord(first + c, 0, 0);
    }
    return;
  }
  if (x == last && y == last) { /* a coarse vertex: the second point of every edge around it, on the edge's first face */
    const int v = ccg->grid_vertex[g];
    const int *edges = ccg->vertex_edges + ccg->vertex_edge_offsets[v];
    const int ne = ccg->vertex_edge_offsets[v + 1] - ccg->vertex_edge_offsets[v];
    const SubdivCCGAdjacentVertex *av = &ccg->adjacent_vertices[v];
    neighbors_reserve(nb, ne, dups ? av->num_adjacent_faces - 1 : 0);
    for (int i = 0; i < ne; i++) {
      const int pt = ccg->edge_vertices[edges[i]][0] == v ? 1 : 2 * gs - 2;
      nb->coords[i] = ccg->adjacent_edges[edges[i]].boundary_coords[0][pt];
    }
    if (dups) {
      int dup_at = ne;
      for (int i = 0; i < av->num_adjacent_faces; i++) {
        if (av->corner_coords[i].grid_index != g) nb->coords[dup_at++] = av->corner_coords[i];
      }
    }
    return;
  }
  if (x == last || y == last) {
    neighbors_coarse_edge(ccg, co, dups, nb);
    return;
  }
  /* on the boundary with a sibling grid: column 0 is row 0 of the previous grid, row 0 is column 0 of the next */
  neighbors_reserve(nb, 4, dups ? 1 : 0);
  if (x == 0) {
    const int prev = first + (corner == 0 ? n - 1 : corner - 1);
    nb->coords[0] = ccg_coord(g, x, y - 1);
    nb->coords[1] = ccg_coord(g, x, y + 1);
    nb->coords[2] = ccg_coord(g, x + 1, y);
    nb->coords[3] = ccg_coord(prev, y, 1);
    if (dups) nb->coords[4] = ccg_coord(prev, y, 0);
  }
  else {
    const int next = first + (corner + 1 == n ? 0 : corner + 1);
    nb->coords[0] = ccg_coord(g, x - 1, y);
    nb->coords[1] = ccg_coord(g, x + 1, y);
    nb->coords[2] = ccg_coord(g, x, y + 1);
    nb->coords[3] = ccg_coord(next, 1, x);
    if (dups) nb->coords[4] = ccg_coord(next, 0, x);
  }
}

SubdivCCGAdjacencyType BKE_subdiv_ccg_coarse_mesh_adjacency_info_get(const SubdivCCG *ccg, const SubdivCCGCoord *co, int *r_v1, int *r_v2)
{
  const int last = ccg->grid_size - 1;
  if (co->x != last && co->y != last) return SUBDIV_CCG_ADJACENT_NONE; /* interior, face centre, sibling boundary */
  const SubdivCCGFace *face = ccg->grid_faces[co->grid_index];
  const int n = face->num_grids, first = face->start_grid_index, corner = co->grid_index - first;
  *r_v1 = *r_v2 = ccg->grid_vertex[co->grid_index];
  if (co->x == last && co->y == last) return SUBDIV_CCG_ADJACENT_VERTEX;
  /* the other end of the coarse edge: the next loop's vertex along x, the previous loop's along y */
  if (co->x == last) *r_v2 = ccg->grid_vertex[first + (corner + 1) % n];
  if (co->y == last) *r_v2 = ccg->grid_vertex[first + (corner + n - 1) % n];
  return SUBDIV_CCG_ADJACENT_EDGE;
}

void DUNE_subdiv_ccg_free(SubdivCCG *ccg)
{
  if (!ccg) return;
  for (int e = 0; e < ccg->num_adjacent_edges; e++) {
    for (int f = 0; f < ccg->adjacent_edges[e].num_adjacent_faces; f++) free(ccg->adjacent_edges[e].boundary_coords[f]);
    free(ccg->adjacent_edges[e].boundary_coords);
  }
  for (int v = 0; v < ccg->num_adjacent_vertices; v++) free(ccg->adjacent_vertices[v].corner_coords);
  free(ccg->adjacent_edges); free(ccg->adjacent_vertices); free(ccg->faces); free(ccg->grid_faces);
  free(ccg->grids); free(ccg->grids_storage); free(ccg->grid_edge); free(ccg->grid_vertex);
  free(ccg->edge_vertices); free(ccg->vertex_edge_offsets); free(ccg->vertex_edges);
  free(ccg);
}

/* the CCG's elements and adjacency as the flat tables of DscGridsDesc, the grids PBVH as DscPbvhDesc */
static int push_leaf_shading(PBVH *pbvh, DscContext *ctx);
int DUNE_pbvh_device_attach_grids(PBVH *pbvh, SubdivCCG *ccg, int device)
{
  return DUNE_pbvh_device_attach_grids_dist(pbvh, ccg, device, 1, 0, NULL);
}

int DUNE_pbvh_device_attach_grids_dist(PBVH *pbvh, SubdivCCG *ccg, int device, int world, int rank, const char *nccl_id)
{
  g_attach_error[0] = 0;
  if (!pbvh || !pbvh->nodes || !pbvh->is_grids || !ccg) return DSC_ERR_INVALID;
  if (pbvh->device) return DSC_OK;
  DscContext *ctx = NULL;
  int r = dsc_ctx_create(device, &ctx);
  if (r != DSC_OK) {
    snprintf(g_attach_error, sizeof(g_attach_error), "%s", dsc_last_error(NULL));
    return r;
  }
  pbvh->dist_world = world > 1 ? world : 0;
  if (world > 1) {
    r = dsc_dist_init(ctx, world, rank, nccl_id);
    if (r != DSC_OK) {
      snprintf(g_attach_error, sizeof(g_attach_error), "%s", dsc_last_error(ctx));
      dsc_ctx_destroy(ctx);
      return r;
    }
  }
  const CCGKey *key = &pbvh->gridkey;
  const int gs = key->grid_size, area = key->grid_area, G = pbvh->totgrid, N = pbvh->totnode;
  const size_t E = (size_t)G * (size_t)area;
  float *co = malloc(sizeof(float[3]) * E), *no = malloc(sizeof(float[3]) * E);
  float *mask = key->has_mask ? malloc(sizeof(float) * E) : NULL;
  bool have_no = false;
  for (int g = 0; g < G; g++) {
    for (int j = 0; j < area; j++) {
      const unsigned char *e = (const unsigned char *)pbvh->grids[g] + (size_t)key->elem_size * (size_t)j;
      const size_t idx = (size_t)g * (size_t)area + (size_t)j;
      memcpy(co + 3 * idx, e, sizeof(float[3]));
      if (key->has_normals) {
        memcpy(no + 3 * idx, e + key->normal_offset, sizeof(float[3]));
        have_no = have_no || no[3 * idx] != 0.0f || no[3 * idx + 1] != 0.0f || no[3 * idx + 2] != 0.0f;
      }
      if (mask) memcpy(mask + idx, e + key->mask_offset, sizeof(float));
    }
  }
  int *face_start = malloc(sizeof(int) * (size_t)(ccg->num_faces + 1)), *face_num = malloc(sizeof(int) * (size_t)(ccg->num_faces + 1));
  for (int f = 0; f < ccg->num_faces; f++) {
    face_start[f] = ccg->faces[f].start_grid_index;
    face_num[f] = ccg->faces[f].num_grids;
  }
  int *edge_off = malloc(sizeof(int) * (size_t)(ccg->num_adjacent_edges + 1));
  edge_off[0] = 0;
  for (int e = 0; e < ccg->num_adjacent_edges; e++) edge_off[e + 1] = edge_off[e] + ccg->adjacent_edges[e].num_adjacent_faces;
  int *edge_elems = malloc(sizeof(int) * ((size_t)edge_off[ccg->num_adjacent_edges] * 2 * (size_t)gs + 1));
  for (int e = 0; e < ccg->num_adjacent_edges; e++) {
    for (int f = 0; f < ccg->adjacent_edges[e].num_adjacent_faces; f++) {
      const SubdivCCGCoord *row = ccg->adjacent_edges[e].boundary_coords[f];
      int *out = edge_elems + (size_t)(edge_off[e] + f) * 2 * (size_t)gs;
      for (int i = 0; i < 2 * gs; i++) out[i] = row[i].grid_index * area + row[i].y * gs + row[i].x;
    }
  }
  int *vert_off = malloc(sizeof(int) * (size_t)(ccg->num_adjacent_vertices + 1));
  vert_off[0] = 0;
  for (int v = 0; v < ccg->num_adjacent_vertices; v++) vert_off[v + 1] = vert_off[v] + ccg->adjacent_vertices[v].num_adjacent_faces;
  int *vert_elems = malloc(sizeof(int) * ((size_t)vert_off[ccg->num_adjacent_vertices] + 1));
  for (int v = 0; v < ccg->num_adjacent_vertices; v++) {
    for (int f = 0; f < ccg->adjacent_vertices[v].num_adjacent_faces; f++) {
      const SubdivCCGCoord *c = &ccg->adjacent_vertices[v].corner_coords[f];
      vert_elems[vert_off[v] + f] = c->grid_index * area + c->y * gs + c->x;
    }
  }
  DscGridsDesc gd = {0};
  gd.totgrid = G;
  gd.grid_size = gs;
  gd.co = co;
  gd.no = have_no ? no : NULL;
  gd.mask = mask;
  gd.totface = ccg->num_faces;
  gd.face_start_grid = face_start;
  gd.face_num_grids = face_num;
  gd.totedge = ccg->num_adjacent_edges;
  gd.edge_offsets = edge_off;
  gd.edge_elems = edge_elems;
  gd.totcvert = ccg->num_adjacent_vertices;
  gd.cvert_offsets = vert_off;
  gd.cvert_elems = vert_elems;
  gd.grid_edge = ccg->grid_edge;
  gd.grid_cvert = ccg->grid_vertex;
  int *rim_nb = NULL;
  unsigned char *rim_bnd = NULL;
  if (ccg->edge_vertices) {
    /* what the smooth brush's neighbour iterator would ask per element per iteration, asked once: the rim elements of
     * every grid through BKE_subdiv_ccg_neighbor_coords_get, flattened to element indices */
    const int rim = 4 * gs - 4;
    unsigned char *vbnd = calloc((size_t)ccg->num_adjacent_vertices + 1, 1); /* boundary vertices of the base mesh */
    for (int e = 0; e < ccg->num_adjacent_edges; e++) {
      if (ccg->adjacent_edges[e].num_adjacent_faces < 2) {
        vbnd[ccg->edge_vertices[e][0]] = 1;
        vbnd[ccg->edge_vertices[e][1]] = 1;
      }
    }
    int width = 4;
    for (int f = 0; f < ccg->num_faces; f++) width = ccg->faces[f].num_grids > width ? ccg->faces[f].num_grids : width;
    for (int e = 0; e < ccg->num_adjacent_edges; e++) {
      const int k = ccg->adjacent_edges[e].num_adjacent_faces + 2;
      width = k > width ? k : width;
    }
    for (int v = 0; v < ccg->num_adjacent_vertices; v++) {
      const int k = ccg->vertex_edge_offsets[v + 1] - ccg->vertex_edge_offsets[v];
      width = k > width ? k : width;
    }
    rim_nb = malloc(sizeof(int) * (size_t)G * (size_t)rim * (size_t)width);
    rim_bnd = calloc((size_t)G * (size_t)rim, 1);
    SubdivCCGNeighbors *nb = malloc(sizeof(SubdivCCGNeighbors));
    for (int g = 0; g < G; g++) {
      for (int b = 0; b < rim; b++) {
        SubdivCCGCoord c;
        c.grid_index = g;
        if (b < gs) { c.x = (short)b; c.y = 0; }
        else if (b < 2 * gs) { c.x = (short)(b - gs); c.y = (short)(gs - 1); }
        else if (b < 3 * gs - 2) { c.x = 0; c.y = (short)(b - 2 * gs + 1); }
        else { c.x = (short)(gs - 1); c.y = (short)(b - (3 * gs - 2) + 1); }
        BKE_subdiv_ccg_neighbor_coords_get(ccg, &c, false, nb);
        int *row = rim_nb + ((size_t)g * (size_t)rim + (size_t)b) * (size_t)width;
        for (int i = 0; i < width; i++) {
          row[i] = i < nb->size ? nb->coords[i].grid_index * area + nb->coords[i].y * gs + nb->coords[i].x : -1;
        }
        if (nb->coords != nb->coords_fixed) MEM_freeN(nb->coords);
        int v1 = 0, v2 = 0;
        switch (BKE_subdiv_ccg_coarse_mesh_adjacency_info_get(ccg, &c, &v1, &v2)) {
          case SUBDIV_CCG_ADJACENT_VERTEX: rim_bnd[(size_t)g * (size_t)rim + (size_t)b] = vbnd[v1]; break;
          case SUBDIV_CCG_ADJACENT_EDGE: rim_bnd[(size_t)g * (size_t)rim + (size_t)b] = vbnd[v1] && vbnd[v2]; break;
          default: break;
        }
      }
    }
    free(nb);
    free(vbnd);
    gd.rim_width = width;
    gd.rim_neighbors = rim_nb;
    gd.rim_boundary = rim_bnd;
  }
  /* grid_hidden (pbvh.c:2527): one BLI_bitmap per grid, NULL for a grid with nothing hidden */
  unsigned char *elem_hidden = NULL;
  if (pbvh->grid_hidden) {
    for (int g = 0; g < G; g++) {
      const BLI_bitmap *gh = pbvh->grid_hidden[g];
      if (!gh) continue;
      if (!elem_hidden) elem_hidden = calloc((size_t)G * (size_t)area, 1);
      for (int i = 0; i < area; i++) elem_hidden[(size_t)g * area + i] = ((gh[i >> 5] >> (i & 31)) & 1u) ? 1 : 0;
    }
  }
  gd.hidden = elem_hidden;
  r = dsc_grids_upload(ctx, &gd);
  free(elem_hidden);
  free(rim_nb);
  free(rim_bnd);
  if (r == DSC_OK && (pbvh->want_draw_buffers & 1)) r = dsc_draw_enable(ctx);
  if (r == DSC_OK && (pbvh->want_draw_buffers & 2)) r = dsc_raycast_enable(ctx);

  float *bb = malloc(sizeof(float[6]) * (size_t)N), *obb = malloc(sizeof(float[6]) * (size_t)N);
  int *child = malloc(sizeof(int) * (size_t)N), *flag = malloc(sizeof(int) * (size_t)N), *prim_off = malloc(sizeof(int) * (size_t)N);
  int *totprim = malloc(sizeof(int) * (size_t)N), *uniq = malloc(sizeof(int) * (size_t)N), *face = malloc(sizeof(int) * (size_t)N);
  for (int n = 0; n < N; n++) {
    const PBVHNode *node = &pbvh->nodes[n];
    memcpy(bb + 6 * (size_t)n, &node->vb, sizeof(float[6]));
    memcpy(obb + 6 * (size_t)n, &node->orig_vb, sizeof(float[6]));
    child[n] = node->children_offset;
    flag[n] = (int)node->flag;
    const bool leaf = (node->flag & PBVH_Leaf) != 0;
    prim_off[n] = leaf ? (int)(node->prim_indices - pbvh->prim_indices) : 0;
    totprim[n] = leaf ? (int)node->totprim : 0;
    uniq[n] = leaf ? (int)node->totprim * area : 0;
    face[n] = 0;
  }
  DscPbvhDesc pd = {0};
  pd.totnode = N;
  pd.node_bb = bb;
  pd.node_orig_bb = obb;
  pd.children_offset = child;
  pd.flag = flag;
  pd.prim_offset = prim_off;
  pd.totprim = totprim;
  pd.prim_indices = pbvh->prim_indices;
  pd.uniq_verts = uniq;
  pd.face_verts = face;
  if (r == DSC_OK) r = dsc_pbvh_upload(ctx, &pd);
  if (r == DSC_OK && (pbvh->want_draw_buffers & 1)) r = push_leaf_shading(pbvh, ctx);
  free(co); free(no); free(mask); free(face_start); free(face_num); free(edge_off); free(edge_elems); free(vert_off);
  free(vert_elems); free(bb); free(obb); free(child); free(flag); free(prim_off); free(totprim); free(uniq); free(face);
  if (r != DSC_OK) {
    snprintf(g_attach_error, sizeof(g_attach_error), "%s", dsc_last_error(ctx));
    dsc_ctx_destroy(ctx);
    return r;
  }
  pbvh->device = ctx;
  pbvh->subdiv_ccg = ccg;
  pbvh->device_dirty = !have_no; /* the device computed the normals */
  return DSC_OK;
}

/* device -> CCGElem storage (co, no, mask) and node boxes / flags */
/* the shading of a leaf's draw buffer: ME_SMOOTH of the poly of its first looptri (gpu_buffers.c:221-222) or of its first grid
 * (grid_flag_mats, gpu_buffers.c:574) */
static int push_leaf_shading(PBVH *pbvh, DscContext *ctx)
{
  if (pbvh->is_grids ? !pbvh->grid_flag_mats : !pbvh->mpoly) return DSC_OK; /* no material flags: the caller names the shading */
  unsigned char *sm = calloc((size_t)pbvh->totnode, 1);
  for (int n = 0; n < pbvh->totnode; n++) {
    const PBVHNode *node = &pbvh->nodes[n];
    if (!(node->flag & PBVH_Leaf) || !node->totprim) continue;
    const int prim = node->prim_indices[0];
    sm[n] = pbvh->is_grids ? (pbvh->grid_flag_mats[prim].flag & ME_SMOOTH) != 0 :
                             (pbvh->mpoly[pbvh->looptri[prim].poly].flag & ME_SMOOTH) != 0;
  }
  const int r = dsc_draw_leaf_shading(ctx, sm);
  free(sm);
  return r;
}

static int push_host_marks(PBVH *pbvh);
static void flags_synced(PBVH *pbvh);

static int sync_grids_to_host(PBVH *pbvh)
{
  const CCGKey *key = &pbvh->gridkey;
  const int area = key->grid_area, G = pbvh->totgrid, N = pbvh->totnode;
  const size_t E = (size_t)G * (size_t)area;
  int r;
  if ((r = push_host_marks(pbvh)) != DSC_OK) return r; /* so the flags that come back carry them */
  /* the grids of a SubdivCCG are one block (grids_storage, subdiv_ccg.c:116-122): whole CCGElem records are packed on
   * the device and land in it by DMA; grids allocated one by one take the layer-by-layer path */
  bool contiguous = G > 0 && key->elem_size % (int)sizeof(float) == 0;
  for (int g = 1; g < G && contiguous; g++) {
    contiguous = (unsigned char *)pbvh->grids[g] == (unsigned char *)pbvh->grids[0] + (size_t)g * (size_t)area * (size_t)key->elem_size;
  }
  if (contiguous && pbvh->dist_world > 1 && !pbvh->gather_whole) {
    /* partitioned: the grids this rank owns */
    r = dsc_download_owned_ccg(pbvh->device, pbvh->grids[0], key->elem_size / (int)sizeof(float),
                               key->has_mask ? key->mask_offset / (int)sizeof(float) : -1,
                               key->has_normals ? key->normal_offset / (int)sizeof(float) : -1);
  }
  else if (contiguous) {
    if (!pbvh->grids_pinned && dsc_host_register(pbvh->device, pbvh->grids[0], E * (size_t)key->elem_size) == DSC_OK) pbvh->grids_pinned = true;
    r = dsc_download_ccg(pbvh->device, pbvh->grids[0], key->elem_size / (int)sizeof(float),
                         key->has_mask ? key->mask_offset / (int)sizeof(float) : -1,
                         key->has_normals ? key->normal_offset / (int)sizeof(float) : -1);
  }
  else {
    float *co = malloc(sizeof(float[3]) * E), *no = malloc(sizeof(float[3]) * E);
    float *mask = key->has_mask ? malloc(sizeof(float) * E) : NULL;
    r = dsc_download_co(pbvh->device, co);
    if (r == DSC_OK) r = dsc_download_no(pbvh->device, no);
    if (r == DSC_OK && mask) r = dsc_download_mask(pbvh->device, mask);
    if (r == DSC_OK) {
      for (int g = 0; g < G; g++) {
        for (int j = 0; j < area; j++) {
          unsigned char *e = (unsigned char *)pbvh->grids[g] + (size_t)key->elem_size * (size_t)j;
          const size_t idx = (size_t)g * (size_t)area + (size_t)j;
          memcpy(e, co + 3 * idx, sizeof(float[3]));
          if (key->has_normals) memcpy(e + key->normal_offset, no + 3 * idx, sizeof(float[3]));
          if (mask) memcpy(e + key->mask_offset, mask + idx, sizeof(float));
        }
      }
    }
    free(co); free(no); free(mask);
  }
  if (r != DSC_OK) return r;
  float *bb = malloc(sizeof(float[6]) * (size_t)N), *obb = malloc(sizeof(float[6]) * (size_t)N);
  int *flag = malloc(sizeof(int) * (size_t)N);
  r = dsc_download_node_bb(pbvh->device, bb, obb);
  if (r == DSC_OK) r = dsc_download_node_flags(pbvh->device, flag);
  if (r == DSC_OK) {
    for (int n = 0; n < N; n++) {
      memcpy(&pbvh->nodes[n].vb, bb + 6 * (size_t)n, sizeof(float[6]));
      memcpy(&pbvh->nodes[n].orig_vb, obb + 6 * (size_t)n, sizeof(float[6]));
      pbvh->nodes[n].flag = (unsigned)flag[n];
    }
    flags_synced(pbvh);
    pbvh->device_dirty = false;
  }
  free(bb); free(obb); free(flag);
  return r;
}

/* kernel/intern/multires_reshape_ccg.c:10-70 + multires_reshape_util.c:417-438 */
bool DUNE_multires_reshape_assign_final_coords(PBVH *pbvh, SubdivCCG *ccg, MDisps *mdisps, GridPaintMask *grid_paint_masks)
{
  if (!ccg || (!mdisps && !grid_paint_masks)) return false;
  const int gs = ccg->grid_size, area = gs * gs, G = ccg->num_grids;
  const bool want_mask = ccg->has_mask && grid_paint_masks != NULL;
  if (pbvh && pbvh->device && pbvh->is_grids) {
    /* the device is authoritative during and after a stroke: element order on the wire is grid, y, x -- a grid's run
     * is its MDisps.disps array */
    const size_t E = (size_t)G * (size_t)area;
    float *co = mdisps ? malloc(sizeof(float[3]) * E) : NULL;
    float *mask = want_mask ? malloc(sizeof(float) * E) : NULL;
    int r = DSC_OK;
    if (co) r = dsc_download_co(pbvh->device, co);
    if (r == DSC_OK && mask) r = dsc_download_mask(pbvh->device, mask);
    if (r == DSC_OK) {
      for (int g = 0; g < G; g++) {
        if (co && mdisps[g].disps) memcpy(mdisps[g].disps, co + 3 * (size_t)g * (size_t)area, sizeof(float[3]) * (size_t)area);
        if (mask && grid_paint_masks[g].data) memcpy(grid_paint_masks[g].data, mask + (size_t)g * (size_t)area, sizeof(float) * (size_t)area);
      }
    }
    free(co);
    free(mask);
    return r == DSC_OK;
  }
  for (int g = 0; g < G; g++) {
    const unsigned char *grid = (const unsigned char *)ccg->grids[g];
    for (int y = 0; y < gs; y++) {
      for (int x = 0; x < gs; x++) {
        const int idx = y * gs + x;
        const unsigned char *e = grid + (size_t)ccg->grid_element_size * (size_t)idx;
        if (mdisps && mdisps[g].disps) memcpy(mdisps[g].disps[idx], e, sizeof(float[3]));
        if (want_mask && grid_paint_masks[g].data) memcpy(&grid_paint_masks[g].data[idx], e + ccg->mask_offset, sizeof(float));
      }
    }
  }
  return true;
}

void BKE_pbvh_free(PBVH *pbvh)
{
  if (!pbvh) return;
  DUNE_pbvh_device_detach(pbvh);
  for (int i = 0; i < pbvh->totnode; i++) {
    PBVHNode *node = &pbvh->nodes[i];
    if (node->flag & PBVH_Leaf) {
      free((void *)node->vert_indices);
      free((void *)node->face_vert_indices);
    }
  }
  if (pbvh->deformed) free(pbvh->verts); /* pbvh.c:2597-2603 */
  if (pbvh->owns_normals) free(pbvh->vert_normals);
  free(pbvh->nodes);
  free(pbvh->prim_indices);
  free(pbvh->vert_bitmap);
  free(pbvh->synced_flag);
  free(pbvh->nb_offsets);
  free(pbvh->nb_indices);
  free(pbvh->boundary);
  free(pbvh);
}

/* ------------------------------------------------------------------------- session tables */

/* vertex -> poly map, count / prefix / fill (kernel/intern/mesh_mapping.c:182-229), then the
 * neighbour list of the sculpt neighbour iterator: per incident poly the previous and next corner
 * (kernel/intern/mesh.c:1566-1589), first occurrence kept.  An edge seen from one poly only makes
 * both its verts boundary verts. */
/* the neighbours of v in iterator order into out[] (distinct, v itself left out), how often each was met into uses[];
 * returns the count.  cap = 2 * (polys at v): enough for every corner's two neighbours */
static int vert_neighbors(const PBVH *pbvh, const int *pm_off, const int *pm_idx, int v, int *out, int *uses)
{
  int n = 0;
  for (int k = pm_off[v]; k < pm_off[v + 1]; k++) {
    const MPoly *mp = &pbvh->mpoly[pm_idx[k]];
    const MLoop *ml = &pbvh->mloop[mp->loopstart];
    int corner = -1;
    for (int j = 0; j < mp->totloop; j++) {
      if ((int)ml[j].v == v) {
        corner = j;
        break;
      }
    }
    if (corner < 0) continue;
    const int adj[2] = {(int)ml[(corner + mp->totloop - 1) % mp->totloop].v, (int)ml[(corner + 1) % mp->totloop].v};
    for (int j = 0; j < 2; j++) {
      if (adj[j] == v) continue;
      int at = -1;
      for (int q = 0; q < n; q++) {
        if (out[q] == adj[j]) {
          at = q;
          break;
        }
      }
      if (at < 0) {
        out[n] = adj[j];
        uses[n++] = 1;
      }
      else {
        uses[at]++;
      }
    }
  }
  return n;
}

static void build_neighbor_tables(PBVH *pbvh)
{
  const int V = pbvh->totvert, P = pbvh->totpoly, Lp = pbvh->totloop;
  /* vertex -> polys (kernel/intern/mesh_mapping.c:182-229): count, prefix, fill in poly order */
  int *pm_off = calloc((size_t)V + 1, sizeof(int));
  int *pm_idx = malloc(sizeof(int) * (size_t)(Lp ? Lp : 1));
  for (int p = 0; p < P; p++) {
    for (int j = 0; j < pbvh->mpoly[p].totloop; j++) pm_off[pbvh->mloop[pbvh->mpoly[p].loopstart + j].v + 1]++;
  }
  for (int v = 0; v < V; v++) pm_off[v + 1] += pm_off[v];
  int *fill = calloc((size_t)V + 1, sizeof(int));
  for (int p = 0; p < P; p++) {
    for (int j = 0; j < pbvh->mpoly[p].totloop; j++) {
      const int v = (int)pbvh->mloop[pbvh->mpoly[p].loopstart + j].v;
      pm_idx[pm_off[v] + fill[v]++] = p;
    }
  }
  free(fill);

  /* a vertex's list depends on nothing but its own polys: count in parallel, prefix, fill in parallel */
  pbvh->nb_offsets = malloc(sizeof(int) * ((size_t)V + 1));
  pbvh->boundary = calloc((size_t)V + 1, 1);
  int max_polys = 1;
  for (int v = 0; v < V; v++) {
    if (pm_off[v + 1] - pm_off[v] > max_polys) max_polys = pm_off[v + 1] - pm_off[v];
  }
#pragma omp parallel
  {
    int *out = malloc(sizeof(int) * (size_t)(2 * max_polys)), *uses = malloc(sizeof(int) * (size_t)(2 * max_polys));
#pragma omp for schedule(static)
    for (int v = 0; v < V; v++) pbvh->nb_offsets[v + 1] = vert_neighbors(pbvh, pm_off, pm_idx, v, out, uses);
    free(out);
    free(uses);
  }
  pbvh->nb_offsets[0] = 0;
  for (int v = 0; v < V; v++) pbvh->nb_offsets[v + 1] += pbvh->nb_offsets[v];
  pbvh->nb_indices = malloc(sizeof(int) * ((size_t)pbvh->nb_offsets[V] + 1));
#pragma omp parallel
  {
    int *uses = malloc(sizeof(int) * (size_t)(2 * max_polys));
#pragma omp for schedule(static)
    for (int v = 0; v < V; v++) {
      int *out = pbvh->nb_indices + pbvh->nb_offsets[v];
      const int n = vert_neighbors(pbvh, pm_off, pm_idx, v, out, uses);
      /* an edge met once is a boundary edge: both ends are boundary verts (a byte set to 1 by several threads is 1) */
      for (int q = 0; q < n; q++) {
        if (uses[q] < 2) {
          pbvh->boundary[v] = 1;
          pbvh->boundary[out[q]] = 1;
        }
      }
    }
    free(uses);
  }
  free(pm_off);
  free(pm_idx);
}

/* -------------------------------------------------------------------------- device hooks */

const char *DUNE_pbvh_device_error(const PBVH *pbvh)
{
  if (pbvh && pbvh->device) return dsc_last_error(pbvh->device);
  return g_attach_error[0] ? g_attach_error : dsc_last_error(NULL);
}

int DUNE_pbvh_device_attach(PBVH *pbvh, int device) { return DUNE_pbvh_device_attach_dist(pbvh, device, 1, 0, NULL); }

int DUNE_pbvh_device_attach_dist(PBVH *pbvh, int device, int world, int rank, const char *nccl_id)
{
  g_attach_error[0] = 0;
  if (!pbvh || !pbvh->nodes) return DSC_ERR_INVALID;
  if (pbvh->device) return DSC_OK;
  if (pbvh->totpoly <= 0 || pbvh->totloop <= 0) {
    snprintf(g_attach_error, sizeof(g_attach_error), "DUNE_pbvh_mesh_sizes_set() was not called");
    return DSC_ERR_STATE;
  }
  DscContext *ctx = NULL;
  int r = dsc_ctx_create(device, &ctx);
  if (r != DSC_OK) {
    snprintf(g_attach_error, sizeof(g_attach_error), "%s", dsc_last_error(NULL));
    return r;
  }
  pbvh->dist_world = world > 1 ? world : 0;
  if (world > 1) {
    r = dsc_dist_init(ctx, world, rank, nccl_id);
    if (r != DSC_OK) {
      snprintf(g_attach_error, sizeof(g_attach_error), "%s", dsc_last_error(ctx));
      dsc_ctx_destroy(ctx);
      return r;
    }
  }
  if (!pbvh->nb_offsets) build_neighbor_tables(pbvh);

  const int V = pbvh->totvert, T = pbvh->totprim, N = pbvh->totnode;
  float *co = malloc(sizeof(float[3]) * (size_t)V);
  for (int v = 0; v < V; v++) memcpy(co + 3 * (size_t)v, pbvh->verts[v].co, sizeof(float[3]));
  int *poly_start = malloc(sizeof(int) * (size_t)pbvh->totpoly), *poly_len = malloc(sizeof(int) * (size_t)pbvh->totpoly);
  for (int p = 0; p < pbvh->totpoly; p++) {
    poly_start[p] = pbvh->mpoly[p].loopstart;
    poly_len[p] = pbvh->mpoly[p].totloop;
  }
  int *loop_v = malloc(sizeof(int) * (size_t)pbvh->totloop);
  for (int l = 0; l < pbvh->totloop; l++) loop_v[l] = (int)pbvh->mloop[l].v;
  int *tri_vert = malloc(sizeof(int[3]) * (size_t)T), *tri_poly = malloc(sizeof(int) * (size_t)T);
  for (int t = 0; t < T; t++) {
    for (int j = 0; j < 3; j++) tri_vert[3 * (size_t)t + j] = (int)pbvh->mloop[pbvh->looptri[t].tri[j]].v;
    tri_poly[t] = (int)pbvh->looptri[t].poly;
  }
  /* normals: use the host's if they are populated, else let the device compute them */
  bool have_no = false;
  for (int v = 0; v < V && !have_no; v++) {
    have_no = pbvh->vert_normals[v][0] != 0.0f || pbvh->vert_normals[v][1] != 0.0f || pbvh->vert_normals[v][2] != 0.0f;
  }
  unsigned int *tail = malloc(sizeof(unsigned int) * (size_t)V);
  for (int v = 0; v < V; v++) memcpy(&tail[v], &pbvh->verts[v].flag, 4);
  DscMeshDesc me = {0};
  me.vert_tail = tail;
  me.totvert = V;
  me.co = co;
  me.no = have_no ? (const float *)pbvh->vert_normals : NULL;
  me.mask = pbvh->vmask;
  me.totpoly = pbvh->totpoly;
  me.totloop = pbvh->totloop;
  me.poly_loopstart = poly_start;
  me.poly_totloop = poly_len;
  me.loop_vert = loop_v;
  me.tottri = T;
  me.tri_vert = tri_vert;
  me.tri_poly = tri_poly;
  me.nb_offsets = pbvh->nb_offsets;
  me.nb_indices = pbvh->nb_indices;
  me.boundary = pbvh->boundary;
  r = dsc_mesh_upload(ctx, &me);

  float *bb = malloc(sizeof(float[6]) * (size_t)N), *obb = malloc(sizeof(float[6]) * (size_t)N);
  int *child = malloc(sizeof(int) * (size_t)N), *flag = malloc(sizeof(int) * (size_t)N);
  int *prim_off = malloc(sizeof(int) * (size_t)N), *totprim = malloc(sizeof(int) * (size_t)N);
  int *uniq = malloc(sizeof(int) * (size_t)N), *face = malloc(sizeof(int) * (size_t)N), *vert_off = malloc(sizeof(int) * (size_t)N);
  size_t totvi = 0;
  for (int n = 0; n < N; n++) {
    const PBVHNode *node = &pbvh->nodes[n];
    memcpy(bb + 6 * (size_t)n, &node->vb, sizeof(float[6]));
    memcpy(obb + 6 * (size_t)n, &node->orig_vb, sizeof(float[6]));
    child[n] = node->children_offset;
    flag[n] = (int)node->flag;
    const bool leaf = (node->flag & PBVH_Leaf) != 0;
    prim_off[n] = leaf ? (int)(node->prim_indices - pbvh->prim_indices) : 0;
    totprim[n] = leaf ? (int)node->totprim : 0;
    uniq[n] = leaf ? (int)node->uniq_verts : 0;
    face[n] = leaf ? (int)node->face_verts : 0;
    vert_off[n] = (int)totvi;
    if (leaf) totvi += node->uniq_verts + node->face_verts;
  }
  int *vert_indices = malloc(sizeof(int) * (totvi ? totvi : 1));
  for (int n = 0; n < N; n++) {
    const PBVHNode *node = &pbvh->nodes[n];
    if (node->flag & PBVH_Leaf) {
      memcpy(vert_indices + vert_off[n], node->vert_indices, sizeof(int) * (node->uniq_verts + node->face_verts));
    }
  }
  DscPbvhDesc pd = {0};
  pd.totnode = N;
  pd.node_bb = bb;
  pd.node_orig_bb = obb;
  pd.children_offset = child;
  pd.flag = flag;
  pd.prim_offset = prim_off;
  pd.totprim = totprim;
  pd.prim_indices = pbvh->prim_indices;
  pd.uniq_verts = uniq;
  pd.face_verts = face;
  pd.vert_offset = vert_off;
  pd.vert_indices = vert_indices;
  if (r == DSC_OK && (pbvh->want_draw_buffers & 1)) r = dsc_draw_enable(ctx);
  if (r == DSC_OK && (pbvh->want_draw_buffers & 2)) r = dsc_raycast_enable(ctx);
  if (r == DSC_OK) r = dsc_pbvh_upload(ctx, &pd);
  if (r == DSC_OK && (pbvh->want_draw_buffers & 1)) r = push_leaf_shading(pbvh, ctx);

  free(tail); free(co); free(poly_start); free(poly_len); free(loop_v); free(tri_vert); free(tri_poly);
  free(bb); free(obb); free(child); free(flag); free(prim_off); free(totprim); free(uniq); free(face);
  free(vert_off); free(vert_indices);
  if (r != DSC_OK) {
    snprintf(g_attach_error, sizeof(g_attach_error), "%s", dsc_last_error(ctx));
    dsc_ctx_destroy(ctx);
    return r;
  }
  pbvh->device = ctx;
  pbvh->device_dirty = !have_no; /* device computed the normals */
  /* page-lock what the stroke-end sync writes: the normals array now, the private MVert copy when it is made */
  if (dsc_host_register(ctx, pbvh->vert_normals, sizeof(float[3]) * (size_t)V) == DSC_OK) pbvh->normals_pinned = true;
  return DSC_OK;
}

void DUNE_pbvh_device_detach(PBVH *pbvh)
{
  if (pbvh && pbvh->device) {
    if (pbvh->normals_pinned) dsc_host_unregister(pbvh->device, pbvh->vert_normals);
    if (pbvh->verts_pinned) dsc_host_unregister(pbvh->device, pbvh->verts);
    if (pbvh->grids_pinned) dsc_host_unregister(pbvh->device, pbvh->grids[0]);
    pbvh->normals_pinned = pbvh->verts_pinned = pbvh->grids_pinned = false;
    dsc_ctx_destroy(pbvh->device);
    pbvh->device = NULL;
  }
}

void DUNE_pbvh_draw_buffers_enable(PBVH *pbvh) { pbvh->want_draw_buffers |= 1; }
void DUNE_pbvh_raycast_enable(PBVH *pbvh) { pbvh->want_draw_buffers |= 2; }

/* BKE_pbvh_raycast (pbvh.c:3915-3928) with the stroke operator's per-node callback folded in: that callback
 * (sculpt_raycast_cb, editors/sculpt_paint/sculpt.c) does nothing but call BKE_pbvh_node_raycast
 * (pbvh.c:4203-4260) on each leaf the ray enters and keep the nearest hit, which is what the device returns. */
bool DUNE_pbvh_raycast_nearest(PBVH *pbvh, const float ray_start[3], const float ray_normal[3], bool original,
                               float max_depth, float *r_depth, int *r_active_vertex_index, int *r_active_face_index,
                               float r_face_normal[3], PBVHNode **r_node)
{
  DscRayHit hit;
  if (!pbvh || !pbvh->device) return false;
  if (dsc_raycast(pbvh->device, ray_start, ray_normal, original ? 1 : 0, max_depth, &hit) != DSC_OK || !hit.hit) return false;
  if (r_depth) *r_depth = hit.depth;
  if (r_active_vertex_index) *r_active_vertex_index = hit.vertex;
  if (r_active_face_index) *r_active_face_index = hit.face;
  if (r_face_normal) memcpy(r_face_normal, hit.face_normal, sizeof(float[3]));
  if (r_node) *r_node = &pbvh->nodes[hit.node];
  return true;
}

int DUNE_pbvh_update_draw_buffers(PBVH *pbvh, int shading, bool show_mask)
{
  if (!pbvh || !pbvh->device) return DSC_ERR_STATE;
  const int pr = push_host_marks(pbvh);
  if (pr != DSC_OK) return pr;
  return dsc_draw_update(pbvh->device, shading, show_mask ? 1 : 0);
}

int DUNE_pbvh_node_draw_buffer(PBVH *pbvh, PBVHNode *node, void **r_device_ptr, int *r_vert_len)
{
  if (!pbvh || !pbvh->device || !node) return DSC_ERR_STATE;
  return dsc_draw_node_buffer(pbvh->device, (int)(node - pbvh->nodes), r_device_ptr, r_vert_len);
}

int DUNE_pbvh_device_checkpoint(PBVH *pbvh)
{
  if (!pbvh || !pbvh->device) return DSC_ERR_STATE;
  return dsc_state_save(pbvh->device);
}

int DUNE_pbvh_device_rollback(PBVH *pbvh)
{
  if (!pbvh || !pbvh->device) return DSC_ERR_STATE;
  const int r = dsc_state_restore(pbvh->device);
  if (r == DSC_OK) pbvh->device_dirty = true;
  return r;
}

int DUNE_pbvh_device_sync_to_host(PBVH *pbvh)
{
  if (!pbvh || !pbvh->device) return DSC_ERR_STATE;
  if (!pbvh->device_dirty) return DSC_OK;
  if (pbvh->is_grids) return sync_grids_to_host(pbvh);
  const int V = pbvh->totvert, N = pbvh->totnode;
  {
    const int pr = push_host_marks(pbvh); /* so the flags that come back carry them */
    if (pr != DSC_OK) return pr;
  }
  if (!pbvh->deformed) {
    /* first write: take a private copy like BKE_pbvh_vert_coords_apply (pbvh.c:4714-4725) */
    MVert *dup = malloc(sizeof(MVert) * (size_t)V);
    memcpy(dup, pbvh->verts, sizeof(MVert) * (size_t)V);
    pbvh->verts = dup;
    pbvh->deformed = true;
    if (dsc_host_register(pbvh->device, dup, sizeof(MVert) * (size_t)V) == DSC_OK) pbvh->verts_pinned = true;
  }
  int r;
  if (pbvh->dist_world > 1 && !pbvh->gather_whole) {
    /* partitioned: the vertices this rank owns (their slots are one run of every device array) */
    r = dsc_download_owned_mvert(pbvh->device, pbvh->verts, (float *)pbvh->vert_normals);
  }
  else {
    /* whole MVert records and the normals array arrive by DMA; no host-side scatter */
    r = dsc_download_mvert(pbvh->device, pbvh->verts);
    if (r == DSC_OK) r = dsc_download_no(pbvh->device, (float *)pbvh->vert_normals);
  }
  if (r != DSC_OK) return r;
  float *bb = malloc(sizeof(float[6]) * (size_t)N), *obb = malloc(sizeof(float[6]) * (size_t)N);
  int *flag = malloc(sizeof(int) * (size_t)N);
  r = dsc_download_node_bb(pbvh->device, bb, obb);
  if (r == DSC_OK) r = dsc_download_node_flags(pbvh->device, flag);
  if (r == DSC_OK) {
    for (int n = 0; n < N; n++) {
      memcpy(&pbvh->nodes[n].vb, bb + 6 * (size_t)n, sizeof(float[6]));
      memcpy(&pbvh->nodes[n].orig_vb, obb + 6 * (size_t)n, sizeof(float[6]));
      pbvh->nodes[n].flag = (unsigned)flag[n];
    }
    flags_synced(pbvh);
    pbvh->device_dirty = false;
  }
  free(bb); free(obb); free(flag);
  return r;
}

int DUNE_pbvh_device_gather(PBVH *pbvh)
{
  if (!pbvh || !pbvh->device) return DSC_ERR_STATE;
  int r = dsc_dist_gather(pbvh->device);
  if (r != DSC_OK) return r;
  pbvh->gather_whole = true;
  pbvh->device_dirty = true;
  r = DUNE_pbvh_device_sync_to_host(pbvh);
  pbvh->gather_whole = false;
  return r;
}

/* ------------------------------------------------------------------------------ traversal */

bool SCULPT_search_sphere_cb(PBVHNode *node, void *data_v)
{
  const SculptSearchSphereData *data = data_v;
  if (data->ignore_fully_ineffective) {
    if (BKE_pbvh_node_fully_hidden_get(node)) return false;
    if (BKE_pbvh_node_fully_masked_get(node)) return false;
  }
  const BB *bb = data->original ? &node->orig_vb : &node->vb;
  float t[3];
  for (int i = 0; i < 3; i++) {
    float nearest = data->center[i];
    if (bb->bmin[i] > data->center[i]) nearest = bb->bmin[i];
    else if (bb->bmax[i] < data->center[i]) nearest = bb->bmax[i];
    t[i] = data->center[i] - nearest;
  }
  return (t[0] * t[0] + t[1] * t[1] + t[2] * t[2]) < data->radius_squared;
}

void BKE_pbvh_search_gather(PBVH *pbvh, BKE_pbvh_SearchCallback scb, void *search_data, PBVHNode ***r_array, int *r_tot)
{
  *r_array = NULL;
  *r_tot = 0;
  if (!pbvh->nodes) return;
  if (pbvh->device && scb == SCULPT_search_sphere_cb) {
    /* device path: flat leaf test + ordered compaction, then node indices -> pointers */
    const SculptSearchSphereData *d = search_data;
    int *idx = malloc(sizeof(int) * (size_t)pbvh->totnode);
    int tot = 0;
    if (dsc_search_sphere(pbvh->device, d->center, d->radius_squared, d->original, d->ignore_fully_ineffective, idx,
                          pbvh->totnode, &tot) == DSC_OK &&
        tot > 0) {
      PBVHNode **array = MEM_mallocN(sizeof(PBVHNode *) * (size_t)tot, __func__);
      for (int i = 0; i < tot; i++) array[i] = &pbvh->nodes[idx[i]];
      *r_array = array;
      *r_tot = tot;
    }
    free(idx);
    return;
  }
  /* host path for arbitrary callbacks: stack DFS, children left first (pbvh.c:2664-2705) */
  if (pbvh->device) DUNE_pbvh_device_sync_to_host(pbvh);
  int cap = 64, top = 0, tot = 0, space = 0;
  int *stack = malloc(sizeof(int) * (size_t)cap);
  PBVHNode **array = NULL;
  stack[top++] = 0;
  while (top) {
    PBVHNode *node = &pbvh->nodes[stack[--top]];
    if (scb && !scb(node, search_data)) continue;
    if (node->flag & PBVH_Leaf) {
      if (tot == space) {
        space = tot ? space * 2 : 32;
        PBVHNode **na = MEM_callocN(sizeof(PBVHNode *) * (size_t)space, __func__);
        if (array) {
          memcpy(na, array, sizeof(PBVHNode *) * (size_t)tot);
          MEM_freeN(array);
        }
        array = na;
      }
      array[tot++] = node;
      continue;
    }
    if (top + 2 > cap) {
      cap *= 2;
      stack = realloc(stack, sizeof(int) * (size_t)cap);
    }
    stack[top++] = node->children_offset + 1;
    stack[top++] = node->children_offset;
  }
  free(stack);
  *r_array = array;
  *r_tot = tot;
}

/* ---------------------------------------------------------------------------- node access */

void BKE_pbvh_node_mark_update(PBVHNode *node)
{
  node->flag |= PBVH_UpdateNormals | PBVH_UpdateBB | PBVH_UpdateOriginalBB | PBVH_UpdateDrawBuffers | PBVH_UpdateRedraw;
}
void BKE_pbvh_vert_mark_update(PBVH *pbvh, int index)
{
  pbvh->vert_bitmap[index >> 5] |= 1u << (index & 31);
  pbvh->host_vert_marks = true;
}
void BKE_pbvh_node_fully_hidden_set(PBVHNode *node, int fully_hidden)
{
  if (fully_hidden) node->flag |= PBVH_FullyHidden;
  else node->flag &= ~(unsigned)PBVH_FullyHidden;
}
bool BKE_pbvh_node_fully_hidden_get(PBVHNode *node) { return (node->flag & PBVH_Leaf) && (node->flag & PBVH_FullyHidden); }
void BKE_pbvh_node_fully_masked_set(PBVHNode *node, int fully_masked)
{
  if (fully_masked) node->flag |= PBVH_FullyMasked;
  else node->flag &= ~(unsigned)PBVH_FullyMasked;
}
bool BKE_pbvh_node_fully_masked_get(PBVHNode *node) { return (node->flag & PBVH_Leaf) && (node->flag & PBVH_FullyMasked); }
void BKE_pbvh_node_get_verts(PBVH *pbvh, PBVHNode *node, const int **r_vert_indices, MVert **r_verts)
{
  if (r_vert_indices) *r_vert_indices = node->vert_indices;
  if (r_verts) *r_verts = pbvh->verts;
}
void BKE_pbvh_node_num_verts(PBVH *pbvh, PBVHNode *node, int *r_uniquevert, int *r_totvert)
{
  (void)pbvh;
  if (r_totvert) *r_totvert = (int)(node->uniq_verts + node->face_verts);
  if (r_uniquevert) *r_uniquevert = (int)node->uniq_verts;
}
void BKE_pbvh_node_get_BB(PBVHNode *node, float bb_min[3], float bb_max[3])
{
  memcpy(bb_min, node->vb.bmin, sizeof(float[3]));
  memcpy(bb_max, node->vb.bmax, sizeof(float[3]));
}
void BKE_pbvh_node_get_original_BB(PBVHNode *node, float bb_min[3], float bb_max[3])
{
  memcpy(bb_min, node->orig_vb.bmin, sizeof(float[3]));
  memcpy(bb_max, node->orig_vb.bmax, sizeof(float[3]));
}

/* -------------------------------------------------------------------------------- updates */

/* Host-side marks made since the last push go down first, in one batch and whether or not the device already holds newer
 * state: node flags changed through the BKE_pbvh_node_* setters (update / redraw / draw-buffer marks, FullyHidden,
 * FullyMasked) as set / clear masks against the flags the device last saw, and the BKE_pbvh_vert_mark_update bitmap. */
static void flags_synced(PBVH *pbvh)
{
  if (!pbvh->synced_flag) pbvh->synced_flag = malloc(sizeof(unsigned) * (size_t)pbvh->totnode);
  for (int n = 0; n < pbvh->totnode; n++) pbvh->synced_flag[n] = pbvh->nodes[n].flag;
}
static int push_host_marks(PBVH *pbvh)
{
  if (!pbvh->device) return DSC_OK;
  int r = DSC_OK;
  if (!pbvh->synced_flag) {
    flags_synced(pbvh); /* the device got the flags at attach */
  }
  else {
    const unsigned forwarded = PBVH_UpdateNormals | PBVH_UpdateBB | PBVH_UpdateOriginalBB | PBVH_UpdateDrawBuffers | PBVH_UpdateRedraw |
                               PBVH_RebuildDrawBuffers | PBVH_FullyHidden | PBVH_FullyMasked | PBVH_UpdateMask | PBVH_UpdateVisibility;
    int count = 0, cap = 0;
    int *nodes = NULL, *set = NULL, *clear = NULL;
    for (int n = 0; n < pbvh->totnode; n++) {
      const unsigned now = pbvh->nodes[n].flag, was = pbvh->synced_flag[n];
      if (now == was) continue;
      const unsigned s_ = now & ~was & forwarded, c_ = was & ~now & forwarded;
      pbvh->synced_flag[n] = now;
      if (!(s_ | c_)) continue;
      if (count == cap) {
        cap = cap ? 2 * cap : 64;
        nodes = realloc(nodes, sizeof(int) * (size_t)cap);
        set = realloc(set, sizeof(int) * (size_t)cap);
        clear = realloc(clear, sizeof(int) * (size_t)cap);
      }
      nodes[count] = n; set[count] = (int)s_; clear[count] = (int)c_;
      count++;
    }
    if (count) r = dsc_node_flags_apply(pbvh->device, count, nodes, set, clear);
    free(nodes); free(set); free(clear);
  }
  if (r == DSC_OK && pbvh->host_vert_marks && !pbvh->is_grids) {
    r = dsc_vert_marks_or(pbvh->device, pbvh->vert_bitmap);
    if (r == DSC_OK) {
      memset(pbvh->vert_bitmap, 0, sizeof(unsigned) * ((size_t)pbvh->totvert / 32 + 1)); /* the device's bitmap owns them now */
      pbvh->host_vert_marks = false;
    }
  }
  return r;
}

void BKE_pbvh_update_normals(PBVH *pbvh, struct SubdivCCG *subdiv_ccg)
{
  (void)subdiv_ccg;
  if (!pbvh->device) return; /* no CPU fallback: without a device the PBVH is a plain container */
  push_host_marks(pbvh);
  dsc_update_normals(pbvh->device);
  pbvh->device_dirty = true;
}

void BKE_pbvh_update_bounds(PBVH *pbvh, int flag)
{
  if (!pbvh->nodes || !pbvh->device) return;
  push_host_marks(pbvh);
  dsc_update_bounds(pbvh->device, flag);
  pbvh->device_dirty = true;
}

/* ----------------------------------------------------------------------------------- sync */

float (*BKE_pbvh_vert_coords_alloc(PBVH *pbvh))[3]
{
  if (!pbvh->verts) return NULL;
  float(*vertCos)[3] = MEM_callocN(3 * (size_t)pbvh->totvert * sizeof(float), "BKE_pbvh_get_vertCoords");
  if (pbvh->device && pbvh->device_dirty) {
    dsc_download_co(pbvh->device, (float *)vertCos);
  }
  else {
    for (int a = 0; a < pbvh->totvert; a++) memcpy(vertCos[a], pbvh->verts[a].co, sizeof(float[3]));
  }
  return vertCos;
}

void BKE_pbvh_vert_coords_apply(PBVH *pbvh, const float (*vertCos)[3], const int totvert)
{
  if (totvert != pbvh->totvert || !pbvh->verts) return;
  if (!pbvh->deformed) {
    MVert *dup = malloc(sizeof(MVert) * (size_t)totvert);
    memcpy(dup, pbvh->verts, sizeof(MVert) * (size_t)totvert);
    pbvh->verts = dup;
    pbvh->deformed = true;
    if (pbvh->device && dsc_host_register(pbvh->device, dup, sizeof(MVert) * (size_t)totvert) == DSC_OK) pbvh->verts_pinned = true;
  }
  for (int a = 0; a < totvert; a++) memcpy(pbvh->verts[a].co, vertCos[a], sizeof(float[3]));
  if (pbvh->device) {
    dsc_upload_co(pbvh->device, (const float *)vertCos);
    pbvh->device_dirty = true;
  }
}

MVert *BKE_pbvh_get_verts(const PBVH *pbvh)
{
  if (pbvh->device && pbvh->device_dirty) DUNE_pbvh_device_sync_to_host((PBVH *)pbvh);
  return pbvh->verts;
}
const float (*BKE_pbvh_get_vert_normals(const PBVH *pbvh))[3]
{
  if (pbvh->device && pbvh->device_dirty) DUNE_pbvh_device_sync_to_host((PBVH *)pbvh);
  return (const float(*)[3])pbvh->vert_normals;
}

/* ----------------------------------------------------------------------------- stroke side */

float DUNE_sculpt_brush_strength(int sculpt_tool, float root_alpha, float pressure, bool dir_in, bool invert, float overlap,
                                 float feather)
{
  /* alpha is squared; direction from BRUSH_DIR_IN and the invert modifier */
  const float alpha = root_alpha * root_alpha;
  const float flip = (dir_in ? -1.0f : 1.0f) * (invert ? -1.0f : 1.0f);
  switch (sculpt_tool) {
    case DSC_TOOL_DRAW:
      return alpha * flip * pressure * overlap * feather;
    case DSC_TOOL_CLAY_STRIPS:
      return alpha * flip * powf(pressure, 1.5f) * overlap * feather * 0.3f;
    case DSC_TOOL_INFLATE:
      return ((flip > 0.0f) ? 0.250f : 0.125f) * alpha * flip * pressure * overlap * feather;
    case DSC_TOOL_SMOOTH:
      return flip * alpha * pressure * feather;
    case DSC_TOOL_GRAB:
      return root_alpha * feather;
  }
  return 0.0f;
}

void DUNE_sculpt_dab_defaults(DscDab *dab, int sculpt_tool)
{
  memset(dab, 0, sizeof(*dab));
  dab->tool = sculpt_tool;
  dab->curve_preset = DSC_CURVE_SMOOTH;
  dab->sculpt_plane = DSC_DIR_AREA;         /* types_brush_defaults.h:28 */
  dab->normal_radius_factor = 0.5f;         /* :23 */
  dab->plane_offset = 0.0f;                 /* :34 */
  dab->plane_trim = 0.5f;                   /* :35 */
  dab->hardness = 0.0f;                     /* :83 */
  dab->tip_roundness = 0.0f;
  dab->scale[0] = dab->scale[1] = dab->scale[2] = 1.0f;
  dab->view_normal[2] = 1.0f;
  dab->radius = 1.0f;
  dab->radius_scale = 1.0f;
  dab->bstrength = DUNE_sculpt_brush_strength(sculpt_tool, 1.0f /* :20 */, 1.0f, false, false, 1.0f, 1.0f);
}

/* SCULPT_is_symmetry_iteration_valid + flip_v3_v3 (paint.h, sculpt.c do_symmetrical_brush_actions): the dab mirrored
 * for every valid combination of the symmetry axes; r_dabs[0] is the dab itself.  Returns the count (1..8). */
int DUNE_sculpt_dab_symmetry(const DscDab *dab, int symm, DscDab r_dabs[8])
{
  int n = 0;
  for (int i = 0; i <= symm; i++) {
    const bool valid = i == 0 || ((symm & i) && (symm != 5 || i != 3) && (symm != 6 || (i != 3 && i != 5)));
    if (!valid) continue;
    DscDab *o = &r_dabs[n++];
    *o = *dab;
    for (int k = 0; k < 3; k++) {
      if (i & (1 << k)) {
        o->location[k] = -o->location[k];
        o->view_normal[k] = -o->view_normal[k];
        o->grab_delta[k] = -o->grab_delta[k];
      }
    }
  }
  return n;
}

int DUNE_sculpt_stroke_begin(PBVH *pbvh, const float *automask)
{
  if (!pbvh->device) return DSC_ERR_STATE;
  int r = push_host_marks(pbvh);
  if (r != DSC_OK) return r;
  r = dsc_stroke_begin(pbvh->device, automask);
  if (r == DSC_OK) pbvh->in_stroke = true;
  return r;
}
int DUNE_sculpt_dab(PBVH *pbvh, const DscDab *dab)
{
  if (!pbvh->device) return DSC_ERR_STATE;
  const int pr = push_host_marks(pbvh);
  if (pr != DSC_OK) return pr;
  pbvh->device_dirty = true;
  return dsc_dab(pbvh->device, dab);
}
int DUNE_sculpt_stroke_end(PBVH *pbvh)
{
  if (!pbvh->device) return DSC_ERR_STATE;
  int r = dsc_stroke_end(pbvh->device);
  pbvh->in_stroke = false;
  if (r != DSC_OK) return r;
  return DUNE_pbvh_device_sync_to_host(pbvh);
}

void DUNE_sculpt_automask_boundary_edges(const PBVH *pbvh, int propagation_steps, float *r_factor)
{
  const int V = pbvh->totvert;
  if (!pbvh->nb_offsets) build_neighbor_tables((PBVH *)pbvh);
  int *dist = malloc(sizeof(int) * (size_t)(V ? V : 1));
  for (int i = 0; i < V; i++) {
    dist[i] = pbvh->boundary[i] ? 0 : -1;
    r_factor[i] = 1.0f;
  }
  for (int it = 0; it < propagation_steps; it++) {
    for (int i = 0; i < V; i++) {
      if (dist[i] != -1) continue;
      for (int q = pbvh->nb_offsets[i]; q < pbvh->nb_offsets[i + 1]; q++) {
        if (dist[pbvh->nb_indices[q]] == it) dist[i] = it + 1;
      }
    }
  }
  for (int i = 0; i < V; i++) {
    if (dist[i] == -1) continue;
    const float p = 1.0f - ((float)dist[i] / (float)propagation_steps);
    r_factor[i] *= (1.0f - p * p);
  }
  free(dist);
}

void DUNE_sculpt_automask_topology(const PBVH *pbvh, int seed_vert, const float location[3], float radius, float *r_factor)
{
  const int V = pbvh->totvert;
  if (!pbvh->nb_offsets) build_neighbor_tables((PBVH *)pbvh);
  for (int i = 0; i < V; i++) r_factor[i] = 0.0f;
  if (seed_vert < 0 || seed_vert >= V) return;
  int *queue = malloc(sizeof(int) * (size_t)V);
  unsigned char *seen = calloc((size_t)V, 1);
  int head = 0, tail = 0;
  queue[tail++] = seed_vert;
  seen[seed_vert] = 1;
  r_factor[seed_vert] = 1.0f;
  const float rsq = radius * radius;
  while (head < tail) {
    const int v = queue[head++];
    for (int q = pbvh->nb_offsets[v]; q < pbvh->nb_offsets[v + 1]; q++) {
      const int u = pbvh->nb_indices[q];
      if (seen[u]) continue;
      seen[u] = 1;
      r_factor[u] = 1.0f;
      bool go_on = true;
      if (radius > 0.0f) {
        const float *co = pbvh->verts[u].co;
        const float dx = co[0] - location[0], dy = co[1] - location[1], dz = co[2] - location[2];
        go_on = (dx * dx + dy * dy + dz * dz) <= rsq;
      }
      if (go_on) queue[tail++] = u;
    }
  }
  free(queue);
  free(seen);
}
