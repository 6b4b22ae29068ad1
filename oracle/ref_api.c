/* oracle/_ref driver -- TEST INFRASTRUCTURE ONLY.
 *
 * Flat-array entry points around the functions oracle/ref_extract.py cuts out of the reference (see ref_shim.h):
 * tests/test_ref_pin.py feeds the same meshes to these and to the oracle and asserts bit equality.  This file holds
 * no algorithm of the path: it allocates the reference's structs, calls the reference's functions, copies results
 * out.  The one exception is the sphere callback (the editor-side SCULPT_search_sphere_cb is absent from the
 * reference): it is the dagger spec, over the reference's own BKE_pbvh_node_get_BB.
 */
#include "ref_shim.h"

int ref_leaf_limit = 10000; /* pbvh.c:1952 */

/* ---- MEM_* ---- */
typedef struct RefMemHead {
  size_t size;
  size_t pad;
} RefMemHead;
void *ref_mem_alloc(size_t size, int zero)
{
  RefMemHead *h = (RefMemHead *)(zero ? calloc(1, sizeof(RefMemHead) + size) : malloc(sizeof(RefMemHead) + size));
  h->size = size;
  return h + 1;
}
void *ref_mem_realloc(void *p, size_t size, int zero)
{
  if (!p) return ref_mem_alloc(size, zero);
  RefMemHead *h = (RefMemHead *)p - 1;
  const size_t old = h->size;
  h = (RefMemHead *)realloc(h, sizeof(RefMemHead) + size);
  h->size = size;
  if (zero && size > old) memset((char *)(h + 1) + old, 0, size - old);
  return h + 1;
}
void ref_mem_free(void *p)
{
  if (p) free((RefMemHead *)p - 1);
}

/* ---- GHash: open addressing, entries kept in insertion order ---- */
struct GHash {
  int cap, n;  /* table capacity (power of two), entries */
  int *table;  /* entry index + 1, 0 = empty */
  void **keys, **vals;
  int ecap;
};
GHash *BLI_ghash_int_new_ex(const char *info, unsigned int reserve)
{
  (void)info;
  GHash *gh = (GHash *)calloc(1, sizeof(GHash));
  gh->cap = 64;
  while ((unsigned)gh->cap < 2 * reserve + 2) gh->cap *= 2;
  gh->table = (int *)calloc((size_t)gh->cap, sizeof(int));
  gh->ecap = (int)reserve + 16;
  gh->keys = (void **)malloc(sizeof(void *) * (size_t)gh->ecap);
  gh->vals = (void **)malloc(sizeof(void *) * (size_t)gh->ecap);
  return gh;
}
static unsigned ref_hash(intptr_t k) { return (unsigned)k * 2654435761u; }
static void ref_ghash_grow(GHash *gh)
{
  free(gh->table);
  gh->cap *= 2;
  gh->table = (int *)calloc((size_t)gh->cap, sizeof(int));
  for (int e = 0; e < gh->n; e++) {
    unsigned s = ref_hash((intptr_t)gh->keys[e]) & (unsigned)(gh->cap - 1);
    while (gh->table[s]) s = (s + 1) & (unsigned)(gh->cap - 1);
    gh->table[s] = e + 1;
  }
}
bool BLI_ghash_ensure_p(GHash *gh, void *key, void ***r_val)
{
  if (2 * (gh->n + 1) > gh->cap) ref_ghash_grow(gh);
  unsigned s = ref_hash((intptr_t)key) & (unsigned)(gh->cap - 1);
  while (gh->table[s]) {
    const int e = gh->table[s] - 1;
    if (gh->keys[e] == key) {
      *r_val = &gh->vals[e];
      return true;
    }
    s = (s + 1) & (unsigned)(gh->cap - 1);
  }
  if (gh->n == gh->ecap) {
    gh->ecap *= 2;
    gh->keys = (void **)realloc(gh->keys, sizeof(void *) * (size_t)gh->ecap);
    gh->vals = (void **)realloc(gh->vals, sizeof(void *) * (size_t)gh->ecap);
  }
  gh->keys[gh->n] = key;
  gh->vals[gh->n] = NULL;
  gh->table[s] = gh->n + 1;
  *r_val = &gh->vals[gh->n];
  gh->n++;
  return false;
}
void BLI_ghash_free(GHash *gh, void *keyfree, void *valfree)
{
  (void)keyfree;
  (void)valfree;
  free(gh->table);
  free(gh->keys);
  free(gh->vals);
  free(gh);
}
void *BLI_ghashIterator_getKey(GHashIterator *ghi) { return ghi->gh->keys[ghi->i]; }
void *BLI_ghashIterator_getValue(GHashIterator *ghi) { return ghi->gh->vals[ghi->i]; }
bool ref_ghash_iter_done(GHashIterator *ghi) { return ghi->i >= ghi->gh->n; }

/* ---- the session the tests hold ---- */
typedef struct RefSession {
  PBVH *pbvh;
  Mesh mesh;
  CustomData vdata;
  MVert *verts;
  MPoly *mpoly;
  MLoop *mloop;
  MLoopTri *looptri;
  float (*vert_normals)[3];
  int totvert, totpoly, totloop, tottri;
  /* grids */
  CCGKey key;
  float *elems; /* totgrid * grid_area * elem floats */
  CCGElem **grids;
  DMFlagMat *flagmats;
  BLI_bitmap **grid_hidden;
  int totgrid;
  SubdivCCG ccg; /* faces / adjacent_edges / adjacent_vertices when ref_grids_set_adjacency was called */
  SubdivCCGCoord *coord_store;
  SubdivCCGCoord **coord_rows;
} RefSession;

/* the reference's own callbacks are static: the generated unit exports them through these (ref_extract.py appends) */
bool ref_update_search_cb(PBVHNode *node, void *data_v);
void ref_faces_update_normals(PBVH *pbvh, PBVHNode **nodes, int totnode);
void ref_ccg_inner_normals(SubdivCCG *ccg, CCGKey *key, int grid_index);
void ref_ccg_average_face(SubdivCCG *ccg, CCGKey *key, int face);
void ref_ccg_average_edge(SubdivCCG *ccg, CCGKey *key, int edge);
void ref_ccg_average_cvert(SubdivCCG *ccg, CCGKey *key, int v);

RefSession *ref_build_mesh(int totvert, const float *co, const unsigned char *vert_flag, int totpoly, const int *poly_start,
                           const int *poly_len, const short *poly_mat, const unsigned char *poly_flag, int totloop,
                           const int *loop_v, int tottri, const int *tri_loop, const int *tri_poly, const float *mask,
                           int leaf_limit)
{
  RefSession *s = (RefSession *)calloc(1, sizeof(RefSession));
  s->totvert = totvert; s->totpoly = totpoly; s->totloop = totloop; s->tottri = tottri;
  s->verts = (MVert *)calloc((size_t)totvert + 1, sizeof(MVert));
  for (int i = 0; i < totvert; i++) {
    memcpy(s->verts[i].co, co + 3 * i, sizeof(float[3]));
    s->verts[i].flag = vert_flag ? (char)vert_flag[i] : 0;
  }
  s->mpoly = (MPoly *)calloc((size_t)totpoly + 1, sizeof(MPoly));
  for (int i = 0; i < totpoly; i++) {
    s->mpoly[i].loopstart = poly_start[i];
    s->mpoly[i].totloop = poly_len[i];
    s->mpoly[i].mat_nr = poly_mat ? poly_mat[i] : 0;
    s->mpoly[i].flag = poly_flag ? (char)poly_flag[i] : 0;
  }
  s->mloop = (MLoop *)calloc((size_t)totloop + 1, sizeof(MLoop));
  for (int i = 0; i < totloop; i++) s->mloop[i].v = (unsigned)loop_v[i];
  /* PBVH owns looptri (pbvh.c:2605-2607); the session frees it itself since BKE_pbvh_free is not extracted */
  s->looptri = (MLoopTri *)calloc((size_t)tottri + 1, sizeof(MLoopTri));
  for (int i = 0; i < tottri; i++) {
    for (int j = 0; j < 3; j++) s->looptri[i].tri[j] = (unsigned)tri_loop[3 * i + j];
    s->looptri[i].poly = (unsigned)tri_poly[i];
  }
  s->vert_normals = (float(*)[3])calloc((size_t)totvert + 1, sizeof(float[3]));
  s->mesh.vert_normals = s->vert_normals;
  if (mask) {
    s->vdata.paint_mask = (float *)malloc(sizeof(float) * (size_t)totvert);
    memcpy(s->vdata.paint_mask, mask, sizeof(float) * (size_t)totvert);
  }
  ref_leaf_limit = leaf_limit > 0 ? leaf_limit : 10000;
  s->pbvh = BKE_pbvh_new();
  BKE_pbvh_build_mesh(s->pbvh, &s->mesh, s->mpoly, s->mloop, s->verts, totvert, &s->vdata, NULL, NULL, s->looptri, tottri);
  ref_leaf_limit = 10000;
  return s;
}

/* elements interleaved co[3], mask (optional), no[3]: subdiv_ccg.c:62-90 */
RefSession *ref_build_grids(int totgrid, int grid_size, const float *co, const float *no, const float *mask,
                            const short *grid_mat, const unsigned char *grid_flag, const unsigned char *hidden /* [totgrid * area] or NULL */,
                            int leaf_limit_prims)
{
  RefSession *s = (RefSession *)calloc(1, sizeof(RefSession));
  const int area = grid_size * grid_size;
  const int ef = 3 + (mask ? 1 : 0) + 3;
  s->totgrid = totgrid;
  s->key.level = 0;
  s->key.elem_size = (int)sizeof(float) * ef;
  s->key.grid_size = grid_size;
  s->key.grid_area = area;
  s->key.grid_bytes = s->key.elem_size * area;
  s->key.has_mask = mask ? 1 : 0;
  s->key.has_normals = 1;
  s->key.mask_offset = mask ? (int)sizeof(float) * 3 : -1;
  s->key.normal_offset = (int)sizeof(float) * (mask ? 4 : 3);
  s->elems = (float *)calloc((size_t)totgrid * (size_t)area * (size_t)ef + 1, sizeof(float));
  s->grids = (CCGElem **)calloc((size_t)totgrid + 1, sizeof(CCGElem *));
  for (int g = 0; g < totgrid; g++) {
    s->grids[g] = (CCGElem *)(s->elems + (size_t)g * (size_t)area * (size_t)ef);
    for (int e = 0; e < area; e++) {
      float *el = s->elems + ((size_t)g * (size_t)area + (size_t)e) * (size_t)ef;
      const size_t i = (size_t)g * (size_t)area + (size_t)e;
      memcpy(el, co + 3 * i, sizeof(float[3]));
      if (mask) el[3] = mask[i];
      if (no) memcpy(el + (mask ? 4 : 3), no + 3 * i, sizeof(float[3]));
    }
  }
  s->flagmats = (DMFlagMat *)calloc((size_t)totgrid + 1, sizeof(DMFlagMat));
  for (int g = 0; g < totgrid; g++) {
    s->flagmats[g].mat_nr = grid_mat ? grid_mat[g] : 0;
    s->flagmats[g].flag = grid_flag ? (char)grid_flag[g] : 0;
  }
  s->grid_hidden = (BLI_bitmap **)calloc((size_t)totgrid + 1, sizeof(BLI_bitmap *));
  if (hidden) {
    for (int g = 0; g < totgrid; g++) {
      int any = 0;
      for (int e = 0; e < area; e++) any |= hidden[(size_t)g * (size_t)area + (size_t)e];
      if (!any) continue;
      s->grid_hidden[g] = BLI_BITMAP_NEW(area, "grid_hidden");
      for (int e = 0; e < area; e++) {
        if (hidden[(size_t)g * (size_t)area + (size_t)e]) BLI_BITMAP_ENABLE(s->grid_hidden[g], e);
      }
    }
  }
  /* leaf_limit = max(LEAF_LIMIT / grid_area, 1) (pbvh.c:2533): choose LEAF_LIMIT to give the asked prims per leaf */
  ref_leaf_limit = leaf_limit_prims > 0 ? leaf_limit_prims * area : 10000;
  s->pbvh = BKE_pbvh_new();
  BKE_pbvh_build_grids(s->pbvh, s->grids, totgrid, &s->key, NULL, s->flagmats, s->grid_hidden);
  ref_leaf_limit = 10000;
  return s;
}

void ref_free(RefSession *s)
{
  if (!s) return;
  PBVH *p = s->pbvh;
  for (int i = 0; i < p->totnode; i++) {
    if (p->nodes[i].flag & PBVH_Leaf) {
      if (p->nodes[i].vert_indices) MEM_freeN(p->nodes[i].vert_indices);
      if (p->nodes[i].face_vert_indices) MEM_freeN(p->nodes[i].face_vert_indices);
    }
  }
  if (p->nodes) MEM_freeN(p->nodes);
  if (p->prim_indices) MEM_freeN(p->prim_indices);
  if (p->vert_bitmap) MEM_freeN(p->vert_bitmap);
  MEM_freeN(p);
  free(s->verts); free(s->mpoly); free(s->mloop); free(s->looptri); free(s->vert_normals); free(s->vdata.paint_mask);
  if (s->grid_hidden) {
    for (int g = 0; g < s->totgrid; g++) {
      if (s->grid_hidden[g]) MEM_freeN(s->grid_hidden[g]);
    }
  }
  free(s->elems); free(s->grids); free(s->flagmats); free(s->grid_hidden);
  free(s->ccg.faces); free(s->ccg.adjacent_edges); free(s->ccg.adjacent_vertices); free(s->coord_store); free(s->coord_rows);
  free(s);
}

int ref_totnode(RefSession *s) { return s->pbvh->totnode; }
int ref_totprim(RefSession *s) { return s->pbvh->totprim; }
int ref_leaf_limit_used(RefSession *s) { return s->pbvh->leaf_limit; }

void ref_nodes(RefSession *s, float *vb, float *orig_vb, int *children_offset, int *flag, int *prim_offset, int *totprim,
               int *uniq_verts, int *face_verts)
{
  PBVH *p = s->pbvh;
  for (int i = 0; i < p->totnode; i++) {
    const PBVHNode *n = &p->nodes[i];
    memcpy(vb + 6 * i, n->vb.bmin, sizeof(float[3]));
    memcpy(vb + 6 * i + 3, n->vb.bmax, sizeof(float[3]));
    memcpy(orig_vb + 6 * i, n->orig_vb.bmin, sizeof(float[3]));
    memcpy(orig_vb + 6 * i + 3, n->orig_vb.bmax, sizeof(float[3]));
    children_offset[i] = n->children_offset;
    flag[i] = (int)n->flag;
    const int leaf = (n->flag & PBVH_Leaf) != 0;
    prim_offset[i] = leaf ? (int)(n->prim_indices - p->prim_indices) : 0;
    totprim[i] = leaf ? (int)n->totprim : 0;
    /* through the reference's accessor (pbvh.c:3749-3781): a grid leaf's count is totprim * grid_area */
    int uniq = 0, tot = 0;
    if (leaf) BKE_pbvh_node_num_verts(p, &p->nodes[i], &uniq, &tot);
    uniq_verts[i] = uniq;
    face_verts[i] = tot - uniq;
  }
}
void ref_prim_indices(RefSession *s, int *out) { memcpy(out, s->pbvh->prim_indices, sizeof(int) * (size_t)s->pbvh->totprim); }
void ref_node_vert_indices(RefSession *s, int node, int *out)
{
  const PBVHNode *n = &s->pbvh->nodes[node];
  memcpy(out, n->vert_indices, sizeof(int) * (size_t)(n->uniq_verts + n->face_verts));
}
void ref_node_face_vert_indices(RefSession *s, int node, int *out)
{
  const PBVHNode *n = &s->pbvh->nodes[node];
  memcpy(out, n->face_vert_indices, sizeof(int[3]) * (size_t)n->totprim);
}

/* dagger: SCULPT_search_sphere_cb over the reference's node accessors */
typedef struct RefSphere {
  float center[3], radius_sq;
  int original, ignore_ineffective;
} RefSphere;
static bool ref_sphere_cb(PBVHNode *node, void *data_v)
{
  RefSphere *d = (RefSphere *)data_v;
  float bb_min[3], bb_max[3], t[3];
  if (d->ignore_ineffective && (node->flag & PBVH_Leaf) && (node->flag & (PBVH_FullyHidden | PBVH_FullyMasked))) return false;
  if (d->original) BKE_pbvh_node_get_original_BB(node, bb_min, bb_max);
  else BKE_pbvh_node_get_BB(node, bb_min, bb_max);
  for (int i = 0; i < 3; i++) {
    float nearest;
    if (bb_min[i] > d->center[i]) nearest = bb_min[i];
    else if (bb_max[i] < d->center[i]) nearest = bb_max[i];
    else nearest = d->center[i];
    t[i] = d->center[i] - nearest;
  }
  return (t[0] * t[0] + t[1] * t[1] + t[2] * t[2]) < d->radius_sq;
}
int ref_gather_sphere(RefSession *s, const float center[3], float radius_sq, int original, int ignore, int *out)
{
  RefSphere d = {{center[0], center[1], center[2]}, radius_sq, original, ignore};
  PBVHNode **nodes = NULL;
  int tot = 0;
  BKE_pbvh_search_gather(s->pbvh, ref_sphere_cb, &d, &nodes, &tot);
  for (int i = 0; i < tot; i++) out[i] = (int)(nodes[i] - s->pbvh->nodes);
  MEM_SAFE_FREE(nodes);
  return tot;
}
int ref_gather_flag(RefSession *s, int flag, int *out)
{
  PBVHNode **nodes = NULL;
  int tot = 0;
  BKE_pbvh_search_gather(s->pbvh, ref_update_search_cb, POINTER_FROM_INT(flag), &nodes, &tot);
  for (int i = 0; i < tot; i++) out[i] = (int)(nodes[i] - s->pbvh->nodes);
  MEM_SAFE_FREE(nodes);
  return tot;
}

void ref_set_co(RefSession *s, const float *co)
{
  if (s->verts) {
    for (int i = 0; i < s->totvert; i++) memcpy(s->verts[i].co, co + 3 * i, sizeof(float[3]));
  }
  else {
    const int ef = s->key.elem_size / (int)sizeof(float);
    const size_t n = (size_t)s->totgrid * (size_t)s->key.grid_area;
    for (size_t i = 0; i < n; i++) memcpy(s->elems + i * (size_t)ef, co + 3 * i, sizeof(float[3]));
  }
}
void ref_get_no(RefSession *s, float *no)
{
  if (s->verts) {
    memcpy(no, s->vert_normals, sizeof(float[3]) * (size_t)s->totvert);
  }
  else {
    const int ef = s->key.elem_size / (int)sizeof(float), off = s->key.normal_offset / (int)sizeof(float);
    const size_t n = (size_t)s->totgrid * (size_t)s->key.grid_area;
    for (size_t i = 0; i < n; i++) memcpy(no + 3 * i, s->elems + i * (size_t)ef + off, sizeof(float[3]));
  }
}
void ref_set_no(RefSession *s, const float *no) { memcpy(s->vert_normals, no, sizeof(float[3]) * (size_t)s->totvert); }
void ref_node_mark_update(RefSession *s, int node) { BKE_pbvh_node_mark_update(&s->pbvh->nodes[node]); }
void ref_node_set_flag(RefSession *s, int node, int flag, int on)
{
  if (on) s->pbvh->nodes[node].flag |= flag;
  else s->pbvh->nodes[node].flag &= ~flag;
}
void ref_vert_mark_update(RefSession *s, int v) { BKE_pbvh_vert_mark_update(s->pbvh, v); }
int ref_vert_marked(RefSession *s, int v) { return BLI_BITMAP_TEST(s->pbvh->vert_bitmap, v) != 0; }

/* BKE_pbvh_update_normals (pbvh.c:4559-4587), PBVH_FACES branch: gather the flagged leaves, pbvh_faces_update_normals */
void ref_update_normals(RefSession *s)
{
  PBVHNode **nodes = NULL;
  int totnode = 0;
  BKE_pbvh_search_gather(s->pbvh, ref_update_search_cb, POINTER_FROM_INT(PBVH_UpdateNormals), &nodes, &totnode);
  if (totnode > 0) ref_faces_update_normals(s->pbvh, nodes, totnode);
  MEM_SAFE_FREE(nodes);
}
void ref_update_bounds(RefSession *s, int flag) { BKE_pbvh_update_bounds(s->pbvh, flag); }

static void ref_ccg_base(RefSession *s)
{
  s->ccg.grid_size = s->key.grid_size;
  s->ccg.num_grids = s->totgrid;
  s->ccg.grids = s->grids;
  s->ccg.has_normal = true;
  s->ccg.has_mask = s->key.has_mask != 0;
}
/* subdiv_ccg_recalc_inner_face_normals + subdiv_ccg_average_inner_face_normals (subdiv_ccg.c:670-740) on every grid */
void ref_grids_inner_normals(RefSession *s)
{
  ref_ccg_base(s);
  for (int g = 0; g < s->totgrid; g++) ref_ccg_inner_normals(&s->ccg, &s->key, g);
}

/* The SubdivCCG adjacency from the flat tables the oracle and the product take (element indices -> SubdivCCGCoord):
 * faces (start_grid_index, num_grids), adjacent_edges (boundary_coords[face][2 * grid_size]), adjacent_vertices
 * (corner_coords[face]) -- the layout subdiv_ccg.c:397-530 builds. */
void ref_grids_set_adjacency(RefSession *s, int totface, const int *face_start, const int *face_num, int totedge, const int *edge_off,
                             const int *edge_elems, int totcvert, const int *cvert_off, const int *cvert_elems)
{
  ref_ccg_base(s);
  const int gs = s->key.grid_size, area = s->key.grid_area, gs2 = 2 * gs;
  s->ccg.num_faces = totface;
  s->ccg.faces = (SubdivCCGFace *)calloc((size_t)totface + 1, sizeof(SubdivCCGFace));
  for (int f = 0; f < totface; f++) {
    s->ccg.faces[f].start_grid_index = face_start[f];
    s->ccg.faces[f].num_grids = face_num[f];
  }
  const size_t ncoord = (size_t)edge_off[totedge] * (size_t)gs2 + (size_t)cvert_off[totcvert];
  s->coord_store = (SubdivCCGCoord *)calloc(ncoord + 1, sizeof(SubdivCCGCoord));
  s->coord_rows = (SubdivCCGCoord **)calloc((size_t)edge_off[totedge] + 1, sizeof(SubdivCCGCoord *));
  size_t at = 0;
  s->ccg.num_adjacent_edges = totedge;
  s->ccg.adjacent_edges = (SubdivCCGAdjacentEdge *)calloc((size_t)totedge + 1, sizeof(SubdivCCGAdjacentEdge));
  for (int e = 0; e < totedge; e++) {
    s->ccg.adjacent_edges[e].num_adjacent_faces = edge_off[e + 1] - edge_off[e];
    s->ccg.adjacent_edges[e].boundary_coords = s->coord_rows + edge_off[e];
    for (int r = edge_off[e]; r < edge_off[e + 1]; r++) {
      s->coord_rows[r] = s->coord_store + at;
      for (int i = 0; i < gs2; i++, at++) {
        const int el = edge_elems[(size_t)r * (size_t)gs2 + (size_t)i];
        s->coord_store[at].grid_index = el / area;
        s->coord_store[at].y = (short)((el % area) / gs);
        s->coord_store[at].x = (short)((el % area) % gs);
      }
    }
  }
  s->ccg.num_adjacent_vertices = totcvert;
  s->ccg.adjacent_vertices = (SubdivCCGAdjacentVertex *)calloc((size_t)totcvert + 1, sizeof(SubdivCCGAdjacentVertex));
  for (int v = 0; v < totcvert; v++) {
    s->ccg.adjacent_vertices[v].num_adjacent_faces = cvert_off[v + 1] - cvert_off[v];
    s->ccg.adjacent_vertices[v].corner_coords = s->coord_store + at;
    for (int r = cvert_off[v]; r < cvert_off[v + 1]; r++, at++) {
      const int el = cvert_elems[r];
      s->coord_store[at].grid_index = el / area;
      s->coord_store[at].y = (short)((el % area) / gs);
      s->coord_store[at].x = (short)((el % area) % gs);
    }
  }
}
/* subdiv_ccg_average_inner_face_grids / _grids_boundary / _grids_corners over everything, in index order: what
 * KERNEL_subdiv_ccg_average_grids (subdiv_ccg.c:1170-1189) runs through its task ranges */
void ref_grids_average_all(RefSession *s)
{
  for (int f = 0; f < s->ccg.num_faces; f++) ref_ccg_average_face(&s->ccg, &s->key, f);
  for (int e = 0; e < s->ccg.num_adjacent_edges; e++) ref_ccg_average_edge(&s->ccg, &s->key, e);
  for (int v = 0; v < s->ccg.num_adjacent_vertices; v++) ref_ccg_average_cvert(&s->ccg, &s->key, v);
}
void ref_get_co(RefSession *s, float *co)
{
  const int ef = s->key.elem_size / (int)sizeof(float);
  const size_t n = (size_t)s->totgrid * (size_t)s->key.grid_area;
  for (size_t i = 0; i < n; i++) memcpy(co + 3 * i, s->elems + i * (size_t)ef, sizeof(float[3]));
}
void ref_get_mask(RefSession *s, float *mask)
{
  const int ef = s->key.elem_size / (int)sizeof(float);
  const size_t n = (size_t)s->totgrid * (size_t)s->key.grid_area;
  for (size_t i = 0; i < n; i++) mask[i] = s->elems[i * (size_t)ef + 3];
}
