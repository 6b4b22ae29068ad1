/* ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Multires grids: the CCG side of the sculpt path.  Restates kernel/intern/subdiv_ccg.c of the
 * reference over flat tables (the OpenSubdiv topology refiner the reference queries for face / edge /
 * vertex incidence, subdiv_ccg.c:402-431, 497-517, 1198-1223, is replaced by tables the caller
 * builds for its base mesh -- SURVEY.md section 8c):
 *   element normals of a grid        subdiv_ccg.c:670-740
 *   inner-face averaging             subdiv_ccg.c:951-984
 *   coarse-edge averaging            subdiv_ccg.c:1010-1048
 *   coarse-vertex averaging          subdiv_ccg.c:1081-1104
 *   update_normals (affected faces)  subdiv_ccg.c:797-866, 1237-1281
 *   stitch after displacement        subdiv_ccg.c:1303-1324  (kernel/intern/multires.c:1171-1196)
 *   full average / recalc            subdiv_ccg.c:1170-1189, 782-790
 *   faces of flagged nodes           pbvh.c:3523-3566
 * The reference runs each of these phases as a parallel loop over faces / edges / vertices whose
 * tasks write disjoint elements, so the result does not depend on the schedule; the phases run in
 * the order of the reference.
 */
#include "oracle_intern.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int elem_index(const OrPbvh *p, int grid, int x, int y)
{
  return grid * p->grid_size * p->grid_size + y * p->grid_size + x;
}

/* lib/intern/math_geom.cc:51-69 normal_quad_v3 + math_vector_inline.c:1165-1181 normalize_v3 */
static void normal_quad(float n[3], const float v1[3], const float v2[3], const float v3[3], const float v4[3])
{
  float n1[3], n2[3];
  n1[0] = v1[0] - v3[0]; n1[1] = v1[1] - v3[1]; n1[2] = v1[2] - v3[2];
  n2[0] = v2[0] - v4[0]; n2[1] = v2[1] - v4[1]; n2[2] = v2[2] - v4[2];
  n[0] = n1[1] * n2[2] - n1[2] * n2[1];
  n[1] = n1[2] * n2[0] - n1[0] * n2[2];
  n[2] = n1[0] * n2[1] - n1[1] * n2[0];
  float d = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
  if (d > 1.0e-35f) {
    d = sqrtf(d);
    const float f = 1.0f / d;
    n[0] = n[0] * f; n[1] = n[1] * f; n[2] = n[2] * f;
  }
  else {
    n[0] = n[1] = n[2] = 0.0f;
  }
}

/* subdiv_ccg.c:670-740: quad normals of the grid, then every element = mean (not renormalised) of
 * its adjacent quads */
static void grid_inner_normals(OrPbvh *p, int grid, float (*fn)[3])
{
  const int gs = p->grid_size, gs1 = gs - 1;
  for (int y = 0; y < gs1; y++) {
    for (int x = 0; x < gs1; x++) {
      normal_quad(fn[y * gs1 + x], p->co[elem_index(p, grid, x, y + 1)], p->co[elem_index(p, grid, x + 1, y + 1)],
                  p->co[elem_index(p, grid, x + 1, y)], p->co[elem_index(p, grid, x, y)]);
    }
  }
  for (int y = 0; y < gs; y++) {
    for (int x = 0; x < gs; x++) {
      float acc[3] = {0.0f, 0.0f, 0.0f};
      int counter = 0;
      if (x < gs1 && y < gs1) {
        const float *f = fn[y * gs1 + x];
        acc[0] += f[0]; acc[1] += f[1]; acc[2] += f[2];
        counter++;
      }
      if (x >= 1) {
        if (y < gs1) {
          const float *f = fn[y * gs1 + (x - 1)];
          acc[0] += f[0]; acc[1] += f[1]; acc[2] += f[2];
          counter++;
        }
        if (y >= 1) {
          const float *f = fn[(y - 1) * gs1 + (x - 1)];
          acc[0] += f[0]; acc[1] += f[1]; acc[2] += f[2];
          counter++;
        }
      }
      if (y >= 1 && x < gs1) {
        const float *f = fn[(y - 1) * gs1 + x];
        acc[0] += f[0]; acc[1] += f[1]; acc[2] += f[2];
        counter++;
      }
      float *no = p->no[elem_index(p, grid, x, y)];
      const float s = 1.0f / (float)counter;
      no[0] = acc[0] * s; no[1] = acc[1] * s; no[2] = acc[2] * s;
    }
  }
}

/* subdiv_ccg.c:873-897 average_grid_element */
static void average_pair(OrPbvh *p, int a, int b)
{
  for (int k = 0; k < 3; k++) {
    float v = p->co[a][k] + p->co[b][k];
    v = v * 0.5f;
    p->co[a][k] = v;
    p->co[b][k] = v;
  }
  for (int k = 0; k < 3; k++) {
    float v = p->no[a][k] + p->no[b][k];
    v = v * 0.5f;
    p->no[a][k] = v;
    p->no[b][k] = v;
  }
  if (p->mask) {
    const float m = (p->mask[a] + p->mask[b]) * 0.5f;
    p->mask[a] = m;
    p->mask[b] = m;
  }
}

/* subdiv_ccg.c:899-949: accumulate in list order, scale by 1 / n, copy to all */
static void average_list(OrPbvh *p, const int *elems, int n, int stride)
{
  float co[3] = {0.0f, 0.0f, 0.0f}, no[3] = {0.0f, 0.0f, 0.0f}, mask = 0.0f;
  for (int i = 0; i < n; i++) {
    const int e = elems[(size_t)i * stride];
    for (int k = 0; k < 3; k++) {
      co[k] += p->co[e][k];
      no[k] += p->no[e][k];
    }
    if (p->mask) mask += p->mask[e];
  }
  const float f = 1.0f / (float)n;
  for (int k = 0; k < 3; k++) {
    co[k] *= f;
    no[k] *= f;
  }
  mask *= f;
  for (int i = 0; i < n; i++) {
    const int e = elems[(size_t)i * stride];
    for (int k = 0; k < 3; k++) {
      p->co[e][k] = co[k];
      p->no[e][k] = no[k];
    }
    if (p->mask) p->mask[e] = mask;
  }
}

/* subdiv_ccg.c:951-984 subdiv_ccg_average_inner_face_grids */
static void average_inner_face_grids(OrPbvh *p, int face)
{
  const int n = p->face_num[face], start = p->face_start[face], gs = p->grid_size;
  int prev_grid = start + n - 1;
  for (int corner = 0; corner < n; corner++) {
    const int grid = start + corner;
    for (int i = 1; i < gs; i++) {
      average_pair(p, elem_index(p, prev_grid, i, 0), elem_index(p, grid, 0, i));
    }
    prev_grid = grid;
  }
  int centers[64];
  int *c = n <= 64 ? centers : malloc(sizeof(int) * (size_t)n);
  for (int corner = 0; corner < n; corner++) c[corner] = elem_index(p, start + corner, 0, 0);
  average_list(p, c, n, 1);
  if (c != centers) free(c);
}

/* subdiv_ccg.c:1010-1048 subdiv_ccg_average_grids_boundary */
static void average_edge(OrPbvh *p, int edge)
{
  const int nf = p->edge_off[edge + 1] - p->edge_off[edge];
  const int gs2 = p->grid_size * 2;
  if (nf == 1) return;
  const int *base = p->edge_elems + (size_t)p->edge_off[edge] * gs2;
  for (int i = 1; i < gs2 - 1; i++) average_list(p, base + i, nf, gs2);
}

/* subdiv_ccg.c:1081-1104 subdiv_ccg_average_grids_corners */
static void average_cvert(OrPbvh *p, int v)
{
  const int nf = p->cvert_off[v + 1] - p->cvert_off[v];
  if (nf == 1) return;
  average_list(p, p->cvert_elems + p->cvert_off[v], nf, 1);
}

/* pbvh.c:3523-3566 BKE_pbvh_get_grid_updates: the faces of the grids of flagged leaves */
int or_grids_get_updates(OrPbvh *p, int clear, int *r_faces)
{
  int tot = 0;
  p->stamp++;
  for (int n = 0; n < p->totnode; n++) {
    OrNode *node = &p->nodes[n];
    if (!(node->flag & OR_PBVH_Leaf) || !(node->flag & OR_PBVH_UpdateNormals)) continue;
    for (int i = 0; i < node->totprim; i++) {
      const int f = p->grid_face[p->prim_indices[node->prim_offset + i]];
      if (p->face_stamp[f] != p->stamp) {
        p->face_stamp[f] = p->stamp;
        r_faces[tot++] = f;
      }
    }
    if (clear) node->flag &= ~(unsigned)OR_PBVH_UpdateNormals;
  }
  return tot;
}

/* subdiv_ccg.c:1237-1281 subdiv_ccg_average_faces_boundaries_and_corners */
static void average_faces_boundaries_and_corners(OrPbvh *p, const int *faces, int totface)
{
  p->stamp++;
  /* boundaries, then corners (two passes over the affected faces, each set de-duplicated) */
  for (int i = 0; i < totface; i++) {
    const int f = faces[i];
    for (int c = 0; c < p->face_num[f]; c++) {
      const int e = p->grid_edge[p->face_start[f] + c];
      if (p->edge_stamp[e] != p->stamp) {
        p->edge_stamp[e] = p->stamp;
        average_edge(p, e);
      }
    }
  }
  for (int i = 0; i < totface; i++) {
    const int f = faces[i];
    for (int c = 0; c < p->face_num[f]; c++) {
      const int v = p->grid_cvert[p->face_start[f] + c];
      if (p->cvert_stamp[v] != p->stamp) {
        p->cvert_stamp[v] = p->stamp;
        average_cvert(p, v);
      }
    }
  }
}

/* subdiv_ccg.c:847-866 KERNEL_subdiv_ccg_update_normals */
void or_grids_update_normals(OrPbvh *p, const int *faces, int totface)
{
  if (totface == 0) return;
  const int gs1 = p->grid_size - 1;
#pragma omp parallel if (or_threads > 1 && totface > 1)
  {
    float(*fn)[3] = malloc(sizeof(float[3]) * (size_t)(gs1 * gs1 > 0 ? gs1 * gs1 : 1));
#pragma omp for schedule(dynamic)
    for (int i = 0; i < totface; i++) {
      const int f = faces[i];
      for (int c = 0; c < p->face_num[f]; c++) grid_inner_normals(p, p->face_start[f] + c, fn);
      average_inner_face_grids(p, f);
    }
    free(fn);
  }
  average_faces_boundaries_and_corners(p, faces, totface);
}

/* subdiv_ccg.c:1303-1324 KERNEL_subdiv_ccg_average_stitch_faces: the affected faces' inner
 * boundaries, then ALL coarse edges and ALL coarse vertices (the TODO at :1321-1323) */
void or_grids_stitch_faces(OrPbvh *p, const int *faces, int totface)
{
#pragma omp parallel for schedule(dynamic) if (or_threads > 1 && totface > 1)
  for (int i = 0; i < totface; i++) average_inner_face_grids(p, faces[i]);
#pragma omp parallel for schedule(dynamic) if (or_threads > 1)
  for (int e = 0; e < p->totedge; e++) average_edge(p, e);
#pragma omp parallel for schedule(dynamic) if (or_threads > 1)
  for (int v = 0; v < p->totcvert; v++) average_cvert(p, v);
}

/* subdiv_ccg.c:1170-1189 KERNEL_subdiv_ccg_average_grids */
void or_grids_average_all(OrPbvh *p)
{
  for (int f = 0; f < p->totface; f++) average_inner_face_grids(p, f);
  for (int e = 0; e < p->totedge; e++) average_edge(p, e);
  for (int v = 0; v < p->totcvert; v++) average_cvert(p, v);
}

void or_grids_inner_normals(OrPbvh *p)
{
  const int gs1 = p->grid_size - 1;
  float(*fn)[3] = malloc(sizeof(float[3]) * (size_t)(gs1 * gs1 > 0 ? gs1 * gs1 : 1));
  for (int g = 0; g < p->totgrid; g++) grid_inner_normals(p, g, fn);
  free(fn);
}

/* subdiv_ccg.c:782-790 KERNEL_subdiv_ccg_recalc_normals */
void or_grids_recalc_normals(OrPbvh *p)
{
  const int gs1 = p->grid_size - 1;
  float(*fn)[3] = malloc(sizeof(float[3]) * (size_t)(gs1 * gs1 > 0 ? gs1 * gs1 : 1));
  for (int g = 0; g < p->totgrid; g++) grid_inner_normals(p, g, fn);
  free(fn);
  or_grids_average_all(p);
}

/* ---- element neighbours (smooth brush on grids) -------------------------
 * KERNEL_subdiv_ccg_neighbor_coords_get (subdiv_ccg.c:1882-1909) with include_duplicates = false, the way
 * the brushes' neighbour iterator calls it, over the flat tables.  The three OpenSubdiv queries it makes
 * -- getFaceEdges / getFaceVertices (grid_edge / grid_cvert, already held), getEdgeVertices and
 * getVertexEdges (subdiv_ccg.c:1558-1582, 1649-1651) -- are tables the caller gives through
 * or_grids_set_topology.  Neighbours come out in the reference's order (the average sums them in it). */
void or_grids_set_topology(OrPbvh *p, const int *edge_verts, const int *cvert_edge_off, const int *cvert_edges)
{
  free(p->edge_verts); free(p->cvert_edge_off); free(p->cvert_edges); free(p->cvert_boundary);
  p->edge_verts = malloc(sizeof(int) * 2 * (size_t)(p->totedge + 1));
  memcpy(p->edge_verts, edge_verts, sizeof(int) * 2 * (size_t)p->totedge);
  p->cvert_edge_off = malloc(sizeof(int) * (size_t)(p->totcvert + 1));
  memcpy(p->cvert_edge_off, cvert_edge_off, sizeof(int) * (size_t)(p->totcvert + 1));
  p->cvert_edges = malloc(sizeof(int) * (size_t)(cvert_edge_off[p->totcvert] + 1));
  memcpy(p->cvert_edges, cvert_edges, sizeof(int) * (size_t)cvert_edge_off[p->totcvert]);
  /* DAGGER boundary vertices of the base mesh (upstream's vertex_info.boundary): both ends of every
   * coarse edge with fewer than two faces */
  p->cvert_boundary = calloc((size_t)p->totcvert + 1, 1);
  p->max_neighbors = 0;
  p->max_neighbors = or_grids_max_neighbors(p);
  if (!p->scratch) { /* Jacobi buffers of the smooth brush */
    p->scratch = malloc(sizeof(float[3]) * ((size_t)p->totvert + 1));
    p->iter_flag = calloc((size_t)p->totvert + 1, 1);
  }
  for (int e = 0; e < p->totedge; e++) {
    if (p->edge_off[e + 1] - p->edge_off[e] < 2) {
      p->cvert_boundary[edge_verts[2 * e]] = 1;
      p->cvert_boundary[edge_verts[2 * e + 1]] = 1;
    }
  }
}

/* subdiv_ccg.c:1448-1471 coord_step_inside_from_boundary, on an element index */
static int step_inside(const OrPbvh *p, int elem)
{
  const int gs = p->grid_size, gs2 = gs * gs, gs1 = gs - 1;
  const int g = elem / gs2, y = (elem % gs2) / gs, x = (elem % gs2) % gs;
  if (x == gs1) return elem_index(p, g, x - 1, y);
  if (y == gs1) return elem_index(p, g, x, y - 1);
  if (x == 0) return elem_index(p, g, x + 1, y);
  return elem_index(p, g, x, y + 1);
}

/* subdiv_ccg.c:1714-1772 neighbor_coords_edge_get: the element lies on a coarse edge (not at a coarse vertex) */
static int neighbors_on_edge(const OrPbvh *p, int g, int x, int y, int *r)
{
  const int gs = p->grid_size, gs1 = gs - 1;
  const int f = p->grid_face[g], start = p->face_start[f], num = p->face_num[f], c = g - start;
  /* adjacent_edge_index_from_coord (1612-1641) */
  const int e = (x == gs1) ? p->grid_edge[start + c] : p->grid_edge[start + (c == 0 ? num - 1 : c - 1)];
  const int nf = p->edge_off[e + 1] - p->edge_off[e];
  /* adjacent_edge_point_index_from_coord (1643-1679) */
  const int v = p->grid_cvert[g];
  int pt, dirv;
  if (x == gs1) {
    pt = gs - y - 1;
    dirv = p->edge_verts[2 * e];
  }
  else {
    pt = gs + x;
    dirv = p->edge_verts[2 * e + 1];
  }
  if (v != dirv) pt = 2 * gs - pt - 1;
  const int next = (pt == gs - 1) ? pt + 2 : pt + 1; /* 1685-1691 */
  const int prev = (pt == gs) ? pt - 2 : pt - 1;     /* 1692-1698 */
  for (int i = 0; i < nf; i++) {
    const int *row = p->edge_elems + (size_t)(p->edge_off[e] + i) * 2 * (size_t)gs;
    r[i + 2] = step_inside(p, row[pt]);
    if (row[pt] / (gs * gs) == g) {
      r[0] = row[prev];
      r[1] = row[next];
    }
  }
  return nf + 2;
}

int or_grids_neighbors(const OrPbvh *p, int elem, int *r)
{
  const int gs = p->grid_size, gs2 = gs * gs, gs1 = gs - 1;
  const int g = elem / gs2, y = (elem % gs2) / gs, x = (elem % gs2) % gs;
  const int f = p->grid_face[g], start = p->face_start[f], num = p->face_num[f];
  const int corner = (x == 0 || x == gs1) && (y == 0 || y == gs1);
  if (corner) {
    if (x == 0 && y == 0) { /* 1496-1520: the face centre */
      for (int c = 0; c < num; c++) r[c] = elem_index(p, start + c, 1, 0);
      return num;
    }
    if (x == gs1 && y == gs1) { /* 1547-1610: a coarse vertex */
      const int v = p->grid_cvert[g];
      const int ne = p->cvert_edge_off[v + 1] - p->cvert_edge_off[v];
      for (int i = 0; i < ne; i++) {
        const int e = p->cvert_edges[p->cvert_edge_off[v] + i];
        const int pt = (p->edge_verts[2 * e] == v) ? 1 : 2 * gs - 2;
        r[i] = p->edge_elems[(size_t)p->edge_off[e] * 2 * (size_t)gs + pt]; /* first face of the edge */
      }
      return ne;
    }
    return neighbors_on_edge(p, g, x, y, r);
  }
  if (x == 0 || y == 0 || x == gs1 || y == gs1) {
    if (x == 0) { /* 1807-1843: boundary between two grids of one face */
      const int prev = start + ((g - start) == 0 ? num - 1 : g - start - 1);
      r[0] = elem_index(p, g, x, y - 1);
      r[1] = elem_index(p, g, x, y + 1);
      r[2] = elem_index(p, g, x + 1, y);
      r[3] = elem_index(p, prev, y, 1);
      return 4;
    }
    if (y == 0) {
      const int next = start + ((g - start + 1) == num ? 0 : g - start + 1);
      r[0] = elem_index(p, g, x - 1, y);
      r[1] = elem_index(p, g, x + 1, y);
      r[2] = elem_index(p, g, x, y + 1);
      r[3] = elem_index(p, next, 1, x);
      return 4;
    }
    return neighbors_on_edge(p, g, x, y, r);
  }
  r[0] = elem_index(p, g, x, y - 1); /* 1870-1880 */
  r[1] = elem_index(p, g, x, y + 1);
  r[2] = elem_index(p, g, x - 1, y);
  r[3] = elem_index(p, g, x + 1, y);
  return 4;
}

/* DAGGER SCULPT_vertex_is_boundary for grids over KERNEL_subdiv_ccg_coarse_mesh_adjacency_info_get
 * (subdiv_ccg.c:1949-2008): an element at a coarse vertex is a boundary element when that vertex is a boundary
 * vertex of the base mesh, one on a coarse edge when both ends of the edge are; everything else is interior */
int or_grids_is_boundary(const OrPbvh *p, int elem)
{
  const int gs = p->grid_size, gs2 = gs * gs, gs1 = gs - 1;
  const int g = elem / gs2, y = (elem % gs2) / gs, x = (elem % gs2) % gs;
  if (x != gs1 && y != gs1) return 0; /* interior, face centre or an inner boundary */
  const int f = p->grid_face[g], start = p->face_start[f], num = p->face_num[f], c = g - start;
  const int v1 = p->grid_cvert[g];
  if (x == gs1 && y == gs1) return p->cvert_boundary[v1];
  int v2 = v1;
  if (x == gs1) v2 = p->grid_cvert[start + (c + 1) % num];
  if (y == gs1) v2 = p->grid_cvert[start + (c + num - 1) % num];
  return p->cvert_boundary[v1] && p->cvert_boundary[v2];
}

int or_grids_max_neighbors(const OrPbvh *p)
{
  if (p->max_neighbors) return p->max_neighbors;
  int w = 4;
  for (int f = 0; f < p->totface; f++) w = p->face_num[f] > w ? p->face_num[f] : w;
  for (int e = 0; e < p->totedge; e++) w = (p->edge_off[e + 1] - p->edge_off[e] + 2) > w ? (p->edge_off[e + 1] - p->edge_off[e] + 2) : w;
  if (p->cvert_edge_off) {
    for (int v = 0; v < p->totcvert; v++) w = (p->cvert_edge_off[v + 1] - p->cvert_edge_off[v]) > w ? (p->cvert_edge_off[v + 1] - p->cvert_edge_off[v]) : w;
  }
  return w;
}
