/* ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h).  Mesh helpers: tessellation and maps. */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* lib/intern/math_geom.cc:5362-5378 is_quad_flip_v3_first_third_fast */
static int quad_flip_first_third(const float v1[3], const float v2[3], const float v3[3], const float v4[3])
{
  float d12[3], d13[3], d14[3], ca[3], cb[3];
  for (int i = 0; i < 3; i++) {
    d12[i] = v2[i] - v1[i];
    d13[i] = v3[i] - v1[i];
    d14[i] = v4[i] - v1[i];
  }
  ca[0] = d12[1] * d13[2] - d12[2] * d13[1];
  ca[1] = d12[2] * d13[0] - d12[0] * d13[2];
  ca[2] = d12[0] * d13[1] - d12[1] * d13[0];
  cb[0] = d14[1] * d13[2] - d14[2] * d13[1];
  cb[1] = d14[2] * d13[0] - d14[0] * d13[2];
  cb[2] = d14[0] * d13[1] - d14[1] * d13[0];
  return (ca[0] * cb[0] + ca[1] * cb[1] + ca[2] * cb[2]) > 0.0f;
}

int or_looptri_count(int totpoly, const int *poly_len)
{
  int n = 0;
  for (int i = 0; i < totpoly; i++) {
    n += poly_len[i] - 2;
  }
  return n;
}

/* kernel/intern/mesh_tessellate.c:420-447: tri -> (0,1,2); quad -> (0,1,2),(0,2,3) with the
 * degenerate flip.  N-gons use a fan here (the reference runs polyfill; no config uses n-gons). */
void or_looptri_calc(int totpoly, const int *poly_start, const int *poly_len, const int *loop_v,
                     const float (*co)[3], int (*tri)[3], int *tri_poly)
{
  int t = 0;
  for (int p = 0; p < totpoly; p++) {
    const int ls = poly_start[p], n = poly_len[p];
    if (n == 3) {
      tri[t][0] = ls; tri[t][1] = ls + 1; tri[t][2] = ls + 2;
      tri_poly[t++] = p;
    }
    else if (n == 4) {
      int a = t, b = t + 1;
      tri[a][0] = ls; tri[a][1] = ls + 1; tri[a][2] = ls + 2;
      tri[b][0] = ls; tri[b][1] = ls + 2; tri[b][2] = ls + 3;
      tri_poly[a] = tri_poly[b] = p;
      if (quad_flip_first_third(co[loop_v[tri[a][0]]], co[loop_v[tri[a][1]]], co[loop_v[tri[a][2]]],
                                co[loop_v[tri[b][2]]])) {
        tri[a][2] = tri[b][2];
        tri[b][0] = tri[a][1];
      }
      t += 2;
    }
    else {
      for (int k = 1; k + 1 < n; k++) {
        tri[t][0] = ls; tri[t][1] = ls + k; tri[t][2] = ls + k + 1;
        tri_poly[t++] = p;
      }
    }
  }
}

/* kernel/intern/mesh_mapping.c:182-229 mesh_vert_poly_or_loop_map_create (do_loops = false):
 * count, prefix, fill -- polys of a vertex come out in ascending poly index. */
void or_vert_poly_map(int totvert, int totpoly, const int *poly_start, const int *poly_len,
                      const int *loop_v, int *off, int *idx)
{
  int *count = calloc((size_t)totvert + 1, sizeof(int));
  for (int i = 0; i < totpoly; i++) {
    for (int j = 0; j < poly_len[i]; j++) {
      count[loop_v[poly_start[i] + j]]++;
    }
  }
  off[0] = 0;
  for (int v = 0; v < totvert; v++) {
    off[v + 1] = off[v] + count[v];
    count[v] = 0;
  }
  for (int i = 0; i < totpoly; i++) {
    for (int j = 0; j < poly_len[i]; j++) {
      int v = loop_v[poly_start[i] + j];
      idx[off[v] + count[v]] = i;
      count[v]++;
    }
  }
  free(count);
}

/* DAGGER vertex neighbours the way the upstream neighbour iterator lists them for a mesh PBVH:
 * for every poly of the vertex in pmap order (mesh_mapping.c:182-229), the previous and the next
 * corner (kernel/intern/mesh.c:1566-1589 poly_get_adj_loops_from_vert), de-duplicated keeping the
 * first occurrence.  Boundary vertex = endpoint of an edge used by fewer than two polys. */
int or_vert_neighbors(int totvert, int totpoly, const int *poly_start, const int *poly_len,
                      const int *loop_v, int *off, int *idx, unsigned char *boundary)
{
  int totloop = 0;
  for (int i = 0; i < totpoly; i++) {
    totloop += poly_len[i];
  }
  int *pm_off = malloc(sizeof(int) * ((size_t)totvert + 1));
  int *pm_idx = malloc(sizeof(int) * (size_t)(totloop > 0 ? totloop : 1));
  or_vert_poly_map(totvert, totpoly, poly_start, poly_len, loop_v, pm_off, pm_idx);

  /* how often each neighbour pair (v -> u) shows up = number of polys using edge (v,u) */
  int *paircount = malloc(sizeof(int) * (size_t)(2 * totloop + 1));
  memset(boundary, 0, (size_t)totvert);
  int n = 0;
  for (int v = 0; v < totvert; v++) {
    off[v] = n;
    for (int k = pm_off[v]; k < pm_off[v + 1]; k++) {
      const int p = pm_idx[k];
      const int ls = poly_start[p], len = poly_len[p];
      int corner = -1;
      for (int j = 0; j < len; j++) {
        if (loop_v[ls + j] == v) {
          corner = j;
          break;
        }
      }
      if (corner == -1) {
        continue;
      }
      int adj[2];
      adj[0] = loop_v[ls + (corner + len - 1) % len];
      adj[1] = loop_v[ls + (corner + 1) % len];
      for (int j = 0; j < 2; j++) {
        if (adj[j] == v) {
          continue;
        }
        int found = -1;
        for (int q = off[v]; q < n; q++) {
          if (idx[q] == adj[j]) {
            found = q;
            break;
          }
        }
        if (found < 0) {
          idx[n] = adj[j];
          paircount[n] = 1;
          n++;
        }
        else {
          paircount[found]++;
        }
      }
    }
    for (int q = off[v]; q < n; q++) {
      if (paircount[q] < 2) {
        boundary[v] = 1;
        boundary[idx[q]] = 1;
      }
    }
  }
  off[totvert] = n;
  free(paircount);
  free(pm_off);
  free(pm_idx);
  return n;
}
