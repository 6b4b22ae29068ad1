/* Multires grids (PBVH_GRIDS) on the device: what runs after the brush when the resident mesh is a
 * SubdivCCG -- the stitch of duplicated boundary elements (kernel/intern/multires.c:1171-1196 ->
 * subdiv_ccg.c:1303-1324), the CCG normal update of the faces of the gathered leaves
 * (pbvh.c:3523-3566, subdiv_ccg.c:670-866, 1237-1281) and the leaf boxes.  Gather, area normal and
 * brush are the mesh kernels: a leaf's grid elements are one contiguous run of slots
 * (slot = first slot of the grid + y * grid_size + x), so to those kernels a grid element is a vertex.
 *
 * The reference runs every averaging phase as a parallel loop over faces / coarse edges / coarse
 * vertices whose tasks write disjoint elements; the phases below keep its order, and every sum runs
 * in the order of the reference's lists, so the results are bit-identical to the CPU path. */
#pragma once
#include "dsc_kernels.cuh"

struct GridCounts {
  int faces, edges, cverts, pad;
};

struct DevGrids {
  int gs, gs2, totgrid, totface, totedge, totcvert;
  const int *grid_slot0;              /* [totgrid] slot of element (0, 0) */
  const int *leaf_gbeg, *leaf_grids;  /* [nleaf + 1] into leaf_grids: the grids of a leaf (PBVHNode.prim_indices) */
  const int *face_start, *face_num, *grid_face, *grid_edge, *grid_cvert;
  const int *edge_off, *edge_slots;   /* SubdivCCGAdjacentEdge.boundary_coords as slots: [edge_off[e] + f][2 * gs] */
  const int *cvert_off, *cvert_slots; /* SubdivCCGAdjacentVertex.corner_coords as slots */
  int *face_stamp, *edge_stamp, *cvert_stamp;
  int *face_list, *edge_list, *cvert_list;
  GridCounts *cnt;
  float *mask; /* per-slot mask layer (averaged along with co / no), or NULL */
  int max_face_grids; /* most corners of a face */
  int has_odd_edges;  /* some coarse edge has more than two faces */
  /* partitioned across GPUs (NULL on one): what this rank averages after the brush -- see plan_grids_rank in dsc_api.cu.
   * face_dom bit 0: a grid of the face is owned (all pairs + centre), bit 1: only middle pairs on edges with an owned
   * half, bit 2: only lists an owned coarse vertex; edge_mine bit h: half h of the edge's points; cvert_mine */
  const unsigned char *face_dom, *edge_mine, *cvert_mine;
  const int *grid_owner;
  int rank;
};

/* Every stage below is a __device__ body over (cta, ncta) so that it runs either as its own kernel or as
 * one phase of k_grid_dab, the fused per-dab kernel (phases separated by grid barriers).  What a phase
 * reads that an earlier phase wrote -- lists, counters, element data -- goes through L2 (__ldcg). */

/* BKE_pbvh_get_grid_updates (pbvh.c:3523-3566): the faces of the grids of the listed leaves, each once; the
 * thread that claims a face also claims its coarse edges and vertices (subdiv_ccg_affected_face_adjacency,
 * subdiv_ccg.c:1191-1235) */
__device__ __forceinline__ void dsc_grid_faces_body(const DevGrids &g, const int *list, const int *count, int seq, int cta, int ncta)
{
  const int n = __ldcg(count);
  for (int h = cta; h < n; h += ncta) {
    const int l = __ldcg(&list[h]);
    const int b = g.leaf_gbeg[l], e = g.leaf_gbeg[l + 1];
    for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
      const int f = g.grid_face[g.leaf_grids[i]];
      const int dom = g.face_dom ? g.face_dom[f] : 1;
      if (!dom) continue;
      if (atomicExch(&g.face_stamp[f], seq) == seq) continue;
      if (dom & 3) g.face_list[atomicAdd(&g.cnt->faces, 1)] = f;
      const int start = g.face_start[f], nc = g.face_num[f];
      for (int c = 0; c < nc; c++) {
        const int ed = g.grid_edge[start + c], v = g.grid_cvert[start + c];
        if ((!g.edge_mine || g.edge_mine[ed]) && atomicExch(&g.edge_stamp[ed], seq) != seq) g.edge_list[atomicAdd(&g.cnt->edges, 1)] = ed;
        if ((!g.cvert_mine || g.cvert_mine[v]) && atomicExch(&g.cvert_stamp[v], seq) != seq) g.cvert_list[atomicAdd(&g.cnt->cverts, 1)] = v;
      }
    }
  }
}

/* average_grid_element (subdiv_ccg.c:873-897) */
__device__ __forceinline__ void dsc_grid_average_pair(const DevMesh &m, const DevGrids &g, int a, int b)
{
  float *arr[6] = {m.cx, m.cy, m.cz, m.nx, m.ny, m.nz};
  float va[6], vb[6];
#pragma unroll
  for (int k = 0; k < 6; k++) {
    va[k] = __ldcg(&arr[k][a]);
    vb[k] = __ldcg(&arr[k][b]);
  }
  float ma = 0.0f, mb = 0.0f;
  if (g.mask) {
    ma = __ldcg(&g.mask[a]);
    mb = __ldcg(&g.mask[b]);
  }
#pragma unroll
  for (int k = 0; k < 6; k++) {
    float v = va[k] + vb[k];
    v = v * 0.5f;
    arr[k][a] = v;
    arr[k][b] = v;
  }
  if (g.mask) {
    const float mk = (ma + mb) * 0.5f;
    g.mask[a] = mk;
    g.mask[b] = mk;
  }
}

/* element_accumulator_* (subdiv_ccg.c:899-949): sum in list order, scale by 1 / n, copy to all.  Up to 4
 * members have their slots, then all their values, loaded together (two round trips, not two per member);
 * the sums still run in list order. */
__device__ __forceinline__ void dsc_grid_average_list(const DevMesh &m, const DevGrids &g, const int *slots, int n, int stride)
{
  float acc[7] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  if (n <= 4) {
    int sl[4];
    float v[4][7];
#pragma unroll
    for (int i = 0; i < 4; i++) sl[i] = (i < n) ? slots[(size_t)i * stride] : 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (i < n) {
        v[i][0] = __ldcg(&m.cx[sl[i]]); v[i][1] = __ldcg(&m.cy[sl[i]]); v[i][2] = __ldcg(&m.cz[sl[i]]);
        v[i][3] = __ldcg(&m.nx[sl[i]]); v[i][4] = __ldcg(&m.ny[sl[i]]); v[i][5] = __ldcg(&m.nz[sl[i]]);
        v[i][6] = g.mask ? __ldcg(&g.mask[sl[i]]) : 0.0f;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (i < n) {
#pragma unroll
        for (int k = 0; k < 7; k++) acc[k] += v[i][k];
      }
    }
    const float f = 1.0f / (float)n;
#pragma unroll
    for (int k = 0; k < 7; k++) acc[k] *= f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (i < n) {
        m.cx[sl[i]] = acc[0]; m.cy[sl[i]] = acc[1]; m.cz[sl[i]] = acc[2];
        m.nx[sl[i]] = acc[3]; m.ny[sl[i]] = acc[4]; m.nz[sl[i]] = acc[5];
        if (g.mask) g.mask[sl[i]] = acc[6];
      }
    }
    return;
  }
  for (int i = 0; i < n; i++) {
    const int s = slots[(size_t)i * stride];
    acc[0] += __ldcg(&m.cx[s]); acc[1] += __ldcg(&m.cy[s]); acc[2] += __ldcg(&m.cz[s]);
    acc[3] += __ldcg(&m.nx[s]); acc[4] += __ldcg(&m.ny[s]); acc[5] += __ldcg(&m.nz[s]);
    if (g.mask) acc[6] += __ldcg(&g.mask[s]);
  }
  const float f = 1.0f / (float)n;
#pragma unroll
  for (int k = 0; k < 7; k++) acc[k] *= f;
  for (int i = 0; i < n; i++) {
    const int s = slots[(size_t)i * stride];
    m.cx[s] = acc[0]; m.cy[s] = acc[1]; m.cz[s] = acc[2];
    m.nx[s] = acc[3]; m.ny[s] = acc[4]; m.nz[s] = acc[5];
    if (g.mask) g.mask[s] = acc[6];
  }
}

/* subdiv_ccg_average_inner_face_grids (subdiv_ccg.c:951-984) of the listed faces: one thread per pair of
 * elements (the pairs of a face are disjoint), one per face centre */
__device__ __forceinline__ void dsc_grid_inner_body(const DevMesh &m, const DevGrids &g, int cta, int ncta)
{
  const int n = __ldcg(&g.cnt->faces);
  const int per = g.gs - 1, per_face = g.max_face_grids * per;
  const long long items = (long long)n * per_face;
  for (long long t = (long long)cta * blockDim.x + threadIdx.x; t < items; t += (long long)ncta * blockDim.x) {
    const int h = (int)(t / per_face), r = (int)(t - (long long)h * per_face);
    const int f = __ldcg(&g.face_list[h]);
    const int nc = g.face_num[f], start = g.face_start[f];
    const int corner = r / per, i = 1 + r % per;
    if (corner >= nc) continue;
    const int grid = start + corner, prev = start + (corner + nc - 1) % nc;
    if (g.face_dom && !(g.face_dom[f] & 1)) {
      /* no grid of this face is ours: only the middle pair of an edge we average (it sits on the edge leaving `prev`) */
      if (i != per || !g.edge_mine[g.grid_edge[prev]]) continue;
    }
    dsc_grid_average_pair(m, g, g.grid_slot0[prev] + i, g.grid_slot0[grid] + i * g.gs);
  }
  for (int h = cta * blockDim.x + threadIdx.x; h < n; h += ncta * blockDim.x) {
    const int f = __ldcg(&g.face_list[h]);
    if (g.face_dom && !(g.face_dom[f] & 1)) continue;
    dsc_grid_average_list(m, g, g.grid_slot0 + g.face_start[f], g.face_num[f], 1);
  }
}

/* subdiv_ccg_average_grids_boundary (subdiv_ccg.c:1010-1048): listed coarse edges (all == 0), every edge
 * (all == 2), or every edge that is not in the list and has more than two faces (all == 1: averaging two
 * equal values is exact, so untouched two-face edges need no pass); one thread per
 * boundary element */
__device__ __forceinline__ void dsc_grid_edges_body(const DevMesh &m, const DevGrids &g, int all, int seq, int cta, int ncta)
{
  const int n = all ? g.totedge : __ldcg(&g.cnt->edges);
  const int gs2 = 2 * g.gs, per = gs2 - 2;
  const long long items = (long long)n * per;
  for (long long t = (long long)cta * blockDim.x + threadIdx.x; t < items; t += (long long)ncta * blockDim.x) {
    const int h = (int)(t / per), i = 1 + (int)(t - (long long)h * per);
    const int e = all ? h : __ldcg(&g.edge_list[h]);
    const int nf = g.edge_off[e + 1] - g.edge_off[e];
    if (nf == 1) continue;
    if (all == 1 && (nf == 2 || g.edge_stamp[e] == seq)) continue;
    if (g.edge_mine && !((g.edge_mine[e] >> (i >= g.gs ? 1 : 0)) & 1)) continue;
    dsc_grid_average_list(m, g, g.edge_slots + (size_t)g.edge_off[e] * gs2 + i, nf, gs2);
  }
}

/* subdiv_ccg_average_grids_corners (subdiv_ccg.c:1081-1104): listed coarse vertices, or all of them */
__device__ __forceinline__ void dsc_grid_cverts_body(const DevMesh &m, const DevGrids &g, int all, int cta, int ncta)
{
  const int n = all ? g.totcvert : __ldcg(&g.cnt->cverts);
  for (int i = cta * blockDim.x + threadIdx.x; i < n; i += ncta * blockDim.x) {
    const int v = all ? i : __ldcg(&g.cvert_list[i]);
    const int nf = g.cvert_off[v + 1] - g.cvert_off[v];
    if (nf == 1) continue;
    if (g.cvert_mine && !g.cvert_mine[v]) continue;
    dsc_grid_average_list(m, g, g.cvert_slots + g.cvert_off[v], nf, 1);
  }
}

/* subdiv_ccg_recalc_inner_face_normals + subdiv_ccg_average_inner_face_normals (subdiv_ccg.c:670-740)
 * of every grid of the listed faces (all == 0) or of all grids: one CTA per grid, positions and quad
 * normals in shared memory */
#define GN_BLOCK 1024
__host__ __device__ inline size_t dsc_grid_normals_smem(int gs) { return sizeof(float) * 3 * ((size_t)gs * gs + (size_t)(gs - 1) * (gs - 1)); }
__device__ __forceinline__ void dsc_grid_normals_body(const DevMesh &m, const DevGrids &g, int all, float *gsm, int cta, int ncta)
{
  const int gs = g.gs, gs1 = gs - 1, gs2 = g.gs2;
  const int bd = blockDim.x;
  float *P = gsm;            /* [gs2][3] */
  float *Fn = gsm + 3 * gs2; /* [gs1 * gs1][3] */
  const int mfg = g.max_face_grids;
  const int units = all ? g.totgrid : __ldcg(&g.cnt->faces) * mfg;
  for (int u = cta; u < units; u += ncta) {
    int grid;
    if (all) {
      grid = u;
    }
    else {
      const int f = __ldcg(&g.face_list[u / mfg]), c = u % mfg;
      if (c >= g.face_num[f]) continue;
      grid = g.face_start[f] + c;
      if (g.grid_owner && g.grid_owner[grid] != g.rank) continue; /* its owner sends the rim normals */
    }
    const int s0 = g.grid_slot0[grid];
    __syncthreads();
    for (int i = threadIdx.x; i < gs2; i += bd) {
      P[3 * i] = __ldcg(&m.cx[s0 + i]); P[3 * i + 1] = __ldcg(&m.cy[s0 + i]); P[3 * i + 2] = __ldcg(&m.cz[s0 + i]);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < gs1 * gs1; q += bd) {
      const int y = q / gs1, x = q - y * gs1;
      /* normal_quad_v3(co(x, y+1), co(x+1, y+1), co(x+1, y), co(x, y)), subdiv_ccg.c:684-698 */
      const float *v1 = P + 3 * ((y + 1) * gs + x), *v2 = P + 3 * ((y + 1) * gs + x + 1);
      const float *v3 = P + 3 * (y * gs + x + 1), *v4 = P + 3 * (y * gs + x);
      const float n1x = v1[0] - v3[0], n1y = v1[1] - v3[1], n1z = v1[2] - v3[2];
      const float n2x = v2[0] - v4[0], n2y = v2[1] - v4[1], n2z = v2[2] - v4[2];
      float ox = n1y * n2z - n1z * n2y;
      float oy = n1z * n2x - n1x * n2z;
      float oz = n1x * n2y - n1y * n2x;
      dsc_normalize(ox, oy, oz);
      Fn[3 * q] = ox; Fn[3 * q + 1] = oy; Fn[3 * q + 2] = oz;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < gs2; i += bd) {
      const int y = i / gs, x = i - y * gs;
      float ax = 0.0f, ay = 0.0f, az = 0.0f;
      int counter = 0;
      if (x < gs1 && y < gs1) {
        const float *f = Fn + 3 * (y * gs1 + x);
        ax += f[0]; ay += f[1]; az += f[2];
        counter++;
      }
      if (x >= 1) {
        if (y < gs1) {
          const float *f = Fn + 3 * (y * gs1 + (x - 1));
          ax += f[0]; ay += f[1]; az += f[2];
          counter++;
        }
        if (y >= 1) {
          const float *f = Fn + 3 * ((y - 1) * gs1 + (x - 1));
          ax += f[0]; ay += f[1]; az += f[2];
          counter++;
        }
      }
      if (y >= 1 && x < gs1) {
        const float *f = Fn + 3 * ((y - 1) * gs1 + x);
        ax += f[0]; ay += f[1]; az += f[2];
        counter++;
      }
      const float sc = 1.0f / (float)counter;
      m.nx[s0 + i] = ax * sc; m.ny[s0 + i] = ay * sc; m.nz[s0 + i] = az * sc;
    }
  }
}

/* The same pass, element-parallel (opt-in experiment, DSC_GRID_NORMALS_FLAT=1; measured slower on B200: 41.8 vs 39.2 ms per
 * C5 stroke -- the IEEE sqrt / div of a quad normal, kept for bit parity, is paid four times): one thread per element, no shared memory.  An element's normal is the mean of the
 * normals of the <= 4 quads around it, each quad normal recomputed from its four corners by every element that uses it
 * (4 x the arithmetic of the staged version, all of it hidden under the loads; the 3 x 3 neighbourhood of positions
 * comes from L1 / L2 -- a grid row is a unit-stride run).  Same expressions in the same order as
 * dsc_grid_normals_body, so the bits are the same; any number of CTAs, no phase barriers, no per-grid serial tail. */
__device__ __forceinline__ void dsc_grid_quad_normal(const DevMesh &m, int s_xy, int gs, float &ox, float &oy, float &oz)
{
  /* normal_quad_v3(co(x, y+1), co(x+1, y+1), co(x+1, y), co(x, y)), subdiv_ccg.c:684-698; s_xy = slot of (x, y) */
  const int a = s_xy + gs, b = s_xy + gs + 1, c = s_xy + 1, d = s_xy;
  const float n1x = __ldg(&m.cx[a]) - __ldg(&m.cx[c]), n1y = __ldg(&m.cy[a]) - __ldg(&m.cy[c]), n1z = __ldg(&m.cz[a]) - __ldg(&m.cz[c]);
  const float n2x = __ldg(&m.cx[b]) - __ldg(&m.cx[d]), n2y = __ldg(&m.cy[b]) - __ldg(&m.cy[d]), n2z = __ldg(&m.cz[b]) - __ldg(&m.cz[d]);
  ox = n1y * n2z - n1z * n2y;
  oy = n1z * n2x - n1x * n2z;
  oz = n1x * n2y - n1y * n2x;
  dsc_normalize(ox, oy, oz);
}
__global__ void __launch_bounds__(256) k_grid_normals_flat(DevMesh m, DevGrids g, int all)
{
  const int gs = g.gs, gs1 = gs - 1, gs2 = g.gs2;
  const int mfg = g.max_face_grids;
  const long long units = all ? g.totgrid : (long long)__ldcg(&g.cnt->faces) * mfg;
  const long long items = units * gs2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < items; t += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(t / gs2), i = (int)(t - (long long)u * gs2);
    int grid;
    if (all) {
      grid = u;
    }
    else {
      const int f = __ldcg(&g.face_list[u / mfg]), c = u % mfg;
      if (c >= g.face_num[f]) continue;
      grid = g.face_start[f] + c;
      if (g.grid_owner && g.grid_owner[grid] != g.rank) continue; /* its owner sends the rim normals */
    }
    const int s0 = g.grid_slot0[grid];
    const int y = i / gs, x = i - y * gs;
    float ax = 0.0f, ay = 0.0f, az = 0.0f, fx, fy, fz;
    int counter = 0;
    if (x < gs1 && y < gs1) {
      dsc_grid_quad_normal(m, s0 + y * gs + x, gs, fx, fy, fz);
      ax += fx; ay += fy; az += fz;
      counter++;
    }
    if (x >= 1) {
      if (y < gs1) {
        dsc_grid_quad_normal(m, s0 + y * gs + (x - 1), gs, fx, fy, fz);
        ax += fx; ay += fy; az += fz;
        counter++;
      }
      if (y >= 1) {
        dsc_grid_quad_normal(m, s0 + (y - 1) * gs + (x - 1), gs, fx, fy, fz);
        ax += fx; ay += fy; az += fz;
        counter++;
      }
    }
    if (y >= 1 && x < gs1) {
      dsc_grid_quad_normal(m, s0 + (y - 1) * gs + x, gs, fx, fy, fz);
      ax += fx; ay += fy; az += fz;
      counter++;
    }
    const float sc = 1.0f / (float)counter;
    m.nx[s0 + i] = ax * sc; m.ny[s0 + i] = ay * sc; m.nz[s0 + i] = az * sc;
  }
}

/* update_node_vb leaf branch for grid leaves (pbvh.c:2033-2041 with PBVH_ITER_ALL over the node's
 * grids): box of every element of the listed leaves; the leaf's vert_bitmap words are cleared on the
 * way (the grid normal pass does not use them) */
__device__ __forceinline__ void dsc_grid_leaf_bb_body(const DevMesh &m, const int *list, const int *count, int cta, int ncta)
{
  __shared__ float red[6][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, bd = blockDim.x, nw = bd >> 5;
  const int tn = m.totnode;
  const int n = __ldcg(count);
  for (int h = cta; h < n; h += ncta) {
    const int l = __ldcg(&list[h]);
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    const int ub = m.leaf_ubeg[l], uc = m.leaf_ucnt[l];
    for (int i = 4 * tid; i < uc; i += 4 * bd) {
      const float4 X = ld4(m.cx, ub + i), Y = ld4(m.cy, ub + i), Z = ld4(m.cz, ub + i);
      const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (i + j < uc) {
          mn[0] = fminf(mn[0], xs[j]); mx[0] = fmaxf(mx[0], xs[j]);
          mn[1] = fminf(mn[1], ys[j]); mx[1] = fmaxf(mx[1], ys[j]);
          mn[2] = fminf(mn[2], zs[j]); mx[2] = fmaxf(mx[2], zs[j]);
        }
      }
    }
    for (int w = tid; w < (uc + 31) / 32; w += bd) m.dirty[(ub >> 5) + w] = 0u;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      for (int o = 16; o > 0; o >>= 1) {
        mn[k] = fminf(mn[k], __shfl_down_sync(0xffffffffu, mn[k], o));
        mx[k] = fmaxf(mx[k], __shfl_down_sync(0xffffffffu, mx[k], o));
      }
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        red[k][warp] = mn[k];
        red[3 + k][warp] = mx[k];
      }
    }
    __syncthreads();
    if (tid < 6) {
      float v = red[tid][0];
      for (int w = 1; w < nw; w++) v = (tid < 3) ? fminf(v, red[tid][w]) : fmaxf(v, red[tid][w]);
      m.bb[tid * tn + l] = v;
    }
  }
}

/* partitioned: the leaves every rank gathered this dab (the all-reduced bitmask) as a list, any order */
__global__ void __launch_bounds__(DSC_BLOCK) k_ghit_expand(DevMesh m, const unsigned *ghit, int *list, int *count)
{
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < m.ghit_words; w += gridDim.x * blockDim.x) {
    unsigned bits = ghit[w];
    if (!bits) continue;
    int at = atomicAdd(count, __popc(bits));
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      list[at++] = 32 * w + b;
    }
  }
}

/* gpu_pbvh_grid_buffers_update (gpu/intern/gpu_buffers.c:548-725) for the listed (flagged) grid leaves, in the record
 * format of k_draw_fill: smooth -- one record per element in (grid, y, x) order with its own normal and mask; flat --
 * four records per quad, corners (x, y), (x+1, y), (x+1, y+1), (x, y+1), the quad normal taken with the corners
 * reversed (gpu_buffers.c:664-666), the mean of the four masks, col = white.  A leaf's records start at
 * leaf_gbeg[leaf] * per_grid.  Byte-exact against the oracle on B200 (tests/test_gpu_grids.py). */
__global__ void __launch_bounds__(DSC_BLOCK) k_grid_draw_fill(DevMesh m, DevGrids g, const int *list, const int *count, int smooth_all,
                                                              const unsigned char *leaf_smooth, const long long *leaf_rec, int show_mask,
                                                              unsigned *vbo)
{
  const int n = *count;
  const int gs = g.gs, gs1 = gs - 1, gs2 = g.gs2;
  for (int h = blockIdx.x; h < n; h += gridDim.x) {
    const int l = list[h];
    /* per leaf: ME_SMOOTH of the leaf's first grid (gpu_buffers.c:574); the leaf's records then start at leaf_rec[l] */
    const int smooth = leaf_smooth ? (int)leaf_smooth[l] : smooth_all;
    const int per_grid = smooth ? gs2 : gs1 * gs1 * 4;
    const int gb = g.leaf_gbeg[l], ge = g.leaf_gbeg[l + 1];
    const int units = smooth ? gs2 : gs1 * gs1;
    const size_t base = leaf_rec ? (size_t)leaf_rec[l] : (size_t)gb * per_grid;
    for (int t = threadIdx.x; t < (ge - gb) * units; t += blockDim.x) {
      const int gi = t / units, u = t - gi * units;
      const int s0 = g.grid_slot0[g.leaf_grids[gb + gi]];
      unsigned *rec = vbo + (base + (size_t)gi * per_grid) * 9;
      if (smooth) {
        const int s = s0 + u;
        rec += (size_t)u * 9;
        unsigned cmask = 0u;
        if (show_mask) cmask = (unsigned)(unsigned char)(int)(g.mask[s] * 255);
        rec[0] = __float_as_uint(m.cx[s]);
        rec[1] = __float_as_uint(m.cy[s]);
        rec[2] = __float_as_uint(m.cz[s]);
        rec[3] = 0u;
        rec[4] = dsc_normal_short(m.nx[s]) | (dsc_normal_short(m.ny[s]) << 16);
        rec[5] = dsc_normal_short(m.nz[s]) | (cmask << 16);
        rec[6] = 0u;
        rec[7] = 0u;
        rec[8] = 0x00ffffffu;
        continue;
      }
      const int y = u / gs1, x = u - y * gs1;
      const int sl[4] = {s0 + y * gs + x, s0 + y * gs + x + 1, s0 + (y + 1) * gs + x + 1, s0 + (y + 1) * gs + x};
      /* normal_quad_v3(fno, co[3], co[2], co[1], co[0]): n1 = co3 - co1, n2 = co2 - co0 */
      const float n1x = m.cx[sl[3]] - m.cx[sl[1]], n1y = m.cy[sl[3]] - m.cy[sl[1]], n1z = m.cz[sl[3]] - m.cz[sl[1]];
      const float n2x = m.cx[sl[2]] - m.cx[sl[0]], n2y = m.cy[sl[2]] - m.cy[sl[0]], n2z = m.cz[sl[2]] - m.cz[sl[0]];
      float fx = n1y * n2z - n1z * n2y;
      float fy = n1z * n2x - n1x * n2z;
      float fz = n1x * n2y - n1y * n2x;
      dsc_normalize(fx, fy, fz);
      const unsigned n01 = dsc_normal_short(fx) | (dsc_normal_short(fy) << 16);
      unsigned cmask = 0u;
      if (show_mask) {
        const float fmask = (g.mask[sl[0]] + g.mask[sl[1]] + g.mask[sl[2]] + g.mask[sl[3]]) * 0.25f;
        cmask = (unsigned)(unsigned char)(int)(fmask * 255);
      }
      const unsigned w5 = dsc_normal_short(fz) | (cmask << 16);
      rec += (size_t)u * 4 * 9;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        rec[9 * j + 0] = __float_as_uint(m.cx[sl[j]]);
        rec[9 * j + 1] = __float_as_uint(m.cy[sl[j]]);
        rec[9 * j + 2] = __float_as_uint(m.cz[sl[j]]);
        rec[9 * j + 3] = 0u;
        rec[9 * j + 4] = n01;
        rec[9 * j + 5] = w5;
        rec[9 * j + 6] = 0xffffffffu;
        rec[9 * j + 7] = 0xffffffffu;
        rec[9 * j + 8] = 0x00ffffffu;
      }
    }
  }
}

/* ---- the stages as kernels of their own (session start, full averages) ---- */
/* the stamp of a dab is its sequence number (ring position + 1, read on the device so that the launch can be replayed
 * from a CUDA graph): faces / edges / vertices claimed by this dab carry it */
__device__ __forceinline__ int dsc_grid_seq(const DevMesh &m, int j) { return m.ring_ctl[0] + j + 1; }
__global__ void __launch_bounds__(DSC_BLOCK) k_grid_faces(DevMesh m, DevGrids g, const int *list, const int *count, int j)
{
  dsc_grid_faces_body(g, list, count, dsc_grid_seq(m, j), blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(128) k_grid_inner(DevMesh m, DevGrids g) { dsc_grid_inner_body(m, g, blockIdx.x, gridDim.x); }
__global__ void __launch_bounds__(128) k_grid_edges(DevMesh m, DevGrids g, int all, int j)
{
  dsc_grid_edges_body(m, g, all, dsc_grid_seq(m, j), blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(DSC_BLOCK) k_grid_cverts(DevMesh m, DevGrids g, int all) { dsc_grid_cverts_body(m, g, all, blockIdx.x, gridDim.x); }
/* partitioned: `n` dabs in a row were out of this rank's reach.  Each of them still averaged every coarse vertex (and every
 * coarse edge with more than two faces), whatever it touched (subdiv_ccg.c:1303-1324) -- a mean of equal floats is not
 * always that float, so the passes are replayed, one after the other, group by group (the groups are disjoint and a thread
 * keeps its groups across the repeats). */
__global__ void __launch_bounds__(DSC_BLOCK) k_grid_skipped(DevMesh m, DevGrids g, int n)
{
  for (int k = 0; k < n; k++) {
    if (g.has_odd_edges) dsc_grid_edges_body(m, g, 1, -1, blockIdx.x, gridDim.x);
    dsc_grid_cverts_body(m, g, 1, blockIdx.x, gridDim.x);
  }
}
/* coarse edges and coarse vertices in one launch: no element is in both a coarse-edge group and a corner group (the
 * edge pass leaves out the two end points of an edge, subdiv_ccg.c:1035), so the two passes -- and the pass over
 * untouched edges with more than two faces -- are independent.  The first `edge_ctas` CTAs take the edges. */
__global__ void __launch_bounds__(128) k_grid_edges_cverts(DevMesh m, DevGrids g, int odd_too, int all_cverts, int j, int edge_ctas,
                                                           const int *nhits)
{
  /* a dab that gathered no leaf does not stitch at all (the stroke step returns on totnode == 0 before multires_stitch_grids):
   * the passes over ALL odd edges / coarse vertices only run when something was hit -- the listed ones are empty then anyway */
  if ((odd_too || all_cverts) && __ldcg(nhits) == 0) return;
  if ((int)blockIdx.x < edge_ctas) {
    const int seq = dsc_grid_seq(m, j);
    dsc_grid_edges_body(m, g, 0, seq, blockIdx.x, edge_ctas);
    if (odd_too) dsc_grid_edges_body(m, g, 1, seq, blockIdx.x, edge_ctas);
  }
  else {
    dsc_grid_cverts_body(m, g, all_cverts, blockIdx.x - edge_ctas, gridDim.x - edge_ctas);
  }
}
__global__ void __launch_bounds__(GN_BLOCK) k_grid_normals(DevMesh m, DevGrids g, int all)
{
  extern __shared__ float gsm[];
  dsc_grid_normals_body(m, g, all, gsm, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(DSC_BLOCK) k_grid_leaf_bb(DevMesh m, const int *list, const int *count)
{
  dsc_grid_leaf_bb_body(m, list, count, blockIdx.x, gridDim.x);
}

/* ---- everything a dab does on grids after the brush, in ONE cooperative launch ----
 * faces of the gathered leaves -> stitch (inner boundaries, coarse edges, all coarse vertices) -> CCG normals
 * of those faces' grids -> their averaging (inner, edges, vertices) -> leaf boxes; the phases are separated by
 * grid barriers (dsc_grid_sync) instead of nine kernel boundaries.  One CTA of GN_BLOCK threads per SM. */
__global__ void __launch_bounds__(GN_BLOCK, 1) k_grid_dab(DevMesh m, DevGrids g, const int *list, const int *count, int j)
{
  extern __shared__ float gsm[];
  const int seq = dsc_grid_seq(m, j);
  const int cta = blockIdx.x, ncta = gridDim.x;
  unsigned target = 0u;
  dsc_grid_faces_body(g, list, count, seq, cta, ncta);
  dsc_grid_sync(m.grid_bar, target, ncta);
  dsc_grid_inner_body(m, g, cta, ncta);
  dsc_grid_sync(m.grid_bar, target, ncta);
  dsc_grid_edges_body(m, g, 0, seq, cta, ncta);
  if (__ldcg(count) > 0) { /* nothing gathered: no stitch (the stroke step returns on totnode == 0) */
    if (g.has_odd_edges) dsc_grid_edges_body(m, g, 1, seq, cta, ncta);
    dsc_grid_cverts_body(m, g, 1, cta, ncta); /* coarse vertices are not on any edge's list: same phase */
  }
  dsc_grid_sync(m.grid_bar, target, ncta);
  dsc_grid_normals_body(m, g, 0, gsm, cta, ncta);
  dsc_grid_sync(m.grid_bar, target, ncta);
  dsc_grid_inner_body(m, g, cta, ncta);
  dsc_grid_sync(m.grid_bar, target, ncta);
  dsc_grid_edges_body(m, g, 0, seq, cta, ncta);
  dsc_grid_cverts_body(m, g, 0, cta, ncta);
  dsc_grid_sync(m.grid_bar, target, ncta);
  dsc_grid_leaf_bb_body(m, list, count, cta, ncta);
}

/* ---- ray-cast on grids: pbvh_grids_node_raycast (pbvh.c:4102-4200) ----
 * One CTA per leaf the ray enters; every quad of the leaf's grids is tested as its two triangles (0, 1, 2) and (0, 2, 3)
 * (ray_face_intersection_quad, pbvh.c:3930-3949).  The reference walks the quads in (grid, y, x) order with a running
 * depth, and looks at the second triangle only when the first is not a nearer hit -- so which of a quad's two distances
 * counts depends on the depth the walk arrives with.  The kernel therefore reports every quad the ray touches with BOTH
 * distances and its place in the walk; the host folds them in the reference's order (a ray touches a handful of quads). */
struct GridRayHit {
  int leaf, grid, quad, order; /* quad = y * (gs - 1) + x; order = its place in the leaf's walk */
  float tmin, d1, d2;          /* leaf entry distance; triangle distances, negative = no intersection */
  int pad;
  float co[4][3];
};
__global__ void __launch_bounds__(DSC_BLOCK) k_grid_raycast(DevMesh m, DevGrids g, RayParams r, int capacity, int *count, GridRayHit *out)
{
  const int gs = g.gs, gs1 = gs - 1, per = gs1 * gs1;
  for (int l = blockIdx.x; l < m.nleaf; l += gridDim.x) {
    float tmin;
    if (!dsc_ray_leaf(m, r, l, tmin)) continue; /* CTA-uniform */
    const bool use_orig = r.original && (m.leaf_state[l] & DSC_LEAF_TOUCHED);
    const float *X = use_orig ? m.ox : m.cx, *Y = use_orig ? m.oy : m.cy, *Z = use_orig ? m.oz : m.cz;
    const int gb = g.leaf_gbeg[l], total = (g.leaf_gbeg[l + 1] - gb) * per;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
      const int gi = t / per, q = t - gi * per;
      const int y = q / gs1, x = q - y * gs1;
      const int grid = g.leaf_grids[gb + gi];
      const int s0 = g.grid_slot0[grid];
      const int sl[4] = {s0 + y * gs + x, s0 + y * gs + x + 1, s0 + (y + 1) * gs + x + 1, s0 + (y + 1) * gs + x};
      if (m.hidden) {
        /* paint_is_grid_face_hidden (paint.c:1234-1241): a quad with a hidden corner is not there */
        bool hid = false;
#pragma unroll
        for (int k = 0; k < 4; k++) hid |= ((m.hidden[sl[k] >> 5] >> (sl[k] & 31)) & 1u) != 0u;
        if (hid) continue;
      }
      float co[4][3];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        co[k][0] = X[sl[k]]; co[k][1] = Y[sl[k]]; co[k][2] = Z[sl[k]];
      }
      float d1 = -1.0f, d2 = -1.0f, lambda;
      if (dsc_ray_tri(r, co[0], co[1], co[2], lambda)) d1 = lambda + 0.0f; /* >= 0 or -0 past the sign test: +0 canonical */
      if (dsc_ray_tri(r, co[0], co[2], co[3], lambda)) d2 = lambda + 0.0f;
      if (d1 < 0.0f && d2 < 0.0f) continue;
      const int at = atomicAdd(count, 1);
      if (at >= capacity) continue; /* counted: the host grows the buffer and asks again */
      GridRayHit &h = out[at];
      h.leaf = l; h.grid = grid; h.quad = q; h.order = t;
      h.tmin = tmin; h.d1 = d1; h.d2 = d2; h.pad = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        h.co[k][0] = co[k][0]; h.co[k][1] = co[k][1]; h.co[k][2] = co[k][2];
      }
    }
  }
}
