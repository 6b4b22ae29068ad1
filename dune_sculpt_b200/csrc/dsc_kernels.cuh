/* Device-side of libdune_sculpt_cuda: resident mesh layout and the per-dab kernels (sm_100a).
 *
 * Layout ("slot order"): leaves are taken in traversal order (ascending PBVHNode.prim_indices
 * offset, the order BKE_pbvh_search_gather emits them, pbvh.c:2664-2705); each leaf's unique
 * vertices (the first uniq_verts entries of its vert_indices, pbvh_intern.h:35-55) occupy one
 * contiguous, 128-byte aligned run of slots.  Every per-vertex array is SoA over slots, so the
 * vertex loop of a leaf is a unit-stride stream instead of verts[vert_indices[i]] gathers.
 * Looptris are stored by position in PBVH.prim_indices, so a leaf's faces are contiguous too.
 *
 * Arithmetic: this file is compiled with -fmad=false and evaluates every expression in the same
 * order as the CPU path, so set membership (float compares) and positions are bit-identical.
 * Reductions that the CPU path runs in undefined order (area normal / centre) use 2^-32
 * fixed-point int64 sums: exact in any order.  Vertex normals are summed per vertex in ascending
 * looptri position, the order a single-threaded pbvh_update_normals_accum_task_cb produces
 * (pbvh.c:2933-2981).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DSC_CHUNK 1024        /* slots per work item, multiple of 32 */
#define DSC_BLOCK 256
#define DSC_LEAF_HIT 1u
#define DSC_LEAF_FIRST 2u
#define DSC_LEAF_TOUCHED 4u

enum {
  F_Leaf = 1 << 0, F_UpdateNormals = 1 << 1, F_UpdateBB = 1 << 2, F_UpdateOriginalBB = 1 << 3,
  F_UpdateDrawBuffers = 1 << 4, F_UpdateRedraw = 1 << 5, F_FullyHidden = 1 << 10, F_FullyMasked = 1 << 11,
};

struct DabState {
  int hit_count;
  int search_count;
  unsigned long long vd_total, hits_total, moved_total, dabs;
  long long acc[16]; /* nos[2][3], cos[2][3], count_no[2], count_co[2] */
  float area_no[3], area_co[3];
};

struct DevMesh {
  /* per slot */
  float *cx, *cy, *cz;    /* MVert.co */
  float *nx, *ny, *nz;    /* vert_normals */
  float *ox, *oy, *oz;    /* undo snapshot: orig_co */
  float *onx, *ony, *onz; /* undo snapshot: orig_no */
  float *tx, *ty, *tz;    /* Jacobi scratch (smooth) */
  const float *mask, *automask;
  unsigned *dirty;      /* PBVH.vert_bitmap, one bit per slot */
  unsigned *iter_moved; /* smooth: moved in this iteration */
  unsigned *capture;    /* debug: verts the current dab marked (NULL when capture is off) */
  const unsigned char *boundary;
  const unsigned *nb_off;
  const int *nb_idx;
  const unsigned *vt_off; /* slot -> incident looptri positions, ascending */
  const unsigned *vt_idx;
  /* per looptri position: slots of the verts of its poly; [3] = -1 triangle, <= -2 n-gon id */
  const int *pv0, *pv1, *pv2, *pv3;
  const int *poly_off, *poly_slots; /* n-gons only */
  const int *tri_leaf;              /* leaf (traversal index) holding the looptri position */
  /* leaves, traversal order */
  int nleaf;
  const int *leaf_node, *leaf_ubeg, *leaf_ucnt, *leaf_sbeg, *leaf_scnt, *leaf_pbeg, *leaf_pcnt;
  const int *shared_slots;
  unsigned *leaf_state;
  /* work items */
  int nchunk;
  const int *chunk_leaf, *chunk_beg, *chunk_cnt;
  /* nodes */
  int totnode;
  float *bb, *obb; /* [6][totnode] */
  int *node_flag;
  const int *node_child, *node_parent;
  int *node_mark;
  int nlevel;
  const int *level_off, *level_nodes; /* inner nodes by depth, root first */
  /* per-dab state */
  DabState *st;
  int *hit_list, *search_list;
  const float *curve; /* 257-entry LUT or NULL */
};

struct DabParams {
  int tool, curve_preset, flags, sculpt_plane;
  float loc[3], radius, view_n[3], bstrength, scale[3], hardness;
  float normal_radius_factor, plane_offset, plane_trim, tip_roundness, grab_delta[3], radius_scale;
};

/* ---------------------------------------------------------------- math, same order as the CPU */
__device__ __forceinline__ float dsc_normalize(float &x, float &y, float &z)
{
  /* lib/intern/math_vector_inline.c:1165-1181 */
  float d = x * x + y * y + z * z;
  if (d > 1.0e-35f) {
    d = sqrtf(d);
    const float f = 1.0f / d;
    x = x * f; y = y * f; z = z * f;
  }
  else {
    x = y = z = 0.0f;
    d = 0.0f;
  }
  return d;
}

__device__ __forceinline__ float dsc_clamp(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ long long dsc_fix32(float q) { return __float2ll_rn(q * 4294967296.0f); }

__device__ __forceinline__ float dsc_curve_lut(const float *t, float value)
{
  /* kernel/intern/colortools.c:942-965 */
  const float fi = value * 256.0f;
  const int i = (int)fi;
  if (fi < 0.0f || i < 0) return t[0];
  if (i >= 256) return t[256];
  const float w = fi - (float)i;
  return (1.0f - w) * t[i] + w * t[i + 1];
}

/* KERNEL_brush_curve_strength (kernel/intern/brush.h:87-91), presets types_brush_enums.h:176-187 */
__device__ __forceinline__ float dsc_curve_strength(const DevMesh &m, int preset, float p, float len)
{
  if (p >= len) return 0.0f;
  p = p / len;
  p = 1.0f - p;
  switch (preset) {
    case 0: return m.curve ? dsc_curve_lut(m.curve, 1.0f - p) : p;
    case 4: return p * p;
    case 1: return 3.0f * p * p - 2.0f * p * p * p;
    case 9: return (p * p * p) * (p * (p * 6.0f - 15.0f) + 10.0f);
    case 3: return sqrtf(p);
    case 5: return p;
    case 8: return 1.0f;
    case 2: return sqrtf(2.0f * p - p * p);
    case 6: return p * p * p * p;
    case 7: return p * (2.0f - p);
  }
  return 1.0f;
}

/* hardness remap, falloff, front-face, mask, automask */
__device__ __forceinline__ float dsc_strength_factor(const DevMesh &m, const DabParams &d, float len, float vnx,
                                                     float vny, float vnz, int s)
{
  float avg = 1.0f;
  float final_len = len;
  float q = len / d.radius;
  if (q < d.hardness) {
    final_len = 0.0f;
  }
  else if (d.hardness == 1.0f) {
    final_len = d.radius;
  }
  else {
    q = (q - d.hardness) / (1.0f - d.hardness);
    final_len = q * d.radius;
  }
  avg *= dsc_curve_strength(m, d.curve_preset, final_len, d.radius);
  if (d.flags & 1) {
    const float dot = vnx * d.view_n[0] + vny * d.view_n[1] + vnz * d.view_n[2];
    avg *= (dot > 0.0f) ? dot : 0.0f;
  }
  const float mk = m.mask ? m.mask[s] : 0.0f;
  avg *= 1.0f - mk;
  if (m.automask) avg *= m.automask[s];
  return avg;
}

/* BKE_mesh_calc_poly_normal of the poly of looptri position `pos`
 * (kernel/intern/mesh_evaluate.c:39-86, lib/intern/math_geom.cc:31-69) */
__device__ __forceinline__ void dsc_poly_normal(const DevMesh &m, unsigned pos, float &fx, float &fy, float &fz)
{
  const int a = m.pv0[pos], b = m.pv1[pos], c = m.pv2[pos], e = m.pv3[pos];
  if (e >= 0) {
    const float n1x = m.cx[a] - m.cx[c], n1y = m.cy[a] - m.cy[c], n1z = m.cz[a] - m.cz[c];
    const float n2x = m.cx[b] - m.cx[e], n2y = m.cy[b] - m.cy[e], n2z = m.cz[b] - m.cz[e];
    fx = n1y * n2z - n1z * n2y;
    fy = n1z * n2x - n1x * n2z;
    fz = n1x * n2y - n1y * n2x;
    dsc_normalize(fx, fy, fz);
  }
  else if (e == -1) {
    const float bx = m.cx[b], by = m.cy[b], bz = m.cz[b];
    const float n1x = m.cx[a] - bx, n1y = m.cy[a] - by, n1z = m.cz[a] - bz;
    const float n2x = bx - m.cx[c], n2y = by - m.cy[c], n2z = bz - m.cz[c];
    fx = n1y * n2z - n1z * n2y;
    fy = n1z * n2x - n1x * n2z;
    fz = n1x * n2y - n1y * n2x;
    dsc_normalize(fx, fy, fz);
  }
  else {
    /* Newell, mesh_evaluate.c:39-60 */
    const int p = -(e + 2);
    const int beg = m.poly_off[p], end = m.poly_off[p + 1];
    int sp = m.poly_slots[end - 1];
    float px = m.cx[sp], py = m.cy[sp], pz = m.cz[sp];
    fx = fy = fz = 0.0f;
    for (int i = beg; i < end; i++) {
      const int sc = m.poly_slots[i];
      const float qx = m.cx[sc], qy = m.cy[sc], qz = m.cz[sc];
      fx += (py - qy) * (pz + qz);
      fy += (pz - qz) * (px + qx);
      fz += (px - qx) * (py + qy);
      px = qx; py = qy; pz = qz;
    }
    if (dsc_normalize(fx, fy, fz) == 0.0f) fz = 1.0f;
  }
}

/* ------------------------------------------------------------------------------ K1 gather */
/* One CTA.  Flat leaf test in traversal order + order-preserving ballot compaction.  A leaf passes
 * BKE_pbvh_search_gather's DFS iff it passes the callback itself, because every inner AABB is the
 * union of its children (pbvh.c:2040-2043) and the sphere test is monotone in the box.
 * mark != 0: also does the per-node part of the dab: undo-node membership (first touch) and
 * BKE_pbvh_node_mark_update (pbvh.c:3641-3645). */
__global__ void __launch_bounds__(1024) k_gather(DevMesh m, float cx, float cy, float cz, float radius_sq,
                                                 int original, int ignore_ineffective, int mark)
{
  __shared__ int warp_cnt[32];
  __shared__ int s_base;
  __shared__ unsigned long long s_vd;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (mark && tid < 16) m.st->acc[tid] = 0;
  if (tid == 0) {
    s_base = 0;
    s_vd = 0ull;
  }
  __syncthreads();
  const float *bbs = original ? m.obb : m.bb;
  const int tn = m.totnode;
  int *out = mark ? m.hit_list : m.search_list;
  unsigned long long vd = 0;
  for (int l0 = 0; l0 < m.nleaf; l0 += 1024) {
    const int l = l0 + tid;
    bool hit = false;
    int node = -1;
    if (l < m.nleaf) {
      node = m.leaf_node[l];
      const int flag = m.node_flag[node];
      const bool skip = ignore_ineffective && (flag & (F_FullyHidden | F_FullyMasked));
      if (!skip) {
        const float c[3] = {cx, cy, cz};
        float t[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const float bmin = bbs[i * tn + node], bmax = bbs[(3 + i) * tn + node];
          float nearest;
          if (bmin > c[i]) nearest = bmin;
          else if (bmax < c[i]) nearest = bmax;
          else nearest = c[i];
          t[i] = c[i] - nearest;
        }
        hit = (t[0] * t[0] + t[1] * t[1] + t[2] * t[2]) < radius_sq;
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 32; w++) {
      const int c = warp_cnt[w];
      if (w < warp) before += c;
      total += c;
    }
    if (hit) {
      const int pos = s_base + before + __popc(bal & ((1u << lane) - 1u));
      out[pos] = l;
    }
    if (mark && l < m.nleaf) {
      const unsigned st = m.leaf_state[l];
      if (hit) {
        m.leaf_state[l] = DSC_LEAF_HIT | DSC_LEAF_TOUCHED | ((st & DSC_LEAF_TOUCHED) ? 0u : DSC_LEAF_FIRST);
        m.node_flag[node] |= F_UpdateNormals | F_UpdateBB | F_UpdateOriginalBB | F_UpdateDrawBuffers | F_UpdateRedraw;
        vd += (unsigned long long)m.leaf_ucnt[l];
      }
      else {
        m.leaf_state[l] = st & DSC_LEAF_TOUCHED;
      }
    }
    __syncthreads();
    if (tid == 0) s_base += total;
    __syncthreads();
  }
  if (mark) {
    for (int o = 16; o > 0; o >>= 1) vd += __shfl_down_sync(0xffffffffu, vd, o);
    if (lane == 0 && vd) atomicAdd(&s_vd, vd);
    __syncthreads();
    if (tid == 0) {
      /* what dsc_last_area reports when the tool samples no plane: zero normal, brush location */
      m.st->area_no[0] = m.st->area_no[1] = m.st->area_no[2] = 0.0f;
      m.st->area_co[0] = cx; m.st->area_co[1] = cy; m.st->area_co[2] = cz;
      m.st->hit_count = s_base;
      m.st->vd_total += s_vd;
      m.st->hits_total += (unsigned long long)s_base;
      m.st->dabs += 1ull;
    }
  }
  else if (tid == 0) {
    m.st->search_count = s_base;
  }
}

/* ------------------------------------------------------------------- K3a area normal / centre */
/* SURVEY.md 8a row a15.  Unique verts of hit leaves inside radius * normal_radius_factor; two
 * buckets by the sign of dot(view_normal, no); smoothstep weight; exact int64 sums. */
__global__ void __launch_bounds__(DSC_BLOCK) k_area(DevMesh m, DabParams d, int use_cos)
{
  __shared__ unsigned long long sacc[16];
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < 16) sacc[tid] = 0ull;
  __syncthreads();
  float test_radius = sqrtf(d.radius * d.radius);
  test_radius *= d.normal_radius_factor;
  const float radius_sq = test_radius * test_radius;
  long long n0x = 0, n0y = 0, n0z = 0, n1x = 0, n1y = 0, n1z = 0;
  long long c0x = 0, c0y = 0, c0z = 0, c1x = 0, c1y = 0, c1z = 0;
  long long cnt0 = 0, cnt1 = 0;
  for (int c = blockIdx.x; c < m.nchunk; c += gridDim.x) {
    if (!(m.leaf_state[m.chunk_leaf[c]] & DSC_LEAF_HIT)) continue;
    const int beg = m.chunk_beg[c], cnt = m.chunk_cnt[c];
    for (int i = tid; i < cnt; i += DSC_BLOCK) {
      const int s = beg + i;
      const float dx = m.cx[s] - d.loc[0], dy = m.cy[s] - d.loc[1], dz = m.cz[s] - d.loc[2];
      const float distsq = dx * dx + dy * dy + dz * dz;
      if (distsq > radius_sq) continue;
      const float vx = m.nx[s], vy = m.ny[s], vz = m.nz[s];
      const bool flip = (d.view_n[0] * vx + d.view_n[1] * vy + d.view_n[2] * vz) <= 0.0f;
      const float q = 1.0f - (sqrtf(distsq) / test_radius);
      const float f = dsc_clamp(3.0f * q * q - 2.0f * q * q * q, 0.0f, 1.0f);
      if (use_cos) {
        const float w = 1.0f - f;
        const long long ax = dsc_fix32((dx * w) / test_radius);
        const long long ay = dsc_fix32((dy * w) / test_radius);
        const long long az = dsc_fix32((dz * w) / test_radius);
        if (flip) { c1x += ax; c1y += ay; c1z += az; }
        else { c0x += ax; c0y += ay; c0z += az; }
      }
      const long long bx = dsc_fix32(vx * f), by = dsc_fix32(vy * f), bz = dsc_fix32(vz * f);
      if (flip) { n1x += bx; n1y += by; n1z += bz; cnt1++; }
      else { n0x += bx; n0y += by; n0z += bz; cnt0++; }
    }
  }
  long long v[16] = {n0x, n0y, n0z, n1x, n1y, n1z, c0x, c0y, c0z, c1x, c1y, c1z, cnt0, cnt1,
                     use_cos ? cnt0 : 0, use_cos ? cnt1 : 0};
#pragma unroll
  for (int k = 0; k < 16; k++) {
    long long x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0 && x != 0) atomicAdd(&sacc[k], (unsigned long long)x);
  }
  __syncthreads();
  if (tid < 16 && sacc[tid] != 0ull) atomicAdd((unsigned long long *)&m.st->acc[tid], sacc[tid]);
}

/* finalisation of the sums, same float/double steps as the CPU path */
__device__ __forceinline__ void dsc_area_finalize(const DabState *st, const DabParams &d, bool use_cos, float no[3],
                                                  float co[3])
{
  no[0] = no[1] = no[2] = 0.0f;
  for (int i = 0; i < 2; i++) {
    float tx = (float)((double)st->acc[i * 3 + 0] * (1.0 / 4294967296.0));
    float ty = (float)((double)st->acc[i * 3 + 1] * (1.0 / 4294967296.0));
    float tz = (float)((double)st->acc[i * 3 + 2] * (1.0 / 4294967296.0));
    if (dsc_normalize(tx, ty, tz) != 0.0f) {
      no[0] = tx; no[1] = ty; no[2] = tz;
      break;
    }
  }
  co[0] = d.loc[0]; co[1] = d.loc[1]; co[2] = d.loc[2];
  if (use_cos) {
    float test_radius = sqrtf(d.radius * d.radius);
    test_radius *= d.normal_radius_factor;
    for (int i = 0; i < 2; i++) {
      const long long cnt = st->acc[14 + i];
      if (cnt == 0) continue;
      for (int k = 0; k < 3; k++) {
        const double mean = (double)st->acc[6 + i * 3 + k] / ((double)cnt * 4294967296.0);
        co[k] = (float)((double)d.loc[k] + (double)test_radius * mean);
      }
      break;
    }
  }
}

__device__ __forceinline__ void dsc_sculpt_normal(const DabState *st, const DabParams &d, float no[3])
{
  float co[3];
  switch (d.sculpt_plane) {
    case 1: no[0] = d.view_n[0]; no[1] = d.view_n[1]; no[2] = d.view_n[2]; break;
    case 2: no[0] = 1.0f; no[1] = 0.0f; no[2] = 0.0f; break;
    case 3: no[0] = 0.0f; no[1] = 1.0f; no[2] = 0.0f; break;
    case 4: no[0] = 0.0f; no[1] = 0.0f; no[2] = 1.0f; break;
    default: dsc_area_finalize(st, d, false, no, co); break;
  }
}

/* ------------------------------------------------------------- K2 + K3 brush (fused snapshot) */
struct BrushDerived {
  float offset[3];
  /* clay strips */
  float origin[3], ax[3][3], sc[3], plane_no[3], plane_d, trim_sq, bstrength;
  int flip, skip;
};

/* Draw / inflate / grab / clay strips over the unique verts of hit leaves (SURVEY.md 8a rows
 * a11-a19).  First touch of a leaf in the stroke snapshots co/no into orig_co/orig_no before the
 * vertex is moved (row a9).  Displaced verts get their vert_bitmap bit (pbvh.c:3729). */
__global__ void __launch_bounds__(DSC_BLOCK) k_brush(DevMesh m, DabParams d)
{
  __shared__ BrushDerived D;
  __shared__ unsigned s_moved;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) {
    s_moved = 0;
    D.skip = 0;
    float an[3] = {0, 0, 0}, ac[3] = {d.loc[0], d.loc[1], d.loc[2]};
    if (d.tool == 1) {
      dsc_sculpt_normal(m.st, d, an);
      for (int k = 0; k < 3; k++) {
        float o = an[k] * d.radius;
        o = o * d.scale[k];
        o = o * d.bstrength;
        D.offset[k] = o;
      }
    }
    else if (d.tool == 18) {
      D.flip = (d.bstrength < 0.0f);
      const float radius = D.flip ? -d.radius : d.radius;
      const float displace = radius * (0.18f + d.plane_offset);
      D.bstrength = D.flip ? -d.bstrength : d.bstrength;
      float area_no[3];
      if (d.sculpt_plane == 0) {
        dsc_area_finalize(m.st, d, true, an, ac);
        area_no[0] = an[0]; area_no[1] = an[1]; area_no[2] = an[2];
      }
      else {
        dsc_sculpt_normal(m.st, d, an);
        dsc_area_finalize(m.st, d, true, area_no, ac);
      }
      const float area_co0[3] = {ac[0], ac[1], ac[2]};
      if ((d.flags & 4) || (d.grab_delta[0] == 0.0f && d.grab_delta[1] == 0.0f && d.grab_delta[2] == 0.0f)) {
        D.skip = 1;
      }
      float area_co[3];
      for (int k = 0; k < 3; k++) {
        const float t = (an[k] * d.scale[k]) * displace;
        area_co[k] = area_co0[k] + t;
        D.origin[k] = area_co[k] + area_no[k] * (-radius * 0.7f);
        D.plane_no[k] = an[k];
      }
      float a0[3], a1[3];
      a0[0] = area_no[1] * d.grab_delta[2] - area_no[2] * d.grab_delta[1];
      a0[1] = area_no[2] * d.grab_delta[0] - area_no[0] * d.grab_delta[2];
      a0[2] = area_no[0] * d.grab_delta[1] - area_no[1] * d.grab_delta[0];
      a1[0] = area_no[1] * a0[2] - area_no[2] * a0[1];
      a1[1] = area_no[2] * a0[0] - area_no[0] * a0[2];
      a1[2] = area_no[0] * a0[1] - area_no[1] * a0[0];
      float a2[3] = {area_no[0], area_no[1], area_no[2]};
      dsc_normalize(a0[0], a0[1], a0[2]);
      dsc_normalize(a1[0], a1[1], a1[2]);
      dsc_normalize(a2[0], a2[1], a2[2]);
      for (int k = 0; k < 3; k++) {
        D.ax[0][k] = a0[k]; D.ax[1][k] = a1[k]; D.ax[2][k] = a2[k];
      }
      D.sc[0] = d.radius; D.sc[1] = d.radius; D.sc[2] = d.radius * 1.25f;
      D.plane_d = -(an[0] * area_co[0] + an[1] * area_co[1] + an[2] * area_co[2]);
      D.trim_sq = (d.radius * d.radius) * (d.plane_trim * d.plane_trim);
      /* what the host reads back as the plane: centre before the offset */
      ac[0] = area_co0[0]; ac[1] = area_co0[1]; ac[2] = area_co0[2];
    }
    if (blockIdx.x == 0) {
      for (int k = 0; k < 3; k++) {
        m.st->area_no[k] = an[k];
        m.st->area_co[k] = ac[k];
      }
    }
  }
  __syncthreads();
  const int tool = d.tool;
  const float radius_sq = d.radius * d.radius;
  const bool need_no = (tool == 4) || (d.flags & 1);
  unsigned moved_cnt = 0;
  for (int c = blockIdx.x; c < m.nchunk; c += gridDim.x) {
    const unsigned lst = m.leaf_state[m.chunk_leaf[c]];
    if (!(lst & DSC_LEAF_HIT)) continue;
    const bool first = (lst & DSC_LEAF_FIRST) != 0;
    const int beg = m.chunk_beg[c], cnt = m.chunk_cnt[c];
    const int cnt32 = (cnt + 31) & ~31;
    for (int i = tid; i < cnt32; i += DSC_BLOCK) {
      const int s = beg + i;
      bool moved = false;
      if (i < cnt) {
        const float x = m.cx[s], y = m.cy[s], z = m.cz[s];
        float vnx = 0.0f, vny = 0.0f, vnz = 0.0f;
        if (first) {
          vnx = m.nx[s]; vny = m.ny[s]; vnz = m.nz[s];
          m.ox[s] = x; m.oy[s] = y; m.oz[s] = z;
          m.onx[s] = vnx; m.ony[s] = vny; m.onz[s] = vnz;
        }
        if (!D.skip) {
          if (tool == 18) {
            /* clay strips: brush-local cube test + plane */
            const float rx = x - D.origin[0], ry = y - D.origin[1], rz = z - D.origin[2];
            float local[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
              local[k] = fabsf((rx * D.ax[k][0] + ry * D.ax[k][1] + rz * D.ax[k][2]) / D.sc[k]);
            }
            const float side = 1.0f;
            if (local[0] <= side && local[1] <= side && local[2] <= side) {
              const float roundness = d.tip_roundness;
              const float hardness = 1.0f - roundness;
              const float constant_side = hardness * side;
              const float falloff_side = roundness * side;
              float dist;
              const float mn = local[0] < local[1] ? local[0] : local[1];
              const float mx = local[0] > local[1] ? local[0] : local[1];
              if (mn > constant_side) {
                const float ex = local[0] - constant_side, ey = local[1] - constant_side;
                dist = sqrtf(ex * ex + ey * ey) / falloff_side;
              }
              else if (mx > constant_side) {
                dist = (mx - constant_side) / falloff_side;
              }
              else {
                dist = 0.0f;
              }
              float side_d = (x * D.plane_no[0] + y * D.plane_no[1] + z * D.plane_no[2]) + D.plane_d;
              if (D.flip) side_d = -side_d;
              if (side_d <= 0.0f) {
                const float pd = (D.plane_no[0] * x + D.plane_no[1] * y + D.plane_no[2] * z) + D.plane_d;
                const float ix = x + D.plane_no[0] * (-pd), iy = y + D.plane_no[1] * (-pd), iz = z + D.plane_no[2] * (-pd);
                const float vx = ix - x, vy = iy - y, vz = iz - z;
                if (!(d.flags & 2) || ((vx * vx + vy * vy + vz * vz) <= D.trim_sq)) {
                  if (!first && (d.flags & 1)) { vnx = m.nx[s]; vny = m.ny[s]; vnz = m.nz[s]; }
                  const float fade = D.bstrength * dsc_strength_factor(m, d, d.radius * dist, vnx, vny, vnz, s);
                  const float px = vx * fade, py = vy * fade, pz = vz * fade;
                  m.cx[s] = x + px; m.cy[s] = y + py; m.cz[s] = z + pz;
                  moved = true;
                }
              }
            }
          }
          else {
            /* sphere test: grab tests the stroke-start coordinates */
            float tx = x, ty = y, tz = z;
            if (tool == 5 && !first) { tx = m.ox[s]; ty = m.oy[s]; tz = m.oz[s]; }
            const float dx = tx - d.loc[0], dy = ty - d.loc[1], dz = tz - d.loc[2];
            const float distsq = dx * dx + dy * dy + dz * dz;
            if (!(distsq > radius_sq)) {
              if (!first && need_no) {
                if (tool == 5) { vnx = m.onx[s]; vny = m.ony[s]; vnz = m.onz[s]; }
                else { vnx = m.nx[s]; vny = m.ny[s]; vnz = m.nz[s]; }
              }
              float fade = dsc_strength_factor(m, d, sqrtf(distsq), vnx, vny, vnz, s);
              if (tool == 1) {
                const float px = D.offset[0] * fade, py = D.offset[1] * fade, pz = D.offset[2] * fade;
                m.cx[s] = x + px; m.cy[s] = y + py; m.cz[s] = z + pz;
              }
              else if (tool == 4) {
                fade = d.bstrength * fade;
                const float sc = fade * d.radius;
                const float px = (vnx * sc) * d.scale[0], py = (vny * sc) * d.scale[1], pz = (vnz * sc) * d.scale[2];
                m.cx[s] = x + px; m.cy[s] = y + py; m.cz[s] = z + pz;
              }
              else {
                fade = d.bstrength * fade;
                const float px = d.grab_delta[0] * fade, py = d.grab_delta[1] * fade, pz = d.grab_delta[2] * fade;
                m.cx[s] = tx + px; m.cy[s] = ty + py; m.cz[s] = tz + pz;
              }
              moved = true;
            }
          }
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, moved);
      if (bal && lane == 0) {
        m.dirty[s >> 5] |= bal;
        if (m.capture) m.capture[s >> 5] |= bal;
        moved_cnt += __popc(bal);
      }
    }
  }
  if (moved_cnt) atomicAdd(&s_moved, moved_cnt);
  __syncthreads();
  if (tid == 0 && s_moved) atomicAdd(&m.st->moved_total, (unsigned long long)s_moved);
}

/* snapshot only (smooth brush: first touch, before iteration 0) */
__global__ void __launch_bounds__(DSC_BLOCK) k_snapshot(DevMesh m)
{
  for (int c = blockIdx.x; c < m.nchunk; c += gridDim.x) {
    const unsigned lst = m.leaf_state[m.chunk_leaf[c]];
    if ((lst & (DSC_LEAF_HIT | DSC_LEAF_FIRST)) != (DSC_LEAF_HIT | DSC_LEAF_FIRST)) continue;
    const int beg = m.chunk_beg[c], cnt = m.chunk_cnt[c];
    for (int i = threadIdx.x; i < cnt; i += DSC_BLOCK) {
      const int s = beg + i;
      m.ox[s] = m.cx[s]; m.oy[s] = m.cy[s]; m.oz[s] = m.cz[s];
      m.onx[s] = m.nx[s]; m.ony[s] = m.ny[s]; m.onz[s] = m.nz[s];
    }
  }
}

/* ------------------------------------------------------------------------------- K4 smooth */
/* One Jacobi iteration, part A: new position of every unique vert of a hit leaf inside the
 * sphere = co + (neighbour average - co) * fade, into the scratch arrays (SURVEY.md 8a row a20:
 * interior verts average all edge neighbours, boundary verts only boundary neighbours, boundary
 * verts with <= 2 neighbours stay). */
__global__ void __launch_bounds__(DSC_BLOCK) k_smooth_a(DevMesh m, DabParams d, float strength)
{
  __shared__ unsigned s_moved;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) s_moved = 0;
  __syncthreads();
  const float radius_sq = d.radius * d.radius;
  unsigned moved_cnt = 0;
  for (int c = blockIdx.x; c < m.nchunk; c += gridDim.x) {
    if (!(m.leaf_state[m.chunk_leaf[c]] & DSC_LEAF_HIT)) continue;
    const int beg = m.chunk_beg[c], cnt = m.chunk_cnt[c];
    const int cnt32 = (cnt + 31) & ~31;
    for (int i = tid; i < cnt32; i += DSC_BLOCK) {
      const int s = beg + i;
      bool moved = false;
      if (i < cnt) {
        const float x = m.cx[s], y = m.cy[s], z = m.cz[s];
        const float dx = x - d.loc[0], dy = y - d.loc[1], dz = z - d.loc[2];
        const float distsq = dx * dx + dy * dy + dz * dz;
        if (!(distsq > radius_sq)) {
          float vnx = 0.0f, vny = 0.0f, vnz = 0.0f;
          if (d.flags & 1) { vnx = m.nx[s]; vny = m.ny[s]; vnz = m.nz[s]; }
          const float fade = strength * dsc_strength_factor(m, d, sqrtf(distsq), vnx, vny, vnz, s);
          float ax = 0.0f, ay = 0.0f, az = 0.0f;
          int total = 0;
          const unsigned qb = m.nb_off[s], qe = m.nb_off[s + 1];
          const int neighbor_count = (int)(qe - qb);
          const bool is_boundary = m.boundary[s] != 0;
          for (unsigned q = qb; q < qe; q++) {
            const int u = m.nb_idx[q];
            if (!is_boundary || m.boundary[u]) {
              ax += m.cx[u]; ay += m.cy[u]; az += m.cz[u];
              total++;
            }
          }
          float rx, ry, rz;
          if ((neighbor_count <= 2 && is_boundary) || total == 0) {
            rx = x; ry = y; rz = z;
          }
          else {
            const float f = 1.0f / (float)total;
            rx = ax * f; ry = ay * f; rz = az * f;
          }
          const float vx = rx - x, vy = ry - y, vz = rz - z;
          m.tx[s] = x + vx * fade;
          m.ty[s] = y + vy * fade;
          m.tz[s] = z + vz * fade;
          moved = true;
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, moved);
      if (lane == 0) {
        m.iter_moved[s >> 5] = bal;
        if (bal) {
          m.dirty[s >> 5] |= bal;
          if (m.capture) m.capture[s >> 5] |= bal;
          moved_cnt += __popc(bal);
        }
      }
    }
  }
  if (moved_cnt) atomicAdd(&s_moved, moved_cnt);
  __syncthreads();
  if (tid == 0 && s_moved) atomicAdd(&m.st->moved_total, (unsigned long long)s_moved);
}

/* part B: commit the scratch positions */
__global__ void __launch_bounds__(DSC_BLOCK) k_smooth_b(DevMesh m)
{
  for (int c = blockIdx.x; c < m.nchunk; c += gridDim.x) {
    if (!(m.leaf_state[m.chunk_leaf[c]] & DSC_LEAF_HIT)) continue;
    const int beg = m.chunk_beg[c], cnt = m.chunk_cnt[c];
    for (int i = threadIdx.x; i < cnt; i += DSC_BLOCK) {
      const int s = beg + i;
      if ((m.iter_moved[s >> 5] >> (s & 31)) & 1u) {
        m.cx[s] = m.tx[s]; m.cy[s] = m.ty[s]; m.cz[s] = m.tz[s];
      }
    }
  }
}

/* ------------------------------------------------------------------------------ K5 normals */
/* BKE_pbvh_update_normals for PBVH_FACES (pbvh.c:2912-3036): for every dirty unique vert of a
 * leaf flagged UpdateNormals, normal = normalize(sum of the poly normals of its looptris), summed
 * in ascending looptri position; then the dirty bit is cleared.  Like the reference's accumulate
 * pass, looptris of leaves that are not flagged contribute nothing (pbvh.c:2943). */
__global__ void __launch_bounds__(DSC_BLOCK) k_normals(DevMesh m)
{
  const int tid = threadIdx.x, lane = tid & 31;
  for (int c = blockIdx.x; c < m.nchunk; c += gridDim.x) {
    const int leaf = m.chunk_leaf[c];
    const int node = m.leaf_node[leaf];
    if (!(m.node_flag[node] & F_UpdateNormals)) continue;
    const unsigned pb = (unsigned)m.leaf_pbeg[leaf], pe = pb + (unsigned)m.leaf_pcnt[leaf];
    const int beg = m.chunk_beg[c], cnt = m.chunk_cnt[c];
    const int cnt32 = (cnt + 31) & ~31;
    for (int i = tid; i < cnt32; i += DSC_BLOCK) {
      const int s = beg + i;
      const unsigned word = m.dirty[s >> 5];
      if (word == 0u) continue; /* warp-uniform */
      if ((word >> (s & 31)) & 1u) {
        float sx = 0.0f, sy = 0.0f, sz = 0.0f;
        const unsigned qb = m.vt_off[s], qe = m.vt_off[s + 1];
        for (unsigned q = qb; q < qe; q++) {
          const unsigned pos = m.vt_idx[q];
          if (pos < pb || pos >= pe) {
            if (!(m.node_flag[m.leaf_node[m.tri_leaf[pos]]] & F_UpdateNormals)) continue;
          }
          float fx, fy, fz;
          dsc_poly_normal(m, pos, fx, fy, fz);
          sx += fx; sy += fy; sz += fz;
        }
        dsc_normalize(sx, sy, sz);
        m.nx[s] = sx; m.ny[s] = sy; m.nz[s] = sz;
      }
      __syncwarp();
      if (lane == 0) m.dirty[s >> 5] = 0u;
    }
  }
}

/* ------------------------------------------------------------------------------ K6 leaf BB */
/* update_node_vb leaf branch (pbvh.c:2033-2041): min/max over ALL verts of the leaf, unique
 * (stream) and shared (gather), for leaves flagged UpdateBB. */
__global__ void __launch_bounds__(DSC_BLOCK) k_leaf_bb(DevMesh m)
{
  __shared__ float red[6][DSC_BLOCK / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tn = m.totnode;
  for (int l = blockIdx.x; l < m.nleaf; l += gridDim.x) {
    const int node = m.leaf_node[l];
    if (!(m.node_flag[node] & F_UpdateBB)) continue;
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    const int ub = m.leaf_ubeg[l], uc = m.leaf_ucnt[l];
    for (int i = tid; i < uc; i += DSC_BLOCK) {
      const int s = ub + i;
      const float x = m.cx[s], y = m.cy[s], z = m.cz[s];
      mn[0] = fminf(mn[0], x); mx[0] = fmaxf(mx[0], x);
      mn[1] = fminf(mn[1], y); mx[1] = fmaxf(mx[1], y);
      mn[2] = fminf(mn[2], z); mx[2] = fmaxf(mx[2], z);
    }
    const int sb = m.leaf_sbeg[l], sc = m.leaf_scnt[l];
    for (int i = tid; i < sc; i += DSC_BLOCK) {
      const int s = m.shared_slots[sb + i];
      const float x = m.cx[s], y = m.cy[s], z = m.cz[s];
      mn[0] = fminf(mn[0], x); mx[0] = fmaxf(mx[0], x);
      mn[1] = fminf(mn[1], y); mx[1] = fmaxf(mx[1], y);
      mn[2] = fminf(mn[2], z); mx[2] = fmaxf(mx[2], z);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      for (int o = 16; o > 0; o >>= 1) {
        mn[k] = fminf(mn[k], __shfl_down_sync(0xffffffffu, mn[k], o));
        mx[k] = fmaxf(mx[k], __shfl_down_sync(0xffffffffu, mx[k], o));
      }
    }
    __syncthreads(); /* red[] reuse across leaves */
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
      for (int w = 1; w < DSC_BLOCK / 32; w++) v = (tid < 3) ? fminf(v, red[tid][w]) : fmaxf(v, red[tid][w]);
      m.bb[tid * tn + node] = v;
    }
  }
}

/* ------------------------------------------------------------------------------ K7 BB flush */
/* pbvh_flush_bb (pbvh.c:3287-3317), bottom-up by depth in one CTA.  Every inner node is
 * recomputed: an inner node none of whose leaves changed already equals the union of its
 * children, so the result is the reference's.  Then the leaf flags in clear_mask are dropped. */
__global__ void __launch_bounds__(1024) k_flush(DevMesh m, int clear_mask)
{
  const int tn = m.totnode;
  for (int lev = m.nlevel - 1; lev >= 0; lev--) {
    const int b = m.level_off[lev], e = m.level_off[lev + 1];
    for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
      const int node = m.level_nodes[i];
      const int c0 = m.node_child[node], c1 = c0 + 1;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        m.bb[k * tn + node] = fminf(__ldcg(&m.bb[k * tn + c0]), __ldcg(&m.bb[k * tn + c1]));
        m.bb[(3 + k) * tn + node] = fmaxf(__ldcg(&m.bb[(3 + k) * tn + c0]), __ldcg(&m.bb[(3 + k) * tn + c1]));
      }
    }
    __threadfence_block();
    __syncthreads();
  }
  if (clear_mask) {
    for (int l = threadIdx.x; l < m.nleaf; l += blockDim.x) {
      const int node = m.leaf_node[l];
      const int f = m.node_flag[node];
      if (f & clear_mask) m.node_flag[node] = f & ~clear_mask;
    }
  }
}

/* PBVH_UpdateOriginalBB flush (pbvh.c:3139-3141, 3298-3314): flagged leaves copy vb to orig_vb and
 * tag their ancestors; the second kernel copies for tagged inner nodes. */
__global__ void __launch_bounds__(DSC_BLOCK) k_orig_leaves(DevMesh m)
{
  const int tn = m.totnode;
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < m.nleaf; l += gridDim.x * blockDim.x) {
    const int node = m.leaf_node[l];
    const int f = m.node_flag[node];
    if (!(f & F_UpdateOriginalBB)) continue;
    m.node_flag[node] = f & ~F_UpdateOriginalBB;
    for (int k = 0; k < 6; k++) m.obb[k * tn + node] = m.bb[k * tn + node];
    int p = m.node_parent[node];
    while (p >= 0 && atomicExch(&m.node_mark[p], 1) == 0) p = m.node_parent[p];
  }
}
__global__ void __launch_bounds__(DSC_BLOCK) k_orig_inner(DevMesh m)
{
  const int tn = m.totnode;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < tn; n += gridDim.x * blockDim.x) {
    if (m.node_mark[n]) {
      m.node_mark[n] = 0;
      for (int k = 0; k < 6; k++) m.obb[k * tn + n] = m.bb[k * tn + n];
    }
  }
}

/* ------------------------------------------------------------------------------ misc */
__global__ void k_mark_all(DevMesh m, int flags, int all_dirty, int nwords)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m.nleaf) m.node_flag[m.leaf_node[i]] |= flags;
  if (all_dirty) {
    for (int w = i; w < nwords; w += gridDim.x * blockDim.x) m.dirty[w] = 0xffffffffu;
  }
}

/* slot order -> original vertex order, AoS float3 */
__global__ void k_export3(float *__restrict__ out, const float *__restrict__ ax, const float *__restrict__ ay,
                          const float *__restrict__ az, const int *__restrict__ slot_of, int totvert)
{
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < totvert; v += gridDim.x * blockDim.x) {
    const int s = slot_of[v];
    out[3 * v + 0] = ax[s];
    out[3 * v + 1] = ay[s];
    out[3 * v + 2] = az[s];
  }
}
__global__ void k_import3(const float *__restrict__ in, float *__restrict__ ax, float *__restrict__ ay,
                          float *__restrict__ az, const int *__restrict__ slot_of, unsigned *dirty, int totvert)
{
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < totvert; v += gridDim.x * blockDim.x) {
    const int s = slot_of[v];
    const float x = in[3 * v + 0], y = in[3 * v + 1], z = in[3 * v + 2];
    /* BKE_pbvh_vert_coords_apply marks only verts whose coordinates changed (pbvh.c:4731-4736) */
    if (ax[s] != x || ay[s] != y || az[s] != z) {
      ax[s] = x; ay[s] = y; az[s] = z;
      atomicOr(&dirty[s >> 5], 1u << (s & 31));
    }
  }
}
__global__ void k_export_bits(int *__restrict__ out_count, int *__restrict__ out, const unsigned *__restrict__ bits,
                              const int *__restrict__ slot_of, int totvert)
{
  /* not order preserving; the host sorts */
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < totvert; v += gridDim.x * blockDim.x) {
    const int s = slot_of[v];
    if ((bits[s >> 5] >> (s & 31)) & 1u) out[atomicAdd(out_count, 1)] = v;
  }
}
