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

#define DSC_CHUNK 1024        /* slots per work item = DSC_BLOCK threads x 4 slots */
#define DSC_TILE 1024         /* most unique verts of a tile */
#define DSC_ENT_FIRST 1       /* tile-list entry bits: first touch of the leaf in this stroke */
#define DSC_ENT_NORMALS 2     /* the leaf updates its normals with this list */
#define DSC_ENT_BOUNDS 4      /* ... its box */
#define DSC_ENT_IBND_SHIFT 8  /* bits 8..18: where the tile's boundary run starts (unique verts other tiles read come last) */
#define DSC_BLOCK 256
#define DSC_LEAF_HIT 1u
#define DSC_LEAF_FIRST 2u
#define DSC_LEAF_TOUCHED 4u

enum {
  F_Leaf = 1 << 0, F_UpdateNormals = 1 << 1, F_UpdateBB = 1 << 2, F_UpdateOriginalBB = 1 << 3,
  F_UpdateDrawBuffers = 1 << 4, F_UpdateRedraw = 1 << 5, F_FullyHidden = 1 << 10, F_FullyMasked = 1 << 11,
};

#define DSC_SLOTS 4 /* per-dab state ring: dab i uses slot i & 3 (the side stream may lag two dabs) */

struct DabState {
  int hit_count;   /* leaves gathered by the dab (hit list of this slot) */
  int area_count;  /* hit leaves that also reach the normal-sampling sphere */
  int tile_count;  /* tiles of the gathered leaves (tile list of this slot) */
  int atile_count; /* tiles of the leaves that reach the normal-sampling sphere (area tile list) */
  long long acc[16]; /* nos[2][3], cos[2][3], count_no[2], count_co[2] */
  float area_no[3], area_co[3];
};

struct StrokeTotals {
  unsigned long long vd_total, hits_total, moved_total, dabs;
  unsigned long long area_vd_total;     /* unique verts of the leaves the area pass walks (its U') */
  unsigned long long area_inside_total; /* verts inside the sampling sphere (its M') */
  unsigned long long all_total;         /* unique + shared verts of the gathered leaves (A of SURVEY.md 8d) */
  unsigned long long prim_total;        /* their looptris / grids (T) */
  unsigned long long first_total;       /* unique + shared verts of the leaves first touched in the stroke (undo snapshot) */
  unsigned long long refit_total;       /* inner nodes the bottom-up refit rewrote */
  int search_count; /* leaves found by the last stand-alone search (search_list) */
  int flag_count;   /* leaves collected by k_collect_flagged (flag_list) */
  int flag_tiles;   /* their tiles (flag_tile_list) */
};

/* Nodes are renumbered on the device: ids [0, nleaf) are the leaves in traversal order, ids
 * [nleaf, totnode) the inner nodes, root first (breadth-first).  The host translates. */
struct DabEntry;
struct DevMesh {
  const DabEntry *ring; /* [DSC_RING] queued dabs */
  int *ring_ctl;        /* [0] sequence number of the running batch's first dab, [1] of the next batch */
  /* per slot */
  float *cx, *cy, *cz;    /* MVert.co */
  float *nx, *ny, *nz;    /* vert_normals */
  float *ox, *oy, *oz;    /* undo snapshot: orig_co */
  float *onx, *ony, *onz; /* undo snapshot: orig_no */
  float *tx, *ty, *tz;    /* Jacobi scratch (smooth) */
  const float *mask, *automask;
  const unsigned *hidden; /* bit per slot: MVert.flag & ME_HIDE / grid_hidden -- the vertex iterator skips it (NULL: none) */
  unsigned *dirty;      /* PBVH.vert_bitmap, one bit per slot */
  unsigned *iter_moved; /* smooth: moved in this iteration */
  unsigned *capture;    /* debug: verts the current dab marked (NULL when capture is off) */
  const unsigned char *boundary;
  const unsigned *nb_off;
  const int *nb_idx;
  /* general normals path (any poly size / leaf size); NULL when every leaf takes the smem path */
  const unsigned *vt_off; /* slot -> incident looptri positions, ascending */
  const unsigned *vt_idx;
  /* per looptri position: slots of the verts of its poly; [3] = -1 triangle, <= -2 n-gon id */
  const int *pv0, *pv1, *pv2, *pv3;
  const int *poly_off, *poly_slots; /* n-gons only */
  const int *tri_leaf;              /* leaf holding the looptri position */
  /* Tiles: every leaf's run of unique-vertex slots is cut into spatially compact runs of at most
   * DSC_TILE slots (32-aligned); a tile is the work unit of the per-vertex kernels and of the
   * shared-memory normals kernel.  Per tile: a local vertex list (unique | staged) and local poly
   * entries (own leaf | other leaves); per unique vert the entries of its looptris in ascending
   * looptri position. */
  int ntile;
  const int *leaf_tile0;      /* [nleaf + 1] */
  const int2 *tile_range;     /* {first slot, unique verts | start of the boundary run << 16} */
  const int4 *tile_meta;      /* 3 per tile, see TileMeta */
  const int *stage_slots;     /* per tile: slots of the staged verts (counted in the leaf box first) */
  const ushort4 *e_pv;        /* local vertex indices of the entry's poly, w = 0xffff: triangle */
  const int *e_halo_leaf;     /* per other-leaf entry: the leaf that holds its looptri */
  /* vertex -> entry lists, sliced ELL: per group of 32 slots `width` row pairs of 32 words, word j
   * of a lane = entries of its looptris 2j (low half) and 2j + 1 (high half), ascending position;
   * padding points at the tile's zero entry */
  const unsigned *v2_goff;    /* [slots / 32 + 1], in words */
  const unsigned *v2_idx;
  const unsigned char *leaf_fast;
  /* byte offsets of the tile kernel's shared-memory regions (sized for the largest tile of the mesh, so a region
   * never moves between tiles): positions at 0, then poly normals, poly entries, index words, other-leaf switches */
  int sm_off_f, sm_off_e, sm_off_v2, sm_off_h;
  /* leaves */
  int nleaf;
  int max_chunks; /* ceil(max uniq_verts / DSC_CHUNK) */
  const int *leaf_ubeg, *leaf_ucnt, *leaf_sbeg, *leaf_scnt, *leaf_pbeg, *leaf_pcnt;
  const int *leaf_sslots; /* per leaf: slots of its shared verts (general path only) */
  unsigned *leaf_state;
  /* nodes, device numbering */
  int totnode;
  float *bb, *obb; /* [6][totnode] */
  int *node_flag;
  const int4 *topo; /* x parent (-1 root), y side bit of this node in its parent (1 / 2), z sibling */
  const int *child0, *child1;
  int *pending, *arrived; /* [2][totnode] bottom-up refit: which children will arrive / have arrived (two sets: the
                             batch kernel tags the next dab while the previous one is still refitted) */
  unsigned *grid_bar;     /* arrival counter of the batch kernel's grid barrier */
  int *node_mark;
  int nlevel;
  const int *level_off, *level_nodes; /* inner nodes by depth, root first */
  /* multi-GPU: this rank owns leaves [own_lo, own_hi); ghit = per-dab bitmask of gathered leaves,
   * one ring slot per dab, all-reduced across ranks (NULL on one GPU) */
  int own_lo, own_hi;
  unsigned *ghit; /* [DSC_SLOTS][ghit_words] leaves gathered by the dab = leaves whose normals update */
  int ghit_words;
  unsigned *flag_mask; /* the same for the leaf list k_collect_flagged builds */
  /* per-dab state */
  DabState *st;       /* [DSC_SLOTS] */
  StrokeTotals *tot;
  int *hit_list, *area_list; /* [DSC_SLOTS][nleaf] */
  int4 *tile_list, *atile_list; /* [DSC_SLOTS][ntile] {tile, first slot, unique verts, DSC_ENT_* bits} */
  int *search_list, *flag_list;
  int4 *flag_tile_list;
  const float *curve; /* 257-entry LUT or NULL */
};

struct DabParams {
  int tool, curve_preset, flags, sculpt_plane;
  float loc[3], radius, view_n[3], bstrength, scale[3], hardness;
  float normal_radius_factor, plane_offset, plane_trim, tip_roundness, grab_delta[3], radius_scale;
  int falloff_shape;      /* 1: tube (distance to the view line through the location) */
  int clip_flags;         /* bits 0-2 locked axes, bits 3-5 mirror clipping */
  float clip_tol[3], normal_weight;
};

/* One queued dab on the device: the descriptor plus what the host derives from it.  The per-dab
 * kernels read their dab from this ring (DevMesh.ring) instead of from kernel arguments, so the launch
 * sequence of a dab is the same for every dab and can be replayed as a CUDA graph: entry
 * (ring_ctl[0] + j) of the ring for the j-th dab of the running batch.  The slot of the per-dab state
 * ring (DSC_SLOTS) is a kernel argument: batches start at multiples of DSC_SLOTS. */
#define DSC_RING 1024
struct DabEntry {
  DabParams d;
  float radius_sq;      /* gather sphere: (radius * radius_scale)^2 */
  float area_radius_sq; /* normal-sampling sphere */
  float smooth_last;    /* strength of the last smoothing iteration */
  int original;         /* gather tests the stroke-start boxes (grab) */
  int set_flags;        /* node flags the gather sets on hit leaves */
  int ent_bits;         /* DSC_ENT_NORMALS | DSC_ENT_BOUNDS */
  int use_cos;          /* the area pass also samples the centre (clay strips) */
  int peers;            /* partitioned PBVH: bit per rank the dab can reach (the ranks that exchange it); 0 on one GPU */
  int gather_resets;    /* the gather empties the boxes of the leaves it hits (the tile kernel accumulates into them); the inner
                           nodes are refitted once, when somebody reads them (stroke end) */
  int pad;
};

__device__ __forceinline__ const DabEntry &dsc_dab_entry(const DevMesh &m, int j)
{
  return m.ring[(m.ring_ctl[0] + j) & (DSC_RING - 1)];
}

/* Programmatic dependent launch: the per-dab kernels of the main stream are launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization, so a kernel may start while its predecessor
 * still runs.  Everything a kernel reads or writes that the predecessor touches comes after
 * dsc_pdl_wait(); each kernel releases its own dependent right after its wait, so by the time a
 * kernel starts, the launch before its predecessor has completed -- its outputs (e.g. the gather's
 * lists for the brush and the tile kernel) may be read before the wait.  Both are no-ops for a
 * kernel launched the ordinary way. */
__device__ __forceinline__ void dsc_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void dsc_pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

/* first node of a batch of `count` dabs: the batch's base sequence number */
__global__ void k_batch_begin(DevMesh m, int count)
{
  const int base = m.ring_ctl[1];
  m.ring_ctl[0] = base;
  m.ring_ctl[1] = base + count;
  *m.grid_bar = 0u;
}

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
    /* (q - 0) / (1 - 0) is q exactly: skip the IEEE divide when there is no hardness */
    if (d.hardness != 0.0f) q = (q - d.hardness) / (1.0f - d.hardness);
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

/* row a11 brush test: squared distance to the brush location, or (tube) to the view line through it */
__device__ __forceinline__ float dsc_test_distsq(const DabParams &d, float x, float y, float z)
{
  if (d.falloff_shape == 1) {
    const float plane_d = -(d.view_n[0] * d.loc[0] + d.view_n[1] * d.loc[1] + d.view_n[2] * d.loc[2]);
    const float side = (d.view_n[0] * x + d.view_n[1] * y + d.view_n[2] * z) + plane_d;
    const float qx = (x + d.view_n[0] * (-side)) - d.loc[0];
    const float qy = (y + d.view_n[1] * (-side)) - d.loc[1];
    const float qz = (z + d.view_n[2] * (-side)) - d.loc[2];
    return qx * qx + qy * qy + qz * qz;
  }
  const float dx = x - d.loc[0], dy = y - d.loc[1], dz = z - d.loc[2];
  return dx * dx + dy * dy + dz * dz;
}
/* row a11 clipping: locked axes keep their coordinate, a vertex within the tolerance of a clipping mirror plane stays on it */
__device__ __forceinline__ void dsc_clip(const DabParams &d, float &x, float &y, float &z, float vx, float vy, float vz)
{
  if (!d.clip_flags) {
    x = vx; y = vy; z = vz;
    return;
  }
  if (!(d.clip_flags & 8)) x = ((d.clip_flags & 1) && fabsf(x) <= d.clip_tol[0]) ? 0.0f : vx;
  if (!(d.clip_flags & 16)) y = ((d.clip_flags & 2) && fabsf(y) <= d.clip_tol[1]) ? 0.0f : vy;
  if (!(d.clip_flags & 32)) z = ((d.clip_flags & 4) && fabsf(z) <= d.clip_tol[2]) ? 0.0f : vz;
}
/* the tube falloff's node test: squared distance from the line loc + t n to the box (0 when it crosses it, else the least
 * distance to one of the 12 edges); same arithmetic as the host's and the oracle's */
__device__ __forceinline__ float dsc_line_aabb_distsq(const float loc[3], const float n[3], const float bmin[3], const float bmax[3])
{
  float tmin = -3.402823466e+38f, tmax = 3.402823466e+38f;
  bool inside = true;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (n[k] != 0.0f) {
      const float t1 = (bmin[k] - loc[k]) / n[k], t2 = (bmax[k] - loc[k]) / n[k];
      const float lo = t1 < t2 ? t1 : t2, hi = t1 < t2 ? t2 : t1;
      if (lo > tmin) tmin = lo;
      if (hi < tmax) tmax = hi;
    }
    else if (loc[k] < bmin[k] || loc[k] > bmax[k]) {
      inside = false;
    }
  }
  if (inside && tmin <= tmax) return 0.0f;
  const float a = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
  float best = 3.402823466e+38f;
#pragma unroll
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
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float df = (w[k] + e[k] * sp) - n[k] * tp;
        dist += df * df;
      }
      if (dist < best) best = dist;
    }
  }
  return best;
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

/* Before the tile kernel accumulates tile boxes into them (dsc_red_min / dsc_red_max), the boxes of
 * the listed leaves that take the tile path start from the empty box. */
__device__ __forceinline__ void dsc_reset_leaf_box(const DevMesh &m, int leaf)
{
  if (!m.leaf_fast[leaf]) return;
  const int tn = m.totnode;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    m.bb[k * tn + leaf] = 3.402823466e+38f;
    m.bb[(3 + k) * tn + leaf] = -3.402823466e+38f;
  }
}
/* ------------------------------------------------------------------------------ K1 gather */
/* One thread per leaf, any number of CTAs.  A leaf passes BKE_pbvh_search_gather's DFS iff it
 * passes the callback itself, because every inner AABB is the union of its children
 * (pbvh.c:2040-2043) and the sphere test is monotone in the box -- so the tree walk collapses to a
 * flat test.  Hits are appended with one warp-aggregated atomic; leaf ids ARE traversal ranks, so
 * the host gets BKE_pbvh_search_gather's order back by sorting the ids it reads.
 * mark != 0 is the dab: undo-node membership (first touch), BKE_pbvh_node_mark_update
 * (pbvh.c:3641-3645), the sub-list of leaves that reach the (smaller) normal-sampling sphere, and
 * the reset of the next slot of the per-dab state ring. */
__device__ __forceinline__ void dsc_gather_body(const DevMesh &m, int slot, float cx, float cy, float cz, float radius_sq,
                                                float area_radius_sq, int original, int ignore_ineffective, int mark,
                                                int set_flags, int ent_bits, int bidx, int tag_parity, const float *tube_n = nullptr)
{
  const int tid = threadIdx.x, lane = tid & 31;
  const int l = bidx * DSC_BLOCK + tid;
  DabState *st = m.st + slot;
  if (mark && bidx == 0) {
    DabState *nx = m.st + ((slot + 1) & (DSC_SLOTS - 1));
    if (tid < 16) nx->acc[tid] = 0;
    if (tid == 16) nx->hit_count = 0;
    if (tid == 17) nx->area_count = 0;
    if (tid == 19) nx->tile_count = 0;
    if (tid == 20) nx->atile_count = 0;
    if (tid == 18) {
      /* what dsc_last_area reports when the tool samples no plane: zero normal, brush location */
      st->area_no[0] = st->area_no[1] = st->area_no[2] = 0.0f;
      st->area_co[0] = cx; st->area_co[1] = cy; st->area_co[2] = cz;
      m.tot->dabs += 1ull;
    }
  }
  bool hit = false, ahit = false;
  int flag = 0;
  unsigned lst = 0;
  if (l < m.nleaf && (!mark || (l >= m.own_lo && l < m.own_hi))) {
    const float *bbs = original ? m.obb : m.bb;
    const int tn = m.totnode;
    flag = __ldcg(&m.node_flag[l]);
    if (mark) lst = __ldcg(&m.leaf_state[l]);
    const float c[3] = {cx, cy, cz};
    float t[3], lo[3], hi[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float bmin = __ldcg(&bbs[i * tn + l]), bmax = __ldcg(&bbs[(3 + i) * tn + l]);
      lo[i] = bmin;
      hi[i] = bmax;
      float nearest;
      if (bmin > c[i]) nearest = bmin;
      else if (bmax < c[i]) nearest = bmax;
      else nearest = c[i];
      t[i] = c[i] - nearest;
    }
    float dist = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
    if (tube_n) dist = dsc_line_aabb_distsq(c, tube_n, lo, hi); /* tube falloff: the view line through the location */
    const bool skip = ignore_ineffective && (flag & (F_FullyHidden | F_FullyMasked));
    hit = !skip && (dist < radius_sq);
    ahit = hit && mark && (dist <= area_radius_sq);
  }
  const unsigned bal = __ballot_sync(0xffffffffu, hit);
  if (bal) {
    int base = 0;
    if (lane == 0) base = atomicAdd(mark ? &st->hit_count : &m.tot->search_count, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (hit) (mark ? m.hit_list + (size_t)slot * m.nleaf : m.search_list)[base + __popc(bal & ((1u << lane) - 1u))] = l;
  }
  if (!mark) return;
  /* the bitmask of gathered leaves: every warp stores its word, so the ring slot needs no reset */
  if (lane == 0 && (l >> 5) < m.ghit_words) m.ghit[(size_t)slot * m.ghit_words + (l >> 5)] = bal;
  if (!hit && l < m.nleaf && (lst & (DSC_LEAF_HIT | DSC_LEAF_FIRST))) m.leaf_state[l] = lst & DSC_LEAF_TOUCHED;
  if (!bal) return;
  const unsigned abal = __ballot_sync(0xffffffffu, ahit);
  if (abal) {
    int base = 0;
    if (lane == 0) base = atomicAdd(&st->area_count, __popc(abal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (ahit) m.area_list[(size_t)slot * m.nleaf + base + __popc(abal & ((1u << lane) - 1u))] = l;
  }
  /* tile lists: the work units of the per-vertex kernels */
  int t0 = 0, nt = 0;
  if (hit) {
    t0 = m.leaf_tile0[l];
    nt = m.leaf_tile0[l + 1] - t0;
  }
  const int bits = ent_bits | ((lst & DSC_LEAF_TOUCHED) ? 0 : DSC_ENT_FIRST);
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    if (pass == 1 && !abal) break;
    const int mine = (pass == 0 || ahit) ? nt : 0;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31) base = atomicAdd(pass == 0 ? &st->tile_count : &st->atile_count, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    int4 *tl = (pass == 0 ? m.tile_list : m.atile_list) + (size_t)slot * m.ntile + base + (incl - mine);
    for (int k = 0; k < mine; k++) {
      const int2 r = m.tile_range[t0 + k];
      tl[k] = make_int4(t0 + k, r.x, r.y & 0xffff, bits | ((r.y >> 16) << DSC_ENT_IBND_SHIFT));
    }
  }
  unsigned long long vd = 0, avd = 0, all = 0, prims = 0, first = 0;
  if (ahit) avd = (unsigned long long)m.leaf_ucnt[l];
  if (hit) {
    /* this thread is the only reader of the leaf's box in this launch; the next one is the tile kernel, which accumulates */
    if (mark == 2) dsc_reset_leaf_box(m, l);
    m.leaf_state[l] = DSC_LEAF_HIT | DSC_LEAF_TOUCHED | ((lst & DSC_LEAF_TOUCHED) ? 0u : DSC_LEAF_FIRST);
    m.node_flag[l] = flag | set_flags;
    vd = (unsigned long long)m.leaf_ucnt[l];
    all = vd + (unsigned long long)m.leaf_scnt[l];
    prims = (unsigned long long)m.leaf_pcnt[l];
    first = (lst & DSC_LEAF_TOUCHED) ? 0ull : all;
    if (tag_parity >= 0 && (ent_bits & DSC_ENT_BOUNDS)) {
      /* batch kernel: what k_tag_ancestors does on the side stream otherwise (the boxes are emptied later) */
      int *pending = m.pending + (size_t)tag_parity * m.totnode;
      int4 t = m.topo[l];
      while (t.x >= 0) {
        const int4 tp = m.topo[t.x];
        const int old = atomicOr(&pending[t.x], t.y);
        if (old & t.y) break;
        t = tp;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    vd += __shfl_down_sync(0xffffffffu, vd, o);
    all += __shfl_down_sync(0xffffffffu, all, o);
    prims += __shfl_down_sync(0xffffffffu, prims, o);
    first += __shfl_down_sync(0xffffffffu, first, o);
  }
  if (abal) {
    for (int o = 16; o > 0; o >>= 1) avd += __shfl_down_sync(0xffffffffu, avd, o);
    if (lane == 0) atomicAdd(&m.tot->area_vd_total, avd);
  }
  if (lane == 0) {
    atomicAdd(&m.tot->vd_total, vd);
    atomicAdd(&m.tot->hits_total, (unsigned long long)__popc(bal));
    atomicAdd(&m.tot->all_total, all);
    atomicAdd(&m.tot->prim_total, prims);
    if (first) atomicAdd(&m.tot->first_total, first);
  }
}

/* stand-alone search (BKE_pbvh_search_gather outside a stroke) */
__global__ void __launch_bounds__(DSC_BLOCK) k_gather(DevMesh m, float cx, float cy, float cz, float radius_sq, int original,
                                                      int ignore_ineffective)
{
  dsc_gather_body(m, 0, cx, cy, cz, radius_sq, 0.0f, original, ignore_ineffective, 0, 0, 0, blockIdx.x, -1);
}
/* the gather of the j-th dab of the running batch */
__global__ void __launch_bounds__(DSC_BLOCK) k_gather_dab(DevMesh m, int j, int slot)
{
  dsc_pdl_wait(); /* leaf boxes of the previous dab's tile kernel; ring_ctl of the batch head */
  dsc_pdl_launch();
  const DabEntry &e = dsc_dab_entry(m, j);
  const float vn[3] = {e.d.view_n[0], e.d.view_n[1], e.d.view_n[2]};
  dsc_gather_body(m, slot, e.d.loc[0], e.d.loc[1], e.d.loc[2], e.radius_sq, e.area_radius_sq, e.original, 1, e.gather_resets ? 2 : 1,
                  e.set_flags, e.ent_bits, blockIdx.x, -1, e.d.falloff_shape == 1 ? vn : nullptr);
}

/* leaves carrying any of `flags` (update_search_cb, pbvh.c:2891-2900) */
__global__ void __launch_bounds__(1024) k_collect_flagged(DevMesh m, int flags)
{
  __shared__ int s_cnt, s_tiles;
  if (threadIdx.x == 0) s_cnt = s_tiles = 0;
  for (int w = threadIdx.x; w < m.ghit_words; w += blockDim.x) m.flag_mask[w] = 0u;
  __syncthreads();
  for (int l = threadIdx.x; l < m.nleaf; l += blockDim.x) {
    const int f = m.node_flag[l] & flags;
    if (!f) continue;
    m.flag_list[atomicAdd(&s_cnt, 1)] = l;
    if (f & F_UpdateNormals) atomicOr(&m.flag_mask[l >> 5], 1u << (l & 31));
    const int bits = ((f & F_UpdateNormals) ? DSC_ENT_NORMALS : 0) | ((f & F_UpdateBB) ? DSC_ENT_BOUNDS : 0);
    const int t0 = m.leaf_tile0[l], nt = m.leaf_tile0[l + 1] - t0;
    const int base = atomicAdd(&s_tiles, nt);
    for (int k = 0; k < nt; k++) {
      const int2 r = m.tile_range[t0 + k];
      m.flag_tile_list[base + k] = make_int4(t0 + k, r.x, r.y & 0xffff, bits | ((r.y >> 16) << DSC_ENT_IBND_SHIFT));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    m.tot->flag_count = s_cnt;
    m.tot->flag_tiles = s_tiles;
  }
}

__global__ void __launch_bounds__(DSC_BLOCK) k_reset_leaf_boxes(DevMesh m, const int *list, const int *count, int need_flag)
{
  const int n = *count;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int l = list[i];
    if (!need_flag || (m.node_flag[l] & need_flag)) dsc_reset_leaf_box(m, l);
  }
}

/* Bottom-up refit, pass 1: tags the ancestors of the listed leaves.  pending[p] gets bit 1 / bit 2
 * when a leaf below its first / second child is listed.  A walker stops at the first node that
 * already carries its bit (somebody else tagged everything above).  With reset_boxes the leaf boxes
 * are emptied on the way: this kernel is ordered after the previous dab's refit (same stream), which
 * is the last reader of the old boxes, and before this dab's tile kernel (event). */
__global__ void __launch_bounds__(DSC_BLOCK) k_tag_ancestors(DevMesh m, int slot, int reset_boxes)
{
  const int *list = m.hit_list + (size_t)slot * m.nleaf;
  const int n = m.st[slot].hit_count;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int leaf = list[i];
    if (reset_boxes) dsc_reset_leaf_box(m, leaf);
    int4 t = m.topo[leaf];
    while (t.x >= 0) {
      const int4 tp = m.topo[t.x]; /* in flight together with the atomic */
      const int old = atomicOr(&m.pending[t.x], t.y);
      if (old & t.y) break;
      t = tp;
    }
  }
}

/* Bottom-up refit, pass 2 = pbvh_flush_bb (pbvh.c:3287-3317) without a level-by-level sweep: one
 * walker per refreshed leaf carries its box upwards; at every inner node the walker that completes
 * the last pending child merges the sibling's box, stores the node and carries on, the other one
 * retires.  Only nodes above a refreshed leaf are touched, as in the reference.  Runs on the side
 * stream: nothing on the device reads inner boxes (the gather is flat), only the host does. */
__device__ __forceinline__ void dsc_refit_body(const DevMesh &m, int slot, int parity, int gtid, int gthreads)
{
  const int *list = m.hit_list + (size_t)slot * m.nleaf;
  const int n = __ldcg(&m.st[slot].hit_count);
  const int tn = m.totnode;
  int *pending = m.pending + (size_t)parity * tn, *arrived = m.arrived + (size_t)parity * tn;
  for (int i = gtid; i < n; i += gthreads) {
    const int leaf = __ldcg(&list[i]);
    float box[6];
#pragma unroll
    for (int k = 0; k < 6; k++) box[k] = __ldcg(&m.bb[k * tn + leaf]);
    int4 t = m.topo[leaf];
    while (t.x >= 0) {
      const int p = t.x;
      const int4 tp = m.topo[p];
      const int pend = __ldcg(&pending[p]);
      __threadfence(); /* my box (stored below, or by the leaf kernel) is visible before I arrive */
      const int old = atomicOr(&arrived[p], t.y);
      if ((old | t.y) != pend) break; /* the sibling subtree is still on its way */
      /* ordered after the atomic: if the sibling subtree was refreshed it arrived, fenced, before
       * me; if not, its stored box is current */
#pragma unroll
      for (int k = 0; k < 3; k++) {
        box[k] = fminf(box[k], __ldcg(&m.bb[k * tn + t.z]));
        box[3 + k] = fmaxf(box[3 + k], __ldcg(&m.bb[(3 + k) * tn + t.z]));
      }
#pragma unroll
      for (int k = 0; k < 6; k++) __stcg(&m.bb[k * tn + p], box[k]);
      arrived[p] = 0;
      pending[p] = 0;
      atomicAdd(&m.tot->refit_total, 1ull);
      t = tp;
    }
  }
}
__global__ void __launch_bounds__(DSC_BLOCK) k_refit(DevMesh m, int slot)
{
  dsc_refit_body(m, slot, 0, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

/* general path: work unit u of a leaf list: (leaf, chunk) -> slot range; false if the chunk is empty */
__device__ __forceinline__ bool dsc_unit(const DevMesh &m, const int *list, int u, int &leaf, int &beg, int &cnt)
{
  const int h = u / m.max_chunks, c = u - h * m.max_chunks;
  leaf = list[h];
  const int ucnt = m.leaf_ucnt[leaf];
  const int off = c * DSC_CHUNK;
  if (off >= ucnt) return false;
  cnt = min(DSC_CHUNK, ucnt - off);
  beg = m.leaf_ubeg[leaf] + off;
  return true;
}

/* through L2 (ld.global.cg): the persistent batch kernel re-reads arrays other CTAs rewrote since this SM last saw them */
__device__ __forceinline__ float4 ld4(const float *p, int s) { return __ldcg(reinterpret_cast<const float4 *>(p + s)); }
__device__ __forceinline__ void st4(float *p, int s, const float4 &v) { *reinterpret_cast<float4 *>(p + s) = v; }

/* 4-bit per-lane flags -> one 32-bit word per 8 lanes (lane & 7 == 0 holds it) */
__device__ __forceinline__ unsigned dsc_pack_nibbles(unsigned nib, int lane)
{
  unsigned w = nib << (4 * (lane & 7));
  w |= __shfl_xor_sync(0xffffffffu, w, 1);
  w |= __shfl_xor_sync(0xffffffffu, w, 2);
  w |= __shfl_xor_sync(0xffffffffu, w, 4);
  return w;
}

/* the hidden bits of the four slots from s0 (s0 is a multiple of 4) */
__device__ __forceinline__ unsigned dsc_hidden4(const DevMesh &m, int s0)
{
  return m.hidden ? (__ldg(&m.hidden[s0 >> 5]) >> (s0 & 31)) & 0xfu : 0u;
}

/* ------------------------------------------------------------------- K3a area normal / centre */
/* SURVEY.md 8a row a15.  Unique verts of hit leaves inside radius * normal_radius_factor; two
 * buckets by the sign of dot(view_normal, no); smoothstep weight; exact int64 sums.  Streams
 * float4 runs of the SoA position arrays; normals are only fetched for runs with a vert inside. */
__device__ __forceinline__ void dsc_area_body(const DevMesh &m, const DabParams &d, int use_cos, int slot, int cta, int ncta)
{
  const bool use_orig = d.tool == 5; /* grab (normal weight): the stroke-start surface */
  DabState *st = m.st + slot;
  const int4 *alist = m.atile_list + (size_t)slot * m.ntile;
  __shared__ unsigned long long sacc[16];
  const int tid = threadIdx.x, lane = tid & 31;
  __syncthreads(); /* sacc of an earlier call */
  if (tid < 16) sacc[tid] = 0ull;
  __syncthreads();
  float test_radius = sqrtf(d.radius * d.radius);
  test_radius *= d.normal_radius_factor;
  const float radius_sq = test_radius * test_radius;
  long long n0x = 0, n0y = 0, n0z = 0, n1x = 0, n1y = 0, n1z = 0;
  long long c0x = 0, c0y = 0, c0z = 0, c1x = 0, c1y = 0, c1z = 0;
  long long cnt0 = 0, cnt1 = 0;
  const int total = __ldcg(&st->atile_count);
  for (int u = cta; u < total; u += ncta) {
    const int4 ent = __ldcg(&alist[u]);
    const int nvalid = ent.z - 4 * tid;
    if (nvalid <= 0) continue;
    const int s0 = ent.y + 4 * tid;
    /* a leaf the stroke has touched before this dab has its stroke-start state in the snapshot arrays */
    const bool snap = use_orig && !(ent.w & DSC_ENT_FIRST);
    const float4 X = ld4(snap ? m.ox : m.cx, s0), Y = ld4(snap ? m.oy : m.cy, s0), Z = ld4(snap ? m.oz : m.cz, s0);
    const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
    const unsigned hid = dsc_hidden4(m, s0);
    float dxs[4], dys[4], dzs[4], dsq[4];
    bool any = false;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      dxs[j] = xs[j] - d.loc[0]; dys[j] = ys[j] - d.loc[1]; dzs[j] = zs[j] - d.loc[2];
      dsq[j] = dsc_test_distsq(d, xs[j], ys[j], zs[j]);
      any |= (j < nvalid) && !((hid >> j) & 1u) && !(dsq[j] > radius_sq);
    }
    if (!any) continue;
    const float4 NX = ld4(snap ? m.onx : m.nx, s0), NY = ld4(snap ? m.ony : m.ny, s0), NZ = ld4(snap ? m.onz : m.nz, s0);
    const float vxs[4] = {NX.x, NX.y, NX.z, NX.w}, vys[4] = {NY.x, NY.y, NY.z, NY.w}, vzs[4] = {NZ.x, NZ.y, NZ.z, NZ.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (j >= nvalid || ((hid >> j) & 1u) || dsq[j] > radius_sq) continue;
      const float vx = vxs[j], vy = vys[j], vz = vzs[j];
      const bool flip = (d.view_n[0] * vx + d.view_n[1] * vy + d.view_n[2] * vz) <= 0.0f;
      const float q = 1.0f - (sqrtf(dsq[j]) / test_radius);
      const float f = dsc_clamp(3.0f * q * q - 2.0f * q * q * q, 0.0f, 1.0f);
      if (use_cos) {
        const float w = 1.0f - f;
        const long long ax = dsc_fix32((dxs[j] * w) / test_radius);
        const long long ay = dsc_fix32((dys[j] * w) / test_radius);
        const long long az = dsc_fix32((dzs[j] * w) / test_radius);
        if (flip) { c1x += ax; c1y += ay; c1z += az; }
        else { c0x += ax; c0y += ay; c0z += az; }
      }
      const long long bx = dsc_fix32(vx * f), by = dsc_fix32(vy * f), bz = dsc_fix32(vz * f);
      if (flip) { n1x += bx; n1y += by; n1z += bz; cnt1++; }
      else { n0x += bx; n0y += by; n0z += bz; cnt0++; }
    }
  }
  /* most CTAs of a small dab sampled nothing: leave before the reduction */
  if (!__syncthreads_or((cnt0 | cnt1) != 0)) return;
  long long v[16] = {n0x, n0y, n0z, n1x, n1y, n1z, c0x, c0y, c0z, c1x, c1y, c1z, cnt0, cnt1,
                     use_cos ? cnt0 : 0, use_cos ? cnt1 : 0};
#pragma unroll
  for (int k = 0; k < 16; k++) {
    long long x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0 && x != 0) atomicAdd(&sacc[k], (unsigned long long)x);
  }
  __syncthreads();
  if (tid < 16 && sacc[tid] != 0ull) atomicAdd((unsigned long long *)&st->acc[tid], sacc[tid]);
}
__global__ void __launch_bounds__(DSC_BLOCK) k_area(DevMesh m, int j, int slot)
{
  dsc_pdl_wait(); /* the gather's area tile list */
  dsc_pdl_launch();
  const DabEntry &e = dsc_dab_entry(m, j);
  const DabParams d = e.d;
  dsc_area_body(m, d, e.use_cos, slot, blockIdx.x, gridDim.x);
}

/* finalisation of the sums, same float/double steps as the CPU path */
__device__ __forceinline__ void dsc_area_finalize(const DabState *st, const DabParams &d, bool use_cos, float no[3],
                                                  float co[3])
{
  no[0] = no[1] = no[2] = 0.0f;
  for (int i = 0; i < 2; i++) {
    float tx = (float)((double)__ldcg(&st->acc[i * 3 + 0]) * (1.0 / 4294967296.0));
    float ty = (float)((double)__ldcg(&st->acc[i * 3 + 1]) * (1.0 / 4294967296.0));
    float tz = (float)((double)__ldcg(&st->acc[i * 3 + 2]) * (1.0 / 4294967296.0));
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
      const long long cnt = __ldcg(&st->acc[14 + i]);
      if (cnt == 0) continue;
      for (int k = 0; k < 3; k++) {
        const double mean = (double)__ldcg(&st->acc[6 + i * 3 + k]) / ((double)cnt * 4294967296.0);
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
  float grab[3]; /* grab: the drag, blended towards the sculpt normal when the brush has a normal weight */
  int flip, skip;
};

/* per-dab constants every CTA derives from the dab descriptor and the area sums */
__device__ void dsc_brush_derive(DabState *st, const DabParams &d, BrushDerived &D, bool publish)
{
  D.skip = 0;
  float an[3] = {0, 0, 0}, ac[3] = {d.loc[0], d.loc[1], d.loc[2]};
  if (d.tool == 1) {
    dsc_sculpt_normal(st, d, an);
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
      dsc_area_finalize(st, d, true, an, ac);
      area_no[0] = an[0]; area_no[1] = an[1]; area_no[2] = an[2];
    }
    else {
      dsc_sculpt_normal(st, d, an);
      dsc_area_finalize(st, d, true, area_no, ac);
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
  else if (d.tool == 5) {
    D.grab[0] = d.grab_delta[0]; D.grab[1] = d.grab_delta[1]; D.grab[2] = d.grab_delta[2];
    if (d.normal_weight > 0.0f) {
      /* row a19, sculpt_project_v3_normal_align: the drag blended towards the sculpt normal, scaled to follow the cursor */
      dsc_sculpt_normal(st, d, an);
      const float len_signed = an[0] * D.grab[0] + an[1] * D.grab[1] + an[2] * D.grab[2];
      const float fac = an[0] * d.view_n[0] + an[1] * d.view_n[1] + an[2] * d.view_n[2];
      const float vax = an[0] - d.view_n[0] * fac, vay = an[1] - d.view_n[1] * fac, vaz = an[2] - d.view_n[2] * fac;
      float lvs = fabsf(vax * an[0] + vay * an[1] + vaz * an[2]);
      lvs = (lvs > 1.1920929e-07f) ? 1.0f / lvs : 1.0f;
      const float w = (len_signed * d.normal_weight) * lvs;
      for (int k = 0; k < 3; k++) {
        D.grab[k] = D.grab[k] * (1.0f - d.normal_weight);
        D.grab[k] = D.grab[k] + an[k] * w;
      }
    }
  }
  if (publish) {
    for (int k = 0; k < 3; k++) {
      st->area_no[k] = an[k];
      st->area_co[k] = ac[k];
    }
  }
}

/* one vertex of the brush loop; returns true if it was displaced (new position in x, y, z) */
template<int TOOL>
__device__ __forceinline__ bool dsc_brush_vertex(const DevMesh &m, const DabParams &d, const BrushDerived &D, int s,
                                                 float &x, float &y, float &z, float tx, float ty, float tz, float vnx,
                                                 float vny, float vnz, float radius_sq)
{
  constexpr int tool = TOOL;
  if (tool == 18) {
    /* clay strips: brush-local cube test + plane */
    const float rx = x - D.origin[0], ry = y - D.origin[1], rz = z - D.origin[2];
    float local[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      local[k] = fabsf((rx * D.ax[k][0] + ry * D.ax[k][1] + rz * D.ax[k][2]) / D.sc[k]);
    }
    const float side = 1.0f;
    if (!(local[0] <= side && local[1] <= side && local[2] <= side)) return false;
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
    if (!(side_d <= 0.0f)) return false;
    const float pd = (D.plane_no[0] * x + D.plane_no[1] * y + D.plane_no[2] * z) + D.plane_d;
    const float ix = x + D.plane_no[0] * (-pd), iy = y + D.plane_no[1] * (-pd), iz = z + D.plane_no[2] * (-pd);
    const float vx = ix - x, vy = iy - y, vz = iz - z;
    if ((d.flags & 2) && !((vx * vx + vy * vy + vz * vz) <= D.trim_sq)) return false;
    const float fade = D.bstrength * dsc_strength_factor(m, d, d.radius * dist, vnx, vny, vnz, s);
    const float px = vx * fade, py = vy * fade, pz = vz * fade;
    dsc_clip(d, x, y, z, x + px, y + py, z + pz);
    return true;
  }
  /* sphere / tube test (row a11); grab tests and offsets the stroke-start coordinates (row a19) */
  const float distsq = dsc_test_distsq(d, tx, ty, tz);
  if (distsq > radius_sq) return false;
  float fade = dsc_strength_factor(m, d, sqrtf(distsq), vnx, vny, vnz, s);
  if (tool == 1) {
    const float px = D.offset[0] * fade, py = D.offset[1] * fade, pz = D.offset[2] * fade;
    dsc_clip(d, x, y, z, x + px, y + py, z + pz);
  }
  else if (tool == 4) {
    fade = d.bstrength * fade;
    const float sc = fade * d.radius;
    const float px = (vnx * sc) * d.scale[0], py = (vny * sc) * d.scale[1], pz = (vnz * sc) * d.scale[2];
    dsc_clip(d, x, y, z, x + px, y + py, z + pz);
  }
  else {
    fade = d.bstrength * fade;
    const float px = D.grab[0] * fade, py = D.grab[1] * fade, pz = D.grab[2] * fade;
    dsc_clip(d, x, y, z, tx + px, ty + py, tz + pz);
  }
  return true;
}

/* Draw / inflate / grab / clay strips over the unique verts of hit leaves (SURVEY.md 8a rows
 * a11-a19), one instantiation per tool.  A CTA takes one tile at a time; each thread owns 4
 * consecutive slots (float4 loads / stores of the SoA arrays).  First touch of a leaf in the stroke
 * snapshots co/no into orig_co/orig_no before the vertex is moved (row a9).  Displaced verts get
 * their vert_bitmap bit (pbvh.c:3729). */
/* four consecutive slots from s0, positions already in X / Y / Z: snapshot on first touch, displacement; returns the
 * moved nibble and the new positions (the caller stores them) */
template<int TOOL>
__device__ __forceinline__ unsigned dsc_brush_quad_core(const DevMesh &m, const DabParams &d, const BrushDerived &D, bool first, int s0,
                                                        int nvalid, float radius_sq, bool need_no, float4 &X, float4 &Y, float4 &Z)
{
  constexpr int tool = TOOL;
  constexpr bool use_orig = (tool == 5);
  unsigned nib = 0;
  float4 NX = make_float4(0, 0, 0, 0), NY = NX, NZ = NX;
  float4 TX = X, TY = Y, TZ = Z;
  if (first) {
    NX = ld4(m.nx, s0); NY = ld4(m.ny, s0); NZ = ld4(m.nz, s0);
    st4(m.ox, s0, X); st4(m.oy, s0, Y); st4(m.oz, s0, Z);
    st4(m.onx, s0, NX); st4(m.ony, s0, NY); st4(m.onz, s0, NZ);
  }
  else if (use_orig) {
    TX = ld4(m.ox, s0); TY = ld4(m.oy, s0); TZ = ld4(m.oz, s0);
  }
  if (tool == 18 && D.skip) return 0u;
  float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
  const float txs[4] = {TX.x, TX.y, TX.z, TX.w}, tys[4] = {TY.x, TY.y, TY.z, TY.w}, tzs[4] = {TZ.x, TZ.y, TZ.z, TZ.w};
  const unsigned hid = dsc_hidden4(m, s0);
  if (need_no && !first) {
    /* only needed for verts inside; one cheap pre-test keeps the streaming case at 12 B/vert */
    bool any = (tool == 18);
#pragma unroll
    for (int j = 0; j < 4; j++) any |= !(dsc_test_distsq(d, txs[j], tys[j], tzs[j]) > radius_sq);
    if (any) {
      if (use_orig) { NX = ld4(m.onx, s0); NY = ld4(m.ony, s0); NZ = ld4(m.onz, s0); }
      else { NX = ld4(m.nx, s0); NY = ld4(m.ny, s0); NZ = ld4(m.nz, s0); }
    }
  }
  const float vxs[4] = {NX.x, NX.y, NX.z, NX.w}, vys[4] = {NY.x, NY.y, NY.z, NY.w}, vzs[4] = {NZ.x, NZ.y, NZ.z, NZ.w};
#pragma unroll
  for (int j = 0; j < 4; j++) {
    if (j < nvalid && !((hid >> j) & 1u) &&
        dsc_brush_vertex<TOOL>(m, d, D, s0 + j, xs[j], ys[j], zs[j], txs[j], tys[j], tzs[j], vxs[j], vys[j], vzs[j], radius_sq)) {
      nib |= 1u << j;
    }
  }
  if (nib) {
    X = make_float4(xs[0], xs[1], xs[2], xs[3]);
    Y = make_float4(ys[0], ys[1], ys[2], ys[3]);
    Z = make_float4(zs[0], zs[1], zs[2], zs[3]);
  }
  return nib;
}
/* the same straight on the global arrays */
template<int TOOL>
__device__ __forceinline__ unsigned dsc_brush_quad(const DevMesh &m, const DabParams &d, const BrushDerived &D, bool first, int s0, int nvalid,
                                                   float radius_sq, bool need_no)
{
  if (nvalid <= 0) return 0u;
  float4 X = ld4(m.cx, s0), Y = ld4(m.cy, s0), Z = ld4(m.cz, s0);
  const unsigned nib = dsc_brush_quad_core<TOOL>(m, d, D, first, s0, nvalid, radius_sq, need_no, X, Y, Z);
  if (nib) {
    st4(m.cx, s0, X);
    st4(m.cy, s0, Y);
    st4(m.cz, s0, Z);
  }
  return nib;
}

/* Draw / inflate / grab / clay strips over the unique verts of hit leaves (SURVEY.md 8a rows
 * a11-a19), one instantiation per tool.  Each thread owns 4 consecutive slots (float4 loads / stores of the SoA
 * arrays).  First touch of a leaf in the stroke snapshots co/no into orig_co/orig_no before the vertex is moved
 * (row a9).  Displaced verts get their vert_bitmap bit (pbvh.c:3729).
 * BOUNDARY = false: a CTA takes one whole tile at a time (the stand-alone brush pass: partitioned PBVHs, which
 * exchange halo positions between the brush and the normals).
 * BOUNDARY = true: a warp takes the boundary run of one tile at a time -- the unique verts other tiles read, which
 * come last in the tile (slots [ibnd, count)); the interiors are displaced inside the fused tile kernel. */
template<int TOOL, bool BOUNDARY>
__device__ __forceinline__ void dsc_brush_body(const DevMesh &m, const DabParams &d, int slot, int cta, int ncta)
{
  DabState *st = m.st + slot;
  const int4 *tl = m.tile_list + (size_t)slot * m.ntile;
  __shared__ BrushDerived D;
  __shared__ unsigned s_moved;
  const int tid = threadIdx.x, lane = tid & 31;
  const int total = __ldcg(&st->tile_count);
  const int first_unit = BOUNDARY ? cta * (DSC_BLOCK / 32) : cta;
  if (first_unit >= total && cta != 0) return; /* CTA 0 publishes the plane even when nothing was gathered */
  __syncthreads(); /* D / s_moved of an earlier call */
  if (tid == 0) {
    s_moved = 0;
    dsc_brush_derive(st, d, D, cta == 0);
    /* statistics: verts the area pass averaged (its M') */
    if (cta == 0) m.tot->area_inside_total += (unsigned long long)(__ldcg(&st->acc[12]) + __ldcg(&st->acc[13]));
  }
  __syncthreads();
  constexpr int tool = TOOL;
  const float radius_sq = d.radius * d.radius;
  const bool need_no = (tool == 4) || (d.flags & 1);
  unsigned moved_cnt = 0;
  if (BOUNDARY) {
    const int nw = DSC_BLOCK / 32;
    for (int u = cta * nw + (tid >> 5); u < total; u += ncta * nw) { /* warp-uniform */
      const int4 ent = __ldcg(&tl[u]);
      const bool first = (ent.w & DSC_ENT_FIRST) != 0;
      const int ibnd = (ent.w >> DSC_ENT_IBND_SHIFT) & 0x7ff;
      for (int o = ibnd; o < ent.z; o += 128) {
        const int s0 = ent.y + o + 4 * lane;
        const unsigned nib = dsc_brush_quad<TOOL>(m, d, D, first, s0, ent.z - o - 4 * lane, radius_sq, need_no);
        const unsigned w = dsc_pack_nibbles(nib, lane);
        if (w && (lane & 7) == 0) {
          atomicOr(&m.dirty[s0 >> 5], w); /* fire-and-forget RED: no round trip for the old word */
          if (m.capture) atomicOr(&m.capture[s0 >> 5], w);
          moved_cnt += __popc(w);
        }
      }
    }
  }
  else {
    for (int u = cta; u < total; u += ncta) {
      const int4 ent = __ldcg(&tl[u]);
      const bool first = (ent.w & DSC_ENT_FIRST) != 0;
      const int s0 = ent.y + 4 * tid;
      const unsigned nib = dsc_brush_quad<TOOL>(m, d, D, first, s0, ent.z - 4 * tid, radius_sq, need_no);
      const unsigned w = dsc_pack_nibbles(nib, lane);
      if (w && (lane & 7) == 0) {
        atomicOr(&m.dirty[s0 >> 5], w);
        if (m.capture) atomicOr(&m.capture[s0 >> 5], w);
        moved_cnt += __popc(w);
      }
    }
  }
  if (moved_cnt) atomicAdd(&s_moved, moved_cnt);
  __syncthreads();
  if (tid == 0 && s_moved) atomicAdd(&m.tot->moved_total, (unsigned long long)s_moved);
}
template<int TOOL> __global__ void __launch_bounds__(DSC_BLOCK) k_brush(DevMesh m, int j, int slot)
{
  dsc_pdl_wait(); /* the area sums (or, without an area pass, the gather) */
  dsc_pdl_launch();
  const DabParams d = dsc_dab_entry(m, j).d;
  dsc_brush_body<TOOL, false>(m, d, slot, blockIdx.x, gridDim.x);
}
/* the boundary runs of the gathered tiles, ahead of the fused tile kernel */
template<int TOOL> __global__ void __launch_bounds__(DSC_BLOCK) k_brush_boundary(DevMesh m, int j, int slot)
{
  const DabParams d = dsc_dab_entry(m, j).d;
  dsc_brush_body<TOOL, true>(m, d, slot, blockIdx.x, gridDim.x);
}

/* snapshot only (smooth brush: first touch, before iteration 0) */
__global__ void __launch_bounds__(DSC_BLOCK) k_snapshot(DevMesh m, int slot)
{
  const DabState *st = m.st + slot;
  const int4 *tl = m.tile_list + (size_t)slot * m.ntile;
  const int tid = threadIdx.x;
  const int total = st->tile_count;
  for (int u = blockIdx.x; u < total; u += gridDim.x) {
    const int4 ent = tl[u];
    if (!(ent.w & DSC_ENT_FIRST)) continue;
    if (ent.z - 4 * tid <= 0) continue;
    const int s0 = ent.y + 4 * tid;
    st4(m.ox, s0, ld4(m.cx, s0)); st4(m.oy, s0, ld4(m.cy, s0)); st4(m.oz, s0, ld4(m.cz, s0));
    st4(m.onx, s0, ld4(m.nx, s0)); st4(m.ony, s0, ld4(m.ny, s0)); st4(m.onz, s0, ld4(m.nz, s0));
  }
}

/* ------------------------------------------------------------------------------- K4 smooth */
/* One Jacobi iteration, part A: new position of every unique vert of a hit leaf inside the
 * sphere = co + (neighbour average - co) * fade, into the scratch arrays (SURVEY.md 8a row a20:
 * interior verts average all edge neighbours, boundary verts only boundary neighbours, boundary
 * verts with <= 2 neighbours stay).  The CSR gather is served by L2.
 * GRIDS: the neighbours of a grid element come from its place in the grid (subdiv_ccg.c:1870-1880: previous row,
 * next row, previous column, next column -- four unit-stride streams, no index traffic) or, on the grid's rim,
 * from the per-grid rim table (KERNEL_subdiv_ccg_neighbor_coords_get asked once at upload, subdiv_ccg.c:1882-1909). */
struct GridNb {
  int gs, gs2, rim_w;
  const int *leaf_gbeg, *leaf_grids;
  const int *rim_nb;              /* [totgrid][4 gs - 4][rim_w] slots, -1 = none */
  const unsigned char *rim_bnd;   /* [totgrid][4 gs - 4] or NULL */
};
template<bool GRIDS> __global__ void __launch_bounds__(DSC_BLOCK) k_smooth_a(DevMesh m, GridNb gn, int j, int slot, int last_iteration)
{
  const DabEntry &ent_ = dsc_dab_entry(m, j);
  const DabParams d = ent_.d;
  const float strength = last_iteration ? ent_.smooth_last : 1.0f;
  const DabState *st = m.st + slot;
  const int4 *tl = m.tile_list + (size_t)slot * m.ntile;
  __shared__ unsigned s_moved;
  /* grids: the rim elements of a tile (about one in sixteen, but one in most warps) are set aside and finished by a
   * few converged warps after the interior ones, instead of dragging every warp through the table lookup */
  __shared__ int s_rim_n;
  __shared__ int s_rim_slot[GRIDS ? DSC_TILE : 1];
  __shared__ float s_rim_fade[GRIDS ? DSC_TILE : 1];
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) s_moved = 0;
  __syncthreads();
  const float radius_sq = d.radius * d.radius;
  unsigned moved_cnt = 0;
  const int total = st->tile_count;
  for (int u = blockIdx.x; u < total; u += gridDim.x) {
    const int4 ent = tl[u];
    const int beg = ent.y, cnt = ent.z;
    const int cnt32 = (cnt + 31) & ~31;
    int leaf = 0, leaf_beg = 0;
    if (GRIDS) {
      leaf = m.tile_meta[3 * ent.x + 2].y;
      leaf_beg = m.leaf_ubeg[leaf];
      if (tid == 0) s_rim_n = 0;
      __syncthreads();
    }
    for (int i = tid; i < cnt32; i += DSC_BLOCK) {
      const int s = beg + i;
      bool moved = false;
      if (i < cnt && !(m.hidden && ((m.hidden[s >> 5] >> (s & 31)) & 1u))) {
        const float x = m.cx[s], y = m.cy[s], z = m.cz[s];
        const float distsq = dsc_test_distsq(d, x, y, z);
        if (!(distsq > radius_sq)) {
          float vnx = 0.0f, vny = 0.0f, vnz = 0.0f;
          if (d.flags & 1) { vnx = m.nx[s]; vny = m.ny[s]; vnz = m.nz[s]; }
          const float fade = strength * dsc_strength_factor(m, d, sqrtf(distsq), vnx, vny, vnz, s);
          float ax = 0.0f, ay = 0.0f, az = 0.0f;
          int tot = 0, neighbor_count = 0;
          bool is_boundary = false, deferred = false;
          if (GRIDS) {
            const int local = s - leaf_beg;
            const int gi = local / gn.gs2, e = local - gi * gn.gs2;
            const int ey = e / gn.gs, ex = e - ey * gn.gs, last = gn.gs - 1;
            if (ex > 0 && ey > 0 && ex < last && ey < last) {
              const int a = s - gn.gs, b = s + gn.gs;
              ax += m.cx[a]; ay += m.cy[a]; az += m.cz[a];
              ax += m.cx[b]; ay += m.cy[b]; az += m.cz[b];
              ax += m.cx[s - 1]; ay += m.cy[s - 1]; az += m.cz[s - 1];
              ax += m.cx[s + 1]; ay += m.cy[s + 1]; az += m.cz[s + 1];
              tot = neighbor_count = 4;
            }
            else {
              const int k = atomicAdd(&s_rim_n, 1);
              s_rim_slot[k] = s;
              s_rim_fade[k] = fade;
              deferred = true;
            }
          }
          else {
            const unsigned qb = m.nb_off[s], qe = m.nb_off[s + 1];
            neighbor_count = (int)(qe - qb);
            is_boundary = m.boundary[s] != 0;
            for (unsigned q = qb; q < qe; q++) {
              const int v = m.nb_idx[q];
              if (!is_boundary || m.boundary[v]) {
                ax += m.cx[v]; ay += m.cy[v]; az += m.cz[v];
                tot++;
              }
            }
          }
          if (!deferred) {
            float rx, ry, rz;
            if ((neighbor_count <= 2 && is_boundary) || tot == 0) {
              rx = x; ry = y; rz = z;
            }
            else {
              const float f = 1.0f / (float)tot;
              rx = ax * f; ry = ay * f; rz = az * f;
            }
            const float vx = rx - x, vy = ry - y, vz = rz - z;
            float cx0 = x, cy0 = y, cz0 = z;
            dsc_clip(d, cx0, cy0, cz0, x + vx * fade, y + vy * fade, z + vz * fade);
            m.tx[s] = cx0;
            m.ty[s] = cy0;
            m.tz[s] = cz0;
          }
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
    if (GRIDS) {
      __syncthreads();
      const int nrim = s_rim_n;
      for (int k = tid; k < nrim; k += DSC_BLOCK) {
        const int s = s_rim_slot[k];
        const float fade = s_rim_fade[k];
        const float x = m.cx[s], y = m.cy[s], z = m.cz[s];
        const int local = s - leaf_beg;
        const int gi = local / gn.gs2, e = local - gi * gn.gs2;
        const int ey = e / gn.gs, ex = e - ey * gn.gs, last = gn.gs - 1;
        const int grid = gn.leaf_grids[gn.leaf_gbeg[leaf] + gi];
        const int rim = 4 * gn.gs - 4;
        const int b = ey == 0 ? ex : (ey == last ? gn.gs + ex : (ex == 0 ? 2 * gn.gs + ey - 1 : 3 * gn.gs - 2 + ey - 1));
        const int *row = gn.rim_nb + ((size_t)grid * rim + b) * gn.rim_w;
        const bool is_boundary = gn.rim_bnd && gn.rim_bnd[(size_t)grid * rim + b] != 0;
        float ax = 0.0f, ay = 0.0f, az = 0.0f;
        int tot = 0, neighbor_count = 0;
        for (int q = 0; q < gn.rim_w; q++) {
          const int v = row[q];
          if (v < 0) break;
          neighbor_count++;
          if (!is_boundary || m.boundary[v]) {
            ax += m.cx[v]; ay += m.cy[v]; az += m.cz[v];
            tot++;
          }
        }
        float rx, ry, rz;
        if ((neighbor_count <= 2 && is_boundary) || tot == 0) {
          rx = x; ry = y; rz = z;
        }
        else {
          const float f = 1.0f / (float)tot;
          rx = ax * f; ry = ay * f; rz = az * f;
        }
        const float vx = rx - x, vy = ry - y, vz = rz - z;
        float cx0 = x, cy0 = y, cz0 = z;
        dsc_clip(d, cx0, cy0, cz0, x + vx * fade, y + vy * fade, z + vz * fade);
        m.tx[s] = cx0;
        m.ty[s] = cy0;
        m.tz[s] = cz0;
      }
      __syncthreads(); /* the list is reset for the next tile */
    }
  }
  if (moved_cnt) atomicAdd(&s_moved, moved_cnt);
  __syncthreads();
  if (tid == 0 && s_moved) atomicAdd(&m.tot->moved_total, (unsigned long long)s_moved);
}

/* part B: commit the scratch positions.  Four slots per thread: a group whose four moved bits are all set is one
 * float4 copy per coordinate (the interior of the brush sphere), a mixed group goes slot by slot -- so a tile is one
 * iteration with three independent 16-byte loads in flight instead of four dependent scalar rounds. */
__global__ void __launch_bounds__(DSC_BLOCK) k_smooth_b(DevMesh m, int slot)
{
  const DabState *st = m.st + slot;
  const int4 *tl = m.tile_list + (size_t)slot * m.ntile;
  const int total = st->tile_count;
  for (int u = blockIdx.x; u < total; u += gridDim.x) {
    const int4 ent = tl[u];
    const int beg = ent.y, cnt = ent.z; /* beg is a multiple of 32 */
    for (int i = 4 * threadIdx.x; i < cnt; i += 4 * DSC_BLOCK) {
      const int s = beg + i;
      const unsigned bits = (m.iter_moved[s >> 5] >> (s & 31)) & 0xfu;
      if (bits == 0u) continue;
      if (bits == 0xfu && i + 4 <= cnt) {
        st4(m.cx, s, ld4(m.tx, s)); st4(m.cy, s, ld4(m.ty, s)); st4(m.cz, s, ld4(m.tz, s));
      }
      else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (i + k < cnt && ((bits >> k) & 1u)) {
            m.cx[s + k] = m.tx[s + k]; m.cy[s + k] = m.ty[s + k]; m.cz[s + k] = m.tz[s + k];
          }
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------ K5 normals */
/* BKE_pbvh_update_normals for PBVH_FACES (pbvh.c:2912-3036), general gather form: for every dirty
 * unique vert of a listed (flagged) leaf, normal = normalize(sum of the poly normals of its
 * looptris), summed in ascending looptri position; then the dirty bit is cleared.  Like the
 * reference's accumulate pass, looptris of leaves that are not flagged contribute nothing
 * (pbvh.c:2943).  Handles any poly size and leaf size; leaves that fit the shared-memory kernel
 * below are skipped when skip_fast is set. */
/* is the leaf flagged for a normals update -- here, or (multi-GPU) gathered this dab by its owner */
__device__ __forceinline__ bool dsc_leaf_updates_normals(const DevMesh &m, const unsigned *ghit, int leaf)
{
  if (m.node_flag[leaf] & F_UpdateNormals) return true;
  return ghit && ((ghit[leaf >> 5] >> (leaf & 31)) & 1u);
}

__global__ void __launch_bounds__(DSC_BLOCK) k_normals(DevMesh m, const int *list, const int *count, int skip_fast,
                                                       const unsigned *ghit)
{
  const int tid = threadIdx.x, lane = tid & 31;
  const int total = *count * m.max_chunks;
  for (int u = blockIdx.x; u < total; u += gridDim.x) {
    int leaf, beg, cnt;
    if (!dsc_unit(m, list, u, leaf, beg, cnt)) continue;
    if (skip_fast && m.leaf_fast[leaf]) continue;
    if (!(m.node_flag[leaf] & F_UpdateNormals)) continue;
    const unsigned pb = (unsigned)m.leaf_pbeg[leaf], pe = pb + (unsigned)m.leaf_pcnt[leaf];
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
            if (!dsc_leaf_updates_normals(m, ghit, m.tri_leaf[pos])) continue;
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

/* ----------------------------------------------------------- K5 + K6 fused, shared-memory form */
#define NT_BLOCK 256 /* compute threads of the tile kernel */
#define NB_NORMALS 1
#define NB_BOUNDS 2
/* TileMeta (3 x int4 per tile), dsc_tile_nloc_a, dsc_tile_smem_bytes and the host-side construction of the tile tables */
#include "dsc_tile_tables.h"

/* --- TMA 1-D bulk copies (cp.async.bulk) completing on an mbarrier --- */
__device__ __forceinline__ unsigned dsc_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dsc_mbar_init(unsigned long long *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dsc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void dsc_mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(dsc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dsc_mbar_arrive(unsigned long long *bar)
{
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(dsc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dsc_mbar_wait(unsigned long long *bar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DSC_WAIT_%=:\n"
      "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DSC_DONE_%=;\n"
      "bra DSC_WAIT_%=;\n"
      "DSC_DONE_%=:\n"
      "}\n" ::"r"(dsc_smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void dsc_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dsc_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(dsc_smem_u32(bar))
               : "memory");
}

/* One CTA per listed tile (a spatially compact run of <= DSC_TILE unique verts of one leaf), several
 * CTAs per SM so the phases of different tiles overlap.
 * Phase 0 reads the tile's dirty words; a tile with no dirty vert only refreshes its box.
 * Phase 1: one thread queues TMA bulk copies of the tile's contiguous inputs -- the three position
 * runs of its unique verts, its poly entries, its index words -- into shared memory, completing on
 * an mbarrier; meanwhile all threads gather the staged verts (corners of the tile's polys owned by
 * other tiles / leaves, and the leaf's shared verts assigned to it for the box: the only gathers
 * left on the path) and arrive on the same mbarrier.  The tile box is reduced from shared memory;
 * the last tile of a leaf to finish merges the tile boxes into the leaf box (update_node_vb,
 * pbvh.c:2033-2041).
 * Phase 2 computes, once, the normal of every poly incident to the tile's unique verts
 * (BKE_mesh_calc_poly_normal), own-leaf entries always, entries of looptris held by another leaf
 * only if that leaf updates its normals too (pbvh.c:2943).
 * Phase 3 sums, per dirty unique vert, the normals of its looptris in ascending looptri position
 * (pbvh.c:2933-2981 run single-threaded), normalises, stores, clears the dirty bit.  Padding of the
 * index rows points at a stored zero vector (adding +0 is exact).
 * Leaves that do not fit (fast == 0) are left to k_normals / k_leaf_bb. */
/* normal of one local poly entry from the staged positions */
template<bool ALLQUAD>
__device__ __forceinline__ float4 dsc_entry_normal(const float *PX, const float *PY, const float *PZ, const ushort4 v)
{
  const float ax = PX[v.x], ay = PY[v.x], az = PZ[v.x];
  const float bx = PX[v.y], by = PY[v.y], bz = PZ[v.y];
  const float cx = PX[v.z], cy = PY[v.z], cz = PZ[v.z];
  float n1x, n1y, n1z, n2x, n2y, n2z;
  if (ALLQUAD || v.w != 0xffffu) {
    /* normal_quad_v3, lib/intern/math_geom.cc:51-69 */
    const float dx = PX[v.w], dy = PY[v.w], dz = PZ[v.w];
    n1x = ax - cx; n1y = ay - cy; n1z = az - cz;
    n2x = bx - dx; n2y = by - dy; n2z = bz - dz;
  }
  else {
    /* normal_tri_v3, lib/intern/math_geom.cc:31-49 */
    n1x = ax - bx; n1y = ay - by; n1z = az - bz;
    n2x = bx - cx; n2y = by - cy; n2z = bz - cz;
  }
  float ox = n1y * n2z - n1z * n2y;
  float oy = n1z * n2x - n1x * n2z;
  float oz = n1x * n2y - n1y * n2x;
  dsc_normalize(ox, oy, oz);
  return make_float4(ox, oy, oz, 0.0f);
}

template<bool ALLQUAD>
__device__ __forceinline__ void dsc_tile_poly_normals(const float *PX, const float *PY, const float *PZ, const ushort4 *E,
                                                      float4 *F, const unsigned char *H, const unsigned *sdirty, int tid,
                                                      int U, int eown, int ne, bool sparse)
{
  if (!sparse) {
#pragma unroll 2
    for (int e = tid; e < eown; e += NT_BLOCK) F[e] = dsc_entry_normal<ALLQUAD>(PX, PY, PZ, E[e]);
    for (int e = eown + tid; e < ne; e += NT_BLOCK) {
      F[e] = H[e - eown] ? dsc_entry_normal<ALLQUAD>(PX, PY, PZ, E[e]) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
  }
  else {
    /* few verts of the tile are dirty: entries that touch no dirty unique vert are skipped (nothing reads them) */
    for (int e = tid; e < ne; e += NT_BLOCK) {
      const ushort4 v = E[e];
      bool need = false;
      const unsigned li[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (li[k] < (unsigned)U) need |= ((sdirty[li[k] >> 5] >> (li[k] & 31)) & 1u) != 0u;
      }
      if (!need) continue;
      F[e] = (e < eown || H[e - eown]) ? dsc_entry_normal<ALLQUAD>(PX, PY, PZ, v) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
  }
}

/* order-free float min / max into global memory (fire-and-forget RED: no return value, no stall) */
__device__ __forceinline__ void dsc_red_min(float *a, float v)
{
  v = v + 0.0f; /* -0 -> +0 */
  if (v >= 0.0f) atomicMin(reinterpret_cast<int *>(a), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned *>(a), __float_as_uint(v));
}
__device__ __forceinline__ void dsc_red_max(float *a, float v)
{
  v = v + 0.0f;
  if (v >= 0.0f) atomicMax(reinterpret_cast<int *>(a), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned *>(a), __float_as_uint(v));
}

#define NT_CONSUMERS 256                 /* 8 compute warps */
#define NT_THREADS (NT_CONSUMERS + 32)   /* + 1 producer warp */
#define NT_UPD_WORDS 512                 /* leaf bitmask words kept in shared memory */
#define NT_ACTIVE 1
#define NT_ANYD 2
#define NT_DO_B 4
#define NT_DO_N 8
#define NT_FIRST 16

/* Producer / consumer form.  Warp 8 is the producer: for the NEXT tile it reads the descriptor, the
 * dirty words and the index-row offsets, publishes them in shared memory (double buffered), queues
 * the TMA bulk copies, gathers the staged verts and resolves the other-leaf switches -- all of it
 * while the 8 consumer warps compute the current tile, so no compute warp ever waits on a global
 * load.  Buffers are handed over with mbarriers: full[0] / empty[0] guard positions + entries +
 * staged verts (consumed by the box reduction and phase 2), full[1] / empty[1] the index words
 * (consumed by phase 3).  The consumers synchronise among themselves with a named barrier.
 *
 * TOOL != 0 is the fused dab kernel: the tile's INTERIOR unique verts (slots [0, ibnd) of the tile: read by no other
 * tile) are displaced by the brush right here, between the arrival of the positions in shared memory and the box /
 * normal phases -- their new positions go to shared memory (phases 2-3 read them there) and to global memory, and
 * their dirty bits live in shared memory only.  The boundary runs [ibnd, count) of all gathered tiles were displaced by
 * k_brush_boundary before this kernel started, so every position this tile reads from another tile (staged verts) and
 * the boundary part of its own bulk copy are post-brush already: same arithmetic, same bits as the two-pass form, one
 * read of the positions less.  With a brush the dirty state of a tile is not known when its loads are queued: the poly
 * entries, index words and other-leaf switches of every gathered tile are loaded. */
template<int TOOL>
__device__ __forceinline__ void dsc_normals_tile_body(const DevMesh &m, const int4 *list, const int *count, int mode,
                                                      const unsigned *upd, const int cta, const int ncta, const bool pdl,
                                                      const DabParams *dab = nullptr, DabState *dst = nullptr)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long s_full[2], s_empty[2];
  __shared__ __align__(16) int4 s_q[2][3];
  __shared__ __align__(16) int4 s_ent[2]; /* {tile, first slot, unique verts | ibnd << 16, NT_* flags | dirty count << 8} */
  __shared__ float red[6][NT_CONSUMERS / 32];
  __shared__ unsigned sdirty[2][DSC_TILE / 32];
  __shared__ unsigned sgoff[2][DSC_TILE / 32 + 1];
  __shared__ unsigned s_upd[NT_UPD_WORDS];
  __shared__ int s_dcount[2];
  __shared__ BrushDerived s_D;
  __shared__ unsigned s_moved;
  constexpr int NW = NT_CONSUMERS / 32;
  constexpr bool BRUSH = TOOL != 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tn = m.totnode;
  /* before the PDL wait: only what the gather (at least two launches back) and the upload wrote */
  const int n = __ldcg(count);
  if (cta >= n) {
    if (pdl) {
      dsc_pdl_wait();
      dsc_pdl_launch();
    }
    return;
  }
  __syncthreads(); /* an earlier call is done with the static shared arrays and the mbarriers */
  const bool upd_shared = m.ghit_words <= NT_UPD_WORDS;
  if (upd_shared) {
    for (int w = tid; w < m.ghit_words; w += NT_THREADS) s_upd[w] = __ldcg(&upd[w]);
  }
  if (tid == 0) {
    dsc_mbar_init(&s_full[0], 32);          /* the producer lanes (+ TMA bytes) */
    dsc_mbar_init(&s_full[1], 1);           /* TMA bytes only */
    dsc_mbar_init(&s_empty[0], NW);         /* one arrival per consumer warp */
    dsc_mbar_init(&s_empty[1], NW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (BRUSH) {
      s_moved = 0u;
      dsc_brush_derive(dst, *dab, s_D, false); /* the plane / offset every vertex of the dab shares */
    }
  }
  /* the producer's first descriptor: tile list entry, tile meta, index-row offsets (constant tables) */
  int4 ent = make_int4(0, 0, 0, 0), q = make_int4(0, 0, 0, 0);
  unsigned go = 0u, glast = 0u;
  if (warp == NW) {
    ent = __ldcg(&list[cta]);
    const int ng0 = (ent.z + 31) >> 5, g00 = ent.y >> 5;
    if (lane < 3) q = m.tile_meta[3 * ent.x + lane];
    go = (lane < ng0) ? m.v2_goff[g00 + lane] : 0u;
    glast = (lane == 0) ? m.v2_goff[g00 + ng0] : 0u;
  }
  if (pdl) {
    dsc_pdl_wait(); /* positions and dirty bits of the brush; the emptied leaf boxes */
    dsc_pdl_launch();
  }
  __syncthreads();

  if (warp == NW) {
    /* ------------------------------------------------------------------ producer warp */
    unsigned par_e0 = 1u, par_e1 = 1u; /* a fresh mbarrier passes a wait on the previous phase */
    int k = 0;
    for (int h = cta; h < n; h += ncta, k++) {
      const int set = k & 1;
      const int hn = h + ncta;
      int4 ent_n = make_int4(0, 0, 0, 0);
      if (hn < n) ent_n = __ldcg(&list[hn]);
      const int tile = ent.x, ub = ent.y, U = ent.z;
      const bool do_n = (mode & NB_NORMALS) && (ent.w & DSC_ENT_NORMALS);
      const bool do_b = (mode & NB_BOUNDS) && (ent.w & DSC_ENT_BOUNDS);
      const int ng = (U + 31) >> 5, G0 = ub >> 5;
      const unsigned dw = ((do_n || BRUSH) && lane < ng) ? __ldcg(&m.dirty[G0 + lane]) : 0u;
      int dcount = __popc(dw);
      for (int o = 16; o > 0; o >>= 1) dcount += __shfl_xor_sync(0xffffffffu, dcount, o);
      const unsigned goff0 = __shfl_sync(0xffffffffu, go, 0);
      glast = __shfl_sync(0xffffffffu, glast, 0);
      const int sb = __shfl_sync(0xffffffffu, q.z, 0), SB = __shfl_sync(0xffffffffu, q.w, 0);
      const int X = __shfl_sync(0xffffffffu, q.x, 1), eb = __shfl_sync(0xffffffffu, q.y, 1);
      const int eown = __shfl_sync(0xffffffffu, q.z, 1), ehalo = __shfl_sync(0xffffffffu, q.w, 1);
      const int hb = __shfl_sync(0xffffffffu, q.x, 2), ntfast = __shfl_sync(0xffffffffu, q.w, 2);
      /* with a brush the interior verts may still become dirty: everything the normal phases need is loaded */
      const bool anyd = BRUSH ? do_n : dcount > 0;
      const bool active = (ntfast & (1 << 16)) && (BRUSH || anyd || do_b);
      /* positions / entries / staged verts / descriptor of tile k - 1 are released */
      dsc_mbar_wait(&s_empty[0], par_e0);
      par_e0 ^= 1u;
      if (lane < 3) s_q[set][lane] = q;
      sdirty[set][lane] = dw;
      sgoff[set][lane] = go;
      if (lane == 0) {
        sgoff[set][ng] = glast;
        s_dcount[set] = 0;
        const int ibnd = BRUSH ? ((ent.w >> DSC_ENT_IBND_SHIFT) & 0x7ff) : 0;
        s_ent[set] = make_int4(tile, ub, U | (ibnd << 16),
                               (active ? NT_ACTIVE : 0) | (anyd ? NT_ANYD : 0) | (do_b ? NT_DO_B : 0) | (do_n ? NT_DO_N : 0) |
                                   ((ent.w & DSC_ENT_FIRST) ? NT_FIRST : 0) | (dcount << 8));
      }
      const int ne = eown + ehalo;
      const int UA = (U + 3) & ~3;
      const int v2w = anyd ? (int)(glast - goff0) : 0;
      const int nloc_a = dsc_tile_nloc_a(U, SB, X);
      float *PX = reinterpret_cast<float *>(smem_raw), *PY = PX + nloc_a, *PZ = PY + nloc_a;
      ushort4 *E = reinterpret_cast<ushort4 *>(smem_raw + m.sm_off_e);
      unsigned *V2 = reinterpret_cast<unsigned *>(smem_raw + m.sm_off_v2);
      unsigned char *H = smem_raw + m.sm_off_h;
      if (active) {
        if (lane == 0) {
          const unsigned pbytes = 4u * (unsigned)UA;
          const unsigned ebytes = anyd ? 8u * (unsigned)((ne + 1) & ~1) : 0u;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          if (3u * pbytes + ebytes) dsc_mbar_expect_tx(&s_full[0], 3u * pbytes + ebytes);
          if (pbytes) {
            dsc_bulk_g2s(PX, m.cx + ub, pbytes, &s_full[0]);
            dsc_bulk_g2s(PY, m.cy + ub, pbytes, &s_full[0]);
            dsc_bulk_g2s(PZ, m.cz + ub, pbytes, &s_full[0]);
          }
          if (ebytes) dsc_bulk_g2s(E, m.e_pv + eb, ebytes, &s_full[0]);
        }
        /* staged verts: 4 x 32 at a time, index loads first, then the 12 position loads, then the stores */
        const int nstaged = SB + (anyd ? X : 0);
        for (int i0 = 0; i0 < nstaged; i0 += 128) {
          int sl[4];
          float x[4], y[4], z[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int i = i0 + u * 32 + lane;
            sl[u] = (i < nstaged) ? m.stage_slots[sb + i] : 0;
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            x[u] = __ldcg(&m.cx[sl[u]]); y[u] = __ldcg(&m.cy[sl[u]]); z[u] = __ldcg(&m.cz[sl[u]]);
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int i = i0 + u * 32 + lane;
            if (i < nstaged) {
              PX[UA + i] = x[u]; PY[UA + i] = y[u]; PZ[UA + i] = z[u];
            }
          }
        }
        if (anyd) {
          for (int i = lane; i < ehalo; i += 32) {
            const int ol = m.e_halo_leaf[hb + i];
            const unsigned w = upd_shared ? s_upd[ol >> 5] : __ldcg(&upd[ol >> 5]);
            H[i] = (unsigned char)((w >> (ol & 31)) & 1u);
          }
        }
      }
      dsc_mbar_arrive(&s_full[0]);
      if (active && anyd) {
        /* index words: their buffer is released when phase 3 of the previous tile that had one is done */
        dsc_mbar_wait(&s_empty[1], par_e1);
        par_e1 ^= 1u;
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          dsc_mbar_expect_tx(&s_full[1], 4u * (unsigned)v2w);
          dsc_bulk_g2s(V2, m.v2_idx + goff0, 4u * (unsigned)v2w, &s_full[1]);
          asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(dsc_smem_u32(&s_full[1])) : "memory");
        }
      }
      /* the next tile's constant tables, one tile ahead */
      ent = ent_n;
      if (hn < n) {
        const int ngn = (ent.z + 31) >> 5, g0n = ent.y >> 5;
        q = make_int4(0, 0, 0, 0);
        if (lane < 3) q = m.tile_meta[3 * ent.x + lane];
        go = (lane < ngn) ? m.v2_goff[g0n + lane] : 0u;
        glast = (lane == 0) ? m.v2_goff[g0n + ngn] : 0u;
      }
    }
    return; /* the caller joins the warps (end of the kernel, or a CTA barrier of the batch kernel) */
  }

  /* -------------------------------------------------------------------- consumer warps */
  unsigned par_f0 = 0u, par_f1 = 0u;
  unsigned moved_cnt = 0u;
  float radius_sq = 0.0f;
  bool need_no = false;
  if (BRUSH) {
    radius_sq = dab->radius * dab->radius;
    need_no = (TOOL == 4) || (dab->flags & 1);
  }
  int k = 0;
  for (int h = cta; h < n; h += ncta, k++) {
    const int set = k & 1;
    dsc_mbar_wait(&s_full[0], par_f0);
    par_f0 ^= 1u;
    const int4 ent = s_ent[set];
    const int flags = ent.w;
    if (!(flags & NT_ACTIVE)) {
      __syncwarp();
      if (lane == 0) dsc_mbar_arrive(&s_empty[0]);
      continue;
    }
    const int4 q0 = s_q[set][0], q1 = s_q[set][1], q2 = s_q[set][2];
    const int ub = ent.y, U = ent.z & 0xffff;
    const bool do_b = (flags & NT_DO_B) != 0;
    const int SB = q0.w, X = q1.x, eown = q1.z, ne = q1.z + q1.w;
    const int ng = (U + 31) >> 5, G0 = ub >> 5;
    const int UA = (U + 3) & ~3;
    const unsigned goff0 = sgoff[set][0];
    const int nloc_a = dsc_tile_nloc_a(U, SB, X);
    float *PX = reinterpret_cast<float *>(smem_raw), *PY = PX + nloc_a, *PZ = PY + nloc_a;
    float4 *F = reinterpret_cast<float4 *>(smem_raw + m.sm_off_f); /* [ne + 1], entry ne = zero */
    const ushort4 *E = reinterpret_cast<const ushort4 *>(smem_raw + m.sm_off_e);
    const unsigned *V2 = reinterpret_cast<const unsigned *>(smem_raw + m.sm_off_v2);
    const unsigned char *H = smem_raw + m.sm_off_h;
    bool anyd = (flags & NT_ANYD) != 0;
    int dcount = flags >> 8;
    /* the thread's four unique verts: brush (interior part), box */
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    if (BRUSH || do_b) {
      const int i0 = 4 * tid;
      unsigned nib = 0u;
      if (i0 < U) {
        float4 x4 = *reinterpret_cast<const float4 *>(PX + i0), y4 = *reinterpret_cast<const float4 *>(PY + i0),
               z4 = *reinterpret_cast<const float4 *>(PZ + i0);
        if (BRUSH && i0 < (ent.z >> 16)) {
          nib = dsc_brush_quad_core<TOOL == 0 ? 1 : TOOL>(m, *dab, s_D, (flags & NT_FIRST) != 0, ub + i0, 4, radius_sq, need_no, x4, y4, z4);
          if (nib) {
            *reinterpret_cast<float4 *>(PX + i0) = x4; *reinterpret_cast<float4 *>(PY + i0) = y4; *reinterpret_cast<float4 *>(PZ + i0) = z4;
            st4(m.cx, ub + i0, x4); st4(m.cy, ub + i0, y4); st4(m.cz, ub + i0, z4);
          }
        }
        if (do_b) {
          const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, ys[4] = {y4.x, y4.y, y4.z, y4.w}, zs[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
          for (int j = 0; j < 4; j++) {
            if (i0 + j < U) {
              mn[0] = fminf(mn[0], xs[j]); mx[0] = fmaxf(mx[0], xs[j]);
              mn[1] = fminf(mn[1], ys[j]); mx[1] = fmaxf(mx[1], ys[j]);
              mn[2] = fminf(mn[2], zs[j]); mx[2] = fmaxf(mx[2], zs[j]);
            }
          }
        }
      }
      if (BRUSH) {
        /* eight lanes own one dirty word: interior bits of this dab on top of the boundary bits the producer read */
        const unsigned w = dsc_pack_nibbles(nib, lane);
        int cnt = 0;
        if ((lane & 7) == 0 && (i0 >> 5) < ng) {
          const unsigned fin = sdirty[set][i0 >> 5] | w;
          if (w) {
            sdirty[set][i0 >> 5] = fin;
            moved_cnt += __popc(w);
            if (m.capture) atomicOr(&m.capture[G0 + (i0 >> 5)], w);
            if (!(flags & NT_DO_N)) atomicOr(&m.dirty[G0 + (i0 >> 5)], w); /* no normal pass in this dab: the bits wait in the bitmap */
          }
          cnt = __popc(fin);
        }
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 8);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 16);
        if (lane == 0 && cnt) atomicAdd(&s_dcount[set], cnt);
        asm volatile("bar.sync 1, 256;" ::: "memory"); /* displaced positions, dirty words and their count are in shared memory */
        dcount = s_dcount[set];
        anyd = (flags & NT_DO_N) && dcount > 0;
      }
    }
    if (do_b) {
      /* the box-counted staged verts, then the tile's share of the leaf box */
      for (int i = tid; i < SB; i += NT_CONSUMERS) {
        mn[0] = fminf(mn[0], PX[UA + i]); mx[0] = fmaxf(mx[0], PX[UA + i]);
        mn[1] = fminf(mn[1], PY[UA + i]); mx[1] = fmaxf(mx[1], PY[UA + i]);
        mn[2] = fminf(mn[2], PZ[UA + i]); mx[2] = fmaxf(mx[2], PZ[UA + i]);
      }
#pragma unroll
      for (int c = 0; c < 3; c++) {
        for (int o = 16; o > 0; o >>= 1) {
          mn[c] = fminf(mn[c], __shfl_down_sync(0xffffffffu, mn[c], o));
          mx[c] = fmaxf(mx[c], __shfl_down_sync(0xffffffffu, mx[c], o));
        }
      }
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
          red[c][warp] = mn[c];
          red[3 + c][warp] = mx[c];
        }
      }
    }
    /* phase 2: poly normals of the local entries */
    if (anyd) {
      if (tid == 0) F[ne] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      const bool sparse = dcount * 2 < U;
      if (q2.w & (1 << 17)) dsc_tile_poly_normals<true>(PX, PY, PZ, E, F, H, sdirty[set], tid, U, eown, ne, sparse);
      else dsc_tile_poly_normals<false>(PX, PY, PZ, E, F, H, sdirty[set], tid, U, eown, ne, sparse);
    }
    __syncwarp();
    if (lane == 0) dsc_mbar_arrive(&s_empty[0]); /* positions, entries, staged verts: the producer may refill */
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (do_b && warp == 1 && lane < 6) {
      /* the tile's share of the leaf box (the gather reset the box of every leaf it hit) */
      float v = red[lane][0];
      for (int w = 1; w < NW; w++) v = (lane < 3) ? fminf(v, red[lane][w]) : fmaxf(v, red[lane][w]);
      float *dstp = &m.bb[lane * tn + q2.y];
      if (lane < 3) dsc_red_min(dstp, v);
      else dsc_red_max(dstp, v);
    }
    if (BRUSH ? (flags & NT_ANYD) != 0 : anyd) {
      /* phase 3: one warp per group of 32 verts.  (With a brush the index words were queued for every tile that may
       * update normals: the hand-over runs even when nothing turned out dirty.) */
      dsc_mbar_wait(&s_full[1], par_f1);
      par_f1 ^= 1u;
      if (anyd) {
        for (int g = warp; g < ng; g += NW) {
          const unsigned word = sdirty[set][g];
          if (word == 0u) continue; /* warp-uniform */
          if ((word >> lane) & 1u) {
            const unsigned off = sgoff[set][g];
            const int wd = (int)((sgoff[set][g + 1] - off) >> 5);
            const unsigned *rp = V2 + (off - goff0) + lane;
            float sx = 0.0f, sy = 0.0f, sz = 0.0f;
            if (wd == 3) {
              const unsigned w0 = rp[0], w1 = rp[32], w2 = rp[64];
              float4 f;
              f = F[w0 & 0xffffu]; sx += f.x; sy += f.y; sz += f.z;
              f = F[w0 >> 16]; sx += f.x; sy += f.y; sz += f.z;
              f = F[w1 & 0xffffu]; sx += f.x; sy += f.y; sz += f.z;
              f = F[w1 >> 16]; sx += f.x; sy += f.y; sz += f.z;
              f = F[w2 & 0xffffu]; sx += f.x; sy += f.y; sz += f.z;
              f = F[w2 >> 16]; sx += f.x; sy += f.y; sz += f.z;
            }
            else {
              for (int j = 0; j < wd; j++) {
                const unsigned w0 = rp[j * 32];
                float4 f;
                f = F[w0 & 0xffffu]; sx += f.x; sy += f.y; sz += f.z;
                f = F[w0 >> 16]; sx += f.x; sy += f.y; sz += f.z;
              }
            }
            dsc_normalize(sx, sy, sz);
            const int sl = ub + g * 32 + lane;
            m.nx[sl] = sx; m.ny[sl] = sy; m.nz[sl] = sz;
          }
          if (lane == 0) m.dirty[G0 + g] = 0u;
        }
      }
      __syncwarp();
      if (lane == 0) dsc_mbar_arrive(&s_empty[1]); /* index words: the producer may refill */
    }
    asm volatile("bar.sync 1, 256;" ::: "memory"); /* poly normals / box partials are free for the next tile */
  }
  if (BRUSH) {
    for (int o = 16; o > 0; o >>= 1) moved_cnt += __shfl_down_sync(0xffffffffu, moved_cnt, o);
    if (lane == 0 && moved_cnt) atomicAdd(&s_moved, moved_cnt);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tid == 0 && s_moved) atomicAdd(&m.tot->moved_total, (unsigned long long)s_moved);
  }
}
__global__ void __launch_bounds__(NT_THREADS, 4) k_normals_tile(DevMesh m, const int4 *list, const int *count, int mode,
                                                                const unsigned *upd)
{
  dsc_normals_tile_body<0>(m, list, count, mode, upd, blockIdx.x, gridDim.x, true);
}
/* the fused dab kernel: interior brush + normals + boxes of the gathered tiles of dab j (ring slot `slot`) */
template<int TOOL>
__global__ void __launch_bounds__(NT_THREADS, 4) k_dab_tile(DevMesh m, int j, int slot, int mode)
{
  __shared__ DabParams s_dab;
  if (threadIdx.x < (int)(sizeof(DabParams) / 4)) {
    reinterpret_cast<int *>(&s_dab)[threadIdx.x] = reinterpret_cast<const int *>(&dsc_dab_entry(m, j).d)[threadIdx.x];
  }
  __syncthreads();
  dsc_normals_tile_body<TOOL>(m, m.tile_list + (size_t)slot * m.ntile, &m.st[slot].tile_count, mode, m.ghit + (size_t)slot * m.ghit_words,
                              blockIdx.x, gridDim.x, false, &s_dab, m.st + slot);
}

/* ------------------------------------------------------------------ the whole dab path, persistent */
/* grid-wide barrier of a cooperative launch (all CTAs resident): arrive on a global counter, spin on
 * an acquire load until every CTA of this round has arrived */
__device__ __forceinline__ void dsc_grid_sync(unsigned *bar, unsigned &target, unsigned ncta)
{
  __syncthreads();
  if (threadIdx.x == 0) {
    target += ncta;
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

/* A run of `batch` dabs of one tool in ONE cooperative launch: gather -> (area) -> brush -> normals +
 * boxes, the stages separated by grid barriers instead of kernel boundaries, the bottom-up refit of a
 * dab riding along with the gather of the next one.  The per-dab cost that is left is the barriers
 * (~1 us each) and each stage's own dependent loads -- no launch gaps, no host calls, no events.
 * CTA shape and shared memory are the tile kernel's; area and brush use its 8 compute warps.  Everything
 * a stage reads that an earlier stage of the same launch wrote goes through L2 (__ldcg / ld4). */
template<int TOOL>
__global__ void __launch_bounds__(NT_THREADS, 4) k_dab_batch(DevMesh m, int batch, int slot0, int needs_area, int refit)
{
  const int cta = blockIdx.x, ncta = gridDim.x, tid = threadIdx.x;
  unsigned target = 0u;
  const int gblocks = (m.nleaf + DSC_BLOCK - 1) / DSC_BLOCK;
  for (int j = 0; j < batch; j++) {
    const int slot = (slot0 + j) & (DSC_SLOTS - 1);
    const DabEntry &e = dsc_dab_entry(m, j);
    const DabParams d = e.d;
    const int mode = ((e.ent_bits & DSC_ENT_NORMALS) ? NB_NORMALS : 0) | ((e.ent_bits & DSC_ENT_BOUNDS) ? NB_BOUNDS : 0);
    /* 1. gather + ancestor tags (first CTAs); the previous dab's refit rides along (last CTAs) */
    if (tid < DSC_BLOCK) {
      const float vn[3] = {d.view_n[0], d.view_n[1], d.view_n[2]};
      for (int b = cta; b < gblocks; b += ncta) {
        dsc_gather_body(m, slot, d.loc[0], d.loc[1], d.loc[2], e.radius_sq, e.area_radius_sq, e.original, 1, 1, e.set_flags,
                        e.ent_bits, b, refit ? (j & 1) : -1, d.falloff_shape == 1 ? vn : nullptr);
      }
    }
    if (refit && j > 0) dsc_refit_body(m, (slot0 + j - 1) & (DSC_SLOTS - 1), (j - 1) & 1, (ncta - 1 - cta) * NT_THREADS + tid, ncta * NT_THREADS);
    dsc_grid_sync(m.grid_bar, target, ncta);
    /* 2. area normal / centre */
    if (needs_area) {
      dsc_area_body(m, d, e.use_cos, slot, cta, ncta);
      dsc_grid_sync(m.grid_bar, target, ncta);
    }
    /* 3. brush; the boxes of the gathered leaves are emptied for the tile stage */
    if (mode & NB_BOUNDS) {
      const int *hl = m.hit_list + (size_t)slot * m.nleaf;
      const int nh = __ldcg(&m.st[slot].hit_count);
      for (int i = cta * NT_THREADS + tid; i < nh; i += ncta * NT_THREADS) dsc_reset_leaf_box(m, __ldcg(&hl[i]));
    }
    dsc_brush_body<TOOL, false>(m, d, slot, cta, ncta);
    dsc_grid_sync(m.grid_bar, target, ncta);
    /* 4. + 5. normals and boxes, tile by tile */
    dsc_normals_tile_body<0>(m, m.tile_list + (size_t)slot * m.ntile, &m.st[slot].tile_count, mode,
                          m.ghit + (size_t)slot * m.ghit_words, cta, ncta, false);
    dsc_grid_sync(m.grid_bar, target, ncta);
  }
  if (refit) dsc_refit_body(m, (slot0 + batch - 1) & (DSC_SLOTS - 1), (batch - 1) & 1, cta * NT_THREADS + tid, ncta * NT_THREADS);
}

/* ------------------------------------------------------------------------------ K6 leaf BB */
/* update_node_vb leaf branch (pbvh.c:2033-2041) on its own: min/max over ALL verts of a listed
 * leaf flagged UpdateBB, unique (float4 stream) and shared (gather). */
__global__ void __launch_bounds__(DSC_BLOCK) k_leaf_bb(DevMesh m, const int *list, const int *count, int skip_fast)
{
  __shared__ float red[6][DSC_BLOCK / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tn = m.totnode;
  const int n = *count;
  for (int h = blockIdx.x; h < n; h += gridDim.x) {
    const int l = list[h];
    if (skip_fast && m.leaf_fast[l]) continue;
    if (!(m.node_flag[l] & F_UpdateBB)) continue;
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    const int ub = m.leaf_ubeg[l], uc = m.leaf_ucnt[l];
    for (int i = 4 * tid; i < uc; i += 4 * DSC_BLOCK) {
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
    const int sb = m.leaf_sbeg[l], sc = m.leaf_scnt[l];
    for (int i = tid; i < sc; i += DSC_BLOCK) {
      const int s = m.leaf_sslots[sb + i];
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
      m.bb[tid * tn + l] = v;
    }
  }
}

/* Whole-tree flush by depth in one CTA (session start, vert_coords_apply, stand-alone
 * BKE_pbvh_update_bounds): every inner node = union of its children. */
__global__ void __launch_bounds__(1024) k_flush(DevMesh m)
{
  const int tn = m.totnode;
  for (int lev = m.nlevel - 1; lev >= 0; lev--) {
    const int b = m.level_off[lev], e = m.level_off[lev + 1];
    for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
      const int node = m.level_nodes[i];
      const int c0 = m.child0[node], c1 = m.child1[node];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        m.bb[k * tn + node] = fminf(__ldcg(&m.bb[k * tn + c0]), __ldcg(&m.bb[k * tn + c1]));
        m.bb[(3 + k) * tn + node] = fmaxf(__ldcg(&m.bb[(3 + k) * tn + c0]), __ldcg(&m.bb[(3 + k) * tn + c1]));
      }
    }
    __threadfence_block();
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(&m.tot->refit_total, (unsigned long long)m.level_off[m.nlevel]); /* inner boxes rewritten */
}

/* drops leaf flags once their stage ran (pbvh.c:3007, 3295) */
__global__ void __launch_bounds__(DSC_BLOCK) k_clear_flags(DevMesh m, const int *list, const int *count, int clear_mask)
{
  const int n = *count;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int l = list[i];
    const int f = m.node_flag[l];
    if (f & clear_mask) m.node_flag[l] = f & ~clear_mask;
  }
}

/* PBVH_UpdateOriginalBB flush (pbvh.c:3139-3141, 3298-3314): flagged leaves copy vb to orig_vb and
 * tag their ancestors; the second kernel copies for tagged inner nodes. */
__global__ void __launch_bounds__(DSC_BLOCK) k_orig_leaves(DevMesh m)
{
  const int tn = m.totnode;
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < m.nleaf; l += gridDim.x * blockDim.x) {
    const int f = m.node_flag[l];
    if (!(f & F_UpdateOriginalBB)) continue;
    m.node_flag[l] = f & ~F_UpdateOriginalBB;
    for (int k = 0; k < 6; k++) m.obb[k * tn + l] = m.bb[k * tn + l];
    int p = m.topo[l].x;
    while (p >= 0 && atomicExch(&m.node_mark[p], 1) == 0) p = m.topo[p].x;
  }
}
__global__ void __launch_bounds__(DSC_BLOCK) k_orig_inner(DevMesh m)
{
  const int tn = m.totnode;
  for (int n = m.nleaf + blockIdx.x * blockDim.x + threadIdx.x; n < tn; n += gridDim.x * blockDim.x) {
    if (m.node_mark[n]) {
      m.node_mark[n] = 0;
      for (int k = 0; k < 6; k++) m.obb[k * tn + n] = m.bb[k * tn + n];
    }
  }
}

/* ------------------------------------------------------------------- ray-cast */
/* BKE_pbvh_raycast (pbvh.c:3896-3928) + pbvh_faces_node_raycast (pbvh.c:4041-4100): nearest intersection of a
 * ray with the looptris of the leaves whose box it enters.  The reference walks the leaves in order of entry
 * distance; inside a leaf it walks the looptris in order and keeps a hit only if strictly nearer, so what a leaf
 * contributes is its nearest depth and the FIRST looptri attaining it.  One CTA per entered leaf reduces exactly
 * that (a min over the key depth-bits:position, depths being >= 0) and appends it with the leaf's entry
 * distance; the walk over the few entered leaves (skip a leaf entered behind the best hit so far) is done on the
 * host in the reference's order.  `original`: stroke-start boxes, and for leaves with an undo node the
 * stroke-start coordinates (a vert the leaf shares: its owner's snapshot if the owner has one). */
struct RayParams {
  float o[3], inv_dir[3];
  int sign[3];
  int kx, ky, kz;
  float sx, sy, sz;
  int original;
};
struct RayLeafHit {
  int pos, leaf, poly, pad;
  float tmin, depth;
  int slot[3];
  float co[3][3];
};

__device__ __forceinline__ bool dsc_ray_leaf(const DevMesh &m, const RayParams &r, int l, float &tmin_out)
{
  /* isect_ray_aabb_v3, lib/intern/math_geom.cc:3032-3080 */
  const float *bbs = r.original ? m.obb : m.bb;
  const int tn = m.totnode;
  float bbox[2][3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    bbox[0][k] = bbs[k * tn + l];
    bbox[1][k] = bbs[(3 + k) * tn + l];
  }
  float tmin = (bbox[r.sign[0]][0] - r.o[0]) * r.inv_dir[0];
  float tmax = (bbox[1 - r.sign[0]][0] - r.o[0]) * r.inv_dir[0];
  const float tymin = (bbox[r.sign[1]][1] - r.o[1]) * r.inv_dir[1];
  const float tymax = (bbox[1 - r.sign[1]][1] - r.o[1]) * r.inv_dir[1];
  if ((tmin > tymax) || (tymin > tmax)) return false;
  if (tymin > tmin) tmin = tymin;
  if (tymax < tmax) tmax = tymax;
  const float tzmin = (bbox[r.sign[2]][2] - r.o[2]) * r.inv_dir[2];
  const float tzmax = (bbox[1 - r.sign[2]][2] - r.o[2]) * r.inv_dir[2];
  if ((tmin > tzmax) || (tzmin > tmax)) return false;
  if (tzmin > tmin) tmin = tzmin;
  tmin_out = tmin;
  return true;
}

__device__ __forceinline__ void dsc_ray_vert(const DevMesh &m, const int *slot_leaf, bool use_orig, int s, float co[3])
{
  if (use_orig && (m.leaf_state[slot_leaf[s >> 5]] & DSC_LEAF_TOUCHED)) {
    co[0] = m.ox[s]; co[1] = m.oy[s]; co[2] = m.oz[s];
  }
  else {
    co[0] = m.cx[s]; co[1] = m.cy[s]; co[2] = m.cz[s];
  }
}

/* isect_ray_tri_watertight_v3, lib/intern/math_geom.cc:1782-1857 */
__device__ __forceinline__ bool dsc_ray_tri(const RayParams &r, const float v0[3], const float v1[3], const float v2[3], float &lambda)
{
  const float a[3] = {v0[0] - r.o[0], v0[1] - r.o[1], v0[2] - r.o[2]};
  const float b[3] = {v1[0] - r.o[0], v1[1] - r.o[1], v1[2] - r.o[2]};
  const float c[3] = {v2[0] - r.o[0], v2[1] - r.o[1], v2[2] - r.o[2]};
  const float a_kx = a[r.kx], a_ky = a[r.ky], a_kz = a[r.kz];
  const float b_kx = b[r.kx], b_ky = b[r.ky], b_kz = b[r.kz];
  const float c_kx = c[r.kx], c_ky = c[r.ky], c_kz = c[r.kz];
  const float ax = a_kx - r.sx * a_kz, ay = a_ky - r.sy * a_kz;
  const float bx = b_kx - r.sx * b_kz, by = b_ky - r.sy * b_kz;
  const float cx = c_kx - r.sx * c_kz, cy = c_ky - r.sy * c_kz;
  const float u = cx * by - cy * bx;
  const float v = ax * cy - ay * cx;
  const float w = bx * ay - by * ax;
  if ((u < 0.0f || v < 0.0f || w < 0.0f) && (u > 0.0f || v > 0.0f || w > 0.0f)) return false;
  const float det = u + v + w;
  if (det == 0.0f || !isfinite(det)) return false;
  const unsigned sign_det = __float_as_uint(det) & 0x80000000u;
  const float t = (u * a_kz + v * b_kz + w * c_kz) * r.sz;
  if (__uint_as_float(__float_as_uint(t) ^ sign_det) < 0.0f) return false;
  const float inv_det = 1.0f / det;
  lambda = t * inv_det;
  return true;
}

__global__ void __launch_bounds__(DSC_BLOCK) k_raycast(DevMesh m, const int4 *tri_slots, const int *slot_leaf, RayParams r, int *count,
                                                       RayLeafHit *out)
{
  __shared__ unsigned long long s_min;
  for (int l = blockIdx.x; l < m.nleaf; l += gridDim.x) {
    float tmin;
    if (!dsc_ray_leaf(m, r, l, tmin)) continue; /* CTA-uniform */
    const bool use_orig = r.original && (m.leaf_state[l] & DSC_LEAF_TOUCHED);
    const int pb = m.leaf_pbeg[l], pe = pb + m.leaf_pcnt[l];
    if (threadIdx.x == 0) s_min = ~0ull;
    __syncthreads();
    unsigned long long best = ~0ull;
    for (int pos = pb + threadIdx.x; pos < pe; pos += blockDim.x) {
      const int4 tv = tri_slots[pos];
      float c0[3], c1[3], c2[3], lambda;
      dsc_ray_vert(m, slot_leaf, use_orig, tv.x, c0);
      dsc_ray_vert(m, slot_leaf, use_orig, tv.y, c1);
      dsc_ray_vert(m, slot_leaf, use_orig, tv.z, c2);
      if (!dsc_ray_tri(r, c0, c1, c2, lambda)) continue;
      /* lambda >= 0 or -0 (the sign test above), never NaN past the det test unless t overflowed: +0 canonical */
      const unsigned long long key = ((unsigned long long)__float_as_uint(lambda + 0.0f) << 32) | (unsigned)pos;
      best = min(best, key);
    }
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_down_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(&s_min, best);
    __syncthreads();
    const unsigned long long win = s_min;
    if (win != ~0ull && threadIdx.x == 0) {
      const int pos = (int)(unsigned)(win & 0xffffffffull);
      const int4 tv = tri_slots[pos];
      RayLeafHit &h = out[atomicAdd(count, 1)];
      h.pos = pos; h.leaf = l; h.poly = tv.w; h.pad = 0;
      h.tmin = tmin;
      h.depth = __uint_as_float((unsigned)(win >> 32));
      h.slot[0] = tv.x; h.slot[1] = tv.y; h.slot[2] = tv.z;
      dsc_ray_vert(m, slot_leaf, use_orig, tv.x, h.co[0]);
      dsc_ray_vert(m, slot_leaf, use_orig, tv.y, h.co[1]);
      dsc_ray_vert(m, slot_leaf, use_orig, tv.z, h.co[2]);
    }
    __syncthreads();
  }
}

/* ------------------------------------------------------------------- draw-buffer fill */
/* GPU_pbvh_mesh_buffers_update (gpu/intern/gpu_buffers.c:174-305) for the listed leaves: one 36-byte record
 * per looptri corner in the vertex format of gpu_pbvh_init (gpu_buffers.c:84-100, offsets from
 * VertexFormat_pack, gpu_vertex_format.cc:300-325): pos f32 x 3 @0, nor i16 x 3 @16, msk u8 @22,
 * col u16 x 4 @24 (not shown on this path: zero), fset u8 x 3 @32 (white).  Flat shading: the poly normal
 * and the mean mask of the looptri; smooth: the vertex normal and mask.  The buffer is indexed by looptri
 * position, so a leaf's VBO is one contiguous run. */
__device__ __forceinline__ unsigned dsc_normal_short(float f) { return (unsigned)(unsigned short)(short)(int)(f * 32767.0f); }
__global__ void __launch_bounds__(DSC_BLOCK) k_draw_fill(DevMesh m, const int4 *tri_slots, const int *list, const int *count, int smooth_all,
                                                         const unsigned char *leaf_smooth, int show_mask, unsigned *vbo)
{
  const int n = *count;
  for (int h = blockIdx.x; h < n; h += gridDim.x) {
    const int l = list[h];
    /* per leaf: ME_SMOOTH of the poly of the leaf's first looptri (gpu_buffers.c:221-222) */
    const int smooth = leaf_smooth ? (int)leaf_smooth[l] : smooth_all;
    const int pb = m.leaf_pbeg[l], pe = pb + m.leaf_pcnt[l];
    for (int pos = pb + threadIdx.x; pos < pe; pos += blockDim.x) {
      const int4 tv = tri_slots[pos];
      const int sl[3] = {tv.x, tv.y, tv.z};
      unsigned n01 = 0u, n2 = 0u, cmask = 0u;
      if (!smooth) {
        float fx, fy, fz;
        dsc_poly_normal(m, (unsigned)pos, fx, fy, fz);
        n01 = dsc_normal_short(fx) | (dsc_normal_short(fy) << 16);
        n2 = dsc_normal_short(fz);
        if (show_mask) {
          const float fmask = (m.mask[sl[0]] + m.mask[sl[1]] + m.mask[sl[2]]) / 3.0f;
          cmask = (unsigned)(unsigned char)(int)(fmask * 255);
        }
      }
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int s = sl[j];
        if (smooth) {
          n01 = dsc_normal_short(m.nx[s]) | (dsc_normal_short(m.ny[s]) << 16);
          n2 = dsc_normal_short(m.nz[s]);
          if (show_mask) cmask = (unsigned)(unsigned char)(int)(m.mask[s] * 255);
        }
        unsigned *rec = vbo + ((size_t)pos * 3 + j) * 9;
        rec[0] = __float_as_uint(m.cx[s]);
        rec[1] = __float_as_uint(m.cy[s]);
        rec[2] = __float_as_uint(m.cz[s]);
        rec[3] = 0u;
        rec[4] = n01;
        rec[5] = n2 | (cmask << 16);
        rec[6] = 0u;
        rec[7] = 0u;
        rec[8] = 0x00ffffffu;
      }
    }
  }
}

/* ------------------------------------------------------------------- multi-GPU halo pack / unpack */
/* positions of the halo slots, [3][n] in the buffer */
__global__ void k_halo_pack(float *__restrict__ buf, const int *__restrict__ idx, int n, const float *__restrict__ ax,
                            const float *__restrict__ ay, const float *__restrict__ az)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = idx[i];
    buf[i] = ax[s];
    buf[n + i] = ay[s];
    buf[2 * n + i] = az[s];
  }
}
__global__ void k_halo_unpack(const float *__restrict__ buf, const int *__restrict__ idx, int n, float *__restrict__ ax,
                              float *__restrict__ ay, float *__restrict__ az)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = idx[i];
    ax[s] = buf[i];
    ay[s] = buf[n + i];
    az[s] = buf[2 * n + i];
  }
}

/* ---- multi-GPU exchanges over peer memory (NVLink): every rank maps one small region of every other rank's HBM
 * (cudaIpc) -- flags, a reduce inbox, a halo inbox, the inboxes double buffered -- and the per-dab exchanges are stores
 * into the peers' inboxes instead of NCCL send / recv pairs.  One exchange = a push kernel and a receive kernel on every
 * rank, numbered by a round counter all ranks advance together (it lives in device memory, so the pair replays from a
 * CUDA graph):
 *   push:    gather the halo elements of this rank from its arrays straight into half (round & 1) of the peer's inbox
 *            (coalesced remote stores), fence, raise "round r delivered" at the peer;
 *   receive: wait for "round r delivered" from every peer, scatter that half of the inbox into the arrays (or sum the
 *            reduce inbox), close the round.
 * No "inbox free" handshake is needed: a rank pushes round r only after its own receive of round r - 1, i.e. after every
 * peer delivered round r - 1, which a peer does only after ITS receive of round r - 2 -- the last reader of the half
 * round r overwrites.  Nothing waits for anything but a delivery that the peer's push raises unconditionally, so the
 * ranks cannot deadlock as long as they queue the same sequence of exchanges (they do: the dab sequence is replicated).
 * A dab that gathers no leaf near a partition cut changes no halo element anywhere; every rank sees that in the
 * all-reduced bitmask of gathered leaves, and the halo exchanges of such a dab return at once on all ranks (`cond`). */
#define DSC_MAX_RANKS 8
#define P2P_DONE 16   /* [+ q]: the last round rank q delivered to me */
#define P2P_COUNT 32  /* [+ q]: CTA counter of a push to rank q */
#define P2P_ERR 48    /* a wait gave up (a peer never arrived): the host reports it at stroke end instead of hanging */
#define P2P_RCOUNT 50 /* CTA counter of the receive kernel */
#define P2P_NEAR 51    /* this dab gathered a leaf near a partition cut: its halo exchanges run */
#define P2P_SKIPPED 53 /* exchanges skipped so far (statistics) */
#define P2P_PROUND 56 /* [+ q]: exchanges completed with rank q so far.  Rounds are counted PER PAIR: a dab is exchanged only
                         among the ranks it can reach (DabEntry.peers), so two ranks advance their common counter exactly on
                         the dabs both take part in; the kernels read it here, so a dab's exchanges replay from a graph */
#define P2P_PDIRTY 40 /* [+ q]: a dab near a cut has run with rank q since the smooth brush's halo was last refreshed from it */
struct PeerLink {
  int world, rank;
  int *flags;                      /* mine, see P2P_* */
  int *peer_flags[DSC_MAX_RANKS];
  float *inbox;                    /* mine: halo inbox, the block from rank q at 3 * recv_off[q] */
  float *peer_inbox[DSC_MAX_RANKS];
  long long *red;                  /* mine: reduce inbox, [q * red_stride] */
  long long *peer_red[DSC_MAX_RANKS];
  int red_stride;                  /* in 8-byte words, per source rank */
  int red_half;                    /* 8-byte words of one half of the reduce inbox */
  int inbox_half;                  /* floats of one half of the halo inbox */
  int send_off[DSC_MAX_RANKS + 1], recv_off[DSC_MAX_RANKS + 1];
  int peer_off[DSC_MAX_RANKS];     /* where this rank's block starts in rank q's inbox (q's recv_off[rank]) */
};
__device__ __forceinline__ void dsc_flag_raise(int *p, int v)
{
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int dsc_flag_read(const int *p)
{
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void dsc_flag_wait(const int *p, int round, int *err)
{
  for (long long spin = 0; dsc_flag_read(p) < round; spin++) {
    __nanosleep(100);
    if (spin > (1ll << 26)) { /* tens of seconds */
      *err = 1;
      return;
    }
  }
}
/* the ranks dab j of the running batch is exchanged with (a bit per rank, this rank's own bit excluded) */
__device__ __forceinline__ unsigned dsc_dab_peers(const DevMesh &m, const PeerLink &L, int j)
{
  return (unsigned)dsc_dab_entry(m, j).peers & ~(1u << L.rank);
}
/* cond 0: always; 1: only when this dab is near a cut; 2: only from peers whose halo is stale for the smooth brush */
__device__ __forceinline__ bool dsc_p2p_skip(const PeerLink &L, int cond, int q)
{
  if (cond == 1) return __ldcg(L.flags + P2P_NEAR) == 0;
  if (cond == 2) return __ldcg(L.flags + P2P_PDIRTY + q) == 0;
  return false;
}
/* grid (ctas_per_peer, world): blockIdx.y = peer */
__global__ void __launch_bounds__(256) k_p2p_halo_push(DevMesh m, int j, PeerLink L, int cond, const int *__restrict__ idx,
                                                       const float *__restrict__ ax, const float *__restrict__ ay, const float *__restrict__ az)
{
  const int q = blockIdx.y;
  if (!((dsc_dab_peers(m, L, j) >> q) & 1u) || dsc_p2p_skip(L, cond, q)) return;
  const int round = __ldcg(L.flags + P2P_PROUND + q) + 1;
  const int n = L.send_off[q + 1] - L.send_off[q];
  const int *id = idx + L.send_off[q];
  float *dst = L.peer_inbox[q] + (size_t)(round & 1) * L.inbox_half + 3 * (size_t)L.peer_off[q];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = id[i];
    dst[i] = ax[s];
    dst[n + i] = ay[s];
    dst[2 * n + i] = az[s];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int arrived = atomicAdd(L.flags + P2P_COUNT + q, 1) + 1;
    if (arrived == (int)gridDim.x) {
      L.flags[P2P_COUNT + q] = 0;
      __threadfence_system();
      dsc_flag_raise(L.peer_flags[q] + P2P_DONE + L.rank, round);
    }
  }
}
__global__ void __launch_bounds__(256) k_p2p_halo_recv(DevMesh m, int j, PeerLink L, int cond, const int *__restrict__ idx, float *__restrict__ ax,
                                                       float *__restrict__ ay, float *__restrict__ az)
{
  const int q = blockIdx.y;
  const unsigned peers = dsc_dab_peers(m, L, j);
  const bool live = ((peers >> q) & 1u) && !dsc_p2p_skip(L, cond, q);
  if (live) {
    const int round = __ldcg(L.flags + P2P_PROUND + q) + 1;
    if (threadIdx.x == 0) dsc_flag_wait(L.flags + P2P_DONE + q, round, L.flags + P2P_ERR);
    __syncthreads();
    const int n = L.recv_off[q + 1] - L.recv_off[q];
    const int *id = idx + L.recv_off[q];
    const float *src = L.inbox + (size_t)(round & 1) * L.inbox_half + 3 * (size_t)L.recv_off[q];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const int s = id[i];
      ax[s] = __ldcg(&src[i]);
      ay[s] = __ldcg(&src[n + i]);
      az[s] = __ldcg(&src[2 * n + i]);
    }
  }
  /* the last CTA to leave closes the rounds (every CTA has read the counters by then) */
  __syncthreads();
  if (threadIdx.x == 0) {
    const int total = (int)(gridDim.x * gridDim.y);
    if (atomicAdd(L.flags + P2P_RCOUNT, 1) + 1 == total) {
      L.flags[P2P_RCOUNT] = 0;
      int skipped = 0;
      for (int r = 0; r < L.world; r++) {
        if (!((peers >> r) & 1u)) continue;
        if (dsc_p2p_skip(L, cond, r)) {
          skipped = 1;
          continue;
        }
        L.flags[P2P_PROUND + r] += 1;
        if (cond == 2) L.flags[P2P_PDIRTY + r] = 0;
      }
      if (skipped) L.flags[P2P_SKIPPED] += 1;
    }
  }
}
/* the per-dab reduce among the ranks the dab can reach: the 16 exact area sums (int64) and the bitmask of gathered
 * leaves.  grid (1, world) */
__global__ void __launch_bounds__(256) k_p2p_reduce_push(DevMesh m, int j, PeerLink L, const long long *__restrict__ acc,
                                                         const unsigned *__restrict__ ghit, int words, int with_area)
{
  const int q = blockIdx.y;
  if (!((dsc_dab_peers(m, L, j) >> q) & 1u)) return;
  const int round = __ldcg(L.flags + P2P_PROUND + q) + 1;
  long long *dst = L.peer_red[q] + (size_t)(round & 1) * L.red_half + (size_t)L.rank * L.red_stride;
  if (with_area && threadIdx.x < 16) dst[threadIdx.x] = acc[threadIdx.x];
  unsigned *dw = reinterpret_cast<unsigned *>(dst + 16);
  for (int w = threadIdx.x; w < words; w += blockDim.x) dw[w] = ghit[w];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) dsc_flag_raise(L.peer_flags[q] + P2P_DONE + L.rank, round);
}
__global__ void __launch_bounds__(256) k_p2p_reduce_recv(DevMesh m, int j, PeerLink L, long long *__restrict__ acc, unsigned *__restrict__ ghit,
                                                         int words, int with_area, const unsigned *__restrict__ near_mask)
{
  const unsigned peers = dsc_dab_peers(m, L, j);
  if (threadIdx.x < L.world && ((peers >> threadIdx.x) & 1u)) {
    dsc_flag_wait(L.flags + P2P_DONE + threadIdx.x, __ldcg(L.flags + P2P_PROUND + threadIdx.x) + 1, L.flags + P2P_ERR);
  }
  __syncthreads();
  int near = 0;
  if (with_area && threadIdx.x < 16) {
    long long sum = acc[threadIdx.x];
    for (int r = 0; r < L.world; r++) {
      if ((peers >> r) & 1u) {
        const long long *red = L.red + (size_t)((__ldcg(L.flags + P2P_PROUND + r) + 1) & 1) * L.red_half;
        sum += __ldcg(&red[(size_t)r * L.red_stride + threadIdx.x]);
      }
    }
    acc[threadIdx.x] = sum;
  }
  for (int w = threadIdx.x; w < words; w += blockDim.x) {
    unsigned bits = ghit[w];
    for (int r = 0; r < L.world; r++) {
      if ((peers >> r) & 1u) {
        const long long *red = L.red + (size_t)((__ldcg(L.flags + P2P_PROUND + r) + 1) & 1) * L.red_half;
        bits |= __ldcg(reinterpret_cast<const unsigned *>(red + (size_t)r * L.red_stride + 16) + w);
      }
    }
    ghit[w] = bits;
    if (near_mask && (bits & near_mask[w])) near = 1;
  }
  near = __syncthreads_or(near);
  if (threadIdx.x == 0) {
    const int is_near = near_mask ? near : 1;
    L.flags[P2P_NEAR] = is_near;
    for (int r = 0; r < L.world; r++) {
      if ((peers >> r) & 1u) {
        L.flags[P2P_PROUND + r] += 1;
        if (is_near) L.flags[P2P_PDIRTY + r] = 1;
      }
    }
  }
}

/* ------------------------------------------------------------------------------ misc */
__global__ void k_mark_all(DevMesh m, int flags, int all_dirty, int nwords)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m.nleaf) m.node_flag[i] |= flags;
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
/* slot order -> MVert records {co[3], tail} in original vertex order */
__global__ void k_export_mvert(float4 *__restrict__ out, const float *__restrict__ ax, const float *__restrict__ ay,
                               const float *__restrict__ az, const int *__restrict__ slot_of,
                               const unsigned *__restrict__ tail, int totvert)
{
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < totvert; v += gridDim.x * blockDim.x) {
    const int s = slot_of[v];
    out[v] = make_float4(ax[s], ay[s], az[s], __uint_as_float(tail ? tail[v] : 0u));
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
