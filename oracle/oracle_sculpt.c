/* ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * DAGGER: the editor-side brush code (sculpt.c & co.) is absent from /root/reference
 * (SURVEY.md section 0 fact 3).  Everything here restates the upstream behaviour listed in
 * SURVEY.md section 8a rows a7, a9, a11-a20 and is the specification by decision.  Data fields and
 * enums cited: types/types_brush.h:138-362, types/types_brush_enums.h, kernel/intern/brush.h:87-91
 * (KERNEL_brush_curve_strength), kernel/intern/colortools.c:942-965 (curve LUT).
 *
 * Two decisions that differ from a float-for-float transcription, both forced by determinism:
 *  1. area normal / centre sums are accumulated in 2^-32 fixed point (int64), so the result does
 *     not depend on summation order (the reference reduces per-thread partials in undefined order);
 *  2. the smooth brush is Jacobi per iteration (the reference updates in place, thread-schedule
 *     dependent).
 */
#include "oracle_intern.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline float clamp_f(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline float min_ff(float a, float b) { return (a < b) ? a : b; }
static inline float max_ff(float a, float b) { return (a > b) ? a : b; }
static inline float dot_v3v3(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline int64_t fix32(float q) { return (int64_t)llrintf(q * 4294967296.0f); }

static float normalize_v3(float n[3])
{
  float d = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
  if (d > 1.0e-35f) {
    d = sqrtf(d);
    const float f = 1.0f / d;
    n[0] = n[0] * f; n[1] = n[1] * f; n[2] = n[2] * f;
  }
  else {
    n[0] = n[1] = n[2] = 0.0f;
    d = 0.0f;
  }
  return d;
}

void or_set_custom_curve(OrPbvh *p, const float *table257)
{
  memcpy(p->curve_table, table257, sizeof(float) * 257);
  p->has_curve_table = 1;
}

/* kernel/intern/colortools.c:942-965 curvemap_evaluateF on a 257-entry table over [0,1]:
 * fi = value * 256; i = (int)fi; lerp(table[i], table[i+1], fi - i); clamped at the ends. */
static float curve_table_eval(const OrPbvh *p, float value)
{
  const float fi = value * 256.0f;
  const int i = (int)fi;
  if (fi < 0.0f || i < 0) {
    return p->curve_table[0];
  }
  if (i >= 256) {
    return p->curve_table[256];
  }
  const float t = fi - (float)i;
  return (1.0f - t) * p->curve_table[i] + t * p->curve_table[i + 1];
}

/* DAGGER kernel/intern/brush.h:87-91 KERNEL_brush_curve_strength(br, p, len), row a12 */
float or_brush_curve_strength(const OrPbvh *pb, int preset, float p, float len)
{
  float strength = 1.0f;
  if (p >= len) {
    return 0.0f;
  }
  p = p / len;
  p = 1.0f - p;
  switch (preset) {
    case OR_CURVE_CUSTOM:
      strength = pb->has_curve_table ? curve_table_eval(pb, 1.0f - p) : p;
      break;
    case OR_CURVE_SHARP:
      strength = p * p;
      break;
    case OR_CURVE_SMOOTH:
      strength = 3.0f * p * p - 2.0f * p * p * p;
      break;
    case OR_CURVE_SMOOTHER:
      strength = (p * p * p) * (p * (p * 6.0f - 15.0f) + 10.0f);
      break;
    case OR_CURVE_ROOT:
      strength = sqrtf(p);
      break;
    case OR_CURVE_LIN:
      strength = p;
      break;
    case OR_CURVE_CONSTANT:
      strength = 1.0f;
      break;
    case OR_CURVE_SPHERE:
      strength = sqrtf(2.0f * p - p * p);
      break;
    case OR_CURVE_POW4:
      strength = p * p * p * p;
      break;
    case OR_CURVE_INVSQUARE:
      strength = p * (2.0f - p);
      break;
  }
  return strength;
}

/* DAGGER strength factor, rows a12-a13: hardness remap, falloff, front-face, mask, automask */
static float strength_factor(const OrPbvh *p, const OrDab *d, float len, const float vno[3], float mask, int v)
{
  float avg = 1.0f;
  float final_len = len;
  const float hardness = d->hardness;
  float q = len / d->radius;
  if (q < hardness) {
    final_len = 0.0f;
  }
  else if (hardness == 1.0f) {
    final_len = d->radius;
  }
  else {
    q = (q - hardness) / (1.0f - hardness);
    final_len = q * d->radius;
  }
  avg *= or_brush_curve_strength(p, d->curve_preset, final_len, d->radius);
  if (d->flags & OR_DAB_FRONTFACE) {
    const float dot = dot_v3v3(vno, d->view_normal);
    avg *= (dot > 0.0f) ? dot : 0.0f;
  }
  avg *= 1.0f - mask;
  if (p->automask) {
    avg *= p->automask[v];
  }
  return avg;
}

/* DAGGER row a10: the vertex iterator with PBVH_ITER_UNIQUE skips hidden vertices (MVert.flag & ME_HIDE) and hidden grid
 * elements (grid_hidden), pbvh.c:4840-4897 + the iterator macro */
static inline int vert_hidden(const OrPbvh *p, int v)
{
  if (p->vert_flag) return (p->vert_flag[v] & 16) != 0;
  if (p->grid_hidden) return p->grid_hidden[v] != 0;
  return 0;
}

/* DAGGER row a11 brush test: squared distance of co to the brush location (sphere) or to the view line through it (tube:
 * co is projected onto the plane through the location whose normal is the view normal) */
static inline float brush_test_distsq(const OrDab *d, const float co[3])
{
  if (d->falloff_shape == 1) {
    const float plane_d = -dot_v3v3(d->view_normal, d->location);
    const float side = dot_v3v3(d->view_normal, co) + plane_d;
    float q[3];
    for (int k = 0; k < 3; k++) {
      const float proj = co[k] + d->view_normal[k] * (-side);
      q[k] = proj - d->location[k];
    }
    return q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
  }
  const float dx = co[0] - d->location[0], dy = co[1] - d->location[1], dz = co[2] - d->location[2];
  return dx * dx + dy * dy + dz * dz;
}

/* DAGGER row a11 clipping: a locked axis keeps its coordinate; with mirror clipping a vertex within the tolerance of the
 * mirror plane is held on it */
static inline void sculpt_clip(const OrDab *d, float co[3], const float val[3])
{
  for (int i = 0; i < 3; i++) {
    if (d->clip_flags & (8 << i)) continue; /* SCULPT_LOCK_X << i */
    if ((d->clip_flags & (1 << i)) && (fabsf(co[i]) <= d->clip_tolerance[i])) co[i] = 0.0f; /* CLIP_X << i */
    else co[i] = val[i];
  }
}

void or_stroke_begin(OrPbvh *p, const float *automask)
{
  free(p->automask);
  p->automask = NULL;
  if (automask) {
    p->automask = malloc(sizeof(float) * (size_t)p->totvert);
    memcpy(p->automask, automask, sizeof(float) * (size_t)p->totvert);
  }
  memset(p->touched, 0, (size_t)p->totnode);
  p->vertex_dabs = 0;
}

/* stroke end: flush the original bounding boxes (PBVH_UpdateOriginalBB) */
void or_stroke_end(OrPbvh *p)
{
  or_update_bounds(p, OR_PBVH_UpdateOriginalBB);
}

/* DAGGER undo snapshot, row a9: first touch of a node copies co / no of all its verts */
static void undo_push_node(OrPbvh *p, int ni)
{
  if (p->touched[ni]) {
    return;
  }
  p->touched[ni] = 1;
  const OrNode *node = &p->nodes[ni];
  /* Only the unique verts are written to the global snapshot: a shared vert is snapshotted by its
   * owner, which is touched no later than any dab that moves the vert. */
  for (int i = 0; i < node->uniq_verts; i++) {
    const int v = node->vert_indices[i];
    memcpy(p->orig_co[v], p->co[v], sizeof(float[3]));
    memcpy(p->orig_no[v], p->no[v], sizeof(float[3]));
  }
}

/* DAGGER area normal / centre, row a15.  Buckets by sign of dot(view_normal, no). */
typedef struct AreaAcc {
  int64_t nos[2][3], cos[2][3];
  int64_t count_no[2], count_co[2];
} AreaAcc;

static void calc_area_normal_and_center(OrPbvh *p, const OrDab *d, const int *nodes, int totnode,
                                        int use_nos, int use_cos, float r_no[3], float r_co[3])
{
  const int use_orig = (d->tool == OR_TOOL_GRAB); /* grab samples the stroke-start surface */
  AreaAcc sum;
  memset(&sum, 0, sizeof(sum));
  float test_radius = sqrtf(d->radius * d->radius);
  test_radius *= d->normal_radius_factor;
  const float radius_sq = test_radius * test_radius;

  const int par = (or_threads > 1 && totnode > 1);
  int64_t *A = &sum.nos[0][0]; /* 16 contiguous int64: integer sums are exact in any order */
#pragma omp parallel for schedule(dynamic) reduction(+ : A[:16]) if (par)
  for (int n = 0; n < totnode; n++) {
    const OrNode *node = &p->nodes[nodes[n]];
    AreaAcc acc;
    memset(&acc, 0, sizeof(acc));
    for (int i = 0; i < node->uniq_verts; i++) {
      const int v = node->vert_indices[i];
      if (vert_hidden(p, v)) continue;
      const float *co = use_orig ? p->orig_co[v] : p->co[v];
      const float dx = co[0] - d->location[0], dy = co[1] - d->location[1], dz = co[2] - d->location[2];
      const float distsq = brush_test_distsq(d, co);
      if (distsq > radius_sq) {
        continue;
      }
      const float *no = use_orig ? p->orig_no[v] : p->no[v];
      const int flip = (dot_v3v3(d->view_normal, no) <= 0.0f);
      const float q = 1.0f - (sqrtf(distsq) / test_radius);
      const float f = clamp_f(3.0f * q * q - 2.0f * q * q * q, 0.0f, 1.0f);
      if (use_cos) {
        /* co weighted towards the centre: location + disp * (1 - f); accumulated relative to the
         * location in units of the test radius so the fixed-point range is [-1, 1] */
        const float w = 1.0f - f;
        acc.cos[flip][0] += fix32((dx * w) / test_radius);
        acc.cos[flip][1] += fix32((dy * w) / test_radius);
        acc.cos[flip][2] += fix32((dz * w) / test_radius);
        acc.count_co[flip] += 1;
      }
      if (use_nos) {
        acc.nos[flip][0] += fix32(no[0] * f);
        acc.nos[flip][1] += fix32(no[1] * f);
        acc.nos[flip][2] += fix32(no[2] * f);
        acc.count_no[flip] += 1;
      }
    }
    const int64_t *L = &acc.nos[0][0];
    for (int k = 0; k < 16; k++) {
      A[k] += L[k];
    }
  }

  if (use_nos) {
    r_no[0] = r_no[1] = r_no[2] = 0.0f;
    for (int i = 0; i < 2; i++) {
      float t[3];
      for (int k = 0; k < 3; k++) {
        t[k] = (float)((double)sum.nos[i][k] * (1.0 / 4294967296.0));
      }
      if (normalize_v3(t) != 0.0f) {
        memcpy(r_no, t, sizeof(t));
        break;
      }
    }
  }
  if (use_cos) {
    int i;
    for (i = 0; i < 2; i++) {
      if (sum.count_co[i] == 0) {
        continue;
      }
      for (int k = 0; k < 3; k++) {
        const double mean = (double)sum.cos[i][k] / ((double)sum.count_co[i] * 4294967296.0);
        r_co[k] = (float)((double)d->location[k] + (double)test_radius * mean);
      }
      break;
    }
    if (i == 2) {
      memcpy(r_co, d->location, sizeof(float[3])); /* no vertex sampled: brush location */
    }
  }
}

static void sculpt_normal(OrPbvh *p, const OrDab *d, const int *nodes, int totnode, float r_no[3])
{
  switch (d->sculpt_plane) {
    case OR_DIR_VIEW:
      memcpy(r_no, d->view_normal, sizeof(float[3]));
      break;
    case OR_DIR_X:
      r_no[0] = 1.0f; r_no[1] = 0.0f; r_no[2] = 0.0f;
      break;
    case OR_DIR_Y:
      r_no[0] = 0.0f; r_no[1] = 1.0f; r_no[2] = 0.0f;
      break;
    case OR_DIR_Z:
      r_no[0] = 0.0f; r_no[1] = 0.0f; r_no[2] = 1.0f;
      break;
    default:
      calc_area_normal_and_center(p, d, nodes, totnode, 1, 0, r_no, NULL);
      break;
  }
}

/* pbvh.c:3729-3733 BKE_pbvh_vert_mark_update, plus the per-dab moved list the parity tests read
 * (single-thread mode only; the timed OpenMP mode only sets the bit) */
static inline void mark_moved(OrPbvh *p, int v, int par)
{
  or_vert_mark_update(p, v);
  if (!par && p->moved_stamp[v] != p->dab_serial) {
    p->moved_stamp[v] = p->dab_serial;
    p->last_moved[p->last_totmoved++] = v;
  }
}

/* DAGGER rows a16 (draw), a17 (inflate), a19 (grab): sphere test on co (grab: orig_co), fade, displace */
static void do_simple_brush(OrPbvh *p, const OrDab *d, const int *nodes, int totnode)
{
  const float radius_sq = d->radius * d->radius;
  float offset[3] = {0, 0, 0};
  if (d->tool == OR_TOOL_DRAW) {
    float eff[3];
    sculpt_normal(p, d, nodes, totnode, eff);
    memcpy(p->last_area_no, eff, sizeof(eff));
    for (int k = 0; k < 3; k++) {
      offset[k] = eff[k] * d->radius;
      offset[k] = offset[k] * d->scale[k];
      offset[k] = offset[k] * d->bstrength;
    }
  }
  float grab_delta[3] = {d->grab_delta[0], d->grab_delta[1], d->grab_delta[2]};
  if (d->tool == OR_TOOL_GRAB && d->normal_weight > 0.0f) {
    /* DAGGER row a19 sculpt_project_v3_normal_align: the drag is blended towards the sculpt normal (the area normal of the
     * stroke-start surface under the brush), scaled so that it still follows the cursor */
    float sn[3];
    sculpt_normal(p, d, nodes, totnode, sn);
    memcpy(p->last_area_no, sn, sizeof(sn));
    const float len_signed = dot_v3v3(sn, grab_delta);
    const float fac = dot_v3v3(sn, d->view_normal);
    float va[3];
    for (int k = 0; k < 3; k++) va[k] = sn[k] - d->view_normal[k] * fac; /* project_plane_v3_v3v3 */
    float len_view_scale = fabsf(dot_v3v3(va, sn));
    len_view_scale = (len_view_scale > FLT_EPSILON) ? 1.0f / len_view_scale : 1.0f;
    const float w = (len_signed * d->normal_weight) * len_view_scale;
    for (int k = 0; k < 3; k++) {
      grab_delta[k] = grab_delta[k] * (1.0f - d->normal_weight);
      grab_delta[k] = grab_delta[k] + sn[k] * w;
    }
  }
  const int par = (or_threads > 1 && totnode > 1);
#pragma omp parallel for schedule(dynamic) if (par)
  for (int n = 0; n < totnode; n++) {
    const OrNode *node = &p->nodes[nodes[n]];
    for (int i = 0; i < node->uniq_verts; i++) {
      const int v = node->vert_indices[i];
      if (vert_hidden(p, v)) continue;
      const float *tco = (d->tool == OR_TOOL_GRAB) ? p->orig_co[v] : p->co[v];
      const float *tno = (d->tool == OR_TOOL_GRAB) ? p->orig_no[v] : p->no[v];
      const float distsq = brush_test_distsq(d, tco);
      if (distsq > radius_sq) {
        continue;
      }
      const float mask = p->mask ? p->mask[v] : 0.0f;
      float fade = strength_factor(p, d, sqrtf(distsq), tno, mask, v);
      float proxy[3], val[3];
      if (d->tool == OR_TOOL_DRAW) {
        for (int k = 0; k < 3; k++) {
          proxy[k] = offset[k] * fade;
          val[k] = p->co[v][k] + proxy[k];
        }
      }
      else if (d->tool == OR_TOOL_INFLATE) {
        fade = d->bstrength * fade;
        const float s = fade * d->radius;
        for (int k = 0; k < 3; k++) {
          const float nv = p->no[v][k] * s;
          proxy[k] = nv * d->scale[k];
          val[k] = p->co[v][k] + proxy[k];
        }
      }
      else { /* grab: co = orig_co + grab_delta * fade */
        fade = d->bstrength * fade;
        for (int k = 0; k < 3; k++) {
          proxy[k] = grab_delta[k] * fade;
          val[k] = p->orig_co[v][k] + proxy[k];
        }
      }
      if (d->clip_flags) sculpt_clip(d, p->co[v], val);
      else memcpy(p->co[v], val, sizeof(val));
      mark_moved(p, v, par);
    }
  }
}

/* DAGGER row a18 clay strips */
static void do_clay_strips_brush(OrPbvh *p, const OrDab *d, const int *nodes, int totnode)
{
  const int flip = (d->bstrength < 0.0f);
  const float radius = flip ? -d->radius : d->radius;
  const float displace = radius * (0.18f + d->plane_offset);
  const float bstrength = flip ? -d->bstrength : d->bstrength;

  float area_no_sp[3], area_no[3], area_co[3];
  if (d->sculpt_plane == OR_DIR_AREA) {
    calc_area_normal_and_center(p, d, nodes, totnode, 1, 1, area_no_sp, area_co);
    memcpy(area_no, area_no_sp, sizeof(area_no));
  }
  else {
    sculpt_normal(p, d, nodes, totnode, area_no_sp);
    calc_area_normal_and_center(p, d, nodes, totnode, 1, 1, area_no, area_co);
  }
  memcpy(p->last_area_no, area_no_sp, sizeof(area_no));
  memcpy(p->last_area_co, area_co, sizeof(area_co));

  /* delay the first dab: the stroke direction (grab_delta) is not known yet */
  if (d->flags & OR_DAB_FIRST_STEP) {
    return;
  }
  if (d->grab_delta[0] == 0.0f && d->grab_delta[1] == 0.0f && d->grab_delta[2] == 0.0f) {
    return;
  }

  for (int k = 0; k < 3; k++) {
    const float t = (area_no_sp[k] * d->scale[k]) * displace;
    area_co[k] = area_co[k] + t;
  }
  float origin[3];
  for (int k = 0; k < 3; k++) {
    origin[k] = area_co[k] + area_no[k] * (-radius * 0.7f);
  }
  /* brush-local frame: x = area_no x stroke dir, y = area_no x x, z = area_no, each normalised;
   * scaled by radius (z by 1.25 radius) -- local = dot(co - origin, axis) / scale */
  float ax[3][3];
  ax[0][0] = area_no[1] * d->grab_delta[2] - area_no[2] * d->grab_delta[1];
  ax[0][1] = area_no[2] * d->grab_delta[0] - area_no[0] * d->grab_delta[2];
  ax[0][2] = area_no[0] * d->grab_delta[1] - area_no[1] * d->grab_delta[0];
  ax[1][0] = area_no[1] * ax[0][2] - area_no[2] * ax[0][1];
  ax[1][1] = area_no[2] * ax[0][0] - area_no[0] * ax[0][2];
  ax[1][2] = area_no[0] * ax[0][1] - area_no[1] * ax[0][0];
  memcpy(ax[2], area_no, sizeof(float[3]));
  normalize_v3(ax[0]);
  normalize_v3(ax[1]);
  normalize_v3(ax[2]);
  const float sc[3] = {d->radius, d->radius, d->radius * 1.25f};
  /* plane through area_co with normal area_no_sp */
  const float plane_d = -dot_v3v3(area_no_sp, area_co);
  const float roundness = d->tip_roundness;
  const float side = 1.0f;
  const float hardness = 1.0f - roundness;
  const float constant_side = hardness * side;
  const float falloff_side = roundness * side;
  const float trim_sq = (d->radius * d->radius) * (d->plane_trim * d->plane_trim);

  const int par = (or_threads > 1 && totnode > 1);
#pragma omp parallel for schedule(dynamic) if (par)
  for (int n = 0; n < totnode; n++) {
    const OrNode *node = &p->nodes[nodes[n]];
    for (int i = 0; i < node->uniq_verts; i++) {
      const int v = node->vert_indices[i];
      if (vert_hidden(p, v)) continue;
      float *co = p->co[v];
      float rel[3] = {co[0] - origin[0], co[1] - origin[1], co[2] - origin[2]};
      float local[3];
      for (int k = 0; k < 3; k++) {
        local[k] = fabsf(dot_v3v3(rel, ax[k]) / sc[k]);
      }
      if (!(local[0] <= side && local[1] <= side && local[2] <= side)) {
        continue;
      }
      float dist;
      if (min_ff(local[0], local[1]) > constant_side) {
        const float ex = local[0] - constant_side, ey = local[1] - constant_side;
        dist = sqrtf(ex * ex + ey * ey) / falloff_side;
      }
      else if (max_ff(local[0], local[1]) > constant_side) {
        dist = (max_ff(local[0], local[1]) - constant_side) / falloff_side;
      }
      else {
        dist = 0.0f;
      }
      /* plane side: only verts below the plane (above when flipped) move */
      float side_d = dot_v3v3(co, area_no_sp) + plane_d;
      if (flip) {
        side_d = -side_d;
      }
      if (!(side_d <= 0.0f)) {
        continue;
      }
      /* closest point on the plane */
      const float pd = dot_v3v3(area_no_sp, co) + plane_d;
      float val[3];
      for (int k = 0; k < 3; k++) {
        const float intr = co[k] + area_no_sp[k] * (-pd);
        val[k] = intr - co[k];
      }
      if ((d->flags & OR_DAB_PLANE_TRIM) && !(dot_v3v3(val, val) <= trim_sq)) {
        continue;
      }
      const float mask = p->mask ? p->mask[v] : 0.0f;
      const float fade = bstrength * strength_factor(p, d, d->radius * dist, p->no[v], mask, v);
      float nv[3];
      for (int k = 0; k < 3; k++) {
        const float proxy = val[k] * fade;
        nv[k] = co[k] + proxy;
      }
      if (d->clip_flags) sculpt_clip(d, co, nv);
      else memcpy(co, nv, sizeof(nv));
      mark_moved(p, v, par);
    }
  }
}

/* DAGGER row a20 neighbour average: interior verts average all neighbours, boundary verts only
 * boundary neighbours, boundary verts with <= 2 neighbours stay */
static void neighbor_average(const OrPbvh *p, const float (*src)[3], int v, float result[3])
{
  float avg[3] = {0.0f, 0.0f, 0.0f};
  int total = 0, neighbor_count = 0;
  /* grids: the neighbour iterator asks KERNEL_subdiv_ccg_neighbor_coords_get (subdiv_ccg.c:1882-1909, no
   * duplicates) and the boundary test goes through the coarse mesh (subdiv_ccg.c:1972-2008) */
  int gnb[64], *nb_heap = NULL;
  const int *nb;
  int nb_count, is_boundary;
  if (p->is_grids) {
    int *dst = gnb;
    if (or_grids_max_neighbors(p) > 64) dst = nb_heap = malloc(sizeof(int) * (size_t)or_grids_max_neighbors(p));
    nb_count = or_grids_neighbors(p, v, dst);
    nb = dst;
    is_boundary = or_grids_is_boundary(p, v);
  }
  else {
    nb = p->nb_idx + p->nb_off[v];
    nb_count = p->nb_off[v + 1] - p->nb_off[v];
    is_boundary = p->boundary[v];
  }
  for (int q = 0; q < nb_count; q++) {
    const int u = nb[q];
    neighbor_count++;
    if (is_boundary) {
      if (p->is_grids ? or_grids_is_boundary(p, u) : p->boundary[u]) {
        avg[0] += src[u][0]; avg[1] += src[u][1]; avg[2] += src[u][2];
        total++;
      }
    }
    else {
      avg[0] += src[u][0]; avg[1] += src[u][1]; avg[2] += src[u][2];
      total++;
    }
  }
  free(nb_heap);
  if ((neighbor_count <= 2 && is_boundary) || total == 0) {
    memcpy(result, src[v], sizeof(float[3]));
    return;
  }
  const float f = 1.0f / (float)total;
  result[0] = avg[0] * f; result[1] = avg[1] * f; result[2] = avg[2] * f;
}

/* DAGGER row a20 smooth: count = (int)(bstrength*4) full iterations + one partial of strength
 * `last`; Jacobi per iteration (see header). */
static void do_smooth_brush(OrPbvh *p, const OrDab *d, const int *nodes, int totnode)
{
  const int max_iterations = 4;
  const float fract = 1.0f / (float)max_iterations;
  const float radius_sq = d->radius * d->radius;
  float bstrength = clamp_f(d->bstrength, 0.0f, 1.0f);
  const int count = (int)(bstrength * (float)max_iterations);
  const float last = (float)max_iterations * (bstrength - (float)count * fract);

  const int par = (or_threads > 1 && totnode > 1);
  for (int iteration = 0; iteration <= count; iteration++) {
    const float strength = clamp_f((iteration != count) ? 1.0f : last, 0.0f, 1.0f);
#pragma omp parallel for schedule(dynamic) if (par)
    for (int n = 0; n < totnode; n++) {
      const OrNode *node = &p->nodes[nodes[n]];
      for (int i = 0; i < node->uniq_verts; i++) {
        const int v = node->vert_indices[i];
        if (vert_hidden(p, v)) continue;
        const float *co = p->co[v];
        const float distsq = brush_test_distsq(d, co);
        if (distsq > radius_sq) {
          continue;
        }
        const float mask = p->mask ? p->mask[v] : 0.0f;
        const float fade = strength * strength_factor(p, d, sqrtf(distsq), p->no[v], mask, v);
        float avg[3];
        neighbor_average(p, (const float(*)[3])p->co, v, avg);
        float nv[3];
        for (int k = 0; k < 3; k++) {
          const float val = avg[k] - co[k];
          nv[k] = co[k] + val * fade;
        }
        memcpy(p->scratch[v], co, sizeof(float[3]));
        if (d->clip_flags) sculpt_clip(d, p->scratch[v], nv);
        else memcpy(p->scratch[v], nv, sizeof(nv));
        p->iter_flag[v] = 1;
      }
    }
#pragma omp parallel for schedule(dynamic) if (par)
    for (int n = 0; n < totnode; n++) {
      const OrNode *node = &p->nodes[nodes[n]];
      for (int i = 0; i < node->uniq_verts; i++) {
        const int v = node->vert_indices[i];
        if (p->iter_flag[v]) {
          p->iter_flag[v] = 0;
          memcpy(p->co[v], p->scratch[v], sizeof(float[3]));
          mark_moved(p, v, par);
        }
      }
    }
  }
}

/* One dab = gather, undo push + mark, brush, normals, bounds */
int or_dab(OrPbvh *p, const OrDab *d)
{
  const int use_original = (d->tool == OR_TOOL_GRAB);
  const float rs = d->radius * d->radius_scale;
  p->last_totmoved = 0;
  p->dab_serial++;
  p->last_area_no[0] = p->last_area_no[1] = p->last_area_no[2] = 0.0f;
  memcpy(p->last_area_co, d->location, sizeof(float[3]));
  if (d->falloff_shape == 1) p->last_tothit = or_gather_tube(p, d->location, d->view_normal, rs * rs, use_original, 1, p->last_hits);
  else p->last_tothit = or_gather_sphere(p, d->location, rs * rs, use_original, 1, p->last_hits);
  const int *nodes = p->last_hits;
  const int totnode = p->last_tothit;
  for (int n = 0; n < totnode; n++) {
    undo_push_node(p, nodes[n]);
    or_node_mark_update(p, nodes[n]);
    p->vertex_dabs += p->nodes[nodes[n]].uniq_verts;
  }
  if (totnode) {
    switch (d->tool) {
      case OR_TOOL_DRAW:
      case OR_TOOL_INFLATE:
      case OR_TOOL_GRAB:
        do_simple_brush(p, d, nodes, totnode);
        break;
      case OR_TOOL_CLAY_STRIPS:
        do_clay_strips_brush(p, d, nodes, totnode);
        break;
      case OR_TOOL_SMOOTH:
        if (p->is_grids && !p->edge_verts) {
          return -1; /* grid neighbours need or_grids_set_topology */
        }
        do_smooth_brush(p, d, nodes, totnode);
        break;
      default:
        return -1;
    }
  }
  if (p->is_grids && totnode) {
    /* multires_stitch_grids (kernel/intern/multires.c:1171-1196): duplicated boundary elements of the
     * faces of flagged nodes are averaged after the displacement */
    int *faces = malloc(sizeof(int) * (size_t)(p->totface + 1));
    const int num_faces = or_grids_get_updates(p, 0, faces);
    if (num_faces) {
      or_grids_stitch_faces(p, faces, num_faces);
    }
    free(faces);
  }
  if (!(d->flags & OR_DAB_NO_NORMALS)) {
    or_update_normals(p);
  }
  if (!(d->flags & OR_DAB_NO_BOUNDS)) {
    or_update_bounds(p, OR_PBVH_UpdateBB);
  }
  return totnode;
}

int or_last_hits(const OrPbvh *p, int *r)
{
  memcpy(r, p->last_hits, sizeof(int) * (size_t)p->last_tothit);
  return p->last_tothit;
}
int or_last_moved(const OrPbvh *p, int *r)
{
  if (r) {
    memcpy(r, p->last_moved, sizeof(int) * (size_t)p->last_totmoved);
  }
  return p->last_totmoved;
}
void or_last_area(const OrPbvh *p, float r_no[3], float r_co[3])
{
  memcpy(r_no, p->last_area_no, sizeof(float[3]));
  memcpy(r_co, p->last_area_co, sizeof(float[3]));
}
int or_touched_nodes(const OrPbvh *p, int *r)
{
  int n = 0;
  for (int i = 0; i < p->totnode; i++) {
    if (p->touched[i]) {
      r[n++] = i;
    }
  }
  return n;
}
int64_t or_stroke_vertex_dabs(const OrPbvh *p) { return p->vertex_dabs; }
