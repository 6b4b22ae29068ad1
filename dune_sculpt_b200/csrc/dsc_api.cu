/* libdune_sculpt_cuda: C ABI + host-side layout construction (see include/dune_sculpt_cuda.h). */
#include "../../include/dune_sculpt_cuda.h"
#include "dsc_kernels.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#define DSC_ABI_VERSION 1

static thread_local std::string g_create_error;

enum { ST_GATHER, ST_AREA, ST_BRUSH, ST_SMOOTH, ST_NORMALS, ST_LEAFBB, ST_FLUSH, ST_OTHER };
static const char *k_stage_names[DSC_NUM_STAGES] = {"gather", "area_normal", "brush", "smooth",
                                                    "normals", "leaf_bb", "bb_flush", "other"};

struct StageEvent {
  int stage;
  cudaEvent_t a, b;
};

struct DscContext {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  std::string error;
  std::vector<void *> allocs;

  /* staged mesh (host copies until the PBVH arrives) */
  bool have_mesh = false, have_pbvh = false, in_stroke = false;
  int totvert = 0, totpoly = 0, totloop = 0, tottri = 0, totnode = 0, vpad = 0, nwords = 0;
  std::vector<float> h_co, h_no, h_mask;
  std::vector<int> h_poly_start, h_poly_len, h_loop_v, h_tri_vert, h_tri_poly, h_nb_off, h_nb_idx;
  std::vector<unsigned char> h_boundary;
  bool has_no = false, has_mask = false, has_nb = false;

  std::vector<int> slot_of;   /* vertex -> slot */
  std::vector<int> leaf_node; /* leaf -> node */
  int *d_slot_of = nullptr;
  float *d_mask = nullptr, *d_automask = nullptr, *d_curve = nullptr;
  float *d_stage3 = nullptr; /* [totvert][3] export/import staging */
  unsigned *d_capture = nullptr;
  int *d_list = nullptr, *d_count = nullptr;
  DevMesh m;
  DabState *h_state = nullptr; /* pinned */
  int *h_list = nullptr;       /* pinned, nleaf ints */

  bool capture = false;
  bool stage_timing = false;
  std::vector<StageEvent> events;
  float stage_ms[DSC_NUM_STAGES] = {0};
  int stage_launches[DSC_NUM_STAGES] = {0};
  long long launches = 0;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  int grid = 148 * 8;
};

static int fail(DscContext *ctx, int code, const char *fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->error = buf;
  else g_create_error = buf;
  return code;
}

#define CU(call) \
  do { \
    cudaError_t e_ = (call); \
    if (e_ != cudaSuccess) return fail(ctx, DSC_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

template<typename T> static int dev_alloc(DscContext *ctx, T **p, size_t n)
{
  *p = nullptr;
  if (n == 0) n = 1;
  CU(cudaMalloc((void **)p, n * sizeof(T)));
  ctx->allocs.push_back((void *)*p);
  return DSC_OK;
}
template<typename T> static int dev_upload(DscContext *ctx, T **p, const std::vector<T> &v)
{
  int r = dev_alloc(ctx, p, v.size());
  if (r) return r;
  if (!v.empty()) CU(cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return DSC_OK;
}
template<typename T> static int dev_zero(DscContext *ctx, T **p, size_t n)
{
  int r = dev_alloc(ctx, p, n);
  if (r) return r;
  CU(cudaMemsetAsync(*p, 0, (n ? n : 1) * sizeof(T), ctx->stream));
  return DSC_OK;
}

/* stage bracket: counts the launch and, when stage timing is on, records an event pair */
struct StageScope {
  DscContext *ctx;
  int stage;
  cudaEvent_t a = nullptr, b = nullptr;
  StageScope(DscContext *c, int s) : ctx(c), stage(s)
  {
    ctx->launches++;
    ctx->stage_launches[stage]++;
    if (ctx->stage_timing) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, ctx->stream);
    }
  }
  ~StageScope()
  {
    if (ctx->stage_timing) {
      cudaEventRecord(b, ctx->stream);
      ctx->events.push_back({stage, a, b});
    }
  }
};

extern "C" {

int dsc_abi_version(void) { return DSC_ABI_VERSION; }

const char *dsc_last_error(const DscContext *ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

const char *dsc_stage_name(int stage) { return (stage >= 0 && stage < DSC_NUM_STAGES) ? k_stage_names[stage] : ""; }

int dsc_ctx_create(int device, DscContext **r_ctx)
{
  DscContext *ctx = nullptr;
  if (!r_ctx) return fail(nullptr, DSC_ERR_INVALID, "r_ctx is NULL");
  *r_ctx = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    return fail(nullptr, DSC_ERR_NO_DEVICE, "no CUDA device: %s (this library has no CPU fallback)",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device < 0 || device >= count) return fail(nullptr, DSC_ERR_INVALID, "device %d out of range (%d devices)", device, count);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, DSC_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, DSC_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major < 10) {
    return fail(nullptr, DSC_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  }
  ctx = new DscContext();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->grid = ctx->num_sms * 8;
  memset(&ctx->m, 0, sizeof(ctx->m));
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&ctx->t0) != cudaSuccess || cudaEventCreate(&ctx->t1) != cudaSuccess ||
      cudaMallocHost((void **)&ctx->h_state, sizeof(DabState)) != cudaSuccess) {
    fail(nullptr, DSC_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return DSC_ERR_CUDA;
  }
  *r_ctx = ctx;
  return DSC_OK;
}

void dsc_ctx_destroy(DscContext *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (void *p : ctx->allocs) cudaFree(p);
  for (auto &ev : ctx->events) {
    cudaEventDestroy(ev.a);
    cudaEventDestroy(ev.b);
  }
  if (ctx->h_state) cudaFreeHost(ctx->h_state);
  if (ctx->h_list) cudaFreeHost(ctx->h_list);
  cudaEventDestroy(ctx->t0);
  cudaEventDestroy(ctx->t1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

void *dsc_stream(DscContext *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int dsc_mesh_upload(DscContext *ctx, const DscMeshDesc *me)
{
  if (!ctx || !me) return fail(ctx, DSC_ERR_INVALID, "NULL argument");
  if (ctx->have_pbvh) return fail(ctx, DSC_ERR_STATE, "a mesh is already resident; create a new context");
  if (me->totvert <= 0 || !me->co || me->totpoly < 0 || me->tottri < 0) return fail(ctx, DSC_ERR_INVALID, "bad mesh sizes");
  if (me->tottri && (!me->tri_vert || !me->tri_poly || !me->poly_loopstart || !me->poly_totloop || !me->loop_vert))
    return fail(ctx, DSC_ERR_INVALID, "missing topology arrays");
  ctx->totvert = me->totvert;
  ctx->totpoly = me->totpoly;
  ctx->totloop = me->totloop;
  ctx->tottri = me->tottri;
  ctx->h_co.assign(me->co, me->co + (size_t)3 * me->totvert);
  ctx->has_no = me->no != nullptr;
  if (me->no) ctx->h_no.assign(me->no, me->no + (size_t)3 * me->totvert);
  ctx->has_mask = me->mask != nullptr;
  if (me->mask) ctx->h_mask.assign(me->mask, me->mask + me->totvert);
  ctx->h_poly_start.assign(me->poly_loopstart, me->poly_loopstart + me->totpoly);
  ctx->h_poly_len.assign(me->poly_totloop, me->poly_totloop + me->totpoly);
  ctx->h_loop_v.assign(me->loop_vert, me->loop_vert + me->totloop);
  ctx->h_tri_vert.assign(me->tri_vert, me->tri_vert + (size_t)3 * me->tottri);
  ctx->h_tri_poly.assign(me->tri_poly, me->tri_poly + me->tottri);
  ctx->has_nb = me->nb_offsets && me->nb_indices;
  if (ctx->has_nb) {
    ctx->h_nb_off.assign(me->nb_offsets, me->nb_offsets + me->totvert + 1);
    ctx->h_nb_idx.assign(me->nb_indices, me->nb_indices + me->nb_offsets[me->totvert]);
    if (me->boundary) ctx->h_boundary.assign(me->boundary, me->boundary + me->totvert);
    else ctx->h_boundary.assign(me->totvert, 0);
  }
  for (int i = 0; i < me->totloop; i++) {
    if (me->loop_vert[i] < 0 || me->loop_vert[i] >= me->totvert) return fail(ctx, DSC_ERR_INVALID, "loop_vert[%d] out of range", i);
  }
  ctx->have_mesh = true;
  return DSC_OK;
}

int dsc_pbvh_upload(DscContext *ctx, const DscPbvhDesc *pb)
{
  if (!ctx || !pb) return fail(ctx, DSC_ERR_INVALID, "NULL argument");
  if (!ctx->have_mesh) return fail(ctx, DSC_ERR_STATE, "dsc_mesh_upload must come first");
  if (ctx->have_pbvh) return fail(ctx, DSC_ERR_STATE, "a PBVH is already resident");
  if (pb->totnode <= 0) return fail(ctx, DSC_ERR_INVALID, "empty PBVH");
  CU(cudaSetDevice(ctx->device));
  const int V = ctx->totvert, T = ctx->tottri, N = pb->totnode;
  ctx->totnode = N;
  DevMesh &m = ctx->m;

  /* leaves in traversal order = ascending prim offset */
  std::vector<int> leaves;
  for (int n = 0; n < N; n++) {
    if (pb->flag[n] & DSC_PBVH_Leaf) leaves.push_back(n);
  }
  std::sort(leaves.begin(), leaves.end(), [&](int a, int b) { return pb->prim_offset[a] < pb->prim_offset[b]; });
  const int L = (int)leaves.size();
  ctx->leaf_node = leaves;

  /* slots */
  ctx->slot_of.assign(V, -1);
  std::vector<int> leaf_ubeg(L), leaf_ucnt(L), leaf_sbeg(L), leaf_scnt(L), leaf_pbeg(L), leaf_pcnt(L);
  long long cur = 0;
  int expect_prim = 0;
  for (int l = 0; l < L; l++) {
    const int n = leaves[l];
    cur = (cur + 31) & ~31ll;
    leaf_ubeg[l] = (int)cur;
    leaf_ucnt[l] = pb->uniq_verts[n];
    leaf_pbeg[l] = pb->prim_offset[n];
    leaf_pcnt[l] = pb->totprim[n];
    if (leaf_pbeg[l] != expect_prim) return fail(ctx, DSC_ERR_INVALID, "leaf prim ranges do not tile prim_indices");
    expect_prim += leaf_pcnt[l];
    const int *vi = pb->vert_indices + pb->vert_offset[n];
    for (int i = 0; i < pb->uniq_verts[n]; i++) {
      const int v = vi[i];
      if (v < 0 || v >= V || ctx->slot_of[v] != -1) return fail(ctx, DSC_ERR_INVALID, "vertex %d is not unique in exactly one leaf", v);
      ctx->slot_of[v] = (int)cur + i;
    }
    cur += pb->uniq_verts[n];
    if (cur > 0x7fffff00ll) return fail(ctx, DSC_ERR_UNSUPPORTED, "more than 2^31 slots");
  }
  if (expect_prim != T) return fail(ctx, DSC_ERR_INVALID, "leaves hold %d looptris, mesh has %d", expect_prim, T);
  for (int v = 0; v < V; v++) {
    if (ctx->slot_of[v] < 0) { /* loose vertex: not in any face; park it after the leaves */
      ctx->slot_of[v] = (int)cur++;
    }
  }
  const int VP = (int)((cur + 31) & ~31ll);
  ctx->vpad = VP;
  ctx->nwords = VP / 32;

  std::vector<int> shared;
  for (int l = 0; l < L; l++) {
    const int n = leaves[l];
    const int *vi = pb->vert_indices + pb->vert_offset[n];
    leaf_sbeg[l] = (int)shared.size();
    leaf_scnt[l] = pb->face_verts[n];
    for (int i = 0; i < pb->face_verts[n]; i++) shared.push_back(ctx->slot_of[vi[pb->uniq_verts[n] + i]]);
  }
  std::vector<int> chunk_leaf, chunk_beg, chunk_cnt;
  for (int l = 0; l < L; l++) {
    for (int off = 0; off < leaf_ucnt[l]; off += DSC_CHUNK) {
      chunk_leaf.push_back(l);
      chunk_beg.push_back(leaf_ubeg[l] + off);
      chunk_cnt.push_back(std::min(DSC_CHUNK, leaf_ucnt[l] - off));
    }
  }

  /* per-slot vertex data */
  auto to_slots = [&](const std::vector<float> &src, int comp, int stride) {
    std::vector<float> out((size_t)VP, 0.0f);
    for (int v = 0; v < V; v++) out[ctx->slot_of[v]] = src[(size_t)stride * v + comp];
    return out;
  };
  int r;
  for (int k = 0; k < 3; k++) {
    float **dst = (k == 0) ? &m.cx : (k == 1) ? &m.cy : &m.cz;
    if ((r = dev_upload(ctx, dst, to_slots(ctx->h_co, k, 3)))) return r;
    float **dn = (k == 0) ? &m.nx : (k == 1) ? &m.ny : &m.nz;
    if (ctx->has_no) {
      if ((r = dev_upload(ctx, dn, to_slots(ctx->h_no, k, 3)))) return r;
    }
    else if ((r = dev_zero(ctx, dn, (size_t)VP))) return r;
  }
  if ((r = dev_zero(ctx, &m.ox, (size_t)VP)) || (r = dev_zero(ctx, &m.oy, (size_t)VP)) || (r = dev_zero(ctx, &m.oz, (size_t)VP)) ||
      (r = dev_zero(ctx, &m.onx, (size_t)VP)) || (r = dev_zero(ctx, &m.ony, (size_t)VP)) || (r = dev_zero(ctx, &m.onz, (size_t)VP)))
    return r;
  if ((r = dev_zero(ctx, &ctx->d_mask, (size_t)VP)) || (r = dev_zero(ctx, &ctx->d_automask, (size_t)VP))) return r;
  if (ctx->has_mask) {
    std::vector<float> ms = to_slots(ctx->h_mask, 0, 1);
    CU(cudaMemcpyAsync(ctx->d_mask, ms.data(), (size_t)VP * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    m.mask = ctx->d_mask;
  }
  if ((r = dev_zero(ctx, &m.dirty, (size_t)ctx->nwords)) || (r = dev_zero(ctx, &m.iter_moved, (size_t)ctx->nwords)) ||
      (r = dev_zero(ctx, &ctx->d_capture, (size_t)ctx->nwords)))
    return r;
  if ((r = dev_upload(ctx, &ctx->d_slot_of, ctx->slot_of))) return r;
  if ((r = dev_alloc(ctx, &ctx->d_stage3, (size_t)3 * V))) return r;
  if ((r = dev_alloc(ctx, &ctx->d_list, (size_t)std::max(V, L))) || (r = dev_zero(ctx, &ctx->d_count, 1))) return r;

  /* smooth adjacency in slot order */
  if (ctx->has_nb) {
    if ((r = dev_zero(ctx, &m.tx, (size_t)VP)) || (r = dev_zero(ctx, &m.ty, (size_t)VP)) || (r = dev_zero(ctx, &m.tz, (size_t)VP))) return r;
    std::vector<int> vert_of((size_t)VP, -1);
    for (int v = 0; v < V; v++) vert_of[ctx->slot_of[v]] = v;
    std::vector<unsigned> off((size_t)VP + 1, 0);
    std::vector<int> idx(ctx->h_nb_idx.size());
    std::vector<unsigned char> bnd((size_t)VP, 0);
    unsigned n = 0;
    for (int s = 0; s < VP; s++) {
      off[s] = n;
      const int v = vert_of[s];
      if (v < 0) continue;
      bnd[s] = ctx->h_boundary[v];
      for (int q = ctx->h_nb_off[v]; q < ctx->h_nb_off[v + 1]; q++) idx[n++] = ctx->slot_of[ctx->h_nb_idx[q]];
    }
    off[VP] = n;
    unsigned *d_off;
    int *d_idx;
    unsigned char *d_b;
    if ((r = dev_upload(ctx, &d_off, off)) || (r = dev_upload(ctx, &d_idx, idx)) || (r = dev_upload(ctx, &d_b, bnd))) return r;
    m.nb_off = d_off;
    m.nb_idx = d_idx;
    m.boundary = d_b;
  }

  /* looptris by position: poly verts as slots, owning leaf; vertex -> looptri CSR */
  {
    std::vector<int> pv[4];
    for (int k = 0; k < 4; k++) pv[k].assign((size_t)std::max(T, 1), -1);
    std::vector<int> tri_leaf((size_t)std::max(T, 1), 0);
    std::vector<int> ngon_id(ctx->totpoly, -1), poly_off(1, 0), poly_slots;
    std::vector<unsigned> deg((size_t)VP + 1, 0);
    for (int l = 0; l < L; l++) {
      for (int pos = leaf_pbeg[l]; pos < leaf_pbeg[l] + leaf_pcnt[l]; pos++) {
        const int t = pb->prim_indices[pos];
        if (t < 0 || t >= T) return fail(ctx, DSC_ERR_INVALID, "prim_indices[%d] out of range", pos);
        tri_leaf[pos] = l;
        const int p = ctx->h_tri_poly[t];
        const int ls = ctx->h_poly_start[p], len = ctx->h_poly_len[p];
        if (len == 3 || len == 4) {
          for (int k = 0; k < len; k++) pv[k][pos] = ctx->slot_of[ctx->h_loop_v[ls + k]];
          if (len == 3) pv[3][pos] = -1;
        }
        else if (len > 4) {
          if (ngon_id[p] < 0) {
            ngon_id[p] = (int)poly_off.size() - 1;
            for (int k = 0; k < len; k++) poly_slots.push_back(ctx->slot_of[ctx->h_loop_v[ls + k]]);
            poly_off.push_back((int)poly_slots.size());
          }
          for (int k = 0; k < 3; k++) pv[k][pos] = ctx->slot_of[ctx->h_tri_vert[(size_t)3 * t + k]];
          pv[3][pos] = -2 - ngon_id[p];
        }
        else {
          return fail(ctx, DSC_ERR_UNSUPPORTED, "poly %d has %d corners", p, len);
        }
        for (int k = 0; k < 3; k++) deg[ctx->slot_of[ctx->h_tri_vert[(size_t)3 * t + k]]]++;
      }
    }
    std::vector<unsigned> vt_off((size_t)VP + 1, 0);
    for (int s = 0; s < VP; s++) vt_off[s + 1] = vt_off[s] + deg[s];
    std::vector<unsigned> vt_idx((size_t)std::max<unsigned>(vt_off[VP], 1u));
    std::fill(deg.begin(), deg.end(), 0u);
    for (int pos = 0; pos < T; pos++) {
      const int t = pb->prim_indices[pos];
      /* the reference adds the face normal for corner j = 2, 1, 0 (pbvh.c:2966); a vertex that is
       * listed twice in one looptri gets it twice -- same here, order within a looptri is moot */
      for (int k = 0; k < 3; k++) {
        const int s = ctx->slot_of[ctx->h_tri_vert[(size_t)3 * t + k]];
        vt_idx[vt_off[s] + deg[s]++] = (unsigned)pos;
      }
    }
    int *d_pv[4], *d_tl, *d_po, *d_ps;
    unsigned *d_vo, *d_vi;
    for (int k = 0; k < 4; k++) {
      if ((r = dev_upload(ctx, &d_pv[k], pv[k]))) return r;
    }
    if ((r = dev_upload(ctx, &d_tl, tri_leaf)) || (r = dev_upload(ctx, &d_po, poly_off)) || (r = dev_upload(ctx, &d_ps, poly_slots)) ||
        (r = dev_upload(ctx, &d_vo, vt_off)) || (r = dev_upload(ctx, &d_vi, vt_idx)))
      return r;
    m.pv0 = d_pv[0]; m.pv1 = d_pv[1]; m.pv2 = d_pv[2]; m.pv3 = d_pv[3];
    m.tri_leaf = d_tl; m.poly_off = d_po; m.poly_slots = d_ps; m.vt_off = d_vo; m.vt_idx = d_vi;
    CU(cudaStreamSynchronize(ctx->stream)); /* host vectors die here */
  }

  /* leaves, chunks */
  {
    int *d;
    if ((r = dev_upload(ctx, &d, leaves))) return r; m.leaf_node = d;
    if ((r = dev_upload(ctx, &d, leaf_ubeg))) return r; m.leaf_ubeg = d;
    if ((r = dev_upload(ctx, &d, leaf_ucnt))) return r; m.leaf_ucnt = d;
    if ((r = dev_upload(ctx, &d, leaf_sbeg))) return r; m.leaf_sbeg = d;
    if ((r = dev_upload(ctx, &d, leaf_scnt))) return r; m.leaf_scnt = d;
    if ((r = dev_upload(ctx, &d, leaf_pbeg))) return r; m.leaf_pbeg = d;
    if ((r = dev_upload(ctx, &d, leaf_pcnt))) return r; m.leaf_pcnt = d;
    if ((r = dev_upload(ctx, &d, shared))) return r; m.shared_slots = d;
    if ((r = dev_upload(ctx, &d, chunk_leaf))) return r; m.chunk_leaf = d;
    if ((r = dev_upload(ctx, &d, chunk_beg))) return r; m.chunk_beg = d;
    if ((r = dev_upload(ctx, &d, chunk_cnt))) return r; m.chunk_cnt = d;
    m.nleaf = L;
    m.nchunk = (int)chunk_leaf.size();
    if ((r = dev_zero(ctx, &m.leaf_state, (size_t)L))) return r;
    if ((r = dev_zero(ctx, &m.hit_list, (size_t)L)) || (r = dev_zero(ctx, &m.search_list, (size_t)L))) return r;
    CU(cudaMallocHost((void **)&ctx->h_list, sizeof(int) * (size_t)std::max(L, 1)));
  }

  /* nodes: SoA boxes, flags, children, parents, inner nodes by depth */
  {
    std::vector<float> bb((size_t)6 * N), obb((size_t)6 * N);
    std::vector<int> flag(N), child(N), parent(N, -1), depth(N, 0);
    for (int n = 0; n < N; n++) {
      for (int k = 0; k < 6; k++) {
        bb[(size_t)k * N + n] = pb->node_bb[(size_t)6 * n + k];
        obb[(size_t)k * N + n] = pb->node_orig_bb[(size_t)6 * n + k];
      }
      flag[n] = pb->flag[n];
      child[n] = pb->children_offset[n];
    }
    std::vector<int> order(1, 0);
    int maxdepth = 0;
    for (size_t i = 0; i < order.size(); i++) {
      const int n = order[i];
      if (flag[n] & DSC_PBVH_Leaf) continue;
      const int c = child[n];
      if (c <= 0 || c + 1 >= N) return fail(ctx, DSC_ERR_INVALID, "node %d has bad children_offset %d", n, c);
      for (int k = 0; k < 2; k++) {
        parent[c + k] = n;
        depth[c + k] = depth[n] + 1;
        maxdepth = std::max(maxdepth, depth[c + k]);
        order.push_back(c + k);
      }
    }
    std::vector<int> level_off(maxdepth + 2, 0), level_nodes;
    for (int dpt = 0; dpt <= maxdepth; dpt++) {
      level_off[dpt] = (int)level_nodes.size();
      for (int n : order) {
        if (!(flag[n] & DSC_PBVH_Leaf) && depth[n] == dpt) level_nodes.push_back(n);
      }
    }
    level_off[maxdepth + 1] = (int)level_nodes.size();
    m.nlevel = maxdepth + 1;
    int *d;
    if ((r = dev_upload(ctx, &m.bb, bb)) || (r = dev_upload(ctx, &m.obb, obb)) || (r = dev_upload(ctx, &m.node_flag, flag))) return r;
    if ((r = dev_upload(ctx, &d, child))) return r; m.node_child = d;
    if ((r = dev_upload(ctx, &d, parent))) return r; m.node_parent = d;
    if ((r = dev_upload(ctx, &d, level_off))) return r; m.level_off = d;
    if ((r = dev_upload(ctx, &d, level_nodes))) return r; m.level_nodes = d;
    if ((r = dev_zero(ctx, &m.node_mark, (size_t)N))) return r;
    m.totnode = N;
    CU(cudaStreamSynchronize(ctx->stream));
  }
  if ((r = dev_zero(ctx, &m.st, 1))) return r;
  if ((r = dev_zero(ctx, &ctx->d_curve, 257))) return r;
  CU(cudaStreamSynchronize(ctx->stream));

  /* the host staging copies are no longer needed */
  std::vector<float>().swap(ctx->h_co);
  std::vector<float>().swap(ctx->h_no);
  std::vector<float>().swap(ctx->h_mask);
  std::vector<int>().swap(ctx->h_tri_vert);
  std::vector<int>().swap(ctx->h_tri_poly);
  std::vector<int>().swap(ctx->h_loop_v);
  std::vector<int>().swap(ctx->h_nb_idx);
  ctx->have_pbvh = true;
  if (!ctx->has_no) return dsc_recalc_normals(ctx);
  return DSC_OK;
}

#define NEED_PBVH() \
  do { \
    if (!ctx) return DSC_ERR_INVALID; \
    if (!ctx->have_pbvh) return fail(ctx, DSC_ERR_STATE, "no PBVH resident (dsc_mesh_upload + dsc_pbvh_upload first)"); \
    CU(cudaSetDevice(ctx->device)); \
  } while (0)

#define LAUNCH_CHECK() CU(cudaGetLastError())

static int run_normals(DscContext *ctx)
{
  StageScope s(ctx, ST_NORMALS);
  k_normals<<<ctx->grid, DSC_BLOCK, 0, ctx->stream>>>(ctx->m);
  LAUNCH_CHECK();
  return DSC_OK;
}
static int run_bounds(DscContext *ctx, int clear_mask)
{
  {
    StageScope s(ctx, ST_LEAFBB);
    k_leaf_bb<<<ctx->grid, DSC_BLOCK, 0, ctx->stream>>>(ctx->m);
    LAUNCH_CHECK();
  }
  {
    StageScope s(ctx, ST_FLUSH);
    k_flush<<<1, 1024, 0, ctx->stream>>>(ctx->m, clear_mask);
    LAUNCH_CHECK();
  }
  return DSC_OK;
}
static int run_flush_only(DscContext *ctx, int clear_mask)
{
  StageScope s(ctx, ST_FLUSH);
  k_flush<<<1, 1024, 0, ctx->stream>>>(ctx->m, clear_mask);
  LAUNCH_CHECK();
  return DSC_OK;
}

int dsc_recalc_normals(DscContext *ctx)
{
  NEED_PBVH();
  {
    StageScope s(ctx, ST_OTHER);
    const int n = std::max(ctx->m.nleaf, 1);
    k_mark_all<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->m, F_UpdateNormals, 1, ctx->nwords);
    LAUNCH_CHECK();
  }
  int r = run_normals(ctx);
  if (r) return r;
  return run_flush_only(ctx, F_UpdateNormals);
}

int dsc_set_custom_curve(DscContext *ctx, const float *table257)
{
  NEED_PBVH();
  if (!table257) {
    ctx->m.curve = nullptr;
    return DSC_OK;
  }
  CU(cudaMemcpyAsync(ctx->d_curve, table257, 257 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->m.curve = ctx->d_curve;
  return DSC_OK;
}

static int upload_per_vertex(DscContext *ctx, float *dst, const float *src)
{
  std::vector<float> tmp((size_t)ctx->vpad, 0.0f);
  for (int v = 0; v < ctx->totvert; v++) tmp[ctx->slot_of[v]] = src[v];
  CU(cudaMemcpyAsync(dst, tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return DSC_OK;
}

int dsc_set_mask(DscContext *ctx, const float *mask)
{
  NEED_PBVH();
  if (!mask) {
    ctx->m.mask = nullptr;
    return DSC_OK;
  }
  int r = upload_per_vertex(ctx, ctx->d_mask, mask);
  if (r) return r;
  ctx->m.mask = ctx->d_mask;
  return DSC_OK;
}

int dsc_node_flag_set(DscContext *ctx, int node, int flag, int on)
{
  NEED_PBVH();
  if (node < 0 || node >= ctx->totnode) return fail(ctx, DSC_ERR_INVALID, "node %d out of range", node);
  int f = 0;
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpy(&f, ctx->m.node_flag + node, sizeof(int), cudaMemcpyDeviceToHost));
  f = on ? (f | flag) : (f & ~flag);
  CU(cudaMemcpy(ctx->m.node_flag + node, &f, sizeof(int), cudaMemcpyHostToDevice));
  return DSC_OK;
}

int dsc_node_mark_update(DscContext *ctx, int node)
{
  return dsc_node_flag_set(ctx, node,
                           F_UpdateNormals | F_UpdateBB | F_UpdateOriginalBB | F_UpdateDrawBuffers | F_UpdateRedraw, 1);
}

int dsc_stroke_begin(DscContext *ctx, const float *automask)
{
  NEED_PBVH();
  if (ctx->in_stroke) return fail(ctx, DSC_ERR_STATE, "stroke already open");
  if (automask) {
    int r = upload_per_vertex(ctx, ctx->d_automask, automask);
    if (r) return r;
    ctx->m.automask = ctx->d_automask;
  }
  else {
    ctx->m.automask = nullptr;
  }
  CU(cudaMemsetAsync(ctx->m.leaf_state, 0, sizeof(unsigned) * (size_t)std::max(ctx->m.nleaf, 1), ctx->stream));
  CU(cudaMemsetAsync(ctx->m.st, 0, sizeof(DabState), ctx->stream));
  ctx->launches = 0;
  ctx->in_stroke = true;
  return DSC_OK;
}

int dsc_dab(DscContext *ctx, const DscDab *dab)
{
  NEED_PBVH();
  if (!dab) return fail(ctx, DSC_ERR_INVALID, "dab is NULL");
  if (!ctx->in_stroke) return fail(ctx, DSC_ERR_STATE, "dsc_stroke_begin first");
  const int tool = dab->tool;
  if (tool != DSC_TOOL_DRAW && tool != DSC_TOOL_SMOOTH && tool != DSC_TOOL_INFLATE && tool != DSC_TOOL_GRAB &&
      tool != DSC_TOOL_CLAY_STRIPS)
    return fail(ctx, DSC_ERR_UNSUPPORTED, "sculpt tool %d is not on the accelerated path", tool);
  if (tool == DSC_TOOL_SMOOTH && !ctx->has_nb) return fail(ctx, DSC_ERR_STATE, "smooth brush needs the neighbour CSR (DscMeshDesc.nb_offsets)");
  if (!(dab->radius > 0.0f)) return fail(ctx, DSC_ERR_INVALID, "radius must be positive");
  DabParams d;
  static_assert(sizeof(DabParams) == sizeof(DscDab), "DabParams mirrors DscDab");
  memcpy(&d, dab, sizeof(d));
  DevMesh &m = ctx->m;
  cudaStream_t st = ctx->stream;

  if (ctx->capture) CU(cudaMemsetAsync(ctx->d_capture, 0, sizeof(unsigned) * (size_t)ctx->nwords, st));
  /* 1. gather + undo membership + node marks */
  {
    StageScope s(ctx, ST_GATHER);
    const float rs = dab->radius * dab->radius_scale;
    k_gather<<<1, 1024, 0, st>>>(m, dab->location[0], dab->location[1], dab->location[2], rs * rs,
                                 tool == DSC_TOOL_GRAB ? 1 : 0, 1, 1);
    LAUNCH_CHECK();
  }
  /* 2.-3. brush */
  if (tool == DSC_TOOL_SMOOTH) {
    {
      StageScope s(ctx, ST_SMOOTH);
      k_snapshot<<<ctx->grid, DSC_BLOCK, 0, st>>>(m);
      LAUNCH_CHECK();
    }
    const int max_iterations = 4;
    const float fract = 1.0f / (float)max_iterations;
    float bstrength = dab->bstrength;
    bstrength = bstrength < 0.0f ? 0.0f : (bstrength > 1.0f ? 1.0f : bstrength);
    const int count = (int)(bstrength * (float)max_iterations);
    const float last = (float)max_iterations * (bstrength - (float)count * fract);
    for (int it = 0; it <= count; it++) {
      float strength = (it != count) ? 1.0f : last;
      strength = strength < 0.0f ? 0.0f : (strength > 1.0f ? 1.0f : strength);
      /* a zero-strength tail iteration moves nothing and marks only verts the previous iteration
       * already marked; skip it */
      if (it == count && count > 0 && strength == 0.0f) break;
      {
        StageScope s(ctx, ST_SMOOTH);
        k_smooth_a<<<ctx->grid, DSC_BLOCK, 0, st>>>(m, d, strength);
        LAUNCH_CHECK();
      }
      {
        StageScope s(ctx, ST_SMOOTH);
        k_smooth_b<<<ctx->grid, DSC_BLOCK, 0, st>>>(m);
        LAUNCH_CHECK();
      }
    }
  }
  else {
    const bool needs_area = (tool == DSC_TOOL_DRAW && dab->sculpt_plane == DSC_DIR_AREA) || tool == DSC_TOOL_CLAY_STRIPS;
    if (needs_area) {
      StageScope s(ctx, ST_AREA);
      k_area<<<ctx->grid, DSC_BLOCK, 0, st>>>(m, d, tool == DSC_TOOL_CLAY_STRIPS ? 1 : 0);
      LAUNCH_CHECK();
    }
    StageScope s(ctx, ST_BRUSH);
    k_brush<<<ctx->grid, DSC_BLOCK, 0, st>>>(m, d);
    LAUNCH_CHECK();
  }
  /* 4. normals, 5. bounds */
  int clear = 0, r;
  if (!(dab->flags & DSC_DAB_NO_NORMALS)) {
    if ((r = run_normals(ctx))) return r;
    clear |= F_UpdateNormals;
  }
  if (!(dab->flags & DSC_DAB_NO_BOUNDS)) {
    if ((r = run_bounds(ctx, clear | F_UpdateBB))) return r;
  }
  else if (clear) {
    if ((r = run_flush_only(ctx, clear))) return r;
  }
  return DSC_OK;
}

static int read_list(DscContext *ctx, const int *d_list, const int *d_count_field, int *r_nodes, int capacity, int *r_tot)
{
  int tot = 0;
  CU(cudaMemcpyAsync(&ctx->h_state->hit_count, d_count_field, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  tot = ctx->h_state->hit_count;
  if (r_tot) *r_tot = tot;
  if (r_nodes && tot > 0) {
    if (capacity < tot) return fail(ctx, DSC_ERR_INVALID, "capacity %d < %d gathered nodes", capacity, tot);
    CU(cudaMemcpyAsync(ctx->h_list, d_list, sizeof(int) * (size_t)tot, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < tot; i++) r_nodes[i] = ctx->leaf_node[ctx->h_list[i]];
  }
  return DSC_OK;
}

int dsc_gather_readback(DscContext *ctx, int *r_nodes, int capacity, int *r_tot)
{
  NEED_PBVH();
  return read_list(ctx, ctx->m.hit_list, &ctx->m.st->hit_count, r_nodes, capacity, r_tot);
}

int dsc_search_sphere(DscContext *ctx, const float center[3], float radius_sq, int original, int ignore_fully_ineffective,
                      int *r_nodes, int capacity, int *r_tot)
{
  NEED_PBVH();
  {
    StageScope s(ctx, ST_GATHER);
    k_gather<<<1, 1024, 0, ctx->stream>>>(ctx->m, center[0], center[1], center[2], radius_sq, original ? 1 : 0,
                                          ignore_fully_ineffective ? 1 : 0, 0);
    LAUNCH_CHECK();
  }
  return read_list(ctx, ctx->m.search_list, &ctx->m.st->search_count, r_nodes, capacity, r_tot);
}

int dsc_last_area(DscContext *ctx, float r_no[3], float r_co[3])
{
  NEED_PBVH();
  CU(cudaMemcpyAsync(ctx->h_state, ctx->m.st, sizeof(DabState), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < 3; k++) {
    if (r_no) r_no[k] = ctx->h_state->area_no[k];
    if (r_co) r_co[k] = ctx->h_state->area_co[k];
  }
  return DSC_OK;
}

int dsc_debug_capture(DscContext *ctx, int on)
{
  if (!ctx) return DSC_ERR_INVALID;
  ctx->capture = on != 0;
  ctx->m.capture = ctx->capture ? ctx->d_capture : nullptr;
  return DSC_OK;
}

int dsc_last_moved(DscContext *ctx, int *r_verts, int capacity, int *r_tot)
{
  NEED_PBVH();
  if (!ctx->capture) return fail(ctx, DSC_ERR_STATE, "dsc_debug_capture(ctx, 1) first");
  CU(cudaMemsetAsync(ctx->d_count, 0, sizeof(int), ctx->stream));
  k_export_bits<<<ctx->grid, 256, 0, ctx->stream>>>(ctx->d_count, ctx->d_list, ctx->d_capture, ctx->d_slot_of, ctx->totvert);
  LAUNCH_CHECK();
  int tot = 0;
  CU(cudaMemcpyAsync(&tot, ctx->d_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (r_tot) *r_tot = tot;
  if (r_verts && tot > 0) {
    if (capacity < tot) return fail(ctx, DSC_ERR_INVALID, "capacity %d < %d moved verts", capacity, tot);
    CU(cudaMemcpy(r_verts, ctx->d_list, sizeof(int) * (size_t)tot, cudaMemcpyDeviceToHost));
    std::sort(r_verts, r_verts + tot);
  }
  return DSC_OK;
}

int dsc_stroke_stats(DscContext *ctx, DscStrokeStats *r)
{
  NEED_PBVH();
  if (!r) return fail(ctx, DSC_ERR_INVALID, "r_stats is NULL");
  CU(cudaMemcpyAsync(ctx->h_state, ctx->m.st, sizeof(DabState), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  r->vertex_dabs = (int64_t)ctx->h_state->vd_total;
  r->node_hits = (int64_t)ctx->h_state->hits_total;
  r->moved_verts = (int64_t)ctx->h_state->moved_total;
  r->dabs = (int64_t)ctx->h_state->dabs;
  r->kernel_launches = ctx->launches;
  return DSC_OK;
}

int dsc_update_normals(DscContext *ctx)
{
  NEED_PBVH();
  int r = run_normals(ctx);
  if (r) return r;
  return run_flush_only(ctx, F_UpdateNormals);
}

static int run_orig_flush(DscContext *ctx)
{
  StageScope s(ctx, ST_OTHER);
  k_orig_leaves<<<ctx->num_sms, DSC_BLOCK, 0, ctx->stream>>>(ctx->m);
  LAUNCH_CHECK();
  k_orig_inner<<<ctx->num_sms, DSC_BLOCK, 0, ctx->stream>>>(ctx->m);
  LAUNCH_CHECK();
  ctx->launches++;
  return DSC_OK;
}

int dsc_update_bounds(DscContext *ctx, int flag)
{
  NEED_PBVH();
  int r;
  if (flag & DSC_PBVH_UpdateBB) {
    if ((r = run_bounds(ctx, F_UpdateBB))) return r;
  }
  if (flag & DSC_PBVH_UpdateOriginalBB) {
    if ((r = run_orig_flush(ctx))) return r;
  }
  return DSC_OK;
}

int dsc_stroke_end(DscContext *ctx)
{
  NEED_PBVH();
  if (!ctx->in_stroke) return fail(ctx, DSC_ERR_STATE, "no stroke open");
  int r = run_orig_flush(ctx);
  if (r) return r;
  ctx->in_stroke = false;
  CU(cudaStreamSynchronize(ctx->stream));
  return DSC_OK;
}

static int export3(DscContext *ctx, float *out, const float *ax, const float *ay, const float *az)
{
  if (!out) return fail(ctx, DSC_ERR_INVALID, "output pointer is NULL");
  k_export3<<<ctx->grid, 256, 0, ctx->stream>>>(ctx->d_stage3, ax, ay, az, ctx->d_slot_of, ctx->totvert);
  LAUNCH_CHECK();
  CU(cudaMemcpyAsync(out, ctx->d_stage3, sizeof(float) * 3 * (size_t)ctx->totvert, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return DSC_OK;
}

int dsc_download_co(DscContext *ctx, float *r_co)
{
  NEED_PBVH();
  return export3(ctx, r_co, ctx->m.cx, ctx->m.cy, ctx->m.cz);
}
int dsc_download_no(DscContext *ctx, float *r_no)
{
  NEED_PBVH();
  return export3(ctx, r_no, ctx->m.nx, ctx->m.ny, ctx->m.nz);
}
int dsc_download_orig_co(DscContext *ctx, float *r_co)
{
  NEED_PBVH();
  return export3(ctx, r_co, ctx->m.ox, ctx->m.oy, ctx->m.oz);
}
int dsc_download_orig_no(DscContext *ctx, float *r_no)
{
  NEED_PBVH();
  return export3(ctx, r_no, ctx->m.onx, ctx->m.ony, ctx->m.onz);
}

int dsc_download_node_bb(DscContext *ctx, float *r_bb, float *r_orig_bb)
{
  NEED_PBVH();
  const int N = ctx->totnode;
  std::vector<float> tmp((size_t)6 * N);
  CU(cudaStreamSynchronize(ctx->stream));
  for (int pass = 0; pass < 2; pass++) {
    float *out = pass ? r_orig_bb : r_bb;
    if (!out) continue;
    CU(cudaMemcpy(tmp.data(), pass ? ctx->m.obb : ctx->m.bb, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int n = 0; n < N; n++) {
      for (int k = 0; k < 6; k++) out[(size_t)6 * n + k] = tmp[(size_t)k * N + n];
    }
  }
  return DSC_OK;
}

int dsc_download_node_flags(DscContext *ctx, int *r_flags)
{
  NEED_PBVH();
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpy(r_flags, ctx->m.node_flag, sizeof(int) * (size_t)ctx->totnode, cudaMemcpyDeviceToHost));
  return DSC_OK;
}

int dsc_download_touched(DscContext *ctx, unsigned char *r_touched)
{
  NEED_PBVH();
  const int L = ctx->m.nleaf;
  std::vector<unsigned> st((size_t)std::max(L, 1));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpy(st.data(), ctx->m.leaf_state, sizeof(unsigned) * (size_t)L, cudaMemcpyDeviceToHost));
  memset(r_touched, 0, (size_t)ctx->totnode);
  for (int l = 0; l < L; l++) {
    if (st[l] & DSC_LEAF_TOUCHED) r_touched[ctx->leaf_node[l]] = 1;
  }
  return DSC_OK;
}

int dsc_upload_co(DscContext *ctx, const float *co)
{
  NEED_PBVH();
  if (!co) return fail(ctx, DSC_ERR_INVALID, "co is NULL");
  CU(cudaMemcpyAsync(ctx->d_stage3, co, sizeof(float) * 3 * (size_t)ctx->totvert, cudaMemcpyHostToDevice, ctx->stream));
  k_import3<<<ctx->grid, 256, 0, ctx->stream>>>(ctx->d_stage3, ctx->m.cx, ctx->m.cy, ctx->m.cz, ctx->d_slot_of, ctx->m.dirty,
                                                ctx->totvert);
  LAUNCH_CHECK();
  /* pbvh.c:4743-4747: every node is marked, bounds (vb and orig_vb) are refreshed */
  const int n = std::max(ctx->m.nleaf, 1);
  k_mark_all<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->m, F_UpdateNormals | F_UpdateBB | F_UpdateOriginalBB | F_UpdateDrawBuffers | F_UpdateRedraw,
                                                       0, ctx->nwords);
  LAUNCH_CHECK();
  int r = run_normals(ctx);
  if (r) return r;
  if ((r = run_bounds(ctx, F_UpdateNormals | F_UpdateBB))) return r;
  if ((r = run_orig_flush(ctx))) return r;
  CU(cudaStreamSynchronize(ctx->stream));
  return DSC_OK;
}

int dsc_synchronize(DscContext *ctx)
{
  if (!ctx) return DSC_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  return DSC_OK;
}

int dsc_timer_start(DscContext *ctx)
{
  if (!ctx) return DSC_ERR_INVALID;
  CU(cudaEventRecord(ctx->t0, ctx->stream));
  return DSC_OK;
}
int dsc_timer_stop(DscContext *ctx, float *r_ms)
{
  if (!ctx) return DSC_ERR_INVALID;
  CU(cudaEventRecord(ctx->t1, ctx->stream));
  CU(cudaEventSynchronize(ctx->t1));
  float ms = 0.0f;
  CU(cudaEventElapsedTime(&ms, ctx->t0, ctx->t1));
  if (r_ms) *r_ms = ms;
  return DSC_OK;
}

int dsc_stage_timing(DscContext *ctx, int enable)
{
  if (!ctx) return DSC_ERR_INVALID;
  CU(cudaStreamSynchronize(ctx->stream));
  for (auto &ev : ctx->events) {
    cudaEventDestroy(ev.a);
    cudaEventDestroy(ev.b);
  }
  ctx->events.clear();
  for (int i = 0; i < DSC_NUM_STAGES; i++) {
    ctx->stage_ms[i] = 0.0f;
    ctx->stage_launches[i] = 0;
  }
  ctx->stage_timing = enable != 0;
  return DSC_OK;
}

int dsc_stage_times(DscContext *ctx, float r_ms[DSC_NUM_STAGES], int r_launches[DSC_NUM_STAGES])
{
  if (!ctx) return DSC_ERR_INVALID;
  CU(cudaStreamSynchronize(ctx->stream));
  for (auto &ev : ctx->events) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, ev.a, ev.b) == cudaSuccess) ctx->stage_ms[ev.stage] += ms;
    cudaEventDestroy(ev.a);
    cudaEventDestroy(ev.b);
  }
  ctx->events.clear();
  for (int i = 0; i < DSC_NUM_STAGES; i++) {
    if (r_ms) r_ms[i] = ctx->stage_ms[i];
    if (r_launches) r_launches[i] = ctx->stage_launches[i];
  }
  return DSC_OK;
}

} /* extern "C" */
