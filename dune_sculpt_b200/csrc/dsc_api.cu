/* libdune_sculpt_cuda: C ABI + host-side layout construction (see include/dune_sculpt_cuda.h). */
#include "../../include/dune_sculpt_cuda.h"
#include "dsc_kernels.cuh"
#include "dsc_grids.cuh"

#include <cupti.h>
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#define DSC_ABI_VERSION 1
#define DSC_REGION_CHUNKS 16

static thread_local std::string g_create_error;

enum { ST_GATHER, ST_AREA, ST_BRUSH, ST_SMOOTH, ST_NORMALS, ST_LEAFBB, ST_FLUSH, ST_OTHER };
static const char *k_stage_names[DSC_NUM_STAGES] = {"gather", "area_normal", "brush", "smooth",
                                                    "normals_bb", "leaf_bb", "bb_refit", "other"};

struct StageEvent {
  int stage;
  cudaEvent_t a, b;
};

struct DabGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int launches = 0;
  int stage_launches[DSC_NUM_STAGES] = {0};
  std::vector<StageEvent> events; /* stage timing inside the replayed graph (dsc_stage_timing mode 2): event-record nodes */
};

struct DscContext {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;  /* the dab pipeline */
  cudaStream_t stream2 = nullptr; /* bottom-up box refit, overlapped with the next dab */
  cudaEvent_t ev_fork = nullptr, ev_bb = nullptr, ev_tag = nullptr, ev_refit[DSC_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  /* the same events for launches recorded into a graph (a captured event cannot be waited on outside its graph) */
  cudaEvent_t cap_fork = nullptr, cap_bb = nullptr, cap_tag = nullptr, cap_join = nullptr, cap_refit[DSC_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  bool side_busy = false; /* something was queued on stream2 since the last join */
  std::string error;
  std::vector<void *> allocs;

  /* staged mesh (host copies until the PBVH arrives) */
  bool have_mesh = false, have_pbvh = false, in_stroke = false;
  int totvert = 0, totpoly = 0, totloop = 0, tottri = 0, totnode = 0, vpad = 0, nwords = 0;
  std::vector<float> h_co, h_no, h_mask;
  std::vector<unsigned> h_tail;
  unsigned *d_tail = nullptr;
  unsigned *d_hidden = nullptr;
  std::vector<int> h_poly_start, h_poly_len, h_loop_v, h_tri_vert, h_tri_poly, h_nb_off, h_nb_idx;
  std::vector<unsigned char> h_boundary;
  bool has_no = false, has_mask = false, has_nb = false;

  /* multires grids (dsc_grids_upload) */
  bool is_grids = false;
  int grid_size = 0, totgrid = 0;
  std::vector<int> h_face_start, h_face_num, h_edge_off, h_edge_elems, h_cvert_off, h_cvert_elems, h_grid_edge, h_grid_cvert;
  DevGrids g;
  GridNb gnb = {}; /* smooth brush on grids: rim neighbour table in slot space */
  std::vector<int> h_rim_nb;
  std::vector<unsigned char> h_rim_bnd;
  std::vector<unsigned char> h_elem_hidden; /* grids: DscGridsDesc.hidden */
  int rim_width = 0;
  size_t gn_smem = 0;
  /* draw-buffer fill (dsc_draw_*) */
  bool want_draw = false, want_raycast = false;
  const int *d_slot_leaf = nullptr; /* [slots / 32] leaf owning the 32-slot group (ray-cast: undo-node lookup per vert) */
  RayLeafHit *d_ray_out = nullptr, *h_ray_out = nullptr; /* [nleaf] on the device, the first DSC_RAY_FIRST pinned */
  GridRayHit *d_gray_out = nullptr;                       /* grids: the quads a ray touches */
  int gray_capacity = 4096;
  int *d_ray_count = nullptr, *h_ray_count = nullptr;
  std::vector<int> vert_of_slot;
  const int4 *d_tri_slots = nullptr;
  unsigned *d_vbo = nullptr; /* [tottri * 3][9] packed vertex records, by looptri position; grids: [totgrid * draw_per_grid][9] */
  int draw_per_grid = 0; /* grids: records per grid of the uniform layout; -1 = the per-leaf layout (leaf_rec) */
  std::vector<unsigned char> h_leaf_smooth; /* dsc_draw_leaf_shading: ME_SMOOTH per leaf */
  std::vector<long long> h_leaf_rec;        /* grids, per-leaf layout: first record of each leaf (+ the total) */
  unsigned char *d_leaf_smooth = nullptr;
  long long *d_leaf_rec = nullptr;
  bool has_odd_edges = false; /* some coarse edge has more than two faces */
  bool grid_normals_flat = false; /* DSC_GRID_NORMALS_FLAT=1: the element-parallel normal pass (measured slower: 4 x the IEEE sqrt / div work) */
  bool grid_fused = false;    /* DSC_GRID_FUSED=1: the stages after the brush as one cooperative kernel instead of nine launches */

  std::vector<int> slot_of;     /* vertex -> slot */
  std::vector<int> leaf_node;   /* leaf (= device id) -> host node index */
  std::vector<int> h_leaf_pbeg, h_leaf_pcnt; /* leaf -> its run of looptri positions */
  std::vector<int> dev_of_node; /* host node index -> device node id */
  std::vector<int> node_of_dev;
  bool stale_flags = false; /* leaves may carry update flags from an earlier dab or from the host */
  bool any_slow_leaf = false;
  size_t nb_smem = 0; /* dynamic shared memory of k_normals_bb_smem */
  int nb_grid = 148;
  long long dab_index = 0;
  int last_slot = 0;

  /* multi-GPU */
  int world = 1, rank = 0;
  ncclComm_t comm = nullptr;
  std::vector<int> leaf_range;            /* [world + 1] traversal-order leaf bounds */
  std::vector<int> slot_range;            /* [world + 1] slot bounds of the owned unique-vert runs */
  std::vector<int> send_off, recv_off;    /* [world + 1] into the index lists below */
  int *d_send_idx = nullptr, *d_recv_idx = nullptr;
  int *d_glist = nullptr, *d_gcount = nullptr; /* partitioned grids: the leaves all ranks gathered this dab */
  /* exchanges over peer memory (see PeerLink in dsc_kernels.cuh); NCCL carries them when the mapping is refused */
  bool p2p = false;
  unsigned *d_near_mask = nullptr; /* bit per leaf: gathering it can change a halo element on some rank */
  PeerLink link = {};
  void *p2p_region = nullptr;
  void *p2p_peer_region[DSC_MAX_RANKS] = {nullptr};
  float *d_send_buf = nullptr, *d_recv_buf = nullptr;
  /* which ranks a dab can reach (dist_dab_mask): conservative boxes of every rank's region -- DSC_REGION_CHUNKS runs of its
   * leaves in traversal order plus the box of the halo elements it reads -- exact at stroke begin, grown by every dab that
   * can move something inside them; all ranks keep the same copy (a pure function of the stroke-start boxes and the dabs) */
  std::vector<float> regions;    /* [world][DSC_REGION_CHUNKS][6] */
  std::vector<unsigned> leaf_reach; /* per leaf: bit per rank that gathering the leaf can affect -- its owner, the ranks that
                                       read one of its elements, on grids the owners of every grid within two face hops (the
                                       stitch and the normal pass reach one hop, a rank reads one hop beyond its own grids) */
  std::vector<std::vector<int>> region_leaves; /* per rank: the leaves that can affect it, ascending */
  int pending_skipped = 0;       /* grids: dabs this rank takes no part in whose all-coarse-vertex averaging is still to run */
  bool last_dab_skipped = false;
  bool subset_exchange = false;  /* peer-memory transport: dabs are exchanged among the ranks they reach only */
  std::vector<int> own_grid_runs; /* partitioned grids: {first grid, count} of every run of consecutive owned grids */
  int own_grid_count = 0;
  int *d_own_grid_pos = nullptr;  /* [totgrid] rank of the grid among the owned ones, -1 */
  float *d_own_pack = nullptr;    /* packed CCGElem records of the owned grids */
  std::vector<void *> registered; /* host arrays this context page-locked and mapped (dsc_download_owned_*) */
  float *h_own = nullptr;        /* pinned staging of the owned slot runs (stroke-end sync of a partitioned PBVH) */
  size_t h_own_floats = 0;
  long long dist_skipped_dabs = 0, dist_local_dabs = 0, dist_exchanged_dabs = 0;

  int *d_slot_of = nullptr;
  float *d_mask = nullptr, *d_automask = nullptr, *d_curve = nullptr;
  float *d_stage3 = nullptr; /* [totvert][3] export/import staging */
  float *d_save_v = nullptr, *d_save_bb = nullptr; /* dsc_state_save: co / no, node + tile boxes */
  int *d_save_flag = nullptr;
  bool have_save = false, save_stale_flags = false;
  unsigned *d_capture = nullptr;
  int *d_list = nullptr, *d_count = nullptr;
  DevMesh m;
  DabState *h_state = nullptr;   /* pinned */
  StrokeTotals *h_tot = nullptr; /* pinned */
  int *h_list = nullptr;         /* pinned, nleaf ints */

  /* dab ring: pinned host staging + device copy; executable graphs by launch-sequence signature */
  DabEntry *h_ring = nullptr, *d_ring = nullptr;
  int *d_ring_ctl = nullptr;
  cudaEvent_t ev_ring[4] = {nullptr, nullptr, nullptr, nullptr};
  std::unordered_map<unsigned, DabGraph> graphs;
  bool use_graphs = true, use_pdl = false, use_batch_kernel = false;
  /* the inner nodes' boxes are refitted when somebody reads them (stroke end, a download), not after every dab: nothing on
   * the device reads them (the gather and the ray-cast are flat over the leaves).  DSC_EAGER_REFIT=1: after every dab, on
   * the side stream, as pbvh_flush_bb would */
  bool lazy_refit = true, refit_pending = false;
  float small_dab_frac = 0.0f, small_dab_radius = 0.0f; /* dabs up to this radius (fraction of the root box diagonal) take the persistent kernel */
  int small_dab_grid = 1 << 30;                         /* ... on this many CTAs */
  int batch_grid[4] = {0, 0, 0, 0}; /* resident CTAs of k_dab_batch<tool> */
  int fused_grid[4] = {0, 0, 0, 0}; /* SMs x resident CTAs of k_dab_tile<tool> */
  bool use_fused = false;           /* DSC_FUSE=1: boundary brush + fused interior brush / normals / boxes kernel */
  long long batch_launches = 0;
  long long graph_launches = 0, ring_seq = 0;

  bool capture = false;
  bool capturing_now = false; /* between cudaStreamBeginCapture and EndCapture of get_graph */
  int stage_timing = 0; /* 1: direct launches, an event pair around every kernel; 2: the same pairs as nodes of the replayed graphs */
  std::vector<StageEvent> events;
  float stage_ms[DSC_NUM_STAGES] = {0};
  int stage_launches[DSC_NUM_STAGES] = {0};
  long long launches = 0;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  int grid = 148 * 8;
  int grid_bnd = 148 * 4;                        /* k_brush_boundary: a warp per tile */
  int grid_area = 148 * 8, grid_brush = 148 * 8; /* per-dab kernels: SMs x a multiple from DSC_GRID_AREA / DSC_GRID_BRUSH */
};

static void invalidate_graphs(DscContext *ctx)
{
  for (auto &kv : ctx->graphs) {
    cudaGraphExecDestroy(kv.second.exec);
    cudaGraphDestroy(kv.second.graph);
  }
  ctx->graphs.clear();
}

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
template<typename T, typename U> static int dev_upload_c(DscContext *ctx, const T **p, const std::vector<U> &v)
{
  U *d = nullptr;
  int r = dev_upload(ctx, &d, v);
  *p = reinterpret_cast<const T *>(d);
  return r;
}
template<typename T> static int dev_zero(DscContext *ctx, T **p, size_t n)
{
  int r = dev_alloc(ctx, p, n);
  if (r) return r;
  CU(cudaMemsetAsync(*p, 0, (n ? n : 1) * sizeof(T), ctx->stream));
  return DSC_OK;
}

/* kernel launch, optionally as a programmatic dependent of the launch before it on the stream (see
 * dsc_pdl_wait in dsc_kernels.cuh) */
template<typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool pdl, Args... args)
{
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3((unsigned)block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

/* stage bracket: counts the launch and, when stage timing is on, records an event pair */
struct StageScope {
  DscContext *ctx;
  int stage;
  cudaStream_t st;
  cudaEvent_t a = nullptr, b = nullptr;
  StageScope(DscContext *c, int s, cudaStream_t stream = nullptr) : ctx(c), stage(s), st(stream ? stream : c->stream)
  {
    ctx->launches++;
    ctx->stage_launches[stage]++;
    if (ctx->stage_timing == 1 || ctx->stage_timing == 2) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      /* inside a capture a plain record is only a dependency marker; an external record becomes a node of the graph */
      if (ctx->capturing_now) cudaEventRecordWithFlags(a, st, cudaEventRecordExternal);
      else cudaEventRecord(a, st);
    }
  }
  ~StageScope()
  {
    if (ctx->stage_timing == 1 || ctx->stage_timing == 2) {
      if (ctx->capturing_now) cudaEventRecordWithFlags(b, st, cudaEventRecordExternal);
      else cudaEventRecord(b, st);
      ctx->events.push_back({stage, a, b});
    }
  }
};

/* the main stream waits for everything queued on the side stream */
static int join_side(DscContext *ctx)
{
  if (ctx->side_busy) {
    CU(cudaEventRecord(ctx->ev_fork, ctx->stream2));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_fork, 0));
    ctx->side_busy = false;
  }
  return DSC_OK;
}
static int sync_all(DscContext *ctx)
{
  CU(cudaStreamSynchronize(ctx->stream2));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->side_busy = false;
  return DSC_OK;
}


/* ------------------------------------------------------------------------------ NCCL, loaded lazily */
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static bool nccl_load(std::string &err)
{
  if (g_nccl.lib) return true;
  /* picks up the libnccl a host process (e.g. torch) already loaded, else the system one */
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    err = std::string("dlopen(libnccl.so.2): ") + dlerror();
    return false;
  }
#define NCCL_SYM(field, name) \
  *(void **)(&g_nccl.field) = dlsym(lib, name); \
  if (!g_nccl.field) { \
    err = std::string("libnccl lacks ") + name; \
    return false; \
  }
  NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
  NCCL_SYM(CommInitRank, "ncclCommInitRank");
  NCCL_SYM(CommDestroy, "ncclCommDestroy");
  NCCL_SYM(AllReduce, "ncclAllReduce");
  NCCL_SYM(Broadcast, "ncclBroadcast");
  NCCL_SYM(Send, "ncclSend");
  NCCL_SYM(Recv, "ncclRecv");
  NCCL_SYM(GroupStart, "ncclGroupStart");
  NCCL_SYM(GroupEnd, "ncclGroupEnd");
  NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
  g_nccl.lib = lib;
  return true;
}

#define NC(call) \
  do { \
    ncclResult_t r_ = (call); \
    if (r_ != ncclSuccess) return fail(ctx, DSC_ERR_NCCL, "%s: %s", #call, g_nccl.GetErrorString(r_)); \
  } while (0)

/* ---- partition: contiguous runs of leaves in traversal order.  With 2^k ranks and a tree at least
 * k deep these are the subtrees k levels below the root (a spatial cut); otherwise runs of equal
 * weight (unique verts + looptris). ---- */
static void plan_partition(const DscPbvhDesc *pb, int world, std::vector<int> &leaves, std::vector<int> &range)
{
  const int N = pb->totnode;
  leaves.clear();
  for (int n = 0; n < N; n++) {
    if (pb->flag[n] & DSC_PBVH_Leaf) leaves.push_back(n);
  }
  std::sort(leaves.begin(), leaves.end(), [&](int a, int b) { return pb->prim_offset[a] < pb->prim_offset[b]; });
  const int L = (int)leaves.size();
  range.assign(world + 1, L);
  range[0] = 0;
  if (world <= 1) return;
  std::vector<int> rank_of(N, -1);
  for (int l = 0; l < L; l++) rank_of[leaves[l]] = l;
  /* leaf span of every node: children have larger indices than their parent (pbvh.c:2387-2388) */
  std::vector<int> lo(N, L), hi(N, -1);
  for (int n = N - 1; n >= 0; n--) {
    if (pb->flag[n] & DSC_PBVH_Leaf) {
      lo[n] = hi[n] = rank_of[n];
    }
    else {
      const int c = pb->children_offset[n];
      lo[n] = std::min(lo[c], lo[c + 1]);
      hi[n] = std::max(hi[c], hi[c + 1]);
    }
  }
  bool done = false;
  if ((world & (world - 1)) == 0) {
    std::vector<int> front(1, 0);
    for (int w = 1; w < world; w *= 2) {
      std::vector<int> next;
      for (int n : front) {
        if (pb->flag[n] & DSC_PBVH_Leaf) next.push_back(n);
        else {
          next.push_back(pb->children_offset[n]);
          next.push_back(pb->children_offset[n] + 1);
        }
      }
      front.swap(next);
    }
    if ((int)front.size() == world) {
      std::sort(front.begin(), front.end(), [&](int a, int b) { return lo[a] < lo[b]; });
      for (int r = 0; r < world; r++) range[r] = lo[front[r]];
      range[world] = L;
      done = true;
    }
  }
  if (!done) {
    double total = 0;
    for (int l = 0; l < L; l++) total += pb->uniq_verts[leaves[l]] + pb->totprim[leaves[l]];
    double acc = 0;
    int r = 1;
    for (int l = 0; l < L && r < world; l++) {
      acc += pb->uniq_verts[leaves[l]] + pb->totprim[leaves[l]];
      if (acc >= total * r / world) range[r++] = l + 1;
    }
    for (; r < world; r++) range[r] = L;
  }
}

/* ---- halo plan in vertex ids: need[q] = vertices rank q reads but does not own: the shared and
 * extra verts of its leaves' polys, the verts of the halo polys around its unique verts, and (smooth)
 * the edge neighbours of its unique verts.  triples are (reader, owner, vertex), sorted, unique. ---- */
struct HaloTriple {
  int reader, owner, vert;
  bool operator<(const HaloTriple &o) const
  {
    return reader != o.reader ? reader < o.reader : owner != o.owner ? owner < o.owner : vert < o.vert;
  }
  bool operator==(const HaloTriple &o) const { return reader == o.reader && owner == o.owner && vert == o.vert; }
};


/* ---- partitioned grids: what a rank computes and what it must be sent ----
 * A rank owns the grids of its leaves.  After the brush it runs the averaging of every group of duplicated elements
 * that holds an element it owns -- the pairs along the boundaries between the grids of a face and the face centre
 * (subdiv_ccg.c:951-984), the points of a coarse edge (1010-1048), the corners at a coarse vertex (1081-1104) --
 * on copies of the other ranks' elements that the halo exchange has made current; every rank that shares a group
 * computes the same value from the same inputs, so nothing has to be sent back.  One dependency crosses phases:
 * the two middle points of a coarse edge are first averaged inside each face (they are the last pair of the boundary
 * between its two grids there), so a rank that averages one half of an edge also runs that pair for every face on
 * the edge.  The plan below is that closure, enumerated over the tables:
 *   face_dom[f]   bit 0: some grid of f is owned -- all its pairs and its centre run; bit 1 (alone): only middle pairs on
 *                 owned edges; bit 2: the face has an owned coarse vertex (it lists it when a dab reaches the face)
 *   edge_mine[e]  bit h: half h of the edge's 2 * grid_size points (one grid per face and half) holds an owned element
 *   cvert_mine[v] some corner at the coarse vertex is owned
 *   need          the elements of other ranks those groups read, plus the neighbours of owned rim elements the smooth
 *                 brush reads (rim table) */
struct GridPlan {
  std::vector<unsigned char> face_dom, edge_mine, cvert_mine;
  std::vector<int> need; /* ascending element indices, none owned by the rank */
};
struct GridTables {
  int totgrid, gs, totface, totedge, totcvert;
  const int *face_start, *face_num, *edge_off, *edge_elems, *cvert_off, *cvert_elems, *grid_edge, *grid_cvert;
  int rim_w;
  const int *rim_nb; /* or NULL */
};
static void plan_grids_rank(const GridTables &t, const std::vector<int> &grid_owner, int rank, GridPlan &pl)
{
  const int gs = t.gs, gs2 = gs * gs, rows2 = 2 * gs;
  auto owner_of = [&](int elem) { return grid_owner[elem / gs2]; };
  pl.face_dom.assign((size_t)t.totface, 0);
  pl.edge_mine.assign((size_t)t.totedge, 0);
  pl.cvert_mine.assign((size_t)t.totcvert, 0);
  std::vector<int> need;
  auto want = [&](int elem) {
    if (owner_of(elem) != rank) need.push_back(elem);
  };
  for (int f = 0; f < t.totface; f++) {
    bool any = false;
    for (int c = 0; c < t.face_num[f]; c++) any = any || grid_owner[t.face_start[f] + c] == rank;
    if (!any) continue;
    pl.face_dom[f] = 1;
    for (int c = 0; c < t.face_num[f]; c++) {
      const int g = t.face_start[f] + c;
      for (int i = 0; i < gs; i++) { /* row y == 0 and column x == 0: every pair of the face and its centre */
        want(g * gs2 + i);
        want(g * gs2 + i * gs);
      }
    }
  }
  for (int e = 0; e < t.totedge; e++) {
    const int nf = t.edge_off[e + 1] - t.edge_off[e];
    for (int h = 0; h < 2; h++) {
      for (int k = 0; k < nf; k++) {
        if (owner_of(t.edge_elems[(size_t)(t.edge_off[e] + k) * rows2 + h * gs]) == rank) pl.edge_mine[e] |= (unsigned char)(1 << h);
      }
    }
    if (!pl.edge_mine[e]) continue;
    for (int k = 0; k < nf; k++) {
      const int *row = t.edge_elems + (size_t)(t.edge_off[e] + k) * rows2;
      for (int h = 0; h < 2; h++) {
        if (!(pl.edge_mine[e] & (1 << h))) continue;
        for (int i = 0; i < gs; i++) want(row[h * gs + i]);
      }
      want(row[gs - 1]); /* the middle pair of every face on the edge */
      want(row[gs]);
    }
  }
  /* faces that only contribute middle pairs: a face with no owned grid on an edge with an owned half */
  for (int f = 0; f < t.totface; f++) {
    if (pl.face_dom[f]) continue;
    for (int c = 0; c < t.face_num[f]; c++) {
      if (pl.edge_mine[t.grid_edge[t.face_start[f] + c]]) pl.face_dom[f] = 2;
    }
  }
  for (int v = 0; v < t.totcvert; v++) {
    bool any = false;
    for (int k = t.cvert_off[v]; k < t.cvert_off[v + 1]; k++) any = any || owner_of(t.cvert_elems[k]) == rank;
    if (!any) continue;
    pl.cvert_mine[v] = 1;
    for (int k = t.cvert_off[v]; k < t.cvert_off[v + 1]; k++) want(t.cvert_elems[k]);
  }
  /* a face none of whose grids or edges is ours still lists our coarse vertices when a dab reaches it */
  for (int f = 0; f < t.totface; f++) {
    for (int c = 0; c < t.face_num[f]; c++) {
      if (pl.cvert_mine[t.grid_cvert[t.face_start[f] + c]]) pl.face_dom[f] |= 4;
    }
  }
  if (t.rim_nb) {
    const int rim = 4 * gs - 4;
    for (int g = 0; g < t.totgrid; g++) {
      if (grid_owner[g] != rank) continue;
      const int *rows = t.rim_nb + (size_t)g * rim * t.rim_w;
      for (int i = 0; i < rim * t.rim_w; i++) {
        if (rows[i] >= 0) want(rows[i]);
      }
    }
  }
  std::sort(need.begin(), need.end());
  need.erase(std::unique(need.begin(), need.end()), need.end());
  pl.need.swap(need);
}
/* owner rank of every grid: the rank of the leaf that holds it */
static void plan_grid_owner(const DscPbvhDesc *pb, int world, int totgrid, std::vector<int> &leaves, std::vector<int> &range,
                            std::vector<int> &grid_owner)
{
  plan_partition(pb, world, leaves, range);
  grid_owner.assign((size_t)totgrid, 0);
  for (int r = 0; r < world; r++) {
    for (int l = range[r]; l < range[r + 1]; l++) {
      const int n = leaves[l];
      for (int k = 0; k < pb->totprim[n]; k++) grid_owner[pb->prim_indices[pb->prim_offset[n] + k]] = r;
    }
  }
}
/* send / receive lists of one rank: per peer, ascending element indices (both sides enumerate the same order) */
static void plan_grids_lists(const GridTables &t, const std::vector<int> &grid_owner, int world, int rank, GridPlan &mine,
                             std::vector<int> &send_off, std::vector<int> &send_elem, std::vector<int> &recv_off,
                             std::vector<int> &recv_elem, std::vector<unsigned char> *halo_grid = nullptr)
{
  const int gs2 = t.gs * t.gs;
  if (halo_grid) halo_grid->assign((size_t)t.totgrid, 0);
  send_off.assign(world + 1, 0);
  recv_off.assign(world + 1, 0);
  send_elem.clear();
  recv_elem.clear();
  plan_grids_rank(t, grid_owner, rank, mine);
  if (halo_grid) {
    for (int e : mine.need) (*halo_grid)[e / gs2] = 1;
  }
  for (int q = 0; q < world; q++) {
    send_off[q] = (int)send_elem.size();
    recv_off[q] = (int)recv_elem.size();
    if (q == rank) continue;
    for (int e : mine.need) {
      if (grid_owner[e / gs2] == q) recv_elem.push_back(e);
    }
    GridPlan theirs;
    plan_grids_rank(t, grid_owner, q, theirs);
    for (int e : theirs.need) {
      if (grid_owner[e / gs2] == rank) send_elem.push_back(e);
      if (halo_grid) (*halo_grid)[e / gs2] = 1;
    }
  }
  send_off[world] = (int)send_elem.size();
  recv_off[world] = (int)recv_elem.size();
}

static void plan_halo(const DscMeshDesc *me, const DscPbvhDesc *pb, int world, const std::vector<int> &leaves,
                      const std::vector<int> &range, std::vector<HaloTriple> &out)
{
  const int V = me->totvert, L = (int)leaves.size();
  std::vector<int> vowner(V, -1), leaf_rank(L, 0);
  for (int r = 0; r < world; r++) {
    for (int l = range[r]; l < range[r + 1]; l++) leaf_rank[l] = r;
  }
  for (int l = 0; l < L; l++) {
    const int n = leaves[l];
    const int *vi = pb->vert_indices + pb->vert_offset[n];
    for (int i = 0; i < pb->uniq_verts[n]; i++) vowner[vi[i]] = leaf_rank[l];
  }
  out.clear();
  for (int l = 0; l < L; l++) {
    const int n = leaves[l], q = leaf_rank[l];
    for (int pos = pb->prim_offset[n]; pos < pb->prim_offset[n] + pb->totprim[n]; pos++) {
      const int t = pb->prim_indices[pos];
      const int p = me->tri_poly[t];
      const int ls = me->poly_loopstart[p], len = me->poly_totloop[p];
      /* the leaf's own polys */
      for (int k = 0; k < len; k++) {
        const int u = me->loop_vert[ls + k];
        if (vowner[u] != q && vowner[u] >= 0) out.push_back({q, vowner[u], u});
      }
      /* this looptri is a halo looptri of the owners of its foreign verts */
      for (int j = 0; j < 3; j++) {
        const int v = me->tri_vert[(size_t)3 * t + j];
        const int q2 = vowner[v];
        if (q2 == q || q2 < 0) continue;
        for (int k = 0; k < len; k++) {
          const int u = me->loop_vert[ls + k];
          if (vowner[u] != q2 && vowner[u] >= 0) out.push_back({q2, vowner[u], u});
        }
      }
    }
  }
  if (me->nb_offsets && me->nb_indices) {
    for (int v = 0; v < V; v++) {
      for (int k = me->nb_offsets[v]; k < me->nb_offsets[v + 1]; k++) {
        const int u = me->nb_indices[k];
        if (vowner[u] != vowner[v] && vowner[u] >= 0 && vowner[v] >= 0) out.push_back({vowner[v], vowner[u], u});
      }
    }
  }
  std::sort(out.begin(), out.end());
  out.erase(std::unique(out.begin(), out.end()), out.end());
}

/* ---- kernel durations of the replayed graphs: CUPTI activity records (dsc_stage_timing mode 3) ----
 * Event pairs around a kernel measure launch latency with it (about 8 us per small kernel inside a graph), so the
 * per-stage times of the timed path are taken from the hardware timestamps CUPTI records for every kernel node.
 * libcupti is looked up at run time; when it cannot be loaded (or a profiler already owns the device) mode 3 is refused
 * and the caller falls back to mode 2. */
struct CuptiApi {
  void *lib = nullptr;
  CUptiResult (*Enable)(CUpti_ActivityKind) = nullptr;
  CUptiResult (*Disable)(CUpti_ActivityKind) = nullptr;
  CUptiResult (*Register)(CUpti_BuffersCallbackRequestFunc, CUpti_BuffersCallbackCompleteFunc) = nullptr;
  CUptiResult (*FlushAll)(uint32_t) = nullptr;
  CUptiResult (*Next)(uint8_t *, size_t, CUpti_Activity **) = nullptr;
  bool registered = false;
};
static CuptiApi g_cupti;
static std::mutex g_cupti_mu;
static double g_cupti_ms[DSC_NUM_STAGES];
static long long g_cupti_n[DSC_NUM_STAGES];

static int stage_of_kernel(const char *name)
{
  if (!name) return ST_OTHER;
  if (strstr(name, "k_gather")) return ST_GATHER;
  if (strstr(name, "k_area")) return ST_AREA;
  if (strstr(name, "k_brush")) return ST_BRUSH;
  if (strstr(name, "k_smooth") || strstr(name, "k_snapshot")) return ST_SMOOTH;
  if (strstr(name, "k_leaf_bb")) return ST_LEAFBB;
  if (strstr(name, "k_normals") || strstr(name, "k_grid_") || strstr(name, "k_dab_tile") || strstr(name, "k_ghit")) return ST_NORMALS;
  if (strstr(name, "k_tag_ancestors") || strstr(name, "k_refit") || strstr(name, "k_flush") || strstr(name, "k_reset_leaf_boxes")) return ST_FLUSH;
  return ST_OTHER;
}
static void CUPTIAPI cupti_buffer_requested(uint8_t **buffer, size_t *size, size_t *max_records)
{
  *size = 8u << 20;
  *buffer = (uint8_t *)aligned_alloc(8, *size);
  *max_records = 0;
}
static void CUPTIAPI cupti_buffer_completed(CUcontext, uint32_t, uint8_t *buffer, size_t, size_t valid)
{
  CUpti_Activity *rec = nullptr;
  std::lock_guard<std::mutex> lk(g_cupti_mu);
  while (valid > 0 && g_cupti.Next(buffer, valid, &rec) == CUPTI_SUCCESS) {
    if (rec->kind == CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL || rec->kind == CUPTI_ACTIVITY_KIND_KERNEL) {
      const CUpti_ActivityKernel9 *k = (const CUpti_ActivityKernel9 *)rec;
      const int st = stage_of_kernel(k->name);
      g_cupti_ms[st] += 1e-6 * (double)(k->end - k->start);
      g_cupti_n[st]++;
    }
  }
  free(buffer);
}
static bool cupti_load()
{
  if (g_cupti.lib) return true;
  const char *names[] = {"/usr/local/cuda/lib64/libcupti.so.12", "/usr/local/cuda/targets/x86_64-linux/lib/libcupti.so.12",
                         "/usr/local/cuda/extras/CUPTI/lib64/libcupti.so.12", "libcupti.so.12", "libcupti.so"};
  void *h = nullptr;
  for (const char *n : names) {
    if ((h = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
  }
  if (!h) return false;
  g_cupti.Enable = (decltype(g_cupti.Enable))dlsym(h, "cuptiActivityEnable");
  g_cupti.Disable = (decltype(g_cupti.Disable))dlsym(h, "cuptiActivityDisable");
  g_cupti.Register = (decltype(g_cupti.Register))dlsym(h, "cuptiActivityRegisterCallbacks");
  g_cupti.FlushAll = (decltype(g_cupti.FlushAll))dlsym(h, "cuptiActivityFlushAll");
  g_cupti.Next = (decltype(g_cupti.Next))dlsym(h, "cuptiActivityGetNextRecord");
  if (!g_cupti.Enable || !g_cupti.Disable || !g_cupti.Register || !g_cupti.FlushAll || !g_cupti.Next) {
    dlclose(h);
    return false;
  }
  g_cupti.lib = h;
  return true;
}
static bool cupti_start()
{
  if (!cupti_load()) return false;
  if (!g_cupti.registered) {
    if (g_cupti.Register(cupti_buffer_requested, cupti_buffer_completed) != CUPTI_SUCCESS) return false;
    g_cupti.registered = true;
  }
  {
    std::lock_guard<std::mutex> lk(g_cupti_mu);
    for (int i = 0; i < DSC_NUM_STAGES; i++) {
      g_cupti_ms[i] = 0.0;
      g_cupti_n[i] = 0;
    }
  }
  return g_cupti.Enable(CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL) == CUPTI_SUCCESS;
}
static void cupti_stop()
{
  if (g_cupti.lib) {
    g_cupti.FlushAll(1);
    g_cupti.Disable(CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL);
  }
}

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
  ctx->grid_area = ctx->grid_brush = ctx->grid;
  memset(&ctx->m, 0, sizeof(ctx->m));
  bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->ev_bb, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->ev_tag, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreate(&ctx->t0) == cudaSuccess && cudaEventCreate(&ctx->t1) == cudaSuccess &&
            cudaMallocHost((void **)&ctx->h_state, sizeof(DabState)) == cudaSuccess &&
            cudaMallocHost((void **)&ctx->h_tot, sizeof(StrokeTotals)) == cudaSuccess &&
            cudaMallocHost((void **)&ctx->h_ring, sizeof(DabEntry) * DSC_RING) == cudaSuccess &&
            cudaMalloc((void **)&ctx->d_ring, sizeof(DabEntry) * DSC_RING) == cudaSuccess &&
            cudaMalloc((void **)&ctx->d_ring_ctl, sizeof(int) * 2) == cudaSuccess &&
            cudaMemset(ctx->d_ring_ctl, 0, sizeof(int) * 2) == cudaSuccess;
  for (int i = 0; ok && i < 4; i++) ok = cudaEventCreateWithFlags(&ctx->ev_ring[i], cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; ok && i < DSC_SLOTS; i++) ok = cudaEventCreateWithFlags(&ctx->cap_refit[i], cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&ctx->cap_fork, cudaEventDisableTiming) == cudaSuccess &&
       cudaEventCreateWithFlags(&ctx->cap_bb, cudaEventDisableTiming) == cudaSuccess &&
       cudaEventCreateWithFlags(&ctx->cap_tag, cudaEventDisableTiming) == cudaSuccess &&
       cudaEventCreateWithFlags(&ctx->cap_join, cudaEventDisableTiming) == cudaSuccess;
  if (ok) {
    ctx->m.ring = ctx->d_ring;
    ctx->m.ring_ctl = ctx->d_ring_ctl;
    ctx->use_graphs = !getenv("DSC_NO_GRAPHS");
    if (getenv("DSC_GRID_AREA")) ctx->grid_area = ctx->num_sms * std::max(1, atoi(getenv("DSC_GRID_AREA")));
    if (getenv("DSC_GRID_BRUSH")) ctx->grid_brush = ctx->num_sms * std::max(1, atoi(getenv("DSC_GRID_BRUSH")));
    /* measured slower than the graph replay (grid barriers cost more than the launch gaps they replace, and the
     * area / brush stages run at the tile kernel's lower occupancy): opt-in */
    ctx->use_batch_kernel = getenv("DSC_BATCH_KERNEL") != nullptr;
    ctx->lazy_refit = getenv("DSC_EAGER_REFIT") == nullptr;
    ctx->small_dab_frac = getenv("DSC_SMALL_DAB_FRAC") ? (float)atof(getenv("DSC_SMALL_DAB_FRAC")) : 1.0e9f;
    if (getenv("DSC_BATCH_GRID")) ctx->small_dab_grid = std::max(1, atoi(getenv("DSC_BATCH_GRID")));
    ctx->use_pdl = getenv("DSC_PDL") != nullptr; /* measured: no gain on small dabs, a loss on large ones (early CTAs hold SM slots) */
  }
  for (int i = 0; ok && i < DSC_SLOTS; i++) ok = cudaEventCreateWithFlags(&ctx->ev_refit[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
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
  cudaStreamSynchronize(ctx->stream2);
  cudaStreamSynchronize(ctx->stream);
  for (int q = 0; q < DSC_MAX_RANKS; q++) {
    if (ctx->p2p_peer_region[q]) cudaIpcCloseMemHandle(ctx->p2p_peer_region[q]);
  }
  if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
  invalidate_graphs(ctx);
  for (void *p : ctx->allocs) cudaFree(p);
  if (ctx->d_ring) cudaFree(ctx->d_ring);
  if (ctx->d_ring_ctl) cudaFree(ctx->d_ring_ctl);
  if (ctx->h_ring) cudaFreeHost(ctx->h_ring);
  for (int i = 0; i < 4; i++) {
    if (ctx->ev_ring[i]) cudaEventDestroy(ctx->ev_ring[i]);
    if (ctx->cap_refit[i]) cudaEventDestroy(ctx->cap_refit[i]);
  }
  if (ctx->cap_fork) cudaEventDestroy(ctx->cap_fork);
  if (ctx->cap_bb) cudaEventDestroy(ctx->cap_bb);
  if (ctx->cap_tag) cudaEventDestroy(ctx->cap_tag);
  if (ctx->cap_join) cudaEventDestroy(ctx->cap_join);
  for (auto &ev : ctx->events) {
    cudaEventDestroy(ev.a);
    cudaEventDestroy(ev.b);
  }
  if (ctx->h_state) cudaFreeHost(ctx->h_state);
  if (ctx->h_tot) cudaFreeHost(ctx->h_tot);
  if (ctx->h_list) cudaFreeHost(ctx->h_list);
  if (ctx->h_own) cudaFreeHost(ctx->h_own);
  for (void *p : ctx->registered) cudaHostUnregister(p);
  if (ctx->h_ray_out) cudaFreeHost(ctx->h_ray_out);
  if (ctx->d_gray_out) cudaFree(ctx->d_gray_out);
  if (ctx->h_ray_count) cudaFreeHost(ctx->h_ray_count);
  cudaEventDestroy(ctx->t0);
  cudaEventDestroy(ctx->t1);
  cudaEventDestroy(ctx->ev_fork);
  cudaEventDestroy(ctx->ev_bb);
  cudaEventDestroy(ctx->ev_tag);
  for (int i = 0; i < DSC_SLOTS; i++) cudaEventDestroy(ctx->ev_refit[i]);
  cudaStreamDestroy(ctx->stream2);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

void *dsc_stream(DscContext *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int dsc_dist_unique_id(char id[DSC_NCCL_ID_BYTES])
{
  DscContext *ctx = nullptr;
  std::string err;
  if (!nccl_load(err)) return fail(nullptr, DSC_ERR_NCCL, "%s", err.c_str());
  static_assert(sizeof(ncclUniqueId) <= DSC_NCCL_ID_BYTES, "id fits");
  ncclUniqueId uid;
  NC(g_nccl.GetUniqueId(&uid));
  memset(id, 0, DSC_NCCL_ID_BYTES);
  memcpy(id, &uid, sizeof(uid));
  return DSC_OK;
}

int dsc_dist_init(DscContext *ctx, int world, int rank, const char id[DSC_NCCL_ID_BYTES])
{
  if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return fail(ctx, DSC_ERR_INVALID, "bad world / rank");
  if (ctx->have_pbvh) return fail(ctx, DSC_ERR_STATE, "dsc_dist_init must precede dsc_pbvh_upload");
  ctx->world = world;
  ctx->rank = rank;
  if (world == 1) return DSC_OK;
  std::string err;
  if (!nccl_load(err)) return fail(ctx, DSC_ERR_NCCL, "%s", err.c_str());
  CU(cudaSetDevice(ctx->device));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  NC(g_nccl.CommInitRank(&ctx->comm, world, uid, rank));
  return DSC_OK;
}

int dsc_dist_partition(const DscPbvhDesc *pb, int world, int *r_leaf_range, int *r_owner)
{
  if (!pb || world < 1) return DSC_ERR_INVALID;
  std::vector<int> leaves, range;
  plan_partition(pb, world, leaves, range);
  if (r_leaf_range) memcpy(r_leaf_range, range.data(), sizeof(int) * (size_t)(world + 1));
  if (r_owner) {
    for (int n = 0; n < pb->totnode; n++) r_owner[n] = -1;
    for (int r = 0; r < world; r++) {
      for (int l = range[r]; l < range[r + 1]; l++) r_owner[leaves[l]] = r;
    }
  }
  return DSC_OK;
}

int dsc_dist_halo_plan(const DscMeshDesc *me, const DscPbvhDesc *pb, int world, int rank, int **r_send_off, int **r_send_vert,
                       int **r_recv_off, int **r_recv_vert)
{
  if (!me || !pb || world < 1 || rank < 0 || rank >= world) return DSC_ERR_INVALID;
  std::vector<int> leaves, range;
  plan_partition(pb, world, leaves, range);
  std::vector<HaloTriple> tr;
  plan_halo(me, pb, world, leaves, range, tr);
  std::vector<int> soff(world + 1, 0), roff(world + 1, 0), sv, rv;
  for (int p = 0; p < world; p++) {
    soff[p] = (int)sv.size();
    roff[p] = (int)rv.size();
    for (const HaloTriple &t : tr) {
      if (t.owner == rank && t.reader == p) sv.push_back(t.vert);
      if (t.reader == rank && t.owner == p) rv.push_back(t.vert);
    }
  }
  soff[world] = (int)sv.size();
  roff[world] = (int)rv.size();
  auto dup = [](const std::vector<int> &v) {
    int *p = (int *)malloc(sizeof(int) * std::max<size_t>(v.size(), 1));
    if (!v.empty()) memcpy(p, v.data(), sizeof(int) * v.size());
    return p;
  };
  if (r_send_off) *r_send_off = dup(soff);
  if (r_send_vert) *r_send_vert = dup(sv);
  if (r_recv_off) *r_recv_off = dup(roff);
  if (r_recv_vert) *r_recv_vert = dup(rv);
  return DSC_OK;
}


static GridTables grid_tables_of(const DscGridsDesc *gr)
{
  GridTables t;
  t.totgrid = gr->totgrid; t.gs = gr->grid_size; t.totface = gr->totface; t.totedge = gr->totedge; t.totcvert = gr->totcvert;
  t.face_start = gr->face_start_grid; t.face_num = gr->face_num_grids; t.edge_off = gr->edge_offsets; t.edge_elems = gr->edge_elems;
  t.cvert_off = gr->cvert_offsets; t.cvert_elems = gr->cvert_elems; t.grid_edge = gr->grid_edge; t.grid_cvert = gr->grid_cvert;
  t.rim_w = gr->rim_width; t.rim_nb = gr->rim_neighbors;
  return t;
}

int dsc_dist_grids_plan(const DscGridsDesc *gr, const DscPbvhDesc *pb, int world, int rank, int *r_grid_owner,
                        unsigned char *r_face_dom, unsigned char *r_edge_mine, unsigned char *r_cvert_mine, int **r_send_off,
                        int **r_send_elem, int **r_recv_off, int **r_recv_elem)
{
  if (!gr || !pb || world < 1 || rank < 0 || rank >= world) return DSC_ERR_INVALID;
  std::vector<int> leaves, range, owner, soff, se, roff, re;
  plan_grid_owner(pb, world, gr->totgrid, leaves, range, owner);
  GridPlan pl;
  plan_grids_lists(grid_tables_of(gr), owner, world, rank, pl, soff, se, roff, re);
  if (r_grid_owner) memcpy(r_grid_owner, owner.data(), sizeof(int) * owner.size());
  if (r_face_dom) memcpy(r_face_dom, pl.face_dom.data(), pl.face_dom.size());
  if (r_edge_mine) memcpy(r_edge_mine, pl.edge_mine.data(), pl.edge_mine.size());
  if (r_cvert_mine) memcpy(r_cvert_mine, pl.cvert_mine.data(), pl.cvert_mine.size());
  auto dup = [](const std::vector<int> &v) {
    int *p = (int *)malloc(sizeof(int) * std::max<size_t>(v.size(), 1));
    if (!v.empty()) memcpy(p, v.data(), sizeof(int) * v.size());
    return p;
  };
  if (r_send_off) *r_send_off = dup(soff);
  if (r_send_elem) *r_send_elem = dup(se);
  if (r_recv_off) *r_recv_off = dup(roff);
  if (r_recv_elem) *r_recv_elem = dup(re);
  return DSC_OK;
}

void dsc_dist_free(void *p) { free(p); }

int dsc_dist_uses_peer_memory(DscContext *ctx) { return ctx && ctx->p2p ? 1 : 0; }
int dsc_dist_exchanges_skipped(DscContext *ctx, int *r_skipped)
{
  if (!ctx || !r_skipped) return DSC_ERR_INVALID;
  *r_skipped = 0;
  if (!ctx->p2p) return DSC_OK;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(r_skipped, ctx->link.flags + P2P_SKIPPED, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return DSC_OK;
}

int dsc_dist_owned_range(DscContext *ctx, int r_range[2])
{
  if (!ctx || !ctx->have_pbvh) return DSC_ERR_STATE;
  r_range[0] = ctx->m.own_lo;
  r_range[1] = ctx->m.own_hi;
  return DSC_OK;
}

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
  if (me->vert_tail) ctx->h_tail.assign(me->vert_tail, me->vert_tail + me->totvert);
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

int dsc_grids_upload(DscContext *ctx, const DscGridsDesc *gr)
{
  if (!ctx || !gr) return fail(ctx, DSC_ERR_INVALID, "NULL argument");
  if (ctx->have_mesh) return fail(ctx, DSC_ERR_STATE, "a mesh is already resident in this context");
  if (gr->totgrid <= 0 || gr->grid_size < 2 || !gr->co) return fail(ctx, DSC_ERR_INVALID, "empty grid set");
  if (!gr->face_start_grid || !gr->face_num_grids || !gr->edge_offsets || !gr->edge_elems || !gr->cvert_offsets ||
      !gr->cvert_elems || !gr->grid_edge || !gr->grid_cvert)
    return fail(ctx, DSC_ERR_INVALID, "grid adjacency tables missing");
  const long long E = (long long)gr->totgrid * gr->grid_size * gr->grid_size;
  if (E > 0x7fff0000ll) return fail(ctx, DSC_ERR_UNSUPPORTED, "more than 2^31 grid elements");
  if (dsc_grid_normals_smem(gr->grid_size) > 200 * 1024)
    return fail(ctx, DSC_ERR_UNSUPPORTED, "grid size %d: one grid does not fit the normal kernel's shared memory", gr->grid_size);
  ctx->is_grids = true;
  ctx->grid_size = gr->grid_size;
  ctx->totgrid = gr->totgrid;
  ctx->totvert = (int)E;
  ctx->h_co.assign(gr->co, gr->co + (size_t)3 * E);
  ctx->has_no = gr->no != nullptr;
  if (gr->no) ctx->h_no.assign(gr->no, gr->no + (size_t)3 * E);
  ctx->has_mask = gr->mask != nullptr;
  if (gr->mask) ctx->h_mask.assign(gr->mask, gr->mask + E);
  ctx->h_face_start.assign(gr->face_start_grid, gr->face_start_grid + gr->totface);
  ctx->h_face_num.assign(gr->face_num_grids, gr->face_num_grids + gr->totface);
  ctx->h_edge_off.assign(gr->edge_offsets, gr->edge_offsets + gr->totedge + 1);
  ctx->h_edge_elems.assign(gr->edge_elems, gr->edge_elems + (size_t)gr->edge_offsets[gr->totedge] * 2 * gr->grid_size);
  ctx->h_cvert_off.assign(gr->cvert_offsets, gr->cvert_offsets + gr->totcvert + 1);
  ctx->h_cvert_elems.assign(gr->cvert_elems, gr->cvert_elems + gr->cvert_offsets[gr->totcvert]);
  ctx->h_grid_edge.assign(gr->grid_edge, gr->grid_edge + gr->totgrid);
  ctx->h_grid_cvert.assign(gr->grid_cvert, gr->grid_cvert + gr->totgrid);
  ctx->tottri = gr->totgrid; /* the PBVH's prims */
  if (gr->hidden) ctx->h_elem_hidden.assign(gr->hidden, gr->hidden + (size_t)gr->totgrid * gr->grid_size * gr->grid_size);
  ctx->has_nb = gr->rim_neighbors != nullptr;
  if (ctx->has_nb) {
    if (gr->rim_width < 4) return fail(ctx, DSC_ERR_INVALID, "rim_width %d: a rim element has up to four neighbours or more", gr->rim_width);
    const size_t rows = (size_t)gr->totgrid * (size_t)(4 * gr->grid_size - 4);
    ctx->rim_width = gr->rim_width;
    ctx->h_rim_nb.assign(gr->rim_neighbors, gr->rim_neighbors + rows * (size_t)gr->rim_width);
    if (gr->rim_boundary) ctx->h_rim_bnd.assign(gr->rim_boundary, gr->rim_boundary + rows);
  }
  ctx->have_mesh = true;
  return DSC_OK;
}

static int dist_p2p_setup(DscContext *ctx);

/* DSC_TIMING=1: where the session start goes, section by section, on stderr */
struct UploadTimer {
  bool on = getenv("DSC_TIMING") != nullptr;
  double t0 = now();
  static double now()
  {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
  }
  void mark(const char *what)
  {
    if (!on) return;
    const double t = now();
    fprintf(stderr, "[dsc upload] %-24s %.3f s\n", what, t - t0);
    t0 = t;
  }
};
#define DSC_TMARK(what) upload_timer.mark(what)

int dsc_pbvh_upload(DscContext *ctx, const DscPbvhDesc *pb)
{
  UploadTimer upload_timer;
  if (!ctx || !pb) return fail(ctx, DSC_ERR_INVALID, "NULL argument");
  if (!ctx->have_mesh) return fail(ctx, DSC_ERR_STATE, "dsc_mesh_upload must come first");
  if (ctx->have_pbvh) return fail(ctx, DSC_ERR_STATE, "a PBVH is already resident");
  if (pb->totnode <= 0) return fail(ctx, DSC_ERR_INVALID, "empty PBVH");
  CU(cudaSetDevice(ctx->device));
  const int V = ctx->totvert, T = ctx->tottri, N = pb->totnode;
  ctx->totnode = N;
  DevMesh &m = ctx->m;
  int r;

  /* leaves in traversal order = ascending prim offset */
  std::vector<int> leaves;
  for (int n = 0; n < N; n++) {
    if (pb->flag[n] & DSC_PBVH_Leaf) leaves.push_back(n);
  }
  std::sort(leaves.begin(), leaves.end(), [&](int a, int b) { return pb->prim_offset[a] < pb->prim_offset[b]; });
  const int L = (int)leaves.size();
  ctx->leaf_node = leaves;
  {
    /* the root box's diagonal scales the small-dab threshold */
    double dd = 0.0;
    for (int k = 0; k < 3; k++) {
      const double e = (double)pb->node_bb[3 + k] - (double)pb->node_bb[k];
      dd += e * e;
    }
    ctx->small_dab_radius = ctx->small_dab_frac >= 1.0e8f ? 3.0e38f : ctx->small_dab_frac * (float)sqrt(dd);
  }

  DSC_TMARK("start");
  /* slots: each leaf's unique verts are one 128-byte aligned run, cut into tiles of <= DSC_TILE
   * slots.  Inside a leaf the order is ours to choose (the host translates through slot_of), so the
   * verts are split by recursive coordinate bisection into spatially compact tiles and ordered in
   * rows inside a tile: a tile's polys then reach few verts of other tiles, and those sit in runs. */
  ctx->slot_of.assign(V, -1);
  std::vector<int> leaf_ubeg(L), leaf_ucnt(L), leaf_scnt(L), leaf_pbeg(L), leaf_pcnt(L), leaf_tile0(L + 1, 0);
  std::vector<int2> tile_range;
  std::vector<int> tile_leaf;
  std::vector<int> tile_ibnd; /* per tile: where its boundary run starts (a multiple of 32); meshes only */
  long long cur = 0;
  int expect_prim = 0;
  std::vector<int> grid_slot0;
  if (ctx->is_grids) {
    /* a leaf's elements: its grids in prim order, each grid's grid_size^2 elements in (y, x) order
     * (the order pbvh_vertex_iter walks them, pbvh.c:4840-4897); tiles = runs of DSC_TILE slots */
    const int gs2 = ctx->grid_size * ctx->grid_size;
    grid_slot0.assign((size_t)ctx->totgrid, -1);
    for (int l = 0; l < L; l++) {
      const int n = leaves[l];
      cur = (cur + 31) & ~31ll;
      leaf_ubeg[l] = (int)cur;
      leaf_pbeg[l] = pb->prim_offset[n];
      leaf_pcnt[l] = pb->totprim[n];
      leaf_ucnt[l] = leaf_pcnt[l] * gs2;
      leaf_scnt[l] = 0;
      if (pb->uniq_verts[n] != leaf_ucnt[l] || pb->face_verts[n] != 0)
        return fail(ctx, DSC_ERR_INVALID, "grid leaf %d: uniq_verts must be totprim * grid_size^2, face_verts 0", n);
      if (leaf_pbeg[l] != expect_prim) return fail(ctx, DSC_ERR_INVALID, "leaf prim ranges do not tile prim_indices");
      expect_prim += leaf_pcnt[l];
      for (int k = 0; k < leaf_pcnt[l]; k++) {
        const int g = pb->prim_indices[leaf_pbeg[l] + k];
        if (g < 0 || g >= ctx->totgrid || grid_slot0[g] != -1) return fail(ctx, DSC_ERR_INVALID, "grid %d is not in exactly one leaf", g);
        grid_slot0[g] = (int)cur + k * gs2;
        for (int j = 0; j < gs2; j++) ctx->slot_of[(size_t)g * gs2 + j] = (int)cur + k * gs2 + j;
      }
      leaf_tile0[l] = (int)tile_range.size();
      for (int o = 0; o < leaf_ucnt[l]; o += DSC_TILE) {
        tile_range.push_back(make_int2((int)cur + o, std::min(DSC_TILE, leaf_ucnt[l] - o)));
        tile_leaf.push_back(l);
      }
      cur += leaf_ucnt[l];
      if (cur > 0x7fffff00ll) return fail(ctx, DSC_ERR_UNSUPPORTED, "more than 2^31 slots");
    }
    leaf_tile0[L] = (int)tile_range.size();
  }
  else {
    /* Pass 1: tiles.  Every leaf's unique verts are bisected into tiles (membership only).
     * Pass 2: a vertex is a BOUNDARY vertex when some other tile reads it -- it is a corner of a poly that has a
     *         corner in another tile, or of a looptri held by another leaf (that leaf lists it as a shared vert).
     * Pass 3: slots.  Inside a tile the interior verts come first, the boundary verts last (each part in rows); the
     *         boundary part starts at a multiple of 32 (`ibnd`, a few interior verts may fall into it).  The dab then
     *         displaces the boundary runs of the gathered tiles first (k_brush_boundary) and the interiors inside the
     *         fused tile kernel, which finds every vertex it reads from another tile already displaced. */
    struct TileSpan { int lo, hi, leaf, base; };
    const float *hco = ctx->h_co.data();
    /* serial, cheap: the leaves' runs, the claim of every unique vertex, where each leaf's verts / tiles go */
    std::vector<long long> ord0((size_t)L + 1, 0);
    for (int l = 0; l < L; l++) {
      const int n = leaves[l];
      cur = (cur + 31) & ~31ll;
      leaf_ubeg[l] = (int)cur;
      leaf_ucnt[l] = pb->uniq_verts[n];
      leaf_scnt[l] = pb->face_verts[n];
      leaf_pbeg[l] = pb->prim_offset[n];
      leaf_pcnt[l] = pb->totprim[n];
      if (leaf_pbeg[l] != expect_prim) return fail(ctx, DSC_ERR_INVALID, "leaf prim ranges do not tile prim_indices");
      expect_prim += leaf_pcnt[l];
      const int *vi = pb->vert_indices + pb->vert_offset[n];
      const int U = pb->uniq_verts[n];
      for (int i = 0; i < U; i++) {
        const int v = vi[i];
        if (v < 0 || v >= V || ctx->slot_of[v] != -1) return fail(ctx, DSC_ERR_INVALID, "vertex %d is not unique in exactly one leaf", v);
        ctx->slot_of[v] = -2; /* claimed; the slot follows */
      }
      ord0[l + 1] = ord0[l] + U;
      leaf_tile0[l + 1] = leaf_tile0[l] + std::max(1, (U + DSC_TILE - 1) / DSC_TILE); /* the bisection makes exactly that many */
      cur += U;
      if (cur > 0x7fffff00ll) return fail(ctx, DSC_ERR_UNSUPPORTED, "more than 2^31 slots");
    }
    std::vector<int> ord_all((size_t)ord0[L]);
    std::vector<TileSpan> spans((size_t)leaf_tile0[L]);
    std::vector<int> tile_of_vert((size_t)V, -1);
    /* pass 1, leaf by leaf (independent): bisection of the leaf's unique verts into tiles */
#pragma omp parallel
    {
      std::vector<int> ord;
      std::vector<TileSpan> made;
      struct Job { int lo, hi, k, base; };
      std::vector<Job> jobs;
#pragma omp for schedule(dynamic, 8)
      for (int l = 0; l < L; l++) {
        const int n = leaves[l];
        const int *vi = pb->vert_indices + pb->vert_offset[n];
        const int U = leaf_ucnt[l];
        ord.assign(vi, vi + U);
        /* bisection of ord[lo, hi) into k tiles, slots from `base` */
        jobs.assign(1, Job{0, U, std::max(1, (U + DSC_TILE - 1) / DSC_TILE), leaf_ubeg[l]});
        made.clear();
        const int g0 = (int)ord0[l];
        while (!jobs.empty()) {
          const Job j = jobs.back();
          jobs.pop_back();
          const int cnt = j.hi - j.lo;
          if (j.k > 1) {
            float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (int i = j.lo; i < j.hi; i++) {
              for (int k = 0; k < 3; k++) {
                const float c = hco[(size_t)3 * ord[i] + k];
                mn[k] = std::min(mn[k], c);
                mx[k] = std::max(mx[k], c);
              }
            }
            int a = 0;
            for (int k = 1; k < 3; k++) {
              if ((mx[k] - mn[k]) > (mx[a] - mn[a])) a = k;
            }
            const int kl = j.k / 2;
            long long nl = ((long long)cnt * kl + j.k - 1) / j.k;
            nl = std::min<long long>((nl + 31) & ~31ll, std::min<long long>((long long)kl * DSC_TILE, cnt));
            std::nth_element(ord.begin() + j.lo, ord.begin() + j.lo + nl, ord.begin() + j.hi, [&](int p, int q) {
              const float cp = hco[(size_t)3 * p + a], cq = hco[(size_t)3 * q + a];
              return cp < cq || (cp == cq && p < q);
            });
            /* right half first on the stack so tiles come out in slot order */
            jobs.push_back(Job{j.lo + (int)nl, j.hi, j.k - kl, j.base + (int)nl});
            jobs.push_back(Job{j.lo, j.lo + (int)nl, kl, j.base});
            continue;
          }
          made.push_back(TileSpan{g0 + j.lo, g0 + j.hi, l, j.base});
        }
        std::sort(made.begin(), made.end(), [](const TileSpan &x, const TileSpan &y) { return x.base < y.base; });
        std::copy(ord.begin(), ord.end(), ord_all.begin() + g0);
        for (size_t k = 0; k < made.size(); k++) {
          const int ti = leaf_tile0[l] + (int)k;
          for (int i = made[k].lo; i < made[k].hi; i++) tile_of_vert[ord_all[i]] = ti;
          spans[(size_t)ti] = made[k];
        }
      }
    }
    /* pass 2 (a byte set to 1 from several threads is the same byte either way) */
    std::vector<unsigned char> is_bnd((size_t)V, 0);
    int bad_prim = -1;
#pragma omp parallel for schedule(dynamic, 8)
    for (int l = 0; l < L; l++) {
      for (int pos = leaf_pbeg[l]; pos < leaf_pbeg[l] + leaf_pcnt[l]; pos++) {
        const int t = pb->prim_indices[pos];
        if (t < 0 || t >= T) {
#pragma omp atomic write
          bad_prim = pos;
          continue;
        }
        const int p = ctx->h_tri_poly[t];
        const int ls = ctx->h_poly_start[p], len = ctx->h_poly_len[p];
        int t0 = -2;
        bool mixed = false;
        for (int k = 0; k < len; k++) {
          const int tv = tile_of_vert[ctx->h_loop_v[ls + k]];
          if (t0 == -2) t0 = tv;
          else if (tv != t0) mixed = true;
        }
        for (int k = 0; k < len; k++) {
          const int v = ctx->h_loop_v[ls + k];
          const int tv = tile_of_vert[v];
          if (mixed || tv < 0 || spans[tv].leaf != l) is_bnd[v] = 1;
        }
      }
    }
    if (bad_prim >= 0) return fail(ctx, DSC_ERR_INVALID, "prim_indices[%d] out of range", bad_prim);
    /* pass 3, tile by tile (independent) */
    tile_ibnd.assign(spans.size(), 0);
    tile_range.assign(spans.size(), make_int2(0, 0));
    tile_leaf.assign(spans.size(), 0);
#pragma omp parallel for schedule(dynamic, 16)
    for (long long ti = 0; ti < (long long)spans.size(); ti++) {
      const TileSpan &sp = spans[(size_t)ti];
      const int cnt = sp.hi - sp.lo;
      float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
      int ninterior = 0;
      for (int i = sp.lo; i < sp.hi; i++) {
        for (int k = 0; k < 3; k++) {
          const float c = hco[(size_t)3 * ord_all[i] + k];
          mn[k] = std::min(mn[k], c);
          mx[k] = std::max(mx[k], c);
        }
        ninterior += is_bnd[ord_all[i]] ? 0 : 1;
      }
      int ax[3] = {0, 1, 2};
      std::sort(ax, ax + 3, [&](int a, int b) { return (mx[a] - mn[a]) > (mx[b] - mn[b]) || ((mx[a] - mn[a]) == (mx[b] - mn[b]) && a < b); });
      /* rows along the second-widest axis, ascending along the widest inside a row */
      const int a = ax[0], b = ax[1];
      const float ea = mx[a] - mn[a], eb2 = mx[b] - mn[b];
      int rows = 1;
      if (ea > 0.0f && eb2 > 0.0f) rows = std::max(1, std::min(cnt, (int)lrintf(sqrtf((float)cnt * eb2 / ea))));
      auto row_of = [&](int v) {
        if (rows <= 1) return 0;
        const int rr = (int)((hco[(size_t)3 * v + b] - mn[b]) / eb2 * (float)rows);
        return std::min(std::max(rr, 0), rows - 1);
      };
      std::sort(ord_all.begin() + sp.lo, ord_all.begin() + sp.hi, [&](int p, int q) {
        if (is_bnd[p] != is_bnd[q]) return is_bnd[p] < is_bnd[q];
        const int rp = row_of(p), rq = row_of(q);
        if (rp != rq) return rp < rq;
        const float cp = hco[(size_t)3 * p + a], cq = hco[(size_t)3 * q + a];
        return cp < cq || (cp == cq && p < q);
      });
      for (int i = sp.lo; i < sp.hi; i++) ctx->slot_of[ord_all[i]] = sp.base + (i - sp.lo);
      tile_range[(size_t)ti] = make_int2(sp.base, cnt);
      tile_leaf[(size_t)ti] = sp.leaf;
      tile_ibnd[(size_t)ti] = ninterior & ~31;
    }
  }
  const int NT = (int)tile_range.size();
  if (expect_prim != T) return fail(ctx, DSC_ERR_INVALID, "leaves hold %d looptris, mesh has %d", expect_prim, T);
  cur = (cur + 31) & ~31ll;
  for (int v = 0; v < V; v++) {
    if (ctx->slot_of[v] < 0) ctx->slot_of[v] = (int)cur++; /* loose vertex: in no face; parked after the leaves */
  }
  const int VP = (int)((cur + 31) & ~31ll) + 32;
  ctx->vpad = VP;
  ctx->nwords = VP / 32;

  DSC_TMARK("slots+tiles");
  /* per-slot vertex data */
  auto to_slots = [&](const std::vector<float> &src, int comp, int stride) {
    std::vector<float> out((size_t)VP, 0.0f);
    const int *so = ctx->slot_of.data();
    const float *sp = src.data();
    float *op = out.data();
#pragma omp parallel for schedule(static)
    for (int v = 0; v < V; v++) op[so[v]] = sp[(size_t)stride * v + comp]; /* slot_of is a permutation: no two writes meet */
    return out;
  };
  for (int k = 0; k < 3; k++) {
    float **dst = (k == 0) ? &m.cx : (k == 1) ? &m.cy : &m.cz;
    if ((r = dev_upload(ctx, dst, to_slots(ctx->h_co, k, 3)))) return r;
    float **dn = (k == 0) ? &m.nx : (k == 1) ? &m.ny : &m.nz;
    if (ctx->has_no) {
      if ((r = dev_upload(ctx, dn, to_slots(ctx->h_no, k, 3)))) return r;
    }
    else if ((r = dev_zero(ctx, dn, (size_t)VP))) return r;
    CU(cudaStreamSynchronize(ctx->stream));
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
  if ((r = dev_alloc(ctx, &ctx->d_stage3, (size_t)4 * V))) return r;
  if (!ctx->h_tail.empty() && (r = dev_upload(ctx, &ctx->d_tail, ctx->h_tail))) return r;
  {
    /* row a10: hidden verts (MVert.flag & ME_HIDE, the low byte of the tail word) / hidden grid elements, bit per slot */
    std::vector<unsigned> hid((size_t)ctx->nwords, 0u);
    bool any = false;
    for (int v = 0; v < V; v++) {
      const bool h = ctx->is_grids ? (!ctx->h_elem_hidden.empty() && ctx->h_elem_hidden[v]) : (!ctx->h_tail.empty() && (ctx->h_tail[v] & 16u));
      if (h) {
        const int sl = ctx->slot_of[v];
        hid[(size_t)sl >> 5] |= 1u << (sl & 31);
        any = true;
      }
    }
    if (any) {
      if ((r = dev_upload(ctx, &ctx->d_hidden, hid))) return r;
      m.hidden = ctx->d_hidden;
    }
    std::vector<unsigned char>().swap(ctx->h_elem_hidden);
  }
  if ((r = dev_alloc(ctx, &ctx->d_list, (size_t)std::max(V, L))) || (r = dev_zero(ctx, &ctx->d_count, 1))) return r;

  DSC_TMARK("vertex data upload");
  /* smooth adjacency in slot order */
  if (ctx->has_nb && ctx->is_grids) {
    /* the rim table, element indices -> slots; per-slot boundary flags only when the base mesh is open */
    if ((r = dev_zero(ctx, &m.tx, (size_t)VP)) || (r = dev_zero(ctx, &m.ty, (size_t)VP)) || (r = dev_zero(ctx, &m.tz, (size_t)VP))) return r;
    const int gs = ctx->grid_size, gs2 = gs * gs, rim = 4 * gs - 4;
    std::vector<int> rim_slots(ctx->h_rim_nb.size());
    for (size_t i = 0; i < rim_slots.size(); i++) {
      const int e = ctx->h_rim_nb[i];
      if (e >= V) return fail(ctx, DSC_ERR_INVALID, "rim neighbour table names element %d", e);
      rim_slots[i] = e < 0 ? -1 : ctx->slot_of[e];
    }
    GridNb &gn = ctx->gnb;
    gn.gs = gs;
    gn.gs2 = gs2;
    gn.rim_w = ctx->rim_width;
    if ((r = dev_upload_c(ctx, &gn.rim_nb, rim_slots))) return r;
    bool any_boundary = false;
    for (unsigned char b : ctx->h_rim_bnd) any_boundary = any_boundary || b;
    if (any_boundary) {
      std::vector<unsigned char> bnd((size_t)VP, 0);
      for (int g = 0; g < ctx->totgrid; g++) {
        for (int b = 0; b < rim; b++) {
          if (!ctx->h_rim_bnd[(size_t)g * rim + b]) continue;
          const int x = b < gs ? b : (b < 2 * gs ? b - gs : (b < 3 * gs - 2 ? 0 : gs - 1));
          const int y = b < gs ? 0 : (b < 2 * gs ? gs - 1 : (b < 3 * gs - 2 ? b - 2 * gs + 1 : b - (3 * gs - 2) + 1));
          bnd[(size_t)ctx->slot_of[(size_t)g * gs2 + (size_t)y * gs + x]] = 1;
        }
      }
      if ((r = dev_upload_c(ctx, &gn.rim_bnd, ctx->h_rim_bnd)) || (r = dev_upload_c(ctx, &m.boundary, bnd))) return r;
    }
    CU(cudaStreamSynchronize(ctx->stream));
  }
  if (ctx->has_nb && !ctx->is_grids) {
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
    if ((r = dev_upload_c(ctx, &m.nb_off, off)) || (r = dev_upload_c(ctx, &m.nb_idx, idx)) || (r = dev_upload_c(ctx, &m.boundary, bnd)))
      return r;
    CU(cudaStreamSynchronize(ctx->stream));
  }

  std::vector<int> leaf_sbeg(L);
  std::vector<unsigned char> leaf_fast(L, 1);
  if (ctx->is_grids) {
    /* grids: no mesh connectivity tables; the tile kernel is not used (leaf_fast = 0 keeps the box
     * reset of the refit tagging off: k_grid_leaf_bb stores whole boxes) */
    std::fill(leaf_fast.begin(), leaf_fast.end(), (unsigned char)0);
    std::vector<TileMeta> tmeta((size_t)std::max(NT, 1));
    for (int t = 0; t < NT; t++) {
      tmeta[t].ubeg = tile_range[t].x;
      tmeta[t].ucnt = tile_range[t].y;
      tmeta[t].leaf = tile_leaf[t];
      tmeta[t].tile0 = leaf_tile0[tile_leaf[t]];
      tmeta[t].ntfast = leaf_tile0[tile_leaf[t] + 1] - leaf_tile0[tile_leaf[t]];
    }
    std::vector<int> one(1, 0);
    std::vector<unsigned> oneu(1, 0u);
    std::vector<unsigned short> e_pv(8, 0);
    std::vector<unsigned> v2_goff((size_t)VP / 32 + 1, 0u);
    if ((r = dev_upload_c(ctx, &m.stage_slots, one)) || (r = dev_upload_c(ctx, &m.e_pv, e_pv)) ||
        (r = dev_upload_c(ctx, &m.e_halo_leaf, one)) || (r = dev_upload_c(ctx, &m.tile_meta, tmeta)) ||
        (r = dev_upload_c(ctx, &m.tile_range, tile_range)) || (r = dev_upload_c(ctx, &m.leaf_tile0, leaf_tile0)) ||
        (r = dev_upload_c(ctx, &m.v2_goff, v2_goff)) || (r = dev_upload_c(ctx, &m.v2_idx, oneu)) ||
        (r = dev_upload_c(ctx, &m.leaf_fast, leaf_fast)))
      return r;
    m.ntile = NT;
    ctx->nb_smem = 1024;
    m.sm_off_f = m.sm_off_e = m.sm_off_v2 = m.sm_off_h = 0;
    if ((r = dev_zero(ctx, &m.tile_list, (size_t)std::max(NT, 1) * DSC_SLOTS)) ||
        (r = dev_zero(ctx, &m.atile_list, (size_t)std::max(NT, 1) * DSC_SLOTS)) ||
        (r = dev_zero(ctx, &m.flag_tile_list, (size_t)std::max(NT, 1))))
      return r;
    /* grid tables, element indices -> slots */
    DevGrids &g = ctx->g;
    memset(&g, 0, sizeof(g));
    g.gs = ctx->grid_size;
    g.gs2 = g.gs * g.gs;
    g.totgrid = ctx->totgrid;
    g.totface = (int)ctx->h_face_start.size();
    g.totedge = (int)ctx->h_edge_off.size() - 1;
    g.totcvert = (int)ctx->h_cvert_off.size() - 1;
    std::vector<int> grid_face((size_t)g.totgrid, 0), leaf_gbeg((size_t)L + 1, 0), leaf_grids((size_t)T);
    for (int f = 0; f < g.totface; f++) {
      for (int c = 0; c < ctx->h_face_num[f]; c++) {
        const int gr = ctx->h_face_start[f] + c;
        if (gr < 0 || gr >= g.totgrid) return fail(ctx, DSC_ERR_INVALID, "face %d names grid %d", f, gr);
        grid_face[gr] = f;
      }
    }
    for (int l = 0; l < L; l++) {
      leaf_gbeg[l] = leaf_pbeg[l];
      for (int k = 0; k < leaf_pcnt[l]; k++) leaf_grids[leaf_pbeg[l] + k] = pb->prim_indices[leaf_pbeg[l] + k];
    }
    leaf_gbeg[L] = T;
    auto to_slot = [&](const std::vector<int> &elems, std::vector<int> &out) -> bool {
      out.resize(elems.size());
      for (size_t i = 0; i < elems.size(); i++) {
        if (elems[i] < 0 || elems[i] >= V) return false;
        out[i] = ctx->slot_of[elems[i]];
      }
      return true;
    };
    std::vector<int> edge_slots, cvert_slots;
    if (!to_slot(ctx->h_edge_elems, edge_slots) || !to_slot(ctx->h_cvert_elems, cvert_slots))
      return fail(ctx, DSC_ERR_INVALID, "grid adjacency names an element out of range");
    ctx->has_odd_edges = false;
    for (int e = 0; e < g.totedge; e++) {
      if (ctx->h_edge_off[e + 1] - ctx->h_edge_off[e] > 2) ctx->has_odd_edges = true;
    }
    if ((r = dev_upload_c(ctx, &g.grid_slot0, grid_slot0)) || (r = dev_upload_c(ctx, &g.leaf_gbeg, leaf_gbeg)) ||
        (r = dev_upload_c(ctx, &g.leaf_grids, leaf_grids)) || (r = dev_upload_c(ctx, &g.face_start, ctx->h_face_start)) ||
        (r = dev_upload_c(ctx, &g.face_num, ctx->h_face_num)) || (r = dev_upload_c(ctx, &g.grid_face, grid_face)) ||
        (r = dev_upload_c(ctx, &g.grid_edge, ctx->h_grid_edge)) || (r = dev_upload_c(ctx, &g.grid_cvert, ctx->h_grid_cvert)) ||
        (r = dev_upload_c(ctx, &g.edge_off, ctx->h_edge_off)) || (r = dev_upload_c(ctx, &g.edge_slots, edge_slots)) ||
        (r = dev_upload_c(ctx, &g.cvert_off, ctx->h_cvert_off)) || (r = dev_upload_c(ctx, &g.cvert_slots, cvert_slots)))
      return r;
    if ((r = dev_zero(ctx, &g.face_stamp, (size_t)g.totface + 1)) || (r = dev_zero(ctx, &g.edge_stamp, (size_t)g.totedge + 1)) ||
        (r = dev_zero(ctx, &g.cvert_stamp, (size_t)g.totcvert + 1)) || (r = dev_zero(ctx, &g.face_list, (size_t)g.totface + 1)) ||
        (r = dev_zero(ctx, &g.edge_list, (size_t)g.totedge + 1)) || (r = dev_zero(ctx, &g.cvert_list, (size_t)g.totcvert + 1)) ||
        (r = dev_zero(ctx, &g.cnt, 1)))
      return r;
    ctx->gnb.leaf_gbeg = g.leaf_gbeg;
    ctx->gnb.leaf_grids = g.leaf_grids;
    g.mask = ctx->has_mask ? ctx->d_mask : nullptr;
    g.has_odd_edges = ctx->has_odd_edges ? 1 : 0;
    g.max_face_grids = 1;
    for (int f = 0; f < g.totface; f++) g.max_face_grids = std::max(g.max_face_grids, ctx->h_face_num[f]);
    ctx->gn_smem = dsc_grid_normals_smem(g.gs);
    ctx->grid_normals_flat = getenv("DSC_GRID_NORMALS_FLAT") != nullptr;
    ctx->grid_fused = getenv("DSC_GRID_FUSED") != nullptr; /* measured slower than the nine launches (119 -> 141 us per C5 dab): opt-in */
    CU(cudaFuncSetAttribute(k_grid_normals, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->gn_smem));
    CU(cudaFuncSetAttribute(k_grid_dab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->gn_smem));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  else
  /* looptris by position; vertex -> looptri CSR; per-tile local tables of the shared-memory normals kernel */
  {
    /* built on the host, leaf by leaf in parallel, appended in leaf order (dsc_tile_tables.h; checked against the serial
     * construction without a GPU by tests/test_tile_tables.py) */
    TileTablesIn tin;
    tin.L = L; tin.VP = VP; tin.T = T; tin.NT = NT; tin.totpoly = ctx->totpoly;
    tin.leaves = leaves.data();
    tin.vert_indices = pb->vert_indices; tin.vert_offset = pb->vert_offset; tin.prim_indices = pb->prim_indices;
    tin.slot_of = ctx->slot_of.data();
    static_assert(sizeof(DscTileRange) == sizeof(int2), "DscTileRange is an int2");
    tin.tile_range = reinterpret_cast<const DscTileRange *>(tile_range.data());
    tin.leaf_tile0 = leaf_tile0.data();
    tin.leaf_ucnt = leaf_ucnt.data(); tin.leaf_scnt = leaf_scnt.data(); tin.leaf_pbeg = leaf_pbeg.data(); tin.leaf_pcnt = leaf_pcnt.data();
    tin.tri_vert = ctx->h_tri_vert.data(); tin.tri_poly = ctx->h_tri_poly.data(); tin.poly_start = ctx->h_poly_start.data();
    tin.poly_len = ctx->h_poly_len.data(); tin.loop_v = ctx->h_loop_v.data();
    TileTablesOut tt;
    {
      int where = -1;
      const int tr = dsc_build_tile_tables(tin, tt, 0, &where);
      if (tr == DSC_TT_BAD_PRIM) return fail(ctx, DSC_ERR_INVALID, "prim_indices[%d] out of range", where);
      if (tr == DSC_TT_TOO_MANY_TILES) return fail(ctx, DSC_ERR_UNSUPPORTED, "leaf %d has too many tiles", where);
    }
    std::vector<int> &tri_leaf = tt.tri_leaf;
    std::vector<unsigned> &vt_off = tt.vt_off, &vt_idx = tt.vt_idx;
    std::vector<int> &leaf_sslots = tt.leaf_sslots, &stage = tt.stage, &e_halo_leaf = tt.e_halo_leaf;
    std::vector<unsigned short> &e_pv = tt.e_pv;
    std::vector<TileMeta> &tmeta = tt.tmeta;
    std::vector<unsigned> &v2_goff = tt.v2_goff, &v2_idx = tt.v2_idx;
    std::vector<TileDims> &tile_dims = tt.tile_dims;
    leaf_sbeg = tt.leaf_sbeg;
    leaf_fast = tt.leaf_fast;
    if (tt.any_slow_leaf) ctx->any_slow_leaf = true;
    {
      /* every shared-memory region is sized for the largest fast tile, so it never moves between tiles */
      int mx_nloc = 4, mx_ne = 1, mx_v2 = 32, mx_h = 16;
      for (const TileDims &t : tile_dims) {
        mx_nloc = std::max(mx_nloc, t.nloc_a);
        mx_ne = std::max(mx_ne, t.ne);
        mx_v2 = std::max(mx_v2, t.v2w);
        mx_h = std::max(mx_h, t.ehalo);
      }
      m.sm_off_f = 12 * mx_nloc;
      m.sm_off_e = m.sm_off_f + 16 * (mx_ne + 1);
      m.sm_off_v2 = m.sm_off_e + 8 * ((mx_ne + 1) & ~1);
      m.sm_off_h = m.sm_off_v2 + 4 * mx_v2;
      ctx->nb_smem = (size_t)m.sm_off_h + (((size_t)mx_h + 15) & ~(size_t)15);
      if (ctx->nb_smem > 220 * 1024) {
        /* the regions' maxima do not fit one SM together: every leaf takes the general path */
        std::fill(leaf_fast.begin(), leaf_fast.end(), (unsigned char)0);
        for (TileMeta &t : tmeta) t.ntfast &= ~(1 << 16);
        ctx->any_slow_leaf = true;
        ctx->nb_smem = 1024;
      }
    }

    static_assert(sizeof(TileMeta) == 3 * sizeof(int4), "TileMeta is three int4");
    /* tile_range.y = unique verts | start of the boundary run << 16; a leaf on the general path has no interior part (its
     * whole tiles are displaced by the boundary kernel, its normals and box come from k_normals / k_leaf_bb) */
    std::vector<int2> tile_range_dev(tile_range);
    for (int t = 0; t < NT; t++) tile_range_dev[t].y |= (leaf_fast[tile_leaf[t]] ? tile_ibnd[t] : 0) << 16;
    if ((r = dev_upload_c(ctx, &m.stage_slots, stage)) || (r = dev_upload_c(ctx, &m.e_pv, e_pv)) ||
        (r = dev_upload_c(ctx, &m.e_halo_leaf, e_halo_leaf)) || (r = dev_upload_c(ctx, &m.tile_meta, tmeta)) ||
        (r = dev_upload_c(ctx, &m.tile_range, tile_range_dev)) || (r = dev_upload_c(ctx, &m.leaf_tile0, leaf_tile0)) ||
        (r = dev_upload_c(ctx, &m.v2_goff, v2_goff)) || (r = dev_upload_c(ctx, &m.v2_idx, v2_idx)) ||
        (r = dev_upload_c(ctx, &m.leaf_fast, leaf_fast)))
      return r;
    m.ntile = NT;
    if ((r = dev_zero(ctx, &m.tile_list, (size_t)std::max(NT, 1) * DSC_SLOTS)) ||
        (r = dev_zero(ctx, &m.atile_list, (size_t)std::max(NT, 1) * DSC_SLOTS)) ||
        (r = dev_zero(ctx, &m.flag_tile_list, (size_t)std::max(NT, 1))))
      return r;
    CU(cudaStreamSynchronize(ctx->stream));

    if (ctx->want_raycast) {
      std::vector<int> slot_leaf((size_t)VP / 32 + 1, 0);
      for (int l = 0; l < L; l++) {
        for (int gq = leaf_ubeg[l] / 32; gq <= (leaf_ubeg[l] + std::max(leaf_ucnt[l], 1) - 1) / 32; gq++) slot_leaf[gq] = l;
      }
      if ((r = dev_upload_c(ctx, &ctx->d_slot_leaf, slot_leaf))) return r;
    }
    if (ctx->want_draw) {
      /* draw-buffer fill: the three corners of every looptri, as slots, by looptri position */
      std::vector<int4> tri_slots((size_t)std::max(T, 1));
      for (int pos = 0; pos < T; pos++) {
        const int t = pb->prim_indices[pos];
        tri_slots[pos] = make_int4(ctx->slot_of[ctx->h_tri_vert[(size_t)3 * t]], ctx->slot_of[ctx->h_tri_vert[(size_t)3 * t + 1]],
                                   ctx->slot_of[ctx->h_tri_vert[(size_t)3 * t + 2]], ctx->h_tri_poly[t]);
      }
      if ((r = dev_upload_c(ctx, &ctx->d_tri_slots, tri_slots))) return r;
    }
    if (ctx->any_slow_leaf || ctx->want_draw) {
      /* general path tables: poly verts as slots per looptri position, n-gon lists */
      std::vector<int> pv[4];
      for (int k = 0; k < 4; k++) pv[k].assign((size_t)std::max(T, 1), -1);
      std::vector<int> ngon_id((size_t)std::max(ctx->totpoly, 1), -1), poly_off(1, 0), poly_slots;
      for (int pos = 0; pos < T; pos++) {
        const int t = pb->prim_indices[pos];
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
      }
      if ((r = dev_upload_c(ctx, &m.pv0, pv[0])) || (r = dev_upload_c(ctx, &m.pv1, pv[1])) || (r = dev_upload_c(ctx, &m.pv2, pv[2])) ||
          (r = dev_upload_c(ctx, &m.pv3, pv[3])) || (r = dev_upload_c(ctx, &m.tri_leaf, tri_leaf)) ||
          (r = dev_upload_c(ctx, &m.poly_off, poly_off)) || (r = dev_upload_c(ctx, &m.poly_slots, poly_slots)) ||
          (r = dev_upload_c(ctx, &m.vt_off, vt_off)) || (r = dev_upload_c(ctx, &m.vt_idx, vt_idx)) ||
          (r = dev_upload_c(ctx, &m.leaf_sslots, leaf_sslots)))
        return r;
      CU(cudaStreamSynchronize(ctx->stream));
    }
  }

  DSC_TMARK("looptri + tile tables");
  /* leaves */
  {
    if ((r = dev_upload_c(ctx, &m.leaf_ubeg, leaf_ubeg)) || (r = dev_upload_c(ctx, &m.leaf_ucnt, leaf_ucnt)) ||
        (r = dev_upload_c(ctx, &m.leaf_sbeg, leaf_sbeg)) || (r = dev_upload_c(ctx, &m.leaf_scnt, leaf_scnt)) ||
        (r = dev_upload_c(ctx, &m.leaf_pbeg, leaf_pbeg)) || (r = dev_upload_c(ctx, &m.leaf_pcnt, leaf_pcnt)))
      return r;
    m.nleaf = L;
    ctx->h_leaf_pbeg = leaf_pbeg;
    ctx->h_leaf_pcnt = leaf_pcnt;
    int max_u = 1;
    for (int l = 0; l < L; l++) max_u = std::max(max_u, leaf_ucnt[l]);
    m.max_chunks = (max_u + DSC_CHUNK - 1) / DSC_CHUNK;
    if ((r = dev_zero(ctx, &m.leaf_state, (size_t)L))) return r;
    if ((r = dev_zero(ctx, &m.hit_list, (size_t)L * DSC_SLOTS)) || (r = dev_zero(ctx, &m.area_list, (size_t)L * DSC_SLOTS)) ||
        (r = dev_zero(ctx, &m.search_list, (size_t)L)) || (r = dev_zero(ctx, &m.flag_list, (size_t)L)))
      return r;
    m.ghit_words = (L + 31) / 32 + 1;
    if ((r = dev_zero(ctx, &m.ghit, (size_t)m.ghit_words * DSC_SLOTS)) || (r = dev_zero(ctx, &m.flag_mask, (size_t)m.ghit_words)))
      return r;
    CU(cudaMallocHost((void **)&ctx->h_list, sizeof(int) * (size_t)std::max(L, 1)));
    CU(cudaStreamSynchronize(ctx->stream));
  }

  DSC_TMARK("leaves");
  /* nodes in device numbering: leaves [0, L) in traversal order, then the inner nodes breadth-first */
  {
    std::vector<int> hflag(N), hchild(N);
    for (int n = 0; n < N; n++) {
      hflag[n] = pb->flag[n];
      hchild[n] = pb->children_offset[n];
    }
    ctx->dev_of_node.assign(N, -1);
    ctx->node_of_dev.assign(N, -1);
    for (int l = 0; l < L; l++) {
      ctx->dev_of_node[leaves[l]] = l;
      ctx->node_of_dev[l] = leaves[l];
    }
    std::vector<int> order(1, 0), depth(N, 0);
    int next = L, maxdepth = 0;
    for (size_t i = 0; i < order.size(); i++) {
      const int n = order[i];
      if (hflag[n] & DSC_PBVH_Leaf) continue;
      ctx->dev_of_node[n] = next;
      ctx->node_of_dev[next] = n;
      next++;
      const int c = hchild[n];
      if (c <= 0 || c + 1 >= N) return fail(ctx, DSC_ERR_INVALID, "node %d has bad children_offset %d", n, c);
      for (int k = 0; k < 2; k++) {
        depth[c + k] = depth[n] + 1;
        maxdepth = std::max(maxdepth, depth[c + k]);
        order.push_back(c + k);
      }
    }
    if (next != N || (int)order.size() != N) return fail(ctx, DSC_ERR_INVALID, "PBVH is not a tree over %d nodes", N);
    std::vector<float> bb((size_t)6 * N), obb((size_t)6 * N);
    std::vector<int> flag(N), child0(N, -1), child1(N, -1);
    std::vector<int4> topo(N, make_int4(-1, 0, -1, 0));
    for (int n = 0; n < N; n++) {
      const int dn = ctx->dev_of_node[n];
      for (int k = 0; k < 6; k++) {
        bb[(size_t)k * N + dn] = pb->node_bb[(size_t)6 * n + k];
        obb[(size_t)k * N + dn] = pb->node_orig_bb[(size_t)6 * n + k];
      }
      flag[dn] = hflag[n];
      if (!(hflag[n] & DSC_PBVH_Leaf)) {
        child0[dn] = ctx->dev_of_node[hchild[n]];
        child1[dn] = ctx->dev_of_node[hchild[n] + 1];
      }
    }
    for (int dn = 0; dn < N; dn++) {
      if (child0[dn] >= 0) {
        topo[child0[dn]] = make_int4(dn, 1, child1[dn], 0);
        topo[child1[dn]] = make_int4(dn, 2, child0[dn], 0);
      }
    }
    std::vector<int> level_off(maxdepth + 2, 0), level_nodes;
    for (int dpt = 0; dpt <= maxdepth; dpt++) {
      level_off[dpt] = (int)level_nodes.size();
      for (int n : order) {
        if (!(hflag[n] & DSC_PBVH_Leaf) && depth[n] == dpt) level_nodes.push_back(ctx->dev_of_node[n]);
      }
    }
    level_off[maxdepth + 1] = (int)level_nodes.size();
    m.nlevel = maxdepth + 1;
    if ((r = dev_upload(ctx, &m.bb, bb)) || (r = dev_upload(ctx, &m.obb, obb)) || (r = dev_upload(ctx, &m.node_flag, flag))) return r;
    if ((r = dev_upload_c(ctx, &m.topo, topo)) || (r = dev_upload_c(ctx, &m.child0, child0)) ||
        (r = dev_upload_c(ctx, &m.child1, child1)) || (r = dev_upload_c(ctx, &m.level_off, level_off)) ||
        (r = dev_upload_c(ctx, &m.level_nodes, level_nodes)))
      return r;
    if ((r = dev_zero(ctx, &m.node_mark, (size_t)N)) || (r = dev_zero(ctx, &m.pending, (size_t)2 * N)) ||
        (r = dev_zero(ctx, &m.arrived, (size_t)2 * N)) || (r = dev_zero(ctx, &m.grid_bar, 1)))
      return r;
    m.totnode = N;
    CU(cudaStreamSynchronize(ctx->stream));
  }
  if ((r = dev_zero(ctx, &m.st, DSC_SLOTS)) || (r = dev_zero(ctx, &m.tot, 1))) return r;
  if ((r = dev_zero(ctx, &ctx->d_curve, 257))) return r;
  CU(cudaStreamSynchronize(ctx->stream));

  DSC_TMARK("nodes");
  /* launch shape of the shared-memory normals kernel */
  CU(cudaFuncSetAttribute(k_normals_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->nb_smem));
  int occ = 1;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_normals_tile, NT_THREADS, ctx->nb_smem));
  ctx->nb_grid = ctx->num_sms * std::max(occ, 1);
  {
    /* the fused dab kernel (interior brush + normals + boxes), one instantiation per tool */
    const void *fn[4] = {(const void *)k_dab_tile<DSC_TOOL_DRAW>, (const void *)k_dab_tile<DSC_TOOL_INFLATE>,
                         (const void *)k_dab_tile<DSC_TOOL_GRAB>, (const void *)k_dab_tile<DSC_TOOL_CLAY_STRIPS>};
    for (int k = 0; k < 4; k++) {
      CU(cudaFuncSetAttribute(fn[k], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->nb_smem));
      int o = 0;
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, fn[k], NT_THREADS, ctx->nb_smem));
      ctx->fused_grid[k] = ctx->num_sms * std::max(o, 1);
    }
    /* opt-in: measured slower than the two passes (C3 sweep 24.6 vs 20.7 ms per stroke): both kernels are bound by
     * instruction issue, not by HBM, so saving the second read of the positions buys nothing while the brush arithmetic
     * runs at the tile kernel's occupancy */
    ctx->use_fused = getenv("DSC_FUSE") != nullptr;
  }
  {
    /* the persistent batch kernel, one instantiation per tool: every CTA must be resident */
    const void *fn[4] = {(const void *)k_dab_batch<DSC_TOOL_DRAW>, (const void *)k_dab_batch<DSC_TOOL_INFLATE>,
                         (const void *)k_dab_batch<DSC_TOOL_GRAB>, (const void *)k_dab_batch<DSC_TOOL_CLAY_STRIPS>};
    for (int k = 0; k < 4; k++) {
      CU(cudaFuncSetAttribute(fn[k], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->nb_smem));
      int o = 0;
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, fn[k], NT_THREADS, ctx->nb_smem));
      ctx->batch_grid[k] = ctx->num_sms * o;
    }
  }

  DSC_TMARK("launch shapes");
  /* multi-GPU: owned leaf run, hit-mask ring, halo index lists */
  m.own_lo = 0;
  m.own_hi = L;
  ctx->leaf_range.assign(2, 0);
  ctx->leaf_range[1] = L;
  if (ctx->world > 1) {
    std::vector<int> pl;
    plan_partition(pb, ctx->world, pl, ctx->leaf_range);
    m.own_lo = ctx->leaf_range[ctx->rank];
    m.own_hi = ctx->leaf_range[ctx->rank + 1];
    ctx->slot_range.assign(ctx->world + 1, 0);
    for (int q = 0; q <= ctx->world; q++) {
      const int l = ctx->leaf_range[q];
      ctx->slot_range[q] = (l < L) ? leaf_ubeg[l] : ((L ? leaf_ubeg[L - 1] + leaf_ucnt[L - 1] + 31 : 0) & ~31);
    }
    std::vector<int> sidx, ridx;
    if (ctx->is_grids) {
      /* partitioned grids: what this rank averages after the brush, and the halo elements (plan_grids_rank) */
      GridTables t;
      t.totgrid = ctx->totgrid; t.gs = ctx->grid_size; t.totface = (int)ctx->h_face_start.size();
      t.totedge = (int)ctx->h_edge_off.size() - 1; t.totcvert = (int)ctx->h_cvert_off.size() - 1;
      t.face_start = ctx->h_face_start.data(); t.face_num = ctx->h_face_num.data(); t.edge_off = ctx->h_edge_off.data();
      t.edge_elems = ctx->h_edge_elems.data(); t.cvert_off = ctx->h_cvert_off.data(); t.cvert_elems = ctx->h_cvert_elems.data();
      t.grid_edge = ctx->h_grid_edge.data(); t.grid_cvert = ctx->h_grid_cvert.data();
      t.rim_w = ctx->rim_width; t.rim_nb = ctx->h_rim_nb.empty() ? nullptr : ctx->h_rim_nb.data();
      std::vector<int> grid_owner((size_t)ctx->totgrid, 0);
      for (int q = 0; q < ctx->world; q++) {
        for (int l = ctx->leaf_range[q]; l < ctx->leaf_range[q + 1]; l++) {
          for (int k = 0; k < leaf_pcnt[l]; k++) grid_owner[pb->prim_indices[leaf_pbeg[l] + k]] = q;
        }
      }
      GridPlan plan;
      std::vector<int> se, re;
      std::vector<unsigned char> halo_grid;
      plan_grids_lists(t, grid_owner, ctx->world, ctx->rank, plan, ctx->send_off, se, ctx->recv_off, re, &halo_grid);
      {
        /* leaves whose gathering can change a halo element on some rank: the dab moves the elements of the leaf's grids,
         * the stitch and the normal pass reach every grid of those grids' faces and the rims of the faces that share a
         * coarse edge or vertex with them -- so a leaf is "near a cut" when one of those faces holds a halo grid */
        const int F = t.totface;
        std::vector<int> grid_face_h((size_t)t.totgrid, 0);
        std::vector<unsigned char> face_halo((size_t)F, 0), edge_halo((size_t)t.totedge, 0), vert_halo((size_t)t.totcvert, 0);
        for (int f = 0; f < F; f++) {
          for (int c = 0; c < t.face_num[f]; c++) {
            grid_face_h[t.face_start[f] + c] = f;
            if (halo_grid[t.face_start[f] + c]) face_halo[f] = 1;
          }
        }
        const int gs2h = t.gs * t.gs;
        for (int e = 0; e < t.totedge; e++) {
          for (int k = t.edge_off[e]; k < t.edge_off[e + 1]; k++) {
            if (face_halo[grid_face_h[t.edge_elems[(size_t)k * 2 * t.gs] / gs2h]]) edge_halo[e] = 1;
          }
        }
        for (int v = 0; v < t.totcvert; v++) {
          for (int k = t.cvert_off[v]; k < t.cvert_off[v + 1]; k++) {
            if (face_halo[grid_face_h[t.cvert_elems[k] / gs2h]]) vert_halo[v] = 1;
          }
        }
        std::vector<unsigned> near_mask((size_t)m.ghit_words, 0u);
        for (int l = 0; l < L; l++) {
          bool near = false;
          for (int k = 0; k < leaf_pcnt[l] && !near; k++) {
            const int f = grid_face_h[pb->prim_indices[leaf_pbeg[l] + k]];
            near = face_halo[f] != 0;
            for (int c = 0; c < t.face_num[f] && !near; c++) {
              near = edge_halo[t.grid_edge[t.face_start[f] + c]] || vert_halo[t.grid_cvert[t.face_start[f] + c]];
            }
          }
          if (near) near_mask[l >> 5] |= 1u << (l & 31);
        }
        if ((r = dev_upload(ctx, &ctx->d_near_mask, near_mask))) return r;
        /* which ranks gathering a leaf can affect: the owners of the grids of every face within two hops */
        std::vector<std::vector<int>> fadj((size_t)F);
        auto link_faces = [&](const std::vector<int> &fs) {
          for (int a : fs) {
            for (int b : fs) {
              if (a != b) fadj[a].push_back(b);
            }
          }
        };
        std::vector<int> fs;
        for (int e = 0; e < t.totedge; e++) {
          fs.clear();
          for (int k = t.edge_off[e]; k < t.edge_off[e + 1]; k++) fs.push_back(grid_face_h[t.edge_elems[(size_t)k * 2 * t.gs] / gs2h]);
          link_faces(fs);
        }
        for (int v = 0; v < t.totcvert; v++) {
          fs.clear();
          for (int k = t.cvert_off[v]; k < t.cvert_off[v + 1]; k++) fs.push_back(grid_face_h[t.cvert_elems[k] / gs2h]);
          link_faces(fs);
        }
        std::vector<unsigned> face_owner((size_t)F, 0u), hop1((size_t)F, 0u), hop2((size_t)F, 0u);
        for (int f = 0; f < F; f++) {
          for (int c = 0; c < t.face_num[f]; c++) face_owner[f] |= 1u << grid_owner[t.face_start[f] + c];
        }
        for (int f = 0; f < F; f++) {
          hop1[f] = face_owner[f];
          for (int b : fadj[f]) hop1[f] |= face_owner[b];
        }
        for (int f = 0; f < F; f++) {
          hop2[f] = hop1[f];
          for (int b : fadj[f]) hop2[f] |= hop1[b];
        }
        ctx->leaf_reach.assign((size_t)L, 0u);
        for (int l = 0; l < L; l++) {
          for (int k = 0; k < leaf_pcnt[l]; k++) ctx->leaf_reach[l] |= hop2[grid_face_h[pb->prim_indices[leaf_pbeg[l] + k]]];
        }
        /* The skip of a dab's halo exchanges must come out the same on every rank, so "near a cut" has to be the same table
         * everywhere: a leaf is near when gathering it can affect any OTHER rank (the halo_grid test above only knows this
         * rank's halo -- with more than two ranks the tables differed and a rank could skip an exchange its peer ran). */
        std::vector<unsigned> near_all((size_t)m.ghit_words, 0u);
        for (int l = 0; l < L; l++) {
          if (ctx->leaf_reach[l] & (ctx->leaf_reach[l] - 1u)) near_all[l >> 5] |= 1u << (l & 31);
        }
        CU(cudaMemcpyAsync(ctx->d_near_mask, near_all.data(), sizeof(unsigned) * near_all.size(), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
      }
      sidx.resize(se.size());
      ridx.resize(re.size());
      for (size_t i = 0; i < se.size(); i++) sidx[i] = ctx->slot_of[se[i]];
      for (size_t i = 0; i < re.size(); i++) ridx[i] = ctx->slot_of[re[i]];
      DevGrids &g = ctx->g;
      if ((r = dev_upload_c(ctx, &g.face_dom, plan.face_dom)) || (r = dev_upload_c(ctx, &g.edge_mine, plan.edge_mine)) ||
          (r = dev_upload_c(ctx, &g.cvert_mine, plan.cvert_mine)) || (r = dev_upload_c(ctx, &g.grid_owner, grid_owner)) ||
          (r = dev_zero(ctx, &ctx->d_glist, (size_t)L + 1)) || (r = dev_zero(ctx, &ctx->d_gcount, 1)))
        return r;
      g.rank = ctx->rank;
      ctx->grid_fused = false;
      if ((r = dev_upload(ctx, &ctx->d_send_idx, sidx)) || (r = dev_upload(ctx, &ctx->d_recv_idx, ridx)) ||
          (r = dev_alloc(ctx, &ctx->d_send_buf, 3 * sidx.size() + 1)) || (r = dev_alloc(ctx, &ctx->d_recv_buf, 3 * ridx.size() + 1)))
        return r;
      CU(cudaStreamSynchronize(ctx->stream));
    }
    else {
    DscMeshDesc me;
    memset(&me, 0, sizeof(me));
    me.totvert = V;
    me.totpoly = ctx->totpoly;
    me.totloop = ctx->totloop;
    me.tottri = T;
    me.poly_loopstart = ctx->h_poly_start.data();
    me.poly_totloop = ctx->h_poly_len.data();
    me.loop_vert = ctx->h_loop_v.data();
    me.tri_vert = ctx->h_tri_vert.data();
    me.tri_poly = ctx->h_tri_poly.data();
    me.nb_offsets = ctx->has_nb ? ctx->h_nb_off.data() : nullptr;
    me.nb_indices = ctx->has_nb ? ctx->h_nb_idx.data() : nullptr;
    std::vector<HaloTriple> tr;
    plan_halo(&me, pb, ctx->world, pl, ctx->leaf_range, tr);
    {
      /* leaves that own a vertex some other rank reads: only a dab that gathers one of them changes a halo vertex */
      std::vector<unsigned> near_mask((size_t)m.ghit_words, 0u);
      for (const HaloTriple &t3 : tr) {
        const int sl = ctx->slot_of[t3.vert];
        const int l = (int)(std::upper_bound(leaf_ubeg.begin(), leaf_ubeg.end(), sl) - leaf_ubeg.begin()) - 1;
        if (l >= 0 && l < L) near_mask[l >> 5] |= 1u << (l & 31);
      }
      if ((r = dev_upload(ctx, &ctx->d_near_mask, near_mask))) return r;
      /* which ranks gathering a leaf can affect: its owner and the ranks that read one of its vertices */
      ctx->leaf_reach.assign((size_t)L, 0u);
      for (int q = 0; q < ctx->world; q++) {
        for (int l = ctx->leaf_range[q]; l < ctx->leaf_range[q + 1]; l++) ctx->leaf_reach[l] |= 1u << q;
      }
      for (const HaloTriple &t3 : tr) {
        const int sl = ctx->slot_of[t3.vert];
        const int l = (int)(std::upper_bound(leaf_ubeg.begin(), leaf_ubeg.end(), sl) - leaf_ubeg.begin()) - 1;
        if (l >= 0 && l < L) ctx->leaf_reach[l] |= 1u << t3.reader;
      }
    }
    ctx->send_off.assign(ctx->world + 1, 0);
    ctx->recv_off.assign(ctx->world + 1, 0);
    for (int q = 0; q < ctx->world; q++) {
      ctx->send_off[q] = (int)sidx.size();
      ctx->recv_off[q] = (int)ridx.size();
      for (const HaloTriple &t : tr) {
        if (t.owner == ctx->rank && t.reader == q) sidx.push_back(ctx->slot_of[t.vert]);
        if (t.reader == ctx->rank && t.owner == q) ridx.push_back(ctx->slot_of[t.vert]);
      }
    }
    ctx->send_off[ctx->world] = (int)sidx.size();
    ctx->recv_off[ctx->world] = (int)ridx.size();
    if ((r = dev_upload(ctx, &ctx->d_send_idx, sidx)) || (r = dev_upload(ctx, &ctx->d_recv_idx, ridx)) ||
        (r = dev_alloc(ctx, &ctx->d_send_buf, 3 * sidx.size() + 1)) || (r = dev_alloc(ctx, &ctx->d_recv_buf, 3 * ridx.size() + 1)))
      return r;
    CU(cudaStreamSynchronize(ctx->stream));
    }
    if ((r = dist_p2p_setup(ctx))) return r;
  }

  DSC_TMARK("multi-GPU tables");
  /* the host staging copies are no longer needed */
  std::vector<int>().swap(ctx->h_rim_nb);
  std::vector<float>().swap(ctx->h_co);
  std::vector<float>().swap(ctx->h_no);
  std::vector<float>().swap(ctx->h_mask);
  std::vector<unsigned>().swap(ctx->h_tail);
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

/* which leaf list a stage walks: the hit list of a dab, or -- when leaves may carry flags from
 * elsewhere -- the list k_collect_flagged builds */
struct LeafList {
  const int *list;
  const int *count;
  const int4 *tiles; /* the tiles of the listed leaves */
  const int *tile_count;
  const unsigned *mask; /* bit per leaf: updates its normals with this list */
};
static LeafList hit_list(DscContext *ctx, int slot)
{
  DevMesh &m = ctx->m;
  return {m.hit_list + (size_t)slot * m.nleaf, &m.st[slot].hit_count, m.tile_list + (size_t)slot * m.ntile, &m.st[slot].tile_count,
          m.ghit + (size_t)slot * m.ghit_words};
}
static LeafList flag_list(DscContext *ctx)
{
  DevMesh &m = ctx->m;
  return {m.flag_list, &m.tot->flag_count, m.flag_tile_list, &m.tot->flag_tiles, m.flag_mask};
}

static int run_collect(DscContext *ctx, int flags)
{
  StageScope s(ctx, ST_OTHER);
  k_collect_flagged<<<1, 1024, 0, ctx->stream>>>(ctx->m, flags);
  LAUNCH_CHECK();
  return DSC_OK;
}
/* normals and/or leaf boxes of the listed leaves (mode: NB_NORMALS | NB_BOUNDS) */
/* the leaves that do not fit the shared-memory tile kernel (n-gons, oversized tiles): general gather kernels */
static int run_slow_leaves(DscContext *ctx, LeafList ll, int mode)
{
  if (mode & NB_NORMALS) {
    StageScope s(ctx, ST_NORMALS);
    k_normals<<<ctx->grid, DSC_BLOCK, 0, ctx->stream>>>(ctx->m, ll.list, ll.count, 1, ll.mask);
    LAUNCH_CHECK();
  }
  if (mode & NB_BOUNDS) {
    StageScope s(ctx, ST_LEAFBB);
    k_leaf_bb<<<ctx->grid, DSC_BLOCK, 0, ctx->stream>>>(ctx->m, ll.list, ll.count, 1);
    LAUNCH_CHECK();
  }
  return DSC_OK;
}
static int run_normals_bounds(DscContext *ctx, LeafList ll, int mode, bool pdl = false)
{
  {
    StageScope s(ctx, ST_NORMALS);
    CU(launch_k(k_normals_tile, ctx->nb_grid, NT_THREADS, ctx->nb_smem, ctx->stream, pdl, ctx->m, ll.tiles, ll.tile_count, mode, ll.mask));
  }
  if (ctx->any_slow_leaf) return run_slow_leaves(ctx, ll, mode);
  return DSC_OK;
}
static int run_clear(DscContext *ctx, LeafList ll, int clear_mask)
{
  StageScope s(ctx, ST_OTHER);
  k_clear_flags<<<ctx->num_sms, DSC_BLOCK, 0, ctx->stream>>>(ctx->m, ll.list, ll.count, clear_mask);
  LAUNCH_CHECK();
  return DSC_OK;
}
static int run_flush_full(DscContext *ctx)
{
  StageScope s(ctx, ST_FLUSH);
  k_flush<<<1, 1024, 0, ctx->stream>>>(ctx->m);
  LAUNCH_CHECK();
  return DSC_OK;
}
/* the deferred refit of the inner nodes: every inner box = the union of its children's, level by level -- the values
 * pbvh_flush_bb (pbvh.c:3287-3317) leaves after each dab, because a box nobody refreshed already is that union */
static int ensure_refit(DscContext *ctx)
{
  if (!ctx->refit_pending) return DSC_OK;
  int r = join_side(ctx);
  if (r) return r;
  ctx->refit_pending = false;
  ctx->launches++;
  return run_flush_full(ctx);
}
/* flagged leaves -> normals / boxes -> whole-tree flush -> clear: every non-dab entry point */
static int run_flagged(DscContext *ctx, int want)
{
  int r;
  if ((r = join_side(ctx))) return r;
  if ((r = run_collect(ctx, want))) return r;
  const int mode = ((want & F_UpdateNormals) ? NB_NORMALS : 0) | ((want & F_UpdateBB) ? NB_BOUNDS : 0);
  if (want & F_UpdateBB) {
    StageScope s(ctx, ST_OTHER);
    k_reset_leaf_boxes<<<ctx->num_sms, DSC_BLOCK, 0, ctx->stream>>>(ctx->m, ctx->m.flag_list, &ctx->m.tot->flag_count, F_UpdateBB);
    LAUNCH_CHECK();
  }
  if ((r = run_normals_bounds(ctx, flag_list(ctx), mode))) return r;
  if ((want & F_UpdateBB) && (r = run_flush_full(ctx))) return r;
  return run_clear(ctx, flag_list(ctx), want);
}


/* maps every peer's exchange region (flags | reduce inbox | halo inbox) into this process: cudaIpc handles travel
 * through NCCL broadcasts, as do the peers' receive offsets.  On any refusal the context keeps using NCCL. */
static int dist_p2p_setup(DscContext *ctx)
{
  const int W = ctx->world;
  ctx->p2p = false;
  if (W < 2 || W > DSC_MAX_RANKS || getenv("DSC_NO_P2P")) return DSC_OK;
  const int words = ctx->m.ghit_words;
  const int red_stride = 16 + (words + 1) / 2 + 2; /* 8-byte words per source rank: 16 sums, the bitmask, padding */
  const size_t flag_bytes = 256, red_half_bytes = ((size_t)W * red_stride * 8 + 255) & ~(size_t)255, red_bytes = 2 * red_half_bytes;
  /* everyone allocates an inbox for the largest receive list so the regions have one layout */
  int my_recv = ctx->recv_off[W];
  int *d_meta = nullptr;
  int r;
  if ((r = dev_zero(ctx, &d_meta, (size_t)W * (W + 2) + 16 * (size_t)W * 2))) return r;
  std::vector<int> meta((size_t)W + 2);
  for (int q = 0; q <= W; q++) meta[q] = ctx->recv_off[q];
  meta[W + 1] = my_recv;
  CU(cudaMemcpyAsync(d_meta + (size_t)ctx->rank * (W + 2), meta.data(), sizeof(int) * (W + 2), cudaMemcpyHostToDevice, ctx->stream));
  NC(g_nccl.GroupStart());
  for (int q = 0; q < W; q++) NC(g_nccl.Broadcast(d_meta + (size_t)q * (W + 2), d_meta + (size_t)q * (W + 2), (size_t)(W + 2), ncclInt32, q, ctx->comm, ctx->stream));
  NC(g_nccl.GroupEnd());
  std::vector<int> all((size_t)W * (W + 2));
  CU(cudaMemcpyAsync(all.data(), d_meta, sizeof(int) * all.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  int max_recv = 0;
  for (int q = 0; q < W; q++) max_recv = std::max(max_recv, all[(size_t)q * (W + 2) + W + 1]);
  const size_t inbox_half_bytes = ((size_t)3 * max_recv * sizeof(float) + 255) & ~(size_t)255, inbox_bytes = 2 * inbox_half_bytes;
  const size_t total = flag_bytes + red_bytes + inbox_bytes + 256;
  char *region = nullptr;
  if (cudaMalloc((void **)&region, total) != cudaSuccess) {
    cudaGetLastError();
    return DSC_OK;
  }
  ctx->allocs.push_back(region);
  ctx->p2p_region = region;
  CU(cudaMemsetAsync(region, 0, total, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  cudaIpcMemHandle_t mine;
  int ok = cudaIpcGetMemHandle(&mine, region) == cudaSuccess ? 1 : 0;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  int *d_handles = d_meta + (size_t)W * (W + 2); /* [W][16 ints] handles, then [W] ok flags packed behind */
  CU(cudaMemcpyAsync(d_handles + 16 * (size_t)ctx->rank, &mine, 64, cudaMemcpyHostToDevice, ctx->stream));
  NC(g_nccl.GroupStart());
  for (int q = 0; q < W; q++) NC(g_nccl.Broadcast(d_handles + 16 * (size_t)q, d_handles + 16 * (size_t)q, 16, ncclInt32, q, ctx->comm, ctx->stream));
  NC(g_nccl.GroupEnd());
  std::vector<cudaIpcMemHandle_t> handles((size_t)W);
  CU(cudaMemcpyAsync(handles.data(), d_handles, 64 * (size_t)W, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int q = 0; q < W && ok; q++) {
    if (q == ctx->rank) continue;
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, handles[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
      break;
    }
    ctx->p2p_peer_region[q] = p;
  }
  /* all ranks or none */
  int *d_ok = d_handles + 16 * (size_t)W;
  CU(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  NC(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt32, ncclSum, ctx->comm, ctx->stream));
  int sum = 0;
  CU(cudaMemcpyAsync(&sum, d_ok, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (sum != W) return DSC_OK;
  PeerLink &L = ctx->link;
  memset(&L, 0, sizeof(L));
  L.world = W;
  L.rank = ctx->rank;
  L.red_stride = red_stride;
  L.red_half = (int)(red_half_bytes / 8);
  L.inbox_half = (int)(inbox_half_bytes / sizeof(float));
  for (int q = 0; q < W; q++) {
    char *base = q == ctx->rank ? region : (char *)ctx->p2p_peer_region[q];
    L.peer_flags[q] = (int *)base;
    L.peer_red[q] = (long long *)(base + flag_bytes);
    L.peer_inbox[q] = (float *)(base + flag_bytes + red_bytes);
    L.peer_off[q] = all[(size_t)q * (W + 2) + ctx->rank];
  }
  L.flags = L.peer_flags[ctx->rank];
  L.red = L.peer_red[ctx->rank];
  L.inbox = L.peer_inbox[ctx->rank];
  for (int q = 0; q <= W; q++) {
    L.send_off[q] = ctx->send_off[q];
    L.recv_off[q] = ctx->recv_off[q];
  }
  ctx->p2p = true;
  /* dabs are exchanged among the ranks they reach only (DSC_DIST_ALL=1: among all ranks, the round-1 protocol) */
  ctx->subset_exchange = getenv("DSC_DIST_ALL") == nullptr;
  return DSC_OK;
}

/* ---- multi-GPU steps of a dab ---- */
/* one all-reduce pair per dab: the exact area sums and the bitmask of gathered leaves (disjoint per
 * rank, so sum == or) */
static int dist_allreduce_dab(DscContext *ctx, int j, int slot, bool with_area)
{
  if (ctx->p2p) {
    unsigned *gh = ctx->m.ghit + (size_t)slot * ctx->m.ghit_words;
    k_p2p_reduce_push<<<dim3(1, ctx->world), 256, 0, ctx->stream>>>(ctx->m, j, ctx->link, ctx->m.st[slot].acc, gh, ctx->m.ghit_words, with_area ? 1 : 0);
    LAUNCH_CHECK();
    k_p2p_reduce_recv<<<1, 256, 0, ctx->stream>>>(ctx->m, j, ctx->link, ctx->m.st[slot].acc, gh, ctx->m.ghit_words, with_area ? 1 : 0, ctx->d_near_mask);
    LAUNCH_CHECK();
    ctx->launches += 2;
    return DSC_OK;
  }
  NC(g_nccl.GroupStart());
  if (with_area) {
    NC(g_nccl.AllReduce(ctx->m.st[slot].acc, ctx->m.st[slot].acc, 16, ncclInt64, ncclSum, ctx->comm, ctx->stream));
  }
  unsigned *gh = ctx->m.ghit + (size_t)slot * ctx->m.ghit_words;
  NC(g_nccl.AllReduce(gh, gh, (size_t)ctx->m.ghit_words, ncclUint32, ncclSum, ctx->comm, ctx->stream));
  NC(g_nccl.GroupEnd());
  ctx->launches++;
  return DSC_OK;
}
/* one-ring halo: owners push the positions other ranks' leaves read */
static int dist_halo_exchange(DscContext *ctx, int j, bool normals = false, int cond = 1)
{
  const int W = ctx->world;
  float *ax = normals ? ctx->m.nx : ctx->m.cx, *ay = normals ? ctx->m.ny : ctx->m.cy, *az = normals ? ctx->m.nz : ctx->m.cz;
  if (ctx->p2p) {
    /* gather from the arrays straight into the peers' inboxes, then scatter what arrived */
    int most = 1;
    for (int q = 0; q < W; q++) most = std::max(most, std::max(ctx->send_off[q + 1] - ctx->send_off[q], ctx->recv_off[q + 1] - ctx->recv_off[q]));
    const int ctas = std::max(1, std::min((most + 1023) / 1024, std::max(1, ctx->num_sms / (2 * W))));
    k_p2p_halo_push<<<dim3(ctas, W), 256, 0, ctx->stream>>>(ctx->m, j, ctx->link, ctx->d_near_mask ? cond : 0, ctx->d_send_idx, ax, ay, az);
    LAUNCH_CHECK();
    k_p2p_halo_recv<<<dim3(ctas, W), 256, 0, ctx->stream>>>(ctx->m, j, ctx->link, ctx->d_near_mask ? cond : 0, ctx->d_recv_idx, ax, ay, az);
    LAUNCH_CHECK();
    ctx->launches += 2;
    return DSC_OK;
  }
  const int ns = ctx->send_off[W], nr = ctx->recv_off[W];
  /* per peer the buffer holds [3][count] */
  for (int q = 0; q < W; q++) {
    const int n = ctx->send_off[q + 1] - ctx->send_off[q];
    if (n) {
      k_halo_pack<<<std::min((n + 255) / 256, ctx->num_sms * 4), 256, 0, ctx->stream>>>(
          ctx->d_send_buf + 3 * (size_t)ctx->send_off[q], ctx->d_send_idx + ctx->send_off[q], n, ax, ay, az);
      LAUNCH_CHECK();
      ctx->launches++;
    }
  }
  if (ns || nr) {
    NC(g_nccl.GroupStart());
    for (int q = 0; q < W; q++) {
      const int n = ctx->send_off[q + 1] - ctx->send_off[q], k = ctx->recv_off[q + 1] - ctx->recv_off[q];
      if (n) NC(g_nccl.Send(ctx->d_send_buf + 3 * (size_t)ctx->send_off[q], 3 * (size_t)n, ncclFloat, q, ctx->comm, ctx->stream));
      if (k) NC(g_nccl.Recv(ctx->d_recv_buf + 3 * (size_t)ctx->recv_off[q], 3 * (size_t)k, ncclFloat, q, ctx->comm, ctx->stream));
    }
    NC(g_nccl.GroupEnd());
  }
  for (int q = 0; q < W; q++) {
    const int k = ctx->recv_off[q + 1] - ctx->recv_off[q];
    if (k) {
      k_halo_unpack<<<std::min((k + 255) / 256, ctx->num_sms * 4), 256, 0, ctx->stream>>>(
          ctx->d_recv_buf + 3 * (size_t)ctx->recv_off[q], ctx->d_recv_idx + ctx->recv_off[q], k, ax, ay, az);
      LAUNCH_CHECK();
      ctx->launches++;
    }
  }
  return DSC_OK;
}
/* every rank broadcasts what it owns, then each replica flushes the whole tree: the BB-root reduction of SURVEY.md section 8e.
 * Stroke end moves the leaf boxes, leaf flags and stroke state only (32 bytes per leaf); the vertex data stays with its
 * owner -- a rank's host side takes its own runs (dsc_download_owned_*) -- unless a whole replica is asked for
 * (dsc_dist_gather: tests, digests). */
static int dist_gather_all(DscContext *ctx, bool vertex_data)
{
  DevMesh &m = ctx->m;
  const int N = ctx->totnode;
  float *vert_arrays[12] = {m.cx, m.cy, m.cz, m.nx, m.ny, m.nz, m.ox, m.oy, m.oz, m.onx, m.ony, m.onz};
  NC(g_nccl.GroupStart());
  for (int q = 0; q < ctx->world; q++) {
    const int s0 = ctx->slot_range[q], sn = ctx->slot_range[q + 1] - s0;
    const int l0 = ctx->leaf_range[q], ln = ctx->leaf_range[q + 1] - l0;
    if (sn > 0 && vertex_data) {
      for (int a = 0; a < 12; a++) NC(g_nccl.Broadcast(vert_arrays[a] + s0, vert_arrays[a] + s0, (size_t)sn, ncclFloat, q, ctx->comm, ctx->stream));
      /* grids: the stitch averages the mask layer too */
      if (ctx->is_grids && ctx->g.mask) NC(g_nccl.Broadcast(ctx->g.mask + s0, ctx->g.mask + s0, (size_t)sn, ncclFloat, q, ctx->comm, ctx->stream));
    }
    if (ln > 0) {
      for (int k = 0; k < 6; k++) NC(g_nccl.Broadcast(m.bb + (size_t)k * N + l0, m.bb + (size_t)k * N + l0, (size_t)ln, ncclFloat, q, ctx->comm, ctx->stream));
      NC(g_nccl.Broadcast(m.node_flag + l0, m.node_flag + l0, (size_t)ln, ncclInt32, q, ctx->comm, ctx->stream));
      NC(g_nccl.Broadcast(m.leaf_state + l0, m.leaf_state + l0, (size_t)ln, ncclUint32, q, ctx->comm, ctx->stream));
    }
  }
  NC(g_nccl.GroupEnd());
  StageScope s(ctx, ST_FLUSH);
  k_flush<<<1, 1024, 0, ctx->stream>>>(ctx->m);
  LAUNCH_CHECK();
  return DSC_OK;
}

/* ---- multires grids: the stages after the brush ---- */
static int grids_reset_counts(DscContext *ctx)
{
  CU(cudaMemsetAsync(ctx->g.cnt, 0, sizeof(GridCounts), ctx->stream));
  return DSC_OK;
}
/* KERNEL_subdiv_ccg_average_grids (subdiv_ccg.c:1170-1189): every face, edge and vertex */
static int grids_average_all(DscContext *ctx)
{
  DevGrids g = ctx->g;
  g.face_dom = g.edge_mine = g.cvert_mine = nullptr; /* every replica averages everything */
  g.grid_owner = nullptr;
  std::vector<int> iota((size_t)g.totface);
  for (int f = 0; f < g.totface; f++) iota[f] = f;
  GridCounts c = {g.totface, 0, 0, 0};
  CU(cudaMemcpyAsync(g.face_list, iota.data(), sizeof(int) * iota.size(), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(g.cnt, &c, sizeof(c), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  k_grid_inner<<<ctx->num_sms * 4, 128, 0, ctx->stream>>>(ctx->m, g);
  LAUNCH_CHECK();
  k_grid_edges<<<ctx->num_sms * 4, 128, 0, ctx->stream>>>(ctx->m, g, 2, 0);
  LAUNCH_CHECK();
  k_grid_cverts<<<ctx->num_sms, DSC_BLOCK, 0, ctx->stream>>>(ctx->m, g, 1);
  LAUNCH_CHECK();
  ctx->launches += 3;
  return DSC_OK;
}
/* KERNEL_subdiv_ccg_recalc_normals (subdiv_ccg.c:782-790) */
static int grids_recalc_normals(DscContext *ctx)
{
  k_grid_normals<<<ctx->num_sms * 2, GN_BLOCK, ctx->gn_smem, ctx->stream>>>(ctx->m, ctx->g, 1);
  LAUNCH_CHECK();
  ctx->launches++;
  return grids_average_all(ctx);
}
/* after the brush of a dab on grids: stitch, CCG normals of the gathered leaves' faces, leaf boxes */
static int grids_after_brush(DscContext *ctx, LeafList hits, int j, bool exch)
{
  DevGrids &g = ctx->g;
  DevMesh &m = ctx->m;
  cudaStream_t st = ctx->stream;
  int r;
  const int seq = j; /* the kernels turn the dab's position in the running batch into its stamp (dsc_grid_seq) */
  if ((r = grids_reset_counts(ctx))) return r;
  StageScope s(ctx, ST_NORMALS);
  if (ctx->grid_fused) {
    /* one cooperative launch, the stages separated by grid barriers */
    CU(cudaMemsetAsync(m.grid_bar, 0, sizeof(unsigned), st));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)ctx->num_sms, 1, 1);
    cfg.blockDim = dim3(GN_BLOCK, 1, 1);
    cfg.dynamicSmemBytes = ctx->gn_smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CU(cudaLaunchKernelEx(&cfg, k_grid_dab, m, g, hits.list, hits.count, seq));
    return DSC_OK;
  }
  const bool dist = ctx->world > 1;
  /* how many leaves the dab gathered (on all ranks): none -> the reference's caller returns before the stitch */
  const int *nhits = dist ? ctx->d_gcount : hits.count;
  if (dist) {
    /* the faces of the leaves ALL ranks gathered (the all-reduced bitmask), restricted to what this rank averages */
    CU(cudaMemsetAsync(ctx->d_gcount, 0, sizeof(int), st));
    k_ghit_expand<<<std::max(1, std::min(ctx->num_sms, (m.ghit_words + DSC_BLOCK - 1) / DSC_BLOCK)), DSC_BLOCK, 0, st>>>(m, hits.mask, ctx->d_glist, ctx->d_gcount);
    LAUNCH_CHECK();
    k_grid_faces<<<ctx->num_sms * 2, DSC_BLOCK, 0, st>>>(m, g, ctx->d_glist, ctx->d_gcount, seq);
    ctx->launches++;
  }
  else k_grid_faces<<<ctx->num_sms * 2, DSC_BLOCK, 0, st>>>(m, g, hits.list, hits.count, seq);
  LAUNCH_CHECK();
  /* multires_stitch_grids (multires.c:1171-1196 -> subdiv_ccg.c:1303-1324): the faces' inner boundaries, then
   * all coarse edges (two-face edges no dab touched average to themselves: only the touched ones and those with
   * more faces run) and all coarse vertices */
  k_grid_inner<<<ctx->num_sms * 4, 128, 0, st>>>(m, g);
  LAUNCH_CHECK();
  k_grid_edges_cverts<<<ctx->num_sms * 5, 128, 0, st>>>(m, g, ctx->has_odd_edges ? 1 : 0, 1, seq, ctx->num_sms * 4, nhits);
  LAUNCH_CHECK();
  /* BKE_pbvh_update_normals, PBVH_GRIDS branch (pbvh.c:4575-4583 -> subdiv_ccg.c:847-866) */
  if (ctx->grid_normals_flat) k_grid_normals_flat<<<ctx->num_sms * 8, 256, 0, st>>>(m, g, 0);
  else k_grid_normals<<<ctx->num_sms * 2, GN_BLOCK, ctx->gn_smem, st>>>(m, g, 0);
  LAUNCH_CHECK();
  if (exch && (r = dist_halo_exchange(ctx, j, true))) return r; /* the new normals of the other ranks' halo elements */
  k_grid_inner<<<ctx->num_sms * 4, 128, 0, st>>>(m, g);
  LAUNCH_CHECK();
  k_grid_edges_cverts<<<ctx->num_sms * 5, 128, 0, st>>>(m, g, 0, 0, seq, ctx->num_sms * 4, nhits);
  LAUNCH_CHECK();
  /* BKE_pbvh_update_bounds: leaf boxes (the refit follows on the side stream) */
  k_grid_leaf_bb<<<ctx->num_sms * 4, DSC_BLOCK, 0, st>>>(m, hits.list, hits.count);
  LAUNCH_CHECK();
  ctx->launches += 7;
  return DSC_OK;
}

int dsc_recalc_normals(DscContext *ctx)
{
  NEED_PBVH();
  if (ctx->is_grids) return grids_recalc_normals(ctx);
  {
    StageScope s(ctx, ST_OTHER);
    const int n = std::max(ctx->m.nleaf, 1);
    k_mark_all<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->m, F_UpdateNormals, 1, ctx->nwords);
    LAUNCH_CHECK();
  }
  return run_flagged(ctx, F_UpdateNormals);
}

int dsc_set_custom_curve(DscContext *ctx, const float *table257)
{
  NEED_PBVH();
  invalidate_graphs(ctx); /* the mesh descriptor is baked into the captured launches */
  if (!table257) {
    ctx->m.curve = nullptr;
    return DSC_OK;
  }
  CU(cudaMemcpyAsync(ctx->d_curve, table257, 257 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->m.curve = ctx->d_curve;
  return DSC_OK;
}

/* a per-vertex float layer (mask, automask factors) into slot order: one H2D copy in vertex order into the staging buffer,
 * permuted on the device (the pad slots of dst stay zero from their allocation) */
__global__ void k_scatter1(float *dst, const float *src, const int *slot_of, int totvert)
{
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < totvert; v += gridDim.x * blockDim.x) dst[slot_of[v]] = src[v];
}
static int upload_per_vertex(DscContext *ctx, float *dst, const float *src)
{
  int r = join_side(ctx);
  if (r) return r;
  /* stream order keeps the staging buffer safe: every earlier export / import through it was queued on this stream */
  CU(cudaMemcpyAsync(ctx->d_stage3, src, sizeof(float) * (size_t)ctx->totvert, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream)); /* the caller's array is free again when the call returns (it may be page-locked) */
  k_scatter1<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(dst, ctx->d_stage3, ctx->d_slot_of, ctx->totvert);
  LAUNCH_CHECK();
  ctx->launches++;
  return DSC_OK;
}

int dsc_set_mask(DscContext *ctx, const float *mask)
{
  NEED_PBVH();
  /* graphs are keyed by the presence of the layer (its device buffer never moves): nothing to invalidate */
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
  const int dn = ctx->dev_of_node[node];
  int r = sync_all(ctx);
  if (r) return r;
  CU(cudaMemcpy(&f, ctx->m.node_flag + dn, sizeof(int), cudaMemcpyDeviceToHost));
  f = on ? (f | flag) : (f & ~flag);
  CU(cudaMemcpy(ctx->m.node_flag + dn, &f, sizeof(int), cudaMemcpyHostToDevice));
  if (on && (flag & (F_UpdateNormals | F_UpdateBB))) ctx->stale_flags = true;
  return DSC_OK;
}

__global__ void k_node_flags_apply(int *node_flag, const int *ops, int count)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int n = ops[3 * i];
  node_flag[n] = (node_flag[n] | ops[3 * i + 1]) & ~ops[3 * i + 2];
}
__global__ void k_vert_marks_or(unsigned *dirty, const unsigned *bitmap, const int *slot_of, int totvert)
{
  const int nwords = (totvert + 31) >> 5;
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += gridDim.x * blockDim.x) {
    unsigned bits = bitmap[w];
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      const int v = (w << 5) + b;
      if (v < totvert) {
        const int s = slot_of[v];
        atomicOr(&dirty[s >> 5], 1u << (s & 31));
      }
    }
  }
}

int dsc_node_flags_apply(DscContext *ctx, int count, const int *nodes, const int *set_bits, const int *clear_bits)
{
  NEED_PBVH();
  if (count <= 0) return DSC_OK;
  if (!nodes || (!set_bits && !clear_bits)) return fail(ctx, DSC_ERR_INVALID, "node list is NULL");
  std::vector<int> ops((size_t)3 * count);
  int any_set = 0;
  for (int i = 0; i < count; i++) {
    if (nodes[i] < 0 || nodes[i] >= ctx->totnode) return fail(ctx, DSC_ERR_INVALID, "node %d out of range", nodes[i]);
    ops[(size_t)3 * i] = ctx->dev_of_node[nodes[i]];
    ops[(size_t)3 * i + 1] = set_bits ? set_bits[i] : 0;
    ops[(size_t)3 * i + 2] = clear_bits ? clear_bits[i] : 0;
    any_set |= ops[(size_t)3 * i + 1];
  }
  int r = join_side(ctx);
  if (r) return r;
  int *d_ops = nullptr;
  CU(cudaMallocAsync((void **)&d_ops, sizeof(int) * ops.size(), ctx->stream));
  /* pageable source: the copy is staged before the call returns, so `ops` may go out of scope */
  CU(cudaMemcpyAsync(d_ops, ops.data(), sizeof(int) * ops.size(), cudaMemcpyHostToDevice, ctx->stream));
  k_node_flags_apply<<<(count + 255) / 256, 256, 0, ctx->stream>>>(ctx->m.node_flag, d_ops, count);
  LAUNCH_CHECK();
  CU(cudaFreeAsync(d_ops, ctx->stream));
  ctx->launches++;
  /* leaves now carry flags no dab of the running stroke set: the per-dab stages walk the flags, not only their own hit list */
  if (any_set & (F_UpdateNormals | F_UpdateBB)) ctx->stale_flags = true;
  return DSC_OK;
}

int dsc_vert_marks_or(DscContext *ctx, const unsigned int *bitmap)
{
  NEED_PBVH();
  if (!bitmap) return fail(ctx, DSC_ERR_INVALID, "bitmap is NULL");
  const size_t nwords = ((size_t)ctx->totvert + 31) >> 5;
  unsigned *d_bits = nullptr;
  CU(cudaMallocAsync((void **)&d_bits, sizeof(unsigned) * nwords, ctx->stream));
  CU(cudaMemcpyAsync(d_bits, bitmap, sizeof(unsigned) * nwords, cudaMemcpyHostToDevice, ctx->stream));
  k_vert_marks_or<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->m.dirty, d_bits, ctx->d_slot_of, ctx->totvert);
  LAUNCH_CHECK();
  CU(cudaFreeAsync(d_bits, ctx->stream));
  ctx->launches++;
  return DSC_OK;
}

int dsc_node_mark_update(DscContext *ctx, int node)
{
  return dsc_node_flag_set(ctx, node,
                           F_UpdateNormals | F_UpdateBB | F_UpdateOriginalBB | F_UpdateDrawBuffers | F_UpdateRedraw, 1);
}

static int dist_regions_refresh(DscContext *ctx);
static int dist_flush_skipped(DscContext *ctx);

int dsc_stroke_begin(DscContext *ctx, const float *automask)
{
  NEED_PBVH();
  if (ctx->in_stroke) return fail(ctx, DSC_ERR_STATE, "stroke already open");
  int r = join_side(ctx);
  if (r) return r;
  if (automask) {
    if ((r = upload_per_vertex(ctx, ctx->d_automask, automask))) return r;
    ctx->m.automask = ctx->d_automask;
  }
  else {
    ctx->m.automask = nullptr;
  }
  CU(cudaMemsetAsync(ctx->m.leaf_state, 0, sizeof(unsigned) * (size_t)std::max(ctx->m.nleaf, 1), ctx->stream));
  CU(cudaMemsetAsync(ctx->m.st, 0, sizeof(DabState) * DSC_SLOTS, ctx->stream));
  CU(cudaMemsetAsync(ctx->m.tot, 0, sizeof(StrokeTotals), ctx->stream));
  if (ctx->world > 1 && ctx->subset_exchange && (r = dist_regions_refresh(ctx))) return r;
  ctx->pending_skipped = 0;
  ctx->last_dab_skipped = false;
  ctx->launches = 0;
  ctx->dab_index = 0;
  ctx->last_slot = 0;
  ctx->in_stroke = true;
  return DSC_OK;
}

/* ---- one dab = a fixed launch sequence over the device-resident dab ring ---- */
/* what decides the launch sequence of a dab (everything else is data in its ring entry) */
struct DabSig {
  int tool, needs_area, do_normals, do_bounds, smooth_iters, smooth_tail;
  int exch; /* partitioned PBVH: the dab is exchanged with other ranks */
  int small; /* a small dab: its whole path runs in the persistent kernel on a few CTAs (launch gaps dominate it otherwise) */
  unsigned key(int batch) const
  {
    return (unsigned)tool | (unsigned)needs_area << 8 | (unsigned)do_normals << 9 | (unsigned)do_bounds << 10 |
           (unsigned)smooth_iters << 11 | (unsigned)smooth_tail << 15 | (unsigned)batch << 16 | (unsigned)small << 26 | (unsigned)exch << 27;
  }
  bool operator==(const DabSig &o) const { return key(0) == o.key(0); }
};

static int make_entry(DscContext *ctx, const DscDab *dab, DabEntry *e, DabSig *sig, unsigned peers = 0u)
{
  if (!dab) return fail(ctx, DSC_ERR_INVALID, "dab is NULL");
  const int tool = dab->tool;
  if (tool != DSC_TOOL_DRAW && tool != DSC_TOOL_SMOOTH && tool != DSC_TOOL_INFLATE && tool != DSC_TOOL_GRAB &&
      tool != DSC_TOOL_CLAY_STRIPS)
    return fail(ctx, DSC_ERR_UNSUPPORTED, "sculpt tool %d is not on the accelerated path", tool);
  if (tool == DSC_TOOL_SMOOTH && !ctx->has_nb)
    return fail(ctx, DSC_ERR_STATE, ctx->is_grids ? "smooth brush needs the rim neighbour table (DscGridsDesc.rim_neighbors)" :
                                                    "smooth brush needs the neighbour CSR (DscMeshDesc.nb_offsets)");
  if (!(dab->radius > 0.0f)) return fail(ctx, DSC_ERR_INVALID, "radius must be positive");
  if (ctx->is_grids) {
    if (dab->flags & (DSC_DAB_NO_NORMALS | DSC_DAB_NO_BOUNDS))
      return fail(ctx, DSC_ERR_UNSUPPORTED, "on grids every dab stitches, updates normals and bounds");
  }
  static_assert(sizeof(DabParams) == sizeof(DscDab), "DabParams mirrors DscDab");
  static_assert(sizeof(DabEntry) == 160, "DabEntry is 160 bytes");
  memset(e, 0, sizeof(*e));
  memcpy(&e->d, dab, sizeof(e->d));
  sig->tool = tool;
  sig->do_normals = !(dab->flags & DSC_DAB_NO_NORMALS);
  sig->do_bounds = !(dab->flags & DSC_DAB_NO_BOUNDS);
  if (dab->falloff_shape != DSC_FALLOFF_SPHERE && dab->falloff_shape != DSC_FALLOFF_TUBE)
    return fail(ctx, DSC_ERR_INVALID, "falloff_shape %d", dab->falloff_shape);
  sig->needs_area = (tool == DSC_TOOL_DRAW && dab->sculpt_plane == DSC_DIR_AREA) || tool == DSC_TOOL_CLAY_STRIPS ||
                    (tool == DSC_TOOL_GRAB && dab->normal_weight > 0.0f && dab->sculpt_plane == DSC_DIR_AREA);
  sig->smooth_iters = 0;
  sig->smooth_tail = 0;
  sig->small = (ctx->use_batch_kernel && ctx->world == 1 && tool != DSC_TOOL_SMOOTH &&
                dab->radius * std::max(dab->radius_scale, 1.0f) <= ctx->small_dab_radius) ? 1 : 0;
  e->gather_resets = (ctx->lazy_refit && ctx->world == 1 && sig->do_bounds && !ctx->stale_flags) ? 1 : 0;
  e->peers = (int)peers;
  sig->exch = (ctx->world > 1 && (peers & ~(1u << ctx->rank))) ? 1 : 0;
  const float rs = dab->radius * dab->radius_scale;
  float ar = sqrtf(dab->radius * dab->radius); /* radius of the normal-sampling sphere, same float steps as k_area */
  ar *= dab->normal_radius_factor;
  e->radius_sq = rs * rs;
  e->area_radius_sq = ar * ar;
  e->original = tool == DSC_TOOL_GRAB ? 1 : 0;
  e->use_cos = tool == DSC_TOOL_CLAY_STRIPS ? 1 : 0;
  e->ent_bits = (sig->do_normals ? DSC_ENT_NORMALS : 0) | (sig->do_bounds ? DSC_ENT_BOUNDS : 0);
  /* flags this dab clears again before it ends are not set in the first place, unless the general
   * path (which reads them) has work or leaves may carry older flags */
  const int all_flags = F_UpdateNormals | F_UpdateBB | F_UpdateOriginalBB | F_UpdateDrawBuffers | F_UpdateRedraw;
  e->set_flags = all_flags;
  if (!ctx->stale_flags && !ctx->any_slow_leaf) {
    e->set_flags &= ~((sig->do_normals ? F_UpdateNormals : 0) | (sig->do_bounds ? F_UpdateBB : 0));
  }
  if (tool == DSC_TOOL_SMOOTH) {
    const int max_iterations = 4;
    const float fract = 1.0f / (float)max_iterations;
    float bstrength = dab->bstrength;
    bstrength = bstrength < 0.0f ? 0.0f : (bstrength > 1.0f ? 1.0f : bstrength);
    const int count = (int)(bstrength * (float)max_iterations);
    float last = (float)max_iterations * (bstrength - (float)count * fract);
    last = last < 0.0f ? 0.0f : (last > 1.0f ? 1.0f : last);
    e->smooth_last = last;
    sig->smooth_iters = count; /* full-strength iterations */
    /* a zero-strength tail iteration moves nothing and marks only verts the previous iteration
     * already marked; it is skipped */
    sig->smooth_tail = !(count > 0 && last == 0.0f);
  }
  return DSC_OK;
}

/* Queues the launch sequence of the j-th dab of a batch (ring entry ring_ctl[0] + j, state slot
 * `slot`).  `capturing`: the calls are being recorded into a CUDA graph. */
static int enqueue_dab(DscContext *ctx, const DabSig &sig, int j, int slot, bool capturing)
{
  DevMesh &m = ctx->m;
  cudaStream_t st = ctx->stream;
  const int tool = sig.tool;
  const bool do_normals = sig.do_normals, do_bounds = sig.do_bounds;
  const LeafList hits = hit_list(ctx, slot);
  /* the stages below walk the hit list unless some leaf may still carry flags of an earlier dab */
  const bool use_hits = !ctx->stale_flags;
  const bool dist = ctx->world > 1;
  const bool exch = dist && sig.exch; /* the dab reaches other ranks: reduce + halo exchanges among them */
  int r;
  cudaEvent_t ev_fork = capturing ? ctx->cap_fork : ctx->ev_fork, ev_bb = capturing ? ctx->cap_bb : ctx->ev_bb;
  cudaEvent_t ev_tag = capturing ? ctx->cap_tag : ctx->ev_tag;
  cudaEvent_t *ev_refit = capturing ? ctx->cap_refit : ctx->ev_refit;
  const bool pdl = ctx->use_pdl && ctx->stage_timing != 1 && ctx->stage_timing != 2 && !ctx->capture;
  /* one pass over the positions for brush + normals + boxes: meshes on one GPU whose leaves carry no older flags */
  const bool fused = ctx->use_fused && !ctx->is_grids && !dist && use_hits && tool != DSC_TOOL_SMOOTH;
  if (ctx->capture) CU(cudaMemsetAsync(ctx->d_capture, 0, sizeof(unsigned) * (size_t)ctx->nwords, st));

  /* 1. gather + undo membership + node marks.  It recycles the ring slot the refit of three dabs ago
   * read (inside a captured batch the first three dabs follow a joined side stream). */
  const bool eager = !ctx->lazy_refit && do_bounds && use_hits && !dist; /* tag + refit of this dab on the side stream */
  if (eager && (!capturing || j >= DSC_SLOTS - 1)) CU(cudaStreamWaitEvent(st, ev_refit[(slot + 1) & (DSC_SLOTS - 1)], 0));
  {
    StageScope s(ctx, ST_GATHER);
    CU(launch_k(k_gather_dab, (m.nleaf + DSC_BLOCK - 1) / DSC_BLOCK, DSC_BLOCK, 0, st, pdl, m, j, slot));
  }
  if (eager) {
    /* side stream: tag the ancestors of the hit leaves for the bottom-up refit while the brush runs */
    CU(cudaEventRecord(ev_fork, st));
    CU(cudaStreamWaitEvent(ctx->stream2, ev_fork, 0));
    {
      StageScope s(ctx, ST_FLUSH, ctx->stream2);
      k_tag_ancestors<<<ctx->num_sms, DSC_BLOCK, 0, ctx->stream2>>>(m, slot, 1);
      LAUNCH_CHECK();
    }
    CU(cudaEventRecord(ev_tag, ctx->stream2)); /* the tile kernel accumulates into the emptied leaf boxes */
    ctx->side_busy = true;
  }
  /* 2.-3. brush */
  if (tool == DSC_TOOL_SMOOTH) {
    /* partitioned grids: the halo elements the averaging groups of this rank hold are kept current by those groups, but
     * the smooth brush also reads neighbours that are in none of them (the next point along a coarse edge in the
     * sibling grid, one step inside another rank's grid): their owners' stitch may have moved them since the last
     * exchange, so the first iteration starts from a fresh halo */
    if (exch && ctx->is_grids && (r = dist_halo_exchange(ctx, j, false, 2))) return r;
    /* the bitmask of gathered leaves first: it tells every rank whether this dab's exchanges carry anything */
    if (exch && (r = dist_allreduce_dab(ctx, j, slot, false))) return r;
    {
      StageScope s(ctx, ST_SMOOTH);
      k_snapshot<<<ctx->grid, DSC_BLOCK, 0, st>>>(m, slot);
      LAUNCH_CHECK();
    }
    const int total = sig.smooth_iters + (sig.smooth_tail ? 1 : 0);
    for (int it = 0; it < total; it++) {
      {
        StageScope s(ctx, ST_SMOOTH);
        if (ctx->is_grids) k_smooth_a<true><<<ctx->grid, DSC_BLOCK, 0, st>>>(m, ctx->gnb, j, slot, it == sig.smooth_iters ? 1 : 0);
        else k_smooth_a<false><<<ctx->grid, DSC_BLOCK, 0, st>>>(m, ctx->gnb, j, slot, it == sig.smooth_iters ? 1 : 0);
        LAUNCH_CHECK();
      }
      {
        StageScope s(ctx, ST_SMOOTH);
        k_smooth_b<<<ctx->grid, DSC_BLOCK, 0, st>>>(m, slot);
        LAUNCH_CHECK();
      }
      if (exch && (r = dist_halo_exchange(ctx, j))) return r; /* the next iteration reads the ring */
    }
  }
  else {
    if (sig.needs_area) {
      StageScope s(ctx, ST_AREA);
      CU(launch_k(k_area, ctx->grid_area, DSC_BLOCK, 0, st, pdl, m, j, slot));
    }
    if (exch && (r = dist_allreduce_dab(ctx, j, slot, sig.needs_area))) return r;
    if (fused) {
      /* the boundary runs of the gathered tiles first: what other tiles read is displaced before the fused kernel starts */
      StageScope s(ctx, ST_BRUSH);
      switch (tool) {
        case DSC_TOOL_DRAW: k_brush_boundary<DSC_TOOL_DRAW><<<ctx->grid_bnd, DSC_BLOCK, 0, st>>>(m, j, slot); break;
        case DSC_TOOL_INFLATE: k_brush_boundary<DSC_TOOL_INFLATE><<<ctx->grid_bnd, DSC_BLOCK, 0, st>>>(m, j, slot); break;
        case DSC_TOOL_GRAB: k_brush_boundary<DSC_TOOL_GRAB><<<ctx->grid_bnd, DSC_BLOCK, 0, st>>>(m, j, slot); break;
        default: k_brush_boundary<DSC_TOOL_CLAY_STRIPS><<<ctx->grid_bnd, DSC_BLOCK, 0, st>>>(m, j, slot); break;
      }
      LAUNCH_CHECK();
    }
    else {
      StageScope s(ctx, ST_BRUSH);
      const bool bpdl = pdl && !dist;
      switch (tool) {
        case DSC_TOOL_DRAW: CU(launch_k(k_brush<DSC_TOOL_DRAW>, ctx->grid_brush, DSC_BLOCK, 0, st, bpdl, m, j, slot)); break;
        case DSC_TOOL_INFLATE: CU(launch_k(k_brush<DSC_TOOL_INFLATE>, ctx->grid_brush, DSC_BLOCK, 0, st, bpdl, m, j, slot)); break;
        case DSC_TOOL_GRAB: CU(launch_k(k_brush<DSC_TOOL_GRAB>, ctx->grid_brush, DSC_BLOCK, 0, st, bpdl, m, j, slot)); break;
        default: CU(launch_k(k_brush<DSC_TOOL_CLAY_STRIPS>, ctx->grid_brush, DSC_BLOCK, 0, st, bpdl, m, j, slot)); break;
      }
    }
    if (exch && (r = dist_halo_exchange(ctx, j))) return r;
  }
  /* 4. normals, 5. bounds */
  if (use_hits) {
    const int mode = (do_normals ? NB_NORMALS : 0) | (do_bounds ? NB_BOUNDS : 0);
    if (do_bounds) {
      if (eager) {
        CU(cudaStreamWaitEvent(st, ev_tag, 0));
      }
      else if (dist) {
        StageScope s(ctx, ST_OTHER);
        k_reset_leaf_boxes<<<ctx->num_sms, DSC_BLOCK, 0, st>>>(m, hits.list, hits.count, 0);
        LAUNCH_CHECK();
      }
    }
    if (ctx->is_grids) {
      if ((r = grids_after_brush(ctx, hits, j, exch))) return r;
    }
    else if (fused) {
      /* interior brush + normals + boxes in one pass over the gathered tiles */
      {
        StageScope s(ctx, ST_NORMALS);
        const size_t sm = ctx->nb_smem;
        switch (tool) {
          case DSC_TOOL_DRAW: k_dab_tile<DSC_TOOL_DRAW><<<ctx->fused_grid[0], NT_THREADS, sm, st>>>(m, j, slot, mode); break;
          case DSC_TOOL_INFLATE: k_dab_tile<DSC_TOOL_INFLATE><<<ctx->fused_grid[1], NT_THREADS, sm, st>>>(m, j, slot, mode); break;
          case DSC_TOOL_GRAB: k_dab_tile<DSC_TOOL_GRAB><<<ctx->fused_grid[2], NT_THREADS, sm, st>>>(m, j, slot, mode); break;
          default: k_dab_tile<DSC_TOOL_CLAY_STRIPS><<<ctx->fused_grid[3], NT_THREADS, sm, st>>>(m, j, slot, mode); break;
        }
        LAUNCH_CHECK();
      }
      if (ctx->any_slow_leaf && (r = run_slow_leaves(ctx, hits, mode))) return r;
    }
    else if (mode) {
      if ((r = run_normals_bounds(ctx, hits, mode, pdl && !dist && tool != DSC_TOOL_SMOOTH))) return r;
    }
    if (eager) {
      /* side stream: carry the refreshed leaf boxes up the tree; overlaps the next dab */
      CU(cudaEventRecord(ev_bb, st));
      CU(cudaStreamWaitEvent(ctx->stream2, ev_bb, 0));
      {
        StageScope s(ctx, ST_FLUSH, ctx->stream2);
        k_refit<<<ctx->num_sms, DSC_BLOCK, 0, ctx->stream2>>>(m, slot);
        LAUNCH_CHECK();
      }
      CU(cudaEventRecord(ev_refit[slot], ctx->stream2));
      ctx->side_busy = true;
    }
    if (mode && ctx->any_slow_leaf) {
      if ((r = run_clear(ctx, hits, (do_normals ? F_UpdateNormals : 0) | (do_bounds ? F_UpdateBB : 0)))) return r;
    }
  }
  else {
    const int want = (do_normals ? F_UpdateNormals : 0) | (do_bounds ? F_UpdateBB : 0);
    if (want && (r = run_flagged(ctx, want))) return r;
  }
  return DSC_OK;
}

/* host ring -> device ring for the dabs [seq, seq + count) */
static int upload_entries(DscContext *ctx, long long seq, int count)
{
  int done = 0;
  while (done < count) {
    const int at = (int)((seq + done) & (DSC_RING - 1));
    const int n = std::min(count - done, DSC_RING - at);
    CU(cudaMemcpyAsync(ctx->d_ring + at, ctx->h_ring + at, sizeof(DabEntry) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    done += n;
  }
  return DSC_OK;
}

/* the host may overwrite a quarter of its pinned ring once the copies queued from it a ring ago ran */
static int ring_reserve(DscContext *ctx, long long seq)
{
  const int q = (int)((seq & (DSC_RING - 1)) / (DSC_RING / 4));
  if ((seq & (DSC_RING / 4 - 1)) == 0 && seq >= DSC_RING) CU(cudaEventSynchronize(ctx->ev_ring[q]));
  return DSC_OK;
}
static int ring_commit(DscContext *ctx, long long seq_end)
{
  /* seq_end: one past the last dab whose copy was just queued */
  if ((seq_end & (DSC_RING / 4 - 1)) == 0) {
    const int q = (int)(((seq_end - 1) & (DSC_RING - 1)) / (DSC_RING / 4));
    CU(cudaEventRecord(ctx->ev_ring[q], ctx->stream));
  }
  return DSC_OK;
}

/* the launch sequence of `batch` dabs of one signature as an executable graph (built on first use) */
static int get_graph(DscContext *ctx, const DabSig &sig, int batch, DabGraph **r_graph)
{
  /* the DevMesh a graph's kernels carry differs by the optional layers: one graph per combination */
  const unsigned gkey = sig.key(batch) | (ctx->stage_timing == 2 ? 1u << 30 : 0u) | (ctx->m.mask ? 1u << 29 : 0u) |
                        (ctx->m.automask ? 1u << 28 : 0u);
  auto it = ctx->graphs.find(gkey);
  if (it != ctx->graphs.end()) {
    *r_graph = &it->second;
    return DSC_OK;
  }
  int r = join_side(ctx);
  if (r) return r;
  DabGraph g;
  const long long launches0 = ctx->launches;
  int stage0[DSC_NUM_STAGES];
  memcpy(stage0, ctx->stage_launches, sizeof(stage0));
  const size_t ev0 = ctx->events.size();
  CU(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  ctx->capturing_now = true;
  k_batch_begin<<<1, 1, 0, ctx->stream>>>(ctx->m, batch);
  r = DSC_OK;
  for (int j = 0; j < batch && r == DSC_OK; j++) r = enqueue_dab(ctx, sig, j, j & (DSC_SLOTS - 1), true);
  if (r == DSC_OK && ctx->side_busy) {
    /* the side stream's work is part of the graph */
    if (cudaEventRecord(ctx->cap_join, ctx->stream2) != cudaSuccess || cudaStreamWaitEvent(ctx->stream, ctx->cap_join, 0) != cudaSuccess)
      r = fail(ctx, DSC_ERR_CUDA, "joining the side stream into the captured batch failed");
  }
  cudaError_t e = cudaStreamEndCapture(ctx->stream, &g.graph);
  ctx->capturing_now = false;
  ctx->side_busy = false;
  /* the event pairs recorded while capturing belong to the graph: read after every replay */
  g.events.assign(ctx->events.begin() + (long)ev0, ctx->events.end());
  ctx->events.resize(ev0);
  if (r != DSC_OK) {
    if (e == cudaSuccess && g.graph) cudaGraphDestroy(g.graph);
    return r;
  }
  if (e != cudaSuccess) return fail(ctx, DSC_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(&g.exec, g.graph, 0);
  if (e != cudaSuccess) {
    cudaGraphDestroy(g.graph);
    return fail(ctx, DSC_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  }
  g.launches = (int)(ctx->launches - launches0) + 1;
  for (int k = 0; k < DSC_NUM_STAGES; k++) g.stage_launches[k] = ctx->stage_launches[k] - stage0[k];
  /* capturing counted the launches once; they are counted per graph launch instead */
  ctx->launches = launches0;
  memcpy(ctx->stage_launches, stage0, sizeof(stage0));
  *r_graph = &ctx->graphs.emplace(gkey, g).first->second;
  return DSC_OK;
}

/* `count` dabs of one launch sequence in one cooperative launch of the persistent batch kernel */
static int launch_batch_kernel(DscContext *ctx, const DabSig &sig, int count, int slot0)
{
  void (*fn)(DevMesh, int, int, int, int) = nullptr;
  int k = 0;
  switch (sig.tool) {
    case DSC_TOOL_DRAW: fn = k_dab_batch<DSC_TOOL_DRAW>; k = 0; break;
    case DSC_TOOL_INFLATE: fn = k_dab_batch<DSC_TOOL_INFLATE>; k = 1; break;
    case DSC_TOOL_GRAB: fn = k_dab_batch<DSC_TOOL_GRAB>; k = 2; break;
    default: fn = k_dab_batch<DSC_TOOL_CLAY_STRIPS>; k = 3; break;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)std::min(ctx->batch_grid[k], ctx->small_dab_grid), 1, 1);
  cfg.blockDim = dim3(NT_THREADS, 1, 1);
  cfg.dynamicSmemBytes = ctx->nb_smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  k_batch_begin<<<1, 1, 0, ctx->stream>>>(ctx->m, count);
  LAUNCH_CHECK();
  CU(cudaLaunchKernelEx(&cfg, fn, ctx->m, count, slot0, sig.needs_area, ctx->lazy_refit ? 0 : 1));
  ctx->launches += 2;
  ctx->batch_launches++;
  return DSC_OK;
}

/* ---- partitioned PBVH: the ranks a dab can reach ----
 * A dab is exchanged among, and executed by, only the ranks whose region it can touch; the others skip it (on grids they
 * still run the averaging of all coarse vertices that every dab does, subdiv_ccg.c:1303-1324) and run ahead.  The decision
 * must come out the same on every rank without talking: it is a pure function of the region boxes all ranks hold at stroke
 * begin and of the dabs so far.  A rank's boxes hold its leaves (shared verts included, as update_node_vb counts them) and
 * the halo elements it reads; a dab that reaches a rank grows that rank's boxes by what it can move: a vertex moves at
 * most `bound` and ends inside the brush's reach. */
static int dist_regions_refresh(DscContext *ctx)
{
  const int W = ctx->world, NB = DSC_REGION_CHUNKS, N = ctx->totnode, L = ctx->m.nleaf;
  ctx->regions.assign((size_t)W * NB * 6, 0.0f);
  if ((int)ctx->leaf_reach.size() != L) {
    ctx->regions.clear(); /* no reach table: every dab goes to every rank */
    return DSC_OK;
  }
  if (ctx->region_leaves.empty()) {
    ctx->region_leaves.resize((size_t)W);
    for (int l = 0; l < L; l++) {
      for (int q = 0; q < W; q++) {
        if ((ctx->leaf_reach[l] >> q) & 1u) ctx->region_leaves[q].push_back(l);
      }
    }
  }
  int r = sync_all(ctx);
  if (r) return r;
  /* leaf boxes: every rank holds all of them (the stroke-end gather keeps them whole) */
  std::vector<float> bb((size_t)6 * N);
  CU(cudaMemcpy(bb.data(), ctx->m.bb, sizeof(float) * bb.size(), cudaMemcpyDeviceToHost));
  for (int q = 0; q < W; q++) {
    const std::vector<int> &ls = ctx->region_leaves[q];
    const int n = (int)ls.size();
    const int per = std::max(1, (n + NB - 1) / NB);
    for (int c = 0; c < NB; c++) {
      float *b = &ctx->regions[((size_t)q * NB + c) * 6];
      for (int k = 0; k < 3; k++) {
        b[k] = FLT_MAX;
        b[3 + k] = -FLT_MAX;
      }
      for (int i = c * per; i < std::min(n, (c + 1) * per); i++) {
        const int l = ls[i];
        for (int k = 0; k < 3; k++) {
          b[k] = std::min(b[k], bb[(size_t)k * N + l]);
          b[3 + k] = std::max(b[3 + k], bb[(size_t)(3 + k) * N + l]);
        }
      }
    }
  }
  return DSC_OK;
}
/* bit per rank the dab can reach; grows the reached ranks' boxes */
static unsigned dist_dab_mask(DscContext *ctx, const DscDab *d)
{
  const int W = ctx->world, NB = DSC_REGION_CHUNKS;
  if (W < 2) return 0u;
  const unsigned all = (1u << W) - 1u;
  if (!ctx->subset_exchange || ctx->regions.empty()) return all;
  /* the tube reaches along the whole view line; a grab's drag blended towards the sculpt normal has no bound known up front */
  if (d->falloff_shape != DSC_FALLOFF_SPHERE || (d->tool == DSC_TOOL_GRAB && d->normal_weight > 0.0f)) return all;
  const float R = d->radius * std::max(d->radius_scale, 1.0f);
  float smax = 0.0f, dl = 0.0f;
  for (int k = 0; k < 3; k++) {
    smax = std::max(smax, fabsf(d->scale[k]));
    dl += d->grab_delta[k] * d->grab_delta[k];
  }
  dl = sqrtf(dl);
  const float bs = fabsf(d->bstrength);
  float reach = R, bound;
  switch (d->tool) {
    case DSC_TOOL_DRAW:
    case DSC_TOOL_INFLATE: bound = d->radius * bs * std::max(smax, 1.0f); break;
    case DSC_TOOL_GRAB: bound = dl * bs; break;
    case DSC_TOOL_CLAY_STRIPS: reach = 4.0f * R; bound = 4.0f * R; break; /* brush-local cube, projection onto its plane */
    default: bound = R; break;                                           /* smooth: towards the neighbour average */
  }
  bound = bound * 1.001f + 1e-6f * R;
  reach = reach * 1.001f + bound;
  unsigned mask = 0u;
  for (int q = 0; q < W; q++) {
    bool hit = false;
    for (int c = 0; c < NB && !hit; c++) {
      const float *b = &ctx->regions[((size_t)q * NB + c) * 6];
      if (b[0] > b[3]) continue;
      float dist = 0.0f;
      for (int k = 0; k < 3; k++) {
        const float c0 = d->location[k];
        const float nearest = c0 < b[k] ? b[k] : (c0 > b[3 + k] ? b[3 + k] : c0);
        dist += (c0 - nearest) * (c0 - nearest);
      }
      hit = dist <= reach * reach;
    }
    if (hit) mask |= 1u << q;
  }
  for (int q = 0; q < W; q++) {
    if (!((mask >> q) & 1u)) continue;
    for (int c = 0; c < NB; c++) {
      float *b = &ctx->regions[((size_t)q * NB + c) * 6];
      if (b[0] > b[3]) continue;
      float lo[3], hi[3];
      bool any = true;
      for (int k = 0; k < 3; k++) {
        lo[k] = std::max(b[k] - bound, d->location[k] - reach);
        hi[k] = std::min(b[3 + k] + bound, d->location[k] + reach);
        any = any && lo[k] <= hi[k];
      }
      if (!any) continue;
      for (int k = 0; k < 3; k++) {
        b[k] = std::min(b[k], lo[k]);
        b[3 + k] = std::max(b[3 + k], hi[k]);
      }
    }
  }
  return mask;
}
/* grids: the dabs this rank skipped still averaged every coarse vertex (and every coarse edge with more than two faces) */
static int dist_flush_skipped(DscContext *ctx)
{
  if (ctx->pending_skipped <= 0) return DSC_OK;
  const int n = ctx->pending_skipped;
  ctx->pending_skipped = 0;
  if (!ctx->is_grids) return DSC_OK;
  k_grid_skipped<<<ctx->num_sms, DSC_BLOCK, 0, ctx->stream>>>(ctx->m, ctx->g, n);
  LAUNCH_CHECK();
  ctx->launches++;
  return DSC_OK;
}

int dsc_dabs(DscContext *ctx, const DscDab *dabs, int count)
{
  NEED_PBVH();
  if (count < 0 || (count > 0 && !dabs)) return fail(ctx, DSC_ERR_INVALID, "bad dab array");
  if (!ctx->in_stroke) return fail(ctx, DSC_ERR_STATE, "dsc_stroke_begin first");
  int r;
  int i = 0;
  /* partitioned: who takes part in which dab (sequential: the region boxes grow with the dabs) */
  std::vector<unsigned> masks;
  if (ctx->world > 1) {
    masks.resize((size_t)count);
    for (int k = 0; k < count; k++) masks[k] = dist_dab_mask(ctx, dabs + k);
  }
  auto mask_of = [&](int k) -> unsigned { return masks.empty() ? 0u : masks[k]; };
  auto mine = [&](int k) -> bool { return masks.empty() || ((masks[k] >> ctx->rank) & 1u); };
  while (i < count) {
    DabEntry e;
    DabSig sig;
    if (!mine(i)) {
      /* out of this dab's reach: nothing of this rank can change, nothing it reads can change */
      if ((r = make_entry(ctx, dabs + i, &e, &sig, mask_of(i)))) return r; /* the argument checks still apply */
      /* a dab no rank's region can reach gathers nothing anywhere: the reference then skips the stitch altogether
       * (its caller returns on totnode == 0), so there is no all-coarse-vertex pass to replay */
      if (mask_of(i)) ctx->pending_skipped++;
      ctx->last_dab_skipped = true;
      ctx->dist_skipped_dabs++;
      i++;
      continue;
    }
    if ((r = dist_flush_skipped(ctx))) return r;
    ctx->last_dab_skipped = false;
    if ((r = make_entry(ctx, dabs + i, &e, &sig, mask_of(i)))) return r;
    if (sig.exch) ctx->dist_exchanged_dabs++;
    else if (ctx->world > 1) ctx->dist_local_dabs++;
    const bool dist = ctx->world > 1;
    if (dist && (ctx->stale_flags || !sig.do_normals || !sig.do_bounds))
      return fail(ctx, DSC_ERR_UNSUPPORTED, "a partitioned PBVH updates normals and bounds with every dab");
    /* how many of the following dabs share the launch sequence */
    const bool graphable = ctx->use_graphs && !(ctx->is_grids && ctx->grid_fused) && ctx->stage_timing != 1 && !ctx->capture && (!dist || ctx->p2p) && !ctx->stale_flags &&
                           !ctx->any_slow_leaf && sig.do_normals && sig.do_bounds;
    int run = 1;
    const long long seq = ctx->ring_seq; /* ring position: runs across strokes; the state slot follows dab_index */
    if ((r = ring_reserve(ctx, seq))) return r;
    ctx->h_ring[seq & (DSC_RING - 1)] = e;
    int batch = 1;
    const bool batchable = graphable && sig.small && ctx->batch_grid[0] > 0;
    if (batchable) {
      while (run < 256 && i + run < count) {
        DabEntry e2;
        DabSig s2;
        if (!mine(i + run) || make_entry(ctx, dabs + i + run, &e2, &s2, mask_of(i + run)) != DSC_OK || !(s2 == sig)) break;
        if ((r = ring_reserve(ctx, seq + run))) return r;
        ctx->h_ring[(seq + run) & (DSC_RING - 1)] = e2;
        run++;
      }
      batch = run;
    }
    else if (graphable && (ctx->dab_index & (DSC_SLOTS - 1)) == 0) {
      const int sizes[3] = {32, 16, 4};
      while (run < 32 && i + run < count) {
        DabEntry e2;
        DabSig s2;
        if (!mine(i + run) || make_entry(ctx, dabs + i + run, &e2, &s2, mask_of(i + run)) != DSC_OK || !(s2 == sig)) break;
        if ((r = ring_reserve(ctx, seq + run))) return r;
        ctx->h_ring[(seq + run) & (DSC_RING - 1)] = e2;
        run++;
      }
      batch = 1;
      for (int k = 0; k < 3; k++) {
        if (run >= sizes[k]) {
          batch = sizes[k];
          break;
        }
      }
    }
    if ((r = upload_entries(ctx, seq, batch))) return r;
    for (int k = 1; k <= batch; k++) {
      if ((r = ring_commit(ctx, seq + k))) return r;
    }
    if (batchable) {
      if ((r = join_side(ctx))) return r;
      if ((r = launch_batch_kernel(ctx, sig, batch, (int)(ctx->dab_index & (DSC_SLOTS - 1))))) return r;
    }
    else if (batch > 1) {
      DabGraph *g = nullptr;
      if ((r = get_graph(ctx, sig, batch, &g))) return r;
      if ((r = join_side(ctx))) return r; /* the graph's first dabs do not wait for an earlier refit themselves */
      CU(cudaGraphLaunch(g->exec, ctx->stream));
      if (ctx->stage_timing == 2) {
        /* kernel durations inside the replayed graph: the kernels still run back to back, only the read-out waits */
        if ((r = sync_all(ctx))) return r;
        for (auto &ev : g->events) {
          float ms = 0.0f;
          if (cudaEventElapsedTime(&ms, ev.a, ev.b) == cudaSuccess) ctx->stage_ms[ev.stage] += ms;
        }
      }
      ctx->launches += g->launches;
      for (int k = 0; k < DSC_NUM_STAGES; k++) ctx->stage_launches[k] += g->stage_launches[k];
      ctx->graph_launches++;
    }
    else {
      const int slot = (int)(ctx->dab_index & (DSC_SLOTS - 1));
      k_batch_begin<<<1, 1, 0, ctx->stream>>>(ctx->m, 1);
      LAUNCH_CHECK();
      ctx->launches++;
      if ((r = enqueue_dab(ctx, sig, 0, slot, false))) return r;
    }
    if (ctx->lazy_refit && sig.do_bounds && !dist) ctx->refit_pending = true;
    /* flag bookkeeping of the sequence that just ran */
    if (!ctx->stale_flags) {
      if (!sig.do_normals || !sig.do_bounds) ctx->stale_flags = true;
    }
    else if (sig.do_normals && sig.do_bounds) {
      ctx->stale_flags = false;
    }
    ctx->dab_index += batch;
    ctx->ring_seq += batch;
    ctx->last_slot = (int)((ctx->dab_index - 1) & (DSC_SLOTS - 1));
    i += batch;
  }
  return DSC_OK;
}

int dsc_dab(DscContext *ctx, const DscDab *dab) { return dsc_dabs(ctx, dab, 1); }

static int read_list(DscContext *ctx, const int *d_list, const int *d_count, int *r_nodes, int capacity, int *r_tot)
{
  int tot = 0;
  int r = sync_all(ctx);
  if (r) return r;
  CU(cudaMemcpy(&tot, d_count, sizeof(int), cudaMemcpyDeviceToHost));
  if (r_tot) *r_tot = tot;
  if (r_nodes && tot > 0) {
    if (capacity < tot) return fail(ctx, DSC_ERR_INVALID, "capacity %d < %d gathered nodes", capacity, tot);
    CU(cudaMemcpy(ctx->h_list, d_list, sizeof(int) * (size_t)tot, cudaMemcpyDeviceToHost));
    /* leaf ids are traversal ranks: ascending id = the order BKE_pbvh_search_gather emits */
    std::sort(ctx->h_list, ctx->h_list + tot);
    for (int i = 0; i < tot; i++) r_nodes[i] = ctx->node_of_dev[ctx->h_list[i]];
  }
  return DSC_OK;
}

int dsc_gather_readback(DscContext *ctx, int *r_nodes, int capacity, int *r_tot)
{
  if (ctx && ctx->have_pbvh && ctx->last_dab_skipped) {
    /* the last dab was out of this rank's reach: it gathered none of its leaves */
    if (r_tot) *r_tot = 0;
    return sync_all(ctx);
  }
  NEED_PBVH();
  const LeafList ll = hit_list(ctx, ctx->last_slot);
  return read_list(ctx, ll.list, ll.count, r_nodes, capacity, r_tot);
}

int dsc_search_sphere(DscContext *ctx, const float center[3], float radius_sq, int original, int ignore_fully_ineffective,
                      int *r_nodes, int capacity, int *r_tot)
{
  NEED_PBVH();
  int r = join_side(ctx);
  if (r) return r;
  CU(cudaMemsetAsync(&ctx->m.tot->search_count, 0, sizeof(int), ctx->stream));
  {
    StageScope s(ctx, ST_GATHER);
    k_gather<<<(ctx->m.nleaf + DSC_BLOCK - 1) / DSC_BLOCK, DSC_BLOCK, 0, ctx->stream>>>(
        ctx->m, center[0], center[1], center[2], radius_sq, original ? 1 : 0, ignore_fully_ineffective ? 1 : 0);
    LAUNCH_CHECK();
  }
  return read_list(ctx, ctx->m.search_list, &ctx->m.tot->search_count, r_nodes, capacity, r_tot);
}

int dsc_last_area(DscContext *ctx, float r_no[3], float r_co[3])
{
  NEED_PBVH();
  int r = sync_all(ctx);
  if (r) return r;
  CU(cudaMemcpy(ctx->h_state, ctx->m.st + ctx->last_slot, sizeof(DabState), cudaMemcpyDeviceToHost));
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
  invalidate_graphs(ctx);
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
  CU(cudaMemcpyAsync(ctx->h_tot, ctx->m.tot, sizeof(StrokeTotals), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  r->vertex_dabs = (int64_t)ctx->h_tot->vd_total;
  r->node_hits = (int64_t)ctx->h_tot->hits_total;
  r->moved_verts = (int64_t)ctx->h_tot->moved_total;
  r->dabs = (int64_t)ctx->h_tot->dabs;
  r->kernel_launches = ctx->launches;
  r->area_verts = (int64_t)ctx->h_tot->area_vd_total;
  r->area_inside = (int64_t)ctx->h_tot->area_inside_total;
  r->all_verts = (int64_t)ctx->h_tot->all_total;
  r->prims = (int64_t)ctx->h_tot->prim_total;
  r->first_touch_verts = (int64_t)ctx->h_tot->first_total;
  r->refit_nodes = (int64_t)ctx->h_tot->refit_total;
  return DSC_OK;
}

int dsc_update_normals(DscContext *ctx)
{
  NEED_PBVH();
  if (ctx->is_grids) return DSC_OK; /* every dab on grids updates its normals; nothing is ever left flagged */
  return run_flagged(ctx, F_UpdateNormals);
}

static int run_orig_flush(DscContext *ctx)
{
  int r = join_side(ctx);
  if (r) return r;
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
  if ((flag & DSC_PBVH_UpdateBB) && !ctx->is_grids) {
    if ((r = run_flagged(ctx, F_UpdateBB))) return r;
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
  int r;
  if ((r = dist_flush_skipped(ctx))) return r;
  if (ctx->world > 1 && (r = dist_gather_all(ctx, false))) return r;
  if (ctx->p2p) {
    int err = 0;
    CU(cudaMemcpyAsync(&err, ctx->link.flags + P2P_ERR, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (err) return fail(ctx, DSC_ERR_NCCL, "a peer-memory exchange timed out: a rank did not queue the same exchanges");
  }
  if ((r = ensure_refit(ctx))) return r;
  r = run_orig_flush(ctx);
  if (r) return r;
  ctx->in_stroke = false;
  return sync_all(ctx);
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

/* every replica whole again: the owners' vertex data (positions, normals, undo snapshot, mask) travel over NCCL */
int dsc_dist_gather(DscContext *ctx)
{
  NEED_PBVH();
  if (ctx->world < 2) return DSC_OK;
  int r = join_side(ctx);
  if (r) return r;
  if ((r = dist_flush_skipped(ctx))) return r;
  if ((r = dist_gather_all(ctx, true))) return r;
  return sync_all(ctx);
}
int dsc_dist_dab_counts(DscContext *ctx, long long r_counts[3])
{
  if (!ctx || !r_counts) return DSC_ERR_INVALID;
  r_counts[0] = ctx->dist_skipped_dabs;
  r_counts[1] = ctx->dist_local_dabs;
  r_counts[2] = ctx->dist_exchanged_dabs;
  return DSC_OK;
}
/* stroke-end sync of a partitioned PBVH: the runs this rank owns only.  Its slots are one contiguous run of every SoA
 * array: six (seven) DMAs into pinned staging, then the host scatters them to vertex / element order. */
static int download_owned_runs(DscContext *ctx, int narr, const float *const *arrs, int *r_s0, int *r_n)
{
  if (ctx->world < 2) return fail(ctx, DSC_ERR_STATE, "not a partitioned PBVH");
  const int s0 = ctx->slot_range[ctx->rank], n = ctx->slot_range[ctx->rank + 1] - s0;
  *r_s0 = s0;
  *r_n = n;
  if (ctx->vert_of_slot.empty()) {
    ctx->vert_of_slot.assign((size_t)ctx->vpad, -1);
    for (int v = 0; v < ctx->totvert; v++) ctx->vert_of_slot[ctx->slot_of[v]] = v;
  }
  if (n <= 0) return DSC_OK;
  const size_t need = (size_t)narr * (size_t)n;
  if (need > ctx->h_own_floats) {
    if (ctx->h_own) cudaFreeHost(ctx->h_own);
    ctx->h_own = nullptr;
    CU(cudaMallocHost((void **)&ctx->h_own, sizeof(float) * need));
    ctx->h_own_floats = need;
  }
  int r = join_side(ctx);
  if (r) return r;
  for (int a = 0; a < narr; a++) {
    CU(cudaMemcpyAsync(ctx->h_own + (size_t)a * n, arrs[a] + s0, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return DSC_OK;
}
/* The owned part straight into the host arrays: the arrays are page-locked and mapped, a kernel walks the vertices (grid
 * elements) in THEIR order and stores the records of those this rank owns through the mapping -- consecutive owned
 * vertices make full-width PCIe writes, every rank uses its own link, no host thread touches the data. */
static void *mapped_host_ptr(DscContext *ctx, void *host, size_t bytes)
{
  void *dp = nullptr;
  if (cudaHostGetDevicePointer(&dp, host, 0) == cudaSuccess && dp) return dp;
  cudaGetLastError();
  if (cudaHostRegister(host, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  ctx->registered.push_back(host);
  if (cudaHostGetDevicePointer(&dp, host, 0) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return dp;
}
__global__ void k_export_owned_mvert(float4 *__restrict__ mv, float *__restrict__ no, const float *__restrict__ cx, const float *__restrict__ cy,
                                     const float *__restrict__ cz, const float *__restrict__ nx, const float *__restrict__ ny,
                                     const float *__restrict__ nz, const int *__restrict__ slot_of, const unsigned *__restrict__ tail, int s0,
                                     int s1, int totvert)
{
  /* a warp takes 32 consecutive vertices: when it owns them all, their normals leave as 24 aligned 16-byte stores (through
   * shared memory) instead of 96 four-byte ones -- stores to mapped host memory cross PCIe one by one */
  __shared__ __align__(16) float sn[8][96];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nround = (totvert + 31) & ~31;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nround; v += gridDim.x * blockDim.x) {
    const int s = v < totvert ? slot_of[v] : -1;
    const bool own = s >= s0 && s < s1;
    float n3[3] = {0.0f, 0.0f, 0.0f};
    if (own) {
      mv[v] = make_float4(cx[s], cy[s], cz[s], __uint_as_float(tail ? tail[v] : 0u));
      if (no) {
        n3[0] = nx[s]; n3[1] = ny[s]; n3[2] = nz[s];
      }
    }
    if (!no) continue;
    if (__all_sync(0xffffffffu, own)) {
      sn[warp][3 * lane] = n3[0]; sn[warp][3 * lane + 1] = n3[1]; sn[warp][3 * lane + 2] = n3[2];
      __syncwarp();
      if (lane < 24) reinterpret_cast<float4 *>(no + 3 * (size_t)(v - lane))[lane] = reinterpret_cast<const float4 *>(sn[warp])[lane];
      __syncwarp();
    }
    else if (own) {
      no[3 * (size_t)v + 0] = n3[0];
      no[3 * (size_t)v + 1] = n3[1];
      no[3 * (size_t)v + 2] = n3[2];
    }
  }
}
/* records of the owned grids, packed: grid g's records start at own_pos[g] * gs2 * ef floats */
__global__ void k_export_owned_ccg(float *__restrict__ out, const float *__restrict__ cx, const float *__restrict__ cy, const float *__restrict__ cz,
                                   const float *__restrict__ nx, const float *__restrict__ ny, const float *__restrict__ nz,
                                   const float *__restrict__ mask, const int *__restrict__ slot_of, const int *__restrict__ own_pos, int gs2,
                                   int totgrid, int ef, int mask_off, int no_off)
{
  for (int g = blockIdx.x; g < totgrid; g += gridDim.x) {
    const int p = own_pos[g];
    if (p < 0) continue;
    const int sl0 = slot_of[(size_t)g * gs2]; /* a grid's elements are consecutive slots */
    float *o = out + (size_t)p * gs2 * ef;
    for (int t = threadIdx.x; t < gs2 * ef; t += blockDim.x) {
      const int v = t / ef, k = t - v * ef;
      const int s = sl0 + v;
      float val = 0.0f;
      if (k < 3) val = (k == 0 ? cx : (k == 1 ? cy : cz))[s];
      else if (k == mask_off) val = mask ? mask[s] : 0.0f;
      else if (no_off >= 0 && k >= no_off && k < no_off + 3) val = (k == no_off ? nx : (k == no_off + 1 ? ny : nz))[s];
      o[t] = val;
    }
  }
}

int dsc_download_owned_mvert(DscContext *ctx, void *r_mvert, float *r_no)
{
  NEED_PBVH();
  if (ctx->is_grids || !r_mvert) return fail(ctx, DSC_ERR_INVALID, "mesh contexts only; r_mvert must not be NULL");
  if (ctx->world > 1 && !getenv("DSC_NO_MAPPED_SYNC")) {
    void *dmv = mapped_host_ptr(ctx, r_mvert, sizeof(float) * 4 * (size_t)ctx->totvert);
    void *dno = r_no ? mapped_host_ptr(ctx, r_no, sizeof(float) * 3 * (size_t)ctx->totvert) : nullptr;
    if (dmv && (dno || !r_no)) {
      int r = join_side(ctx);
      if (r) return r;
      k_export_owned_mvert<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>((float4 *)dmv, (float *)dno, ctx->m.cx, ctx->m.cy, ctx->m.cz, ctx->m.nx, ctx->m.ny,
                                                                     ctx->m.nz, ctx->d_slot_of, ctx->d_tail, ctx->slot_range[ctx->rank],
                                                                     ctx->slot_range[ctx->rank + 1], ctx->totvert);
      LAUNCH_CHECK();
      CU(cudaStreamSynchronize(ctx->stream));
      return DSC_OK;
    }
  }
  const float *arrs[6] = {ctx->m.cx, ctx->m.cy, ctx->m.cz, ctx->m.nx, ctx->m.ny, ctx->m.nz};
  int s0 = 0, n = 0;
  int r = download_owned_runs(ctx, r_no ? 6 : 3, arrs, &s0, &n);
  if (r) return r;
  float *mv = (float *)r_mvert;
  const float *h = ctx->h_own;
  const int *vos = ctx->vert_of_slot.data();
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) {
    const int v = vos[s0 + i];
    if (v < 0) continue;
    mv[(size_t)4 * v + 0] = h[i];
    mv[(size_t)4 * v + 1] = h[(size_t)n + i];
    mv[(size_t)4 * v + 2] = h[(size_t)2 * n + i];
    if (r_no) {
      r_no[(size_t)3 * v + 0] = h[(size_t)3 * n + i];
      r_no[(size_t)3 * v + 1] = h[(size_t)4 * n + i];
      r_no[(size_t)3 * v + 2] = h[(size_t)5 * n + i];
    }
  }
  return DSC_OK;
}
int dsc_download_owned_ccg(DscContext *ctx, void *r_elems, int elem_floats, int mask_offset_floats, int normal_offset_floats)
{
  NEED_PBVH();
  if (!ctx->is_grids || !r_elems || elem_floats < 3 || elem_floats > 16) return fail(ctx, DSC_ERR_INVALID, "grids contexts only; bad CCG element layout");
  if (ctx->world > 1 && !getenv("DSC_NO_MAPPED_SYNC")) {
    /* whole CCGElem records of the owned grids are packed on the device in grid order; every run of consecutive owned grids
     * is one DMA into the (page-locked) CCG storage -- a rank's grids are the grids of a patch of coarse faces, a few
     * hundred runs of megabytes each */
    const int gs2 = ctx->grid_size * ctx->grid_size;
    if (ctx->own_grid_runs.empty()) {
      std::vector<int> own;
      const int s0 = ctx->slot_range[ctx->rank], s1 = ctx->slot_range[ctx->rank + 1];
      for (int g = 0; g < ctx->totgrid; g++) {
        const int sl = ctx->slot_of[(size_t)g * gs2];
        if (sl >= s0 && sl < s1) own.push_back(g);
      }
      std::vector<int> pos((size_t)ctx->totgrid, -1);
      for (size_t i = 0; i < own.size(); i++) pos[own[i]] = (int)i;
      for (size_t i = 0; i < own.size();) {
        size_t j = i + 1;
        while (j < own.size() && own[j] == own[j - 1] + 1) j++;
        ctx->own_grid_runs.push_back(own[i]);
        ctx->own_grid_runs.push_back((int)(j - i));
        i = j;
      }
      ctx->own_grid_count = (int)own.size();
      int r0;
      if ((r0 = dev_upload(ctx, &ctx->d_own_grid_pos, pos))) return r0;
      if ((r0 = dev_alloc(ctx, &ctx->d_own_pack, (size_t)std::max(ctx->own_grid_count, 1) * gs2 * 16))) return r0;
    }
    if (elem_floats <= 16 && mapped_host_ptr(ctx, r_elems, sizeof(float) * (size_t)elem_floats * (size_t)ctx->totvert)) {
      int r = join_side(ctx);
      if (r) return r;
      k_export_owned_ccg<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(ctx->d_own_pack, ctx->m.cx, ctx->m.cy, ctx->m.cz, ctx->m.nx, ctx->m.ny, ctx->m.nz,
                                                                   ctx->m.mask, ctx->d_slot_of, ctx->d_own_grid_pos, gs2, ctx->totgrid, elem_floats,
                                                                   mask_offset_floats, normal_offset_floats);
      LAUNCH_CHECK();
      const size_t grid_bytes = sizeof(float) * (size_t)elem_floats * gs2;
      size_t at = 0;
      for (size_t i = 0; i + 1 < ctx->own_grid_runs.size(); i += 2) {
        const size_t bytes = grid_bytes * (size_t)ctx->own_grid_runs[i + 1];
        CU(cudaMemcpyAsync((char *)r_elems + grid_bytes * (size_t)ctx->own_grid_runs[i], (const char *)ctx->d_own_pack + at, bytes,
                           cudaMemcpyDeviceToHost, ctx->stream));
        at += bytes;
      }
      CU(cudaStreamSynchronize(ctx->stream));
      return DSC_OK;
    }
  }
  const float *arrs[7] = {ctx->m.cx, ctx->m.cy, ctx->m.cz, ctx->m.nx, ctx->m.ny, ctx->m.nz, ctx->m.mask};
  const bool with_mask = ctx->m.mask && mask_offset_floats >= 0;
  int s0 = 0, n = 0;
  int r = download_owned_runs(ctx, with_mask ? 7 : 6, arrs, &s0, &n);
  if (r) return r;
  float *out = (float *)r_elems;
  const float *h = ctx->h_own;
  const int *vos = ctx->vert_of_slot.data();
  const int ef = elem_floats, mo = mask_offset_floats, no = normal_offset_floats;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) {
    const int e = vos[s0 + i];
    if (e < 0) continue;
    float *rec = out + (size_t)ef * e;
    rec[0] = h[i];
    rec[1] = h[(size_t)n + i];
    rec[2] = h[(size_t)2 * n + i];
    if (no >= 0) {
      rec[no] = h[(size_t)3 * n + i];
      rec[no + 1] = h[(size_t)4 * n + i];
      rec[no + 2] = h[(size_t)5 * n + i];
    }
    if (with_mask) rec[mo] = h[(size_t)6 * n + i];
  }
  return DSC_OK;
}

int dsc_download_co(DscContext *ctx, float *r_co)
{
  NEED_PBVH();
  return export3(ctx, r_co, ctx->m.cx, ctx->m.cy, ctx->m.cz);
}
int dsc_download_mvert(DscContext *ctx, void *r_mvert)
{
  NEED_PBVH();
  if (!r_mvert) return fail(ctx, DSC_ERR_INVALID, "output pointer is NULL");
  k_export_mvert<<<ctx->grid, 256, 0, ctx->stream>>>(reinterpret_cast<float4 *>(ctx->d_stage3), ctx->m.cx, ctx->m.cy, ctx->m.cz,
                                                     ctx->d_slot_of, ctx->d_tail, ctx->totvert);
  LAUNCH_CHECK();
  CU(cudaMemcpyAsync(r_mvert, ctx->d_stage3, sizeof(float) * 4 * (size_t)ctx->totvert, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return DSC_OK;
}

/* whole CCGElem records (co, [mask], no interleaved: subdiv_ccg.c:62-90) in element order, packed on the device */
__global__ void k_export_ccg(float *__restrict__ out, const float *__restrict__ cx, const float *__restrict__ cy, const float *__restrict__ cz,
                             const float *__restrict__ nx, const float *__restrict__ ny, const float *__restrict__ nz,
                             const float *__restrict__ mask, const int *__restrict__ slot_of, int first, int count, int ef, int mask_off,
                             int no_off)
{
  const long long items = (long long)count * ef;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < items; t += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(t / ef), k = (int)(t - (long long)v * ef);
    const int s = slot_of[first + v];
    float val = 0.0f;
    if (k < 3) val = (k == 0 ? cx : (k == 1 ? cy : cz))[s];
    else if (k == mask_off) val = mask ? mask[s] : 0.0f;
    else if (no_off >= 0 && k >= no_off && k < no_off + 3) val = (k == no_off ? nx : (k == no_off + 1 ? ny : nz))[s];
    out[t] = val;
  }
}
int dsc_download_ccg(DscContext *ctx, void *r_elems, int elem_floats, int mask_offset_floats, int normal_offset_floats)
{
  NEED_PBVH();
  if (!r_elems || elem_floats < 3 || elem_floats > 16) return fail(ctx, DSC_ERR_INVALID, "bad CCG element layout");
  int r = join_side(ctx);
  if (r) return r;
  const int V = ctx->totvert;
  const int chunk = (int)(((size_t)4 * V) / (size_t)elem_floats); /* elements per pass through the staging array */
  for (int first = 0; first < V; first += chunk) {
    const int count = std::min(chunk, V - first);
    k_export_ccg<<<ctx->grid, 256, 0, ctx->stream>>>(ctx->d_stage3, ctx->m.cx, ctx->m.cy, ctx->m.cz, ctx->m.nx, ctx->m.ny, ctx->m.nz,
                                                     ctx->m.mask, ctx->d_slot_of, first, count, elem_floats, mask_offset_floats,
                                                     normal_offset_floats);
    LAUNCH_CHECK();
    CU(cudaMemcpyAsync((float *)r_elems + (size_t)first * elem_floats, ctx->d_stage3, sizeof(float) * (size_t)count * elem_floats,
                       cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return DSC_OK;
}
int dsc_host_register(DscContext *ctx, void *ptr, size_t bytes)
{
  if (!ctx || !ptr) return DSC_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError(); /* the context page-locked it itself (dsc_download_owned_*): fine */
    return DSC_OK;
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, DSC_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e));
  }
  return DSC_OK;
}
int dsc_host_unregister(DscContext *ctx, void *ptr)
{
  if (!ctx || !ptr) return DSC_ERR_INVALID;
  CU(cudaHostUnregister(ptr));
  return DSC_OK;
}
__global__ void k_export1(float *__restrict__ out, const float *__restrict__ a, const int *__restrict__ slot_of, int totvert)
{
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < totvert; v += gridDim.x * blockDim.x) out[v] = a[slot_of[v]];
}
int dsc_download_mask(DscContext *ctx, float *r_mask)
{
  NEED_PBVH();
  if (!r_mask) return fail(ctx, DSC_ERR_INVALID, "output pointer is NULL");
  if (!ctx->m.mask) return fail(ctx, DSC_ERR_STATE, "no mask layer resident");
  int r = join_side(ctx);
  if (r) return r;
  k_export1<<<ctx->grid, 256, 0, ctx->stream>>>(ctx->d_stage3, ctx->m.mask, ctx->d_slot_of, ctx->totvert);
  LAUNCH_CHECK();
  CU(cudaMemcpyAsync(r_mask, ctx->d_stage3, sizeof(float) * (size_t)ctx->totvert, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return DSC_OK;
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
  int r = ensure_refit(ctx);
  if (r) return r;
  if ((r = sync_all(ctx))) return r;
  for (int pass = 0; pass < 2; pass++) {
    float *out = pass ? r_orig_bb : r_bb;
    if (!out) continue;
    CU(cudaMemcpy(tmp.data(), pass ? ctx->m.obb : ctx->m.bb, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int n = 0; n < N; n++) {
      const int dn = ctx->dev_of_node[n];
      for (int k = 0; k < 6; k++) out[(size_t)6 * n + k] = tmp[(size_t)k * N + dn];
    }
  }
  return DSC_OK;
}

int dsc_download_node_flags(DscContext *ctx, int *r_flags)
{
  NEED_PBVH();
  std::vector<int> tmp((size_t)ctx->totnode);
  int r = sync_all(ctx);
  if (r) return r;
  CU(cudaMemcpy(tmp.data(), ctx->m.node_flag, sizeof(int) * (size_t)ctx->totnode, cudaMemcpyDeviceToHost));
  for (int n = 0; n < ctx->totnode; n++) r_flags[n] = tmp[ctx->dev_of_node[n]];
  return DSC_OK;
}

int dsc_download_touched(DscContext *ctx, unsigned char *r_touched)
{
  NEED_PBVH();
  const int L = ctx->m.nleaf;
  std::vector<unsigned> st((size_t)std::max(L, 1));
  int r = sync_all(ctx);
  if (r) return r;
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
  if (ctx->is_grids) return fail(ctx, DSC_ERR_UNSUPPORTED, "vert_coords_apply is a mesh entry point (pbvh.c:4707 asserts PBVH_FACES data)");
  int r = join_side(ctx);
  if (r) return r;
  CU(cudaMemcpyAsync(ctx->d_stage3, co, sizeof(float) * 3 * (size_t)ctx->totvert, cudaMemcpyHostToDevice, ctx->stream));
  k_import3<<<ctx->grid, 256, 0, ctx->stream>>>(ctx->d_stage3, ctx->m.cx, ctx->m.cy, ctx->m.cz, ctx->d_slot_of, ctx->m.dirty,
                                                ctx->totvert);
  LAUNCH_CHECK();
  /* pbvh.c:4743-4747: every node is marked, bounds (vb and orig_vb) are refreshed */
  const int n = std::max(ctx->m.nleaf, 1);
  k_mark_all<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->m, F_UpdateNormals | F_UpdateBB | F_UpdateOriginalBB | F_UpdateDrawBuffers | F_UpdateRedraw,
                                                       0, ctx->nwords);
  LAUNCH_CHECK();
  if ((r = run_flagged(ctx, F_UpdateNormals | F_UpdateBB))) return r;
  if ((r = run_orig_flush(ctx))) return r;
  return sync_all(ctx);
}

/* ---- ray-cast ---- */
#define DSC_RAY_FIRST 64
int dsc_raycast_enable(DscContext *ctx)
{
  if (!ctx) return DSC_ERR_INVALID;
  if (ctx->have_pbvh) return fail(ctx, DSC_ERR_STATE, "dsc_raycast_enable comes before dsc_pbvh_upload");
  if (!ctx->is_grids) ctx->want_draw = true; /* the looptri corner table (grids: a quad's corners follow from its place) */
  ctx->want_raycast = true;
  return DSC_OK;
}
/* pbvh_grids_node_raycast (pbvh.c:4102-4200) under BKE_pbvh_raycast + the stroke operator's hit callback: the device lists
 * the quads the ray touches, the host folds them in the reference's order */
static int grids_raycast(DscContext *ctx, const RayParams &rp, const float ray_start[3], const float ray_normal[3], float max_depth,
                         DscRayHit *r_hit)
{
  const int L = ctx->m.nleaf;
  int r;
  if (!ctx->d_ray_count && (r = dev_zero(ctx, &ctx->d_ray_count, 1))) return r;
  std::vector<GridRayHit> hits;
  for (int attempt = 0; attempt < 2; attempt++) {
    if (!ctx->d_gray_out) {
      CU(cudaMalloc((void **)&ctx->d_gray_out, sizeof(GridRayHit) * (size_t)ctx->gray_capacity));
    }
    CU(cudaMemsetAsync(ctx->d_ray_count, 0, sizeof(int), ctx->stream));
    k_grid_raycast<<<std::min(L, ctx->num_sms * 8), DSC_BLOCK, 0, ctx->stream>>>(ctx->m, ctx->g, rp, ctx->gray_capacity, ctx->d_ray_count,
                                                                                    ctx->d_gray_out);
    LAUNCH_CHECK();
    ctx->launches += 1;
    int nh = 0;
    CU(cudaMemcpyAsync(&nh, ctx->d_ray_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (nh > ctx->gray_capacity) {
      /* a grazing ray touched more quads than the buffer holds: make room for all of them and ask again */
      CU(cudaFree(ctx->d_gray_out));
      ctx->d_gray_out = nullptr;
      ctx->gray_capacity = nh + nh / 4 + 64;
      continue;
    }
    hits.resize((size_t)std::max(nh, 0));
    if (nh > 0) {
      CU(cudaMemcpyAsync(hits.data(), ctx->d_gray_out, sizeof(GridRayHit) * (size_t)nh, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
    }
    break;
  }
  if (hits.empty()) return DSC_OK;
  /* BKE_pbvh_search_callback_occluded (pbvh.c:2852-2889): leaves by entry distance, ties in traversal order; inside a leaf
   * the quads in (grid, y, x) order */
  std::sort(hits.begin(), hits.end(), [](const GridRayHit &a, const GridRayHit &b) {
    if (a.tmin != b.tmin) return a.tmin < b.tmin;
    if (a.leaf != b.leaf) return a.leaf < b.leaf;
    return a.order < b.order;
  });
  float depth = max_depth, tmin = FLT_MAX;
  const GridRayHit *win = nullptr;
  size_t i = 0;
  while (i < hits.size()) {
    size_t e = i;
    while (e < hits.size() && hits[e].leaf == hits[i].leaf) e++;
    /* sculpt_raycast_cb: a leaf entered behind the best hit is not looked at */
    if (hits[i].tmin < tmin) {
      bool node_hit = false;
      for (size_t k = i; k < e; k++) {
        const GridRayHit &h = hits[k];
        float d;
        /* ray_face_intersection_quad (pbvh.c:3930-3949): the second triangle only when the first is not a nearer hit */
        if (h.d1 >= 0.0f && h.d1 < depth) d = h.d1;
        else if (h.d2 >= 0.0f && h.d2 < depth) d = h.d2;
        else continue;
        depth = d;
        win = &h;
        node_hit = true;
      }
      if (node_hit) tmin = depth;
    }
    i = e;
  }
  if (!win) return DSC_OK;
  const GridRayHit &c = *win;
  r_hit->hit = 1;
  r_hit->depth = depth;
  r_hit->face = c.grid; /* r_active_grid_index */
  r_hit->node = ctx->leaf_node[c.leaf];
  {
    /* normal_quad_v3, lib/intern/math_geom.cc:51-69 */
    float n1[3], n2[3], n[3];
    for (int k = 0; k < 3; k++) {
      n1[k] = c.co[0][k] - c.co[2][k];
      n2[k] = c.co[1][k] - c.co[3][k];
    }
    n[0] = n1[1] * n2[2] - n1[2] * n2[1];
    n[1] = n1[2] * n2[0] - n1[0] * n2[2];
    n[2] = n1[0] * n2[1] - n1[1] * n2[0];
    float d = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    if (d > 1.0e-35f) {
      d = sqrtf(d);
      const float f = 1.0f / d;
      for (int k = 0; k < 3; k++) n[k] = n[k] * f;
    }
    else {
      n[0] = n[1] = n[2] = 0.0f;
    }
    memcpy(r_hit->face_normal, n, sizeof(n));
  }
  {
    /* the corner nearest to the hit point (pbvh.c:4170-4190): r_active_vertex_index as an element index */
    const int gs = ctx->grid_size, gs1 = gs - 1;
    const int y = c.quad / gs1, x = c.quad - y * gs1;
    const int xy[4][2] = {{x, y}, {x + 1, y}, {x + 1, y + 1}, {x, y + 1}};
    float location[3], nearest[3] = {0.0f, 0.0f, 0.0f};
    for (int k = 0; k < 3; k++) location[k] = ray_start[k] + ray_normal[k] * depth;
    int best = 0;
    for (int j = 0; j < 4; j++) {
      float da = 0.0f, db = 0.0f;
      for (int k = 0; k < 3; k++) {
        da += (location[k] - c.co[j][k]) * (location[k] - c.co[j][k]);
        db += (location[k] - nearest[k]) * (location[k] - nearest[k]);
      }
      if (j == 0 || da < db) {
        memcpy(nearest, c.co[j], sizeof(float[3]));
        best = j;
      }
    }
    r_hit->vertex = c.grid * gs * gs + xy[best][1] * gs + xy[best][0];
  }
  return DSC_OK;
}

int dsc_raycast(DscContext *ctx, const float ray_start[3], const float ray_normal[3], int original, float max_depth, DscRayHit *r_hit)
{
  NEED_PBVH();
  if (!ray_start || !ray_normal || !r_hit) return fail(ctx, DSC_ERR_INVALID, "NULL argument");
  if (!ctx->want_raycast) return fail(ctx, DSC_ERR_STATE, "dsc_raycast_enable first");
  int r = join_side(ctx);
  if (r) return r;
  memset(r_hit, 0, sizeof(*r_hit));
  RayParams rp;
  for (int k = 0; k < 3; k++) {
    rp.o[k] = ray_start[k];
    rp.inv_dir[k] = 1.0f / ray_normal[k]; /* isect_ray_aabb_v3_precalc, math_geom.cc:3017-3030 */
    rp.sign[k] = rp.inv_dir[k] < 0.0f;
  }
  {
    /* isect_ray_tri_watertight_v3_precalc, math_geom.cc:1755-1780 */
    const float x = fabsf(ray_normal[0]), y = fabsf(ray_normal[1]), z = fabsf(ray_normal[2]);
    int kz = ((x > y) ? ((x > z) ? 0 : 2) : ((y > z) ? 1 : 2));
    int kx = (kz != 2) ? (kz + 1) : 0;
    int ky = (kx != 2) ? (kx + 1) : 0;
    if (ray_normal[kz] < 0.0f) std::swap(kx, ky);
    const float inv_dir_z = 1.0f / ray_normal[kz];
    rp.sx = ray_normal[kx] * inv_dir_z;
    rp.sy = ray_normal[ky] * inv_dir_z;
    rp.sz = inv_dir_z;
    rp.kx = kx; rp.ky = ky; rp.kz = kz;
  }
  rp.original = original ? 1 : 0;
  const int L = ctx->m.nleaf;
  if (ctx->is_grids) return grids_raycast(ctx, rp, ray_start, ray_normal, max_depth, r_hit);
  if (!ctx->d_ray_out) {
    if ((r = dev_zero(ctx, &ctx->d_ray_out, (size_t)L))) return r;
    if ((r = dev_zero(ctx, &ctx->d_ray_count, 1))) return r;
    CU(cudaMallocHost((void **)&ctx->h_ray_out, sizeof(RayLeafHit) * DSC_RAY_FIRST));
    CU(cudaMallocHost((void **)&ctx->h_ray_count, sizeof(int)));
  }
  CU(cudaMemsetAsync(ctx->d_ray_count, 0, sizeof(int), ctx->stream));
  k_raycast<<<std::min(L, ctx->num_sms * 8), DSC_BLOCK, 0, ctx->stream>>>(ctx->m, ctx->d_tri_slots, ctx->d_slot_leaf, rp, ctx->d_ray_count,
                                                                           ctx->d_ray_out);
  LAUNCH_CHECK();
  ctx->launches += 1;
  /* the entered leaves with a hit: a handful for any real mesh; one copy brings the count and the first few */
  CU(cudaMemcpyAsync(ctx->h_ray_count, ctx->d_ray_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(ctx->h_ray_out, ctx->d_ray_out, sizeof(RayLeafHit) * (size_t)std::min(L, DSC_RAY_FIRST), cudaMemcpyDeviceToHost,
                     ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const int nh = *ctx->h_ray_count;
  if (nh <= 0) return DSC_OK;
  std::vector<RayLeafHit> hits((size_t)nh);
  memcpy(hits.data(), ctx->h_ray_out, sizeof(RayLeafHit) * (size_t)std::min(nh, DSC_RAY_FIRST));
  if (nh > DSC_RAY_FIRST) {
    CU(cudaMemcpyAsync(hits.data() + DSC_RAY_FIRST, ctx->d_ray_out + DSC_RAY_FIRST, sizeof(RayLeafHit) * (size_t)(nh - DSC_RAY_FIRST),
                       cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  /* BKE_pbvh_search_callback_occluded (pbvh.c:2852-2889): leaves by entry distance, ties in traversal order (the
   * looptri ranges of the leaves ascend in traversal order) */
  std::stable_sort(hits.begin(), hits.end(), [](const RayLeafHit &a, const RayLeafHit &b) {
    return a.tmin < b.tmin || (!(b.tmin < a.tmin) && a.pos < b.pos);
  });
  /* the stroke operator's hit callback (sculpt_raycast_cb): a leaf entered behind the best hit is not looked at */
  float depth = max_depth, tmin = FLT_MAX;
  const RayLeafHit *win = nullptr;
  for (const RayLeafHit &h : hits) {
    if (!(h.tmin < tmin)) continue;
    if (h.depth < depth) {
      depth = h.depth;
      tmin = depth;
      win = &h;
    }
  }
  if (!win) return DSC_OK;
  const RayLeafHit &c = *win;
  r_hit->hit = 1;
  r_hit->depth = depth;
  r_hit->face = c.poly;
  r_hit->node = ctx->leaf_node[c.leaf];
  {
    /* normal_tri_v3, lib/intern/math_geom.cc:31-49 */
    float n1[3], n2[3], n[3];
    for (int k = 0; k < 3; k++) {
      n1[k] = c.co[0][k] - c.co[1][k];
      n2[k] = c.co[1][k] - c.co[2][k];
    }
    n[0] = n1[1] * n2[2] - n1[2] * n2[1];
    n[1] = n1[2] * n2[0] - n1[0] * n2[2];
    n[2] = n1[0] * n2[1] - n1[1] * n2[0];
    float d = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    if (d > 1.0e-35f) {
      d = sqrtf(d);
      const float f = 1.0f / d;
      for (int k = 0; k < 3; k++) n[k] = n[k] * f;
    }
    else {
      n[0] = n[1] = n[2] = 0.0f;
    }
    memcpy(r_hit->face_normal, n, sizeof(n));
  }
  {
    /* the corner nearest to the hit point (pbvh.c:4084-4098) */
    float location[3], nearest[3] = {0.0f, 0.0f, 0.0f};
    for (int k = 0; k < 3; k++) location[k] = ray_start[k] + ray_normal[k] * depth;
    int sl = c.slot[0];
    for (int j = 0; j < 3; j++) {
      float da = 0.0f, db = 0.0f;
      for (int k = 0; k < 3; k++) {
        da += (location[k] - c.co[j][k]) * (location[k] - c.co[j][k]);
        db += (location[k] - nearest[k]) * (location[k] - nearest[k]);
      }
      if (j == 0 || da < db) {
        memcpy(nearest, c.co[j], sizeof(float[3]));
        sl = c.slot[j];
      }
    }
    if (ctx->vert_of_slot.empty()) {
      ctx->vert_of_slot.assign((size_t)ctx->vpad, -1);
      for (int v = 0; v < ctx->totvert; v++) ctx->vert_of_slot[ctx->slot_of[v]] = v;
    }
    r_hit->vertex = ctx->vert_of_slot[sl];
  }
  return DSC_OK;
}

/* ---- draw-buffer fill from the device ---- */
int dsc_draw_enable(DscContext *ctx)
{
  if (!ctx) return DSC_ERR_INVALID;
  if (ctx->have_pbvh) return fail(ctx, DSC_ERR_STATE, "dsc_draw_enable comes before dsc_pbvh_upload");
  ctx->want_draw = true;
  return DSC_OK;
}
int dsc_draw_leaf_shading(DscContext *ctx, const unsigned char *node_smooth)
{
  NEED_PBVH();
  if (!ctx->want_draw) return fail(ctx, DSC_ERR_STATE, "dsc_draw_enable first");
  if (!node_smooth) return fail(ctx, DSC_ERR_INVALID, "node_smooth is NULL");
  if (ctx->d_vbo) return fail(ctx, DSC_ERR_STATE, "the per-leaf shading comes before the first dsc_draw_update");
  const int L = ctx->m.nleaf;
  ctx->h_leaf_smooth.assign((size_t)std::max(L, 1), 0);
  for (int l = 0; l < L; l++) ctx->h_leaf_smooth[l] = node_smooth[ctx->leaf_node[l]] ? 1 : 0;
  int r;
  if ((r = dev_upload(ctx, &ctx->d_leaf_smooth, ctx->h_leaf_smooth))) return r;
  if (ctx->is_grids) {
    const int gs = ctx->grid_size;
    ctx->h_leaf_rec.assign((size_t)L + 1, 0);
    for (int l = 0; l < L; l++)
      ctx->h_leaf_rec[l + 1] = ctx->h_leaf_rec[l] + (long long)ctx->h_leaf_pcnt[l] * (ctx->h_leaf_smooth[l] ? gs * gs : (gs - 1) * (gs - 1) * 4);
    if ((r = dev_upload(ctx, &ctx->d_leaf_rec, ctx->h_leaf_rec))) return r;
  }
  return DSC_OK;
}
int dsc_draw_update(DscContext *ctx, int smooth, int show_mask)
{
  NEED_PBVH();
  if (!ctx->want_draw) return fail(ctx, DSC_ERR_STATE, "dsc_draw_enable first");
  const bool per_leaf = smooth < 0;
  if (per_leaf && ctx->h_leaf_smooth.empty()) return fail(ctx, DSC_ERR_STATE, "DSC_DRAW_SHADING_PER_LEAF needs dsc_draw_leaf_shading");
  int r = join_side(ctx);
  if (r) return r;
  const int flags = DSC_PBVH_UpdateDrawBuffers | DSC_PBVH_RebuildDrawBuffers;
  if (ctx->is_grids) {
    /* gpu_pbvh_grid_buffers_update (gpu_buffers.c:548-725): gs^2 records per grid when smooth, 4 (gs - 1)^2 when flat; the
     * shading mode is a property of the mesh (grid_flag_mats) and fixes the layout at the first update */
    const int gs = ctx->grid_size;
    const int per_grid = per_leaf ? -1 : (smooth ? gs * gs : (gs - 1) * (gs - 1) * 4);
    if (ctx->d_vbo && ctx->draw_per_grid != per_grid) return fail(ctx, DSC_ERR_STATE, "the shading mode of a grids context is fixed by its first dsc_draw_update");
    const size_t recs = per_leaf ? (size_t)ctx->h_leaf_rec.back() : (size_t)std::max(ctx->totgrid, 1) * (size_t)per_grid;
    if (!ctx->d_vbo && (r = dev_zero(ctx, &ctx->d_vbo, std::max(recs, (size_t)1) * 9))) return r;
    ctx->draw_per_grid = per_grid;
    if ((r = run_collect(ctx, flags))) return r;
    k_grid_draw_fill<<<ctx->num_sms * 4, DSC_BLOCK, 0, ctx->stream>>>(ctx->m, ctx->g, ctx->m.flag_list, &ctx->m.tot->flag_count, smooth > 0 ? 1 : 0,
                                                                       per_leaf ? ctx->d_leaf_smooth : nullptr, per_leaf ? ctx->d_leaf_rec : nullptr,
                                                                       (show_mask && ctx->g.mask) ? 1 : 0, ctx->d_vbo);
    LAUNCH_CHECK();
    ctx->launches++;
    return run_clear(ctx, flag_list(ctx), flags);
  }
  if (!ctx->d_vbo && (r = dev_zero(ctx, &ctx->d_vbo, (size_t)std::max(ctx->tottri, 1) * 3 * 9))) return r;
  if ((r = run_collect(ctx, flags))) return r;
  k_draw_fill<<<ctx->num_sms * 4, DSC_BLOCK, 0, ctx->stream>>>(ctx->m, ctx->d_tri_slots, ctx->m.flag_list, &ctx->m.tot->flag_count,
                                                                smooth > 0 ? 1 : 0, per_leaf ? ctx->d_leaf_smooth : nullptr,
                                                                (show_mask && ctx->m.mask) ? 1 : 0, ctx->d_vbo);
  LAUNCH_CHECK();
  ctx->launches++;
  return run_clear(ctx, flag_list(ctx), flags); /* pbvh.c:3276 */
}
int dsc_draw_node_buffer(DscContext *ctx, int node, void **r_device_ptr, int *r_vert_len)
{
  NEED_PBVH();
  if (!ctx->d_vbo) return fail(ctx, DSC_ERR_STATE, "dsc_draw_update first");
  if (node < 0 || node >= ctx->totnode || ctx->dev_of_node[node] >= ctx->m.nleaf) return fail(ctx, DSC_ERR_INVALID, "node %d is not a leaf", node);
  const int l = ctx->dev_of_node[node];
  if (ctx->is_grids && ctx->draw_per_grid < 0) {
    if (r_device_ptr) *r_device_ptr = ctx->d_vbo + (size_t)ctx->h_leaf_rec[l] * 9;
    if (r_vert_len) *r_vert_len = (int)(ctx->h_leaf_rec[l + 1] - ctx->h_leaf_rec[l]);
    return DSC_OK;
  }
  const int per_prim = ctx->is_grids ? ctx->draw_per_grid : 3;
  if (r_device_ptr) *r_device_ptr = ctx->d_vbo + (size_t)ctx->h_leaf_pbeg[l] * per_prim * 9;
  if (r_vert_len) *r_vert_len = ctx->h_leaf_pcnt[l] * per_prim;
  return DSC_OK;
}
int dsc_draw_download(DscContext *ctx, int node, void *r_host, size_t capacity_bytes, int *r_vert_len)
{
  void *dp = nullptr;
  int n = 0;
  int r = dsc_draw_node_buffer(ctx, node, &dp, &n);
  if (r) return r;
  if (r_vert_len) *r_vert_len = n;
  if (!r_host) return DSC_OK;
  if (capacity_bytes < (size_t)n * 36) return fail(ctx, DSC_ERR_INVALID, "buffer too small for %d vertices", n);
  if ((r = join_side(ctx))) return r;
  CU(cudaMemcpyAsync(r_host, dp, (size_t)n * 36, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return DSC_OK;
}

/* Checkpoint / rollback of the resident mesh state (positions, normals, node boxes, node flags):
 * device-to-device copies.  What an operator cancel or the undo system's restore does on the host
 * side of the reference (the undo nodes' co / no written back, paint_hide.c:78, then
 * BKE_pbvh_update_bounds), kept on the device so no mesh crosses PCIe. */
int dsc_state_save(DscContext *ctx)
{
  NEED_PBVH();
  if (ctx->in_stroke) return fail(ctx, DSC_ERR_STATE, "no checkpoint inside a stroke");
  int r = join_side(ctx);
  if (r) return r;
  DevMesh &m = ctx->m;
  const size_t VP = (size_t)ctx->vpad, N = (size_t)ctx->totnode;
  if (!ctx->d_save_v) {
    if ((r = dev_alloc(ctx, &ctx->d_save_v, 7 * VP)) || (r = dev_alloc(ctx, &ctx->d_save_bb, 12 * N)) ||
        (r = dev_alloc(ctx, &ctx->d_save_flag, N)))
      return r;
  }
  float *src[6] = {m.cx, m.cy, m.cz, m.nx, m.ny, m.nz};
  for (int k = 0; k < 6; k++) CU(cudaMemcpyAsync(ctx->d_save_v + k * VP, src[k], VP * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  if (ctx->is_grids && m.mask) CU(cudaMemcpyAsync(ctx->d_save_v + 6 * VP, m.mask, VP * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_save_bb, m.bb, 6 * N * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_save_bb + 6 * N, m.obb, 6 * N * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_save_flag, m.node_flag, N * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
  ctx->save_stale_flags = ctx->stale_flags;
  ctx->have_save = true;
  return DSC_OK;
}
int dsc_state_restore(DscContext *ctx)
{
  NEED_PBVH();
  if (ctx->in_stroke) return fail(ctx, DSC_ERR_STATE, "no rollback inside a stroke");
  if (!ctx->have_save) return fail(ctx, DSC_ERR_STATE, "dsc_state_save first");
  int r = join_side(ctx);
  if (r) return r;
  DevMesh &m = ctx->m;
  const size_t VP = (size_t)ctx->vpad, N = (size_t)ctx->totnode;
  float *dst[6] = {m.cx, m.cy, m.cz, m.nx, m.ny, m.nz};
  for (int k = 0; k < 6; k++) CU(cudaMemcpyAsync(dst[k], ctx->d_save_v + k * VP, VP * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  if (ctx->is_grids && m.mask) CU(cudaMemcpyAsync(ctx->d_mask, ctx->d_save_v + 6 * VP, VP * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  CU(cudaMemcpyAsync(m.bb, ctx->d_save_bb, 6 * N * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  CU(cudaMemcpyAsync(m.obb, ctx->d_save_bb + 6 * N, 6 * N * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  CU(cudaMemcpyAsync(m.node_flag, ctx->d_save_flag, N * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
  CU(cudaMemsetAsync(m.dirty, 0, sizeof(unsigned) * (size_t)ctx->nwords, ctx->stream));
  ctx->stale_flags = ctx->save_stale_flags;
  ctx->launches += 11;
  return DSC_OK;
}

int dsc_synchronize(DscContext *ctx)
{
  if (!ctx) return DSC_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  return sync_all(ctx);
}

int dsc_timer_start(DscContext *ctx)
{
  if (!ctx) return DSC_ERR_INVALID;
  int r = join_side(ctx);
  if (r) return r;
  CU(cudaEventRecord(ctx->t0, ctx->stream));
  return DSC_OK;
}
int dsc_timer_stop(DscContext *ctx, float *r_ms)
{
  if (!ctx) return DSC_ERR_INVALID;
  int r = join_side(ctx); /* the timed region ends when the side stream has drained too */
  if (r) return r;
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
  int r = sync_all(ctx);
  if (r) return r;
  for (auto &ev : ctx->events) {
    cudaEventDestroy(ev.a);
    cudaEventDestroy(ev.b);
  }
  ctx->events.clear();
  for (int i = 0; i < DSC_NUM_STAGES; i++) {
    ctx->stage_ms[i] = 0.0f;
    ctx->stage_launches[i] = 0;
  }
  if (ctx->stage_timing == 3) cupti_stop();
  if (enable == 3) {
    if (!cupti_start()) {
      ctx->stage_timing = 0;
      return fail(ctx, DSC_ERR_UNSUPPORTED, "CUPTI kernel tracing is not available (libcupti not found, or a profiler owns the device)");
    }
    ctx->stage_timing = 3;
    return DSC_OK;
  }
  ctx->stage_timing = enable == 2 ? 2 : (enable != 0 ? 1 : 0);
  return DSC_OK;
}

int dsc_stage_times(DscContext *ctx, float r_ms[DSC_NUM_STAGES], int r_launches[DSC_NUM_STAGES])
{
  if (!ctx) return DSC_ERR_INVALID;
  int r = sync_all(ctx);
  if (r) return r;
  for (auto &ev : ctx->events) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, ev.a, ev.b) == cudaSuccess) ctx->stage_ms[ev.stage] += ms;
    cudaEventDestroy(ev.a);
    cudaEventDestroy(ev.b);
  }
  ctx->events.clear();
  if (ctx->stage_timing == 3) {
    /* everything queued has run (sync_all above): pull the records */
    g_cupti.FlushAll(1);
    std::lock_guard<std::mutex> lk(g_cupti_mu);
    for (int i = 0; i < DSC_NUM_STAGES; i++) {
      if (r_ms) r_ms[i] = (float)g_cupti_ms[i];
      if (r_launches) r_launches[i] = (int)g_cupti_n[i];
    }
    return DSC_OK;
  }
  for (int i = 0; i < DSC_NUM_STAGES; i++) {
    if (r_ms) r_ms[i] = ctx->stage_ms[i];
    if (r_launches) r_launches[i] = ctx->stage_launches[i];
  }
  return DSC_OK;
}

} /* extern "C" */
