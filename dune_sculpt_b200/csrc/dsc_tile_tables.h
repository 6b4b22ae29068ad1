/* Host-side construction of the tile kernel's tables (no CUDA calls): the vertex -> looptri CSR in slot space and, per tile,
 * the staged-vert list, the local poly entries, the sliced-ELL index words and the tile descriptor.  Split out of
 * dsc_pbvh_upload so that (a) the leaves are processed in parallel -- a leaf's tables depend on nothing but the mesh and the
 * slot layout; the pieces are then appended in leaf order, which reproduces the serial layout byte for byte -- and (b) the
 * result can be checked without a GPU (tests/test_tile_tables.py builds this header with g++ and compares it with the
 * serial construction it replaced). */
#ifndef DSC_TILE_TABLES_H
#define DSC_TILE_TABLES_H

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__CUDACC__)
#define DSC_HD __host__ __device__
#else
#define DSC_HD
#endif

#ifndef DSC_TILE
#define DSC_TILE 1024 /* most unique verts of a tile */
#endif
#define DSC_TILE_SMEM_BUDGET (96 * 1024) /* shared memory one tile may ask of k_normals_tile; heavier tiles take the general path */

/* tile_meta: 3 x int4 per tile */
struct TileMeta {
  int ubeg, ucnt, sbeg, sbb;     /* unique slot run; staged verts: offset, how many count in the leaf box */
  int xcnt, ebeg, eown, ehalo;   /* further staged verts; entries: offset (even), own-leaf, other-leaf */
  int hbeg, leaf, tile0, ntfast; /* e_halo_leaf offset; leaf; its first tile; tile count | fast << 16 */
};
/* shared-memory regions of the tile kernel: positions SoA [3][nloc_a], poly normals float4 [ne + 1],
 * poly entries ushort4 [ne_a], index words [v2w], other-leaf entry switches [ehalo]; each region is
 * sized for the largest tile of the mesh (DevMesh.sm_off_*) */
DSC_HD inline int dsc_tile_nloc_a(int ucnt, int sbb, int xcnt) { return (((ucnt + 3) & ~3) + sbb + xcnt + 3) & ~3; }
DSC_HD inline size_t dsc_tile_smem_bytes(int nloc_a, int ne, int v2w, int ehalo)
{
  return 12 * (size_t)nloc_a + 16 * ((size_t)ne + 1) + 8 * (size_t)((ne + 1) & ~1) + 4 * (size_t)v2w + (((size_t)ehalo + 15) & ~(size_t)15);
}

struct DscTileRange { int x, y; }; /* first slot, unique verts (layout of int2) */
struct TileDims { int tile, nloc_a, ne, v2w, ehalo; };

struct TileTablesIn {
  int L, VP, T, NT, totpoly;
  const int *leaves;                                   /* [L] node of each leaf, traversal order */
  const int *vert_indices, *vert_offset, *prim_indices; /* DscPbvhDesc */
  const int *slot_of;                                  /* vertex -> slot */
  const DscTileRange *tile_range;                      /* [NT] */
  const int *leaf_tile0;                               /* [L + 1] */
  const int *leaf_ucnt, *leaf_scnt, *leaf_pbeg, *leaf_pcnt;
  const int *tri_vert, *tri_poly, *poly_start, *poly_len, *loop_v;
};
struct TileTablesOut {
  std::vector<int> tri_leaf;             /* [T] looptri position -> leaf */
  std::vector<unsigned> vt_off, vt_idx;  /* slot -> looptri positions, ascending */
  std::vector<int> leaf_sslots, leaf_sbeg; /* per leaf: slots of its shared verts */
  std::vector<int> stage;                /* per tile: staged slots, box-counted ones first */
  std::vector<unsigned short> e_pv;      /* 4 per entry */
  std::vector<int> e_halo_leaf;          /* per other-leaf entry: that leaf */
  std::vector<TileMeta> tmeta;
  std::vector<unsigned> v2_goff, v2_idx;
  std::vector<unsigned char> leaf_fast;
  std::vector<TileDims> tile_dims;
  bool any_slow_leaf = false;
};

enum { DSC_TT_OK = 0, DSC_TT_BAD_PRIM = 1, DSC_TT_TOO_MANY_TILES = 2 };

/* threads <= 0: as many as OpenMP gives.  *r_where: the offending prim position / leaf on error. */
inline int dsc_build_tile_tables(const TileTablesIn &in, TileTablesOut &out, int threads, int *r_where)
{
  const int L = in.L, VP = in.VP, T = in.T, NT = in.NT;
  const bool timing = getenv("DSC_TIMING") != nullptr;
  auto now = []() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
  };
  double t_mark = now();
  auto mark = [&](const char *what) {
    if (!timing) return;
    const double t = now();
    fprintf(stderr, "[dsc tile tables] %-22s %.3f s\n", what, t - t_mark);
    t_mark = t;
  };
  /* ---- looptris by position; vertex -> looptri CSR ----
   * a slot's list holds the positions of its looptris in ascending order (the order a single-threaded
   * pbvh_update_normals_accum_task_cb adds them): counted and filled in parallel with atomic cursors, then every list
   * sorted -- the same table as a serial fill in position order */
#ifdef _OPENMP
  const int nthreads = threads > 0 ? threads : omp_get_max_threads();
#else
  const int nthreads = 1;
  (void)threads;
#endif
  out.tri_leaf.assign((size_t)std::max(T, 1), 0);
  std::vector<int> &tri_leaf = out.tri_leaf;
  std::vector<unsigned> deg((size_t)VP + 1, 0);
  int bad_pos = -1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int l = 0; l < L; l++) {
    for (int pos = in.leaf_pbeg[l]; pos < in.leaf_pbeg[l] + in.leaf_pcnt[l]; pos++) {
      const int t = in.prim_indices[pos];
      if (t < 0 || t >= T) {
#pragma omp atomic write
        bad_pos = pos;
        continue;
      }
      tri_leaf[pos] = l;
      for (int k = 0; k < 3; k++) {
        unsigned *d = &deg[in.slot_of[in.tri_vert[(size_t)3 * t + k]]];
#pragma omp atomic update
        (*d)++;
      }
    }
  }
  if (bad_pos >= 0) {
    if (r_where) *r_where = bad_pos;
    return DSC_TT_BAD_PRIM;
  }
  out.vt_off.assign((size_t)VP + 1, 0);
  std::vector<unsigned> &vt_off = out.vt_off;
  for (int s = 0; s < VP; s++) vt_off[s + 1] = vt_off[s] + deg[s];
  out.vt_idx.assign((size_t)std::max<unsigned>(vt_off[VP], 1u), 0u);
  std::vector<unsigned> &vt_idx = out.vt_idx;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int s = 0; s <= VP; s++) deg[s] = 0u;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int pos = 0; pos < T; pos++) {
    const int t = in.prim_indices[pos];
    /* the reference adds the face normal for corner j = 2, 1, 0 (pbvh.c:2966); a vertex that is
     * listed twice in one looptri gets it twice -- same here, order within a looptri is moot */
    for (int k = 0; k < 3; k++) {
      const int s = in.slot_of[in.tri_vert[(size_t)3 * t + k]];
      unsigned at;
#pragma omp atomic capture
      at = deg[s]++;
      vt_idx[vt_off[s] + at] = (unsigned)pos;
    }
  }
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int s = 0; s < VP; s++) {
    const unsigned b = vt_off[s], e = vt_off[s + 1];
    if (e - b > 1) std::sort(vt_idx.begin() + b, vt_idx.begin() + e);
  }
  std::vector<unsigned>().swap(deg);

  mark("vertex -> looptri CSR");
  /* ---- local tables, leaf by leaf ---- */
  out.tmeta.assign((size_t)std::max(NT, 1), TileMeta());
  out.v2_goff.assign((size_t)VP / 32 + 1, 0u);
  out.leaf_fast.assign((size_t)L, (unsigned char)1);
  out.leaf_sbeg.assign((size_t)L, 0);
  struct LeafOut {
    std::vector<int> leaf_sslots, stage, e_halo_leaf;
    std::vector<unsigned short> e_pv;
    std::vector<unsigned> v2_idx, tile_v2_begin;
    std::vector<TileDims> tile_dims;
    bool slow = false, too_many_tiles = false;
  };
  /* slot / poly stamps: value + 1, so that the zero pages calloc hands out mean "never seen" -- a thread only ever
   * touches the pages of the leaves it is given (contiguous chunks) and their rims */
  struct Scratch {
    int *lstamp, *lidx, *sh_leaf, *sh_done, *pstamp, *pentry;
    std::unordered_map<unsigned long long, int> halo;
    std::vector<int> own_polys, halo_polys, halo_leaves, st_bb, st_x;
    std::vector<unsigned> rows, raw, goff_local; /* entry ids of the current group, [row][lane]; bit 31 = other-leaf entry */
    Scratch(int VP_, int totpoly)
    {
      lstamp = (int *)calloc((size_t)VP_ + 1, sizeof(int));
      lidx = (int *)calloc((size_t)VP_ + 1, sizeof(int));
      sh_leaf = (int *)calloc((size_t)VP_ + 1, sizeof(int));
      sh_done = (int *)calloc((size_t)VP_ + 1, sizeof(int));
      pstamp = (int *)calloc((size_t)std::max(totpoly, 1), sizeof(int));
      pentry = (int *)calloc((size_t)std::max(totpoly, 1), sizeof(int));
    }
    ~Scratch()
    {
      free(lstamp); free(lidx); free(sh_leaf); free(sh_done); free(pstamp); free(pentry);
    }
  };
  std::vector<LeafOut> outs((size_t)L);
  auto process_leaf = [&](int l, Scratch &S_, LeafOut &o) {
      const int n = in.leaves[l];
      const int U_leaf = in.leaf_ucnt[l], S = in.leaf_scnt[l];
      const int pbeg = in.leaf_pbeg[l], pend = pbeg + in.leaf_pcnt[l];
      const int *vi = in.vert_indices + in.vert_offset[n];
      for (int i = 0; i < S; i++) {
        const int sl = in.slot_of[vi[U_leaf + i]];
        o.leaf_sslots.push_back(sl);
        S_.sh_leaf[sl] = l + 1;
      }
      bool ok = true;
      const int t_lo = in.leaf_tile0[l], t_hi = in.leaf_tile0[l + 1];
      for (int tg = t_lo; tg < t_hi; tg++) {
        const int ub = in.tile_range[tg].x, U = in.tile_range[tg].y;
        TileMeta &tm = out.tmeta[tg];
        tm.ubeg = ub;
        tm.ucnt = U;
        tm.sbeg = (int)o.stage.size();
        if ((o.e_pv.size() / 4) & 1) o.e_pv.insert(o.e_pv.end(), 4, (unsigned short)0); /* bulk copies start 16-byte aligned */
        tm.ebeg = (int)(o.e_pv.size() / 4);
        tm.hbeg = (int)o.e_halo_leaf.size();
        tm.leaf = l;
        tm.tile0 = t_lo;
        for (int i = 0; i < U; i++) {
          S_.lstamp[ub + i] = tg + 1;
          S_.lidx[ub + i] = i;
        }
        S_.own_polys.clear();
        S_.halo_polys.clear();
        S_.halo_leaves.clear();
        S_.halo.clear();
        const int G0 = ub / 32, ng = (U + 31) / 32;
        o.tile_v2_begin.push_back((unsigned)o.v2_idx.size());
        S_.raw.clear();
        S_.goff_local.assign((size_t)ng, 0u);
        for (int g = 0; g < ng; g++) {
          const int i0 = g * 32, cntv = std::min(32, U - i0);
          int width = 0;
          for (int i = 0; i < cntv; i++) width = std::max(width, (int)(vt_off[ub + i0 + i + 1] - vt_off[ub + i0 + i]));
          const int wpairs = (width + 1) / 2;
          S_.rows.assign((size_t)64 * std::max(wpairs, 1), 0xffffffffu);
          for (int i = 0; i < cntv; i++) {
            const int sl = ub + i0 + i;
            int r2 = 0;
            for (unsigned q = vt_off[sl]; q < vt_off[sl + 1]; q++, r2++) {
              const int pos = (int)vt_idx[q];
              const int p = in.tri_poly[in.prim_indices[pos]];
              unsigned e;
              if (pos >= pbeg && pos < pend) {
                if (S_.pstamp[p] != tg + 1) {
                  S_.pstamp[p] = tg + 1;
                  S_.pentry[p] = (int)S_.own_polys.size();
                  S_.own_polys.push_back(p);
                }
                e = (unsigned)S_.pentry[p];
              }
              else {
                const int ol = tri_leaf[pos];
                const unsigned long long key = ((unsigned long long)(unsigned)p << 32) | (unsigned)ol;
                auto it = S_.halo.find(key);
                if (it == S_.halo.end()) {
                  it = S_.halo.emplace(key, (int)S_.halo_polys.size()).first;
                  S_.halo_polys.push_back(p);
                  S_.halo_leaves.push_back(ol);
                }
                e = 0x80000000u | (unsigned)it->second;
              }
              S_.rows[(size_t)r2 * 32 + i] = e;
            }
          }
          S_.goff_local[g] = (unsigned)(S_.raw.size() / 2);
          for (int w = 0; w < wpairs; w++) {
            for (int i = 0; i < 32; i++) {
              S_.raw.push_back(S_.rows[(size_t)(2 * w) * 32 + i]);
              S_.raw.push_back(S_.rows[(size_t)(2 * w + 1) * 32 + i]);
            }
          }
        }
        const int eown = (int)S_.own_polys.size(), ehalo = (int)S_.halo_polys.size(), ne = eown + ehalo;
        tm.eown = eown;
        tm.ehalo = ehalo;
        if (ne >= 0xffff) ok = false;
        /* pack the rows two entries to a word; other-leaf ids follow the own ones, padding -> the zero entry `ne` */
        {
          auto fix = [&](unsigned id) -> unsigned {
            if (id == 0xffffffffu) return (unsigned)std::min(ne, 0xffff);
            if (id & 0x80000000u) return (unsigned)std::min(eown + (int)(id & 0x7fffffffu), 0xfffe);
            return (unsigned)std::min((int)id, 0xfffe);
          };
          const unsigned base = (unsigned)o.v2_idx.size();
          for (size_t q = 0; q + 1 < S_.raw.size(); q += 2) o.v2_idx.push_back(fix(S_.raw[q]) | (fix(S_.raw[q + 1]) << 16));
          for (int g = 0; g < ng; g++) out.v2_goff[G0 + g] = base + S_.goff_local[g];
        }
        /* staged verts: corners of the entries that are not unique verts of this tile */
        S_.st_bb.clear();
        S_.st_x.clear();
        auto visit_poly = [&](int p) {
          const int ls = in.poly_start[p], len = in.poly_len[p];
          if (len != 3 && len != 4) {
            ok = false; /* n-gon: this leaf takes the general path */
            return;
          }
          for (int k = 0; k < len; k++) {
            const int sl = in.slot_of[in.loop_v[ls + k]];
            if (S_.lstamp[sl] == tg + 1) continue;
            S_.lstamp[sl] = tg + 1;
            S_.lidx[sl] = -1;
            if (S_.sh_leaf[sl] == l + 1 && S_.sh_done[sl] != l + 1) {
              S_.sh_done[sl] = l + 1;
              S_.st_bb.push_back(sl);
            }
            else {
              S_.st_x.push_back(sl);
            }
          }
        };
        for (int p : S_.own_polys) visit_poly(p);
        for (int p : S_.halo_polys) visit_poly(p);
        if (tg == t_hi - 1) {
          /* shared verts of the leaf that no tile reached through a poly of its unique verts */
          for (int i = 0; i < S; i++) {
            const int sl = o.leaf_sslots[(size_t)i];
            if (S_.sh_done[sl] != l + 1) {
              S_.sh_done[sl] = l + 1;
              if (S_.lstamp[sl] == tg + 1 && S_.lidx[sl] == -1) {
                /* staged here already as a plain corner: move it to the box-counted part */
                S_.st_x.erase(std::find(S_.st_x.begin(), S_.st_x.end(), sl));
              }
              S_.lstamp[sl] = tg + 1;
              S_.lidx[sl] = -1;
              S_.st_bb.push_back(sl);
            }
          }
        }
        std::sort(S_.st_bb.begin(), S_.st_bb.end());
        std::sort(S_.st_x.begin(), S_.st_x.end());
        int nloc = (U + 3) & ~3; /* staged verts follow the 16-byte padded unique run */
        for (int sl : S_.st_bb) {
          S_.lidx[sl] = nloc++;
          o.stage.push_back(sl);
        }
        for (int sl : S_.st_x) {
          S_.lidx[sl] = nloc++;
          o.stage.push_back(sl);
        }
        tm.sbb = (int)S_.st_bb.size();
        tm.xcnt = (int)S_.st_x.size();
        if (nloc > 0xfffe) ok = false;
        bool allquad = true;
        auto emit = [&](int p) {
          const int ls = in.poly_start[p], len = in.poly_len[p];
          if (len != 4) allquad = false;
          unsigned short loc[4] = {0, 0, 0, 0xffff};
          if (len == 3 || len == 4) {
            for (int k = 0; k < len; k++) loc[k] = (unsigned short)std::min(std::max(S_.lidx[in.slot_of[in.loop_v[ls + k]]], 0), 0xfffe);
          }
          o.e_pv.insert(o.e_pv.end(), loc, loc + 4);
        };
        for (int p : S_.own_polys) emit(p);
        for (int p : S_.halo_polys) emit(p);
        o.e_halo_leaf.insert(o.e_halo_leaf.end(), S_.halo_leaves.begin(), S_.halo_leaves.end());
        tm.ntfast = allquad ? 1 << 17 : 0;
        const size_t bytes = dsc_tile_smem_bytes(dsc_tile_nloc_a(U, tm.sbb, tm.xcnt), ne, (int)(S_.raw.size() / 2), ehalo);
        if (bytes > DSC_TILE_SMEM_BUDGET) {
          ok = false; /* a tile this heavy would push the regions past what an SM can hold: general path */
        }
        else {
          o.tile_dims.push_back({tg, dsc_tile_nloc_a(U, tm.sbb, tm.xcnt), ne, (int)(S_.raw.size() / 2), ehalo});
        }
      }
      if (!ok) {
        out.leaf_fast[l] = 0;
        o.slow = true;
        while (!o.tile_dims.empty() && o.tile_dims.back().tile >= t_lo) o.tile_dims.pop_back();
      }
      for (int tg = t_lo; tg < t_hi; tg++) out.tmeta[tg].ntfast |= (t_hi - t_lo) | (ok ? 1 << 16 : 0); /* bit 17: all entries are quads */
      if (t_hi - t_lo > 0xffff) o.too_many_tiles = true;
  };
#pragma omp parallel num_threads(nthreads)
  {
    Scratch S_(VP, in.totpoly);
#pragma omp for schedule(static)
    for (int l = 0; l < L; l++) process_leaf(l, S_, outs[(size_t)l]);
  }
  mark("leaves (parallel)");
  /* ---- the leaves' pieces appended in leaf order: the layout of the serial construction ----
   * where each leaf's pieces start follows from the sizes alone (serial, cheap); the copies and the offset fix-ups of a
   * leaf touch only its own ranges (parallel) */
  struct LeafBase { size_t sslots, stage, e_entries, halo, v2, dims; bool pad; };
  std::vector<LeafBase> base((size_t)L + 1);
  {
    LeafBase c = {0, 0, 0, 0, 0, 0, false};
    for (int l = 0; l < L; l++) {
      const LeafOut &o = outs[(size_t)l];
      if (o.too_many_tiles) {
        if (r_where) *r_where = l;
        return DSC_TT_TOO_MANY_TILES;
      }
      if (o.slow) out.any_slow_leaf = true;
      c.pad = (c.e_entries & 1) != 0; /* the pad the leaf's first tile asks for: bulk copies start 16-byte aligned */
      if (c.pad) c.e_entries++;
      base[(size_t)l] = c;
      c.sslots += o.leaf_sslots.size();
      c.stage += o.stage.size();
      c.e_entries += o.e_pv.size() / 4;
      c.halo += o.e_halo_leaf.size();
      c.v2 += o.v2_idx.size();
      c.dims += o.tile_dims.size();
    }
    c.pad = false;
    base[(size_t)L] = c;
    out.leaf_sslots.assign(c.sslots, 0);
    out.stage.assign(c.stage, 0);
    out.e_pv.assign(c.e_entries * 4, (unsigned short)0); /* the pads stay zero */
    out.e_halo_leaf.assign(c.halo, 0);
    out.v2_idx.assign(c.v2, 0u);
    out.tile_dims.assign(c.dims, TileDims());
  }
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int l = 0; l < L; l++) {
    LeafOut &o = outs[(size_t)l];
    const LeafBase &b = base[(size_t)l];
    out.leaf_sbeg[l] = (int)b.sslots;
    std::copy(o.leaf_sslots.begin(), o.leaf_sslots.end(), out.leaf_sslots.begin() + (long)b.sslots);
    std::copy(o.stage.begin(), o.stage.end(), out.stage.begin() + (long)b.stage);
    std::copy(o.e_pv.begin(), o.e_pv.end(), out.e_pv.begin() + (long)(b.e_entries * 4));
    std::copy(o.e_halo_leaf.begin(), o.e_halo_leaf.end(), out.e_halo_leaf.begin() + (long)b.halo);
    std::copy(o.v2_idx.begin(), o.v2_idx.end(), out.v2_idx.begin() + (long)b.v2);
    std::copy(o.tile_dims.begin(), o.tile_dims.end(), out.tile_dims.begin() + (long)b.dims);
    const int t_lo = in.leaf_tile0[l], t_hi = in.leaf_tile0[l + 1];
    for (int tg = t_lo; tg < t_hi; tg++) {
      TileMeta &tm = out.tmeta[tg];
      tm.sbeg += (int)b.stage;
      tm.ebeg += (int)b.e_entries;
      tm.hbeg += (int)b.halo;
      const int G0 = tm.ubeg / 32, ng = (tm.ucnt + 31) / 32;
      /* the groups of the pad slots between the previous tile (of this leaf or the one before) and this one point at
       * this tile's first word; tile_meta's ubeg / ucnt were final after the leaf pass */
      int g_prev = 0;
      if (tg > 0) g_prev = out.tmeta[tg - 1].ubeg / 32 + (out.tmeta[tg - 1].ucnt + 31) / 32;
      const unsigned first_word = (unsigned)b.v2 + o.tile_v2_begin[(size_t)(tg - t_lo)];
      for (int g = g_prev; g < G0; g++) out.v2_goff[g] = first_word;
      for (int g = 0; g < ng; g++) out.v2_goff[G0 + g] += (unsigned)b.v2;
    }
    std::vector<int>().swap(o.leaf_sslots); /* release as we go */
    std::vector<int>().swap(o.stage);
    std::vector<unsigned short>().swap(o.e_pv);
    std::vector<unsigned>().swap(o.v2_idx);
  }
  int next_group_to_fill = NT > 0 ? out.tmeta[(size_t)NT - 1].ubeg / 32 + (out.tmeta[(size_t)NT - 1].ucnt + 31) / 32 : 0;
  for (; next_group_to_fill <= VP / 32; next_group_to_fill++) out.v2_goff[next_group_to_fill] = (unsigned)out.v2_idx.size();
  if (out.v2_idx.empty()) out.v2_idx.push_back(0u);
  out.e_pv.insert(out.e_pv.end(), 8, (unsigned short)0); /* the last tile's bulk copy may read one entry past its own */
  mark("append in leaf order");
  return DSC_TT_OK;
}

#endif /* DSC_TILE_TABLES_H */
