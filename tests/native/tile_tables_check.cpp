/* CPU check of dune_sculpt_b200/csrc/dsc_tile_tables.h (no CUDA): builds a slot layout for a flattened mesh + PBVH and
 * hashes every table the tile kernel reads, through the serial reference construction (mode 0) or the shipped one with
 * `mode` threads.  tests/test_tile_tables.py compares the hashes. */
#include "../../include/dune_sculpt_cuda.h"
#include "../../dune_sculpt_b200/csrc/dsc_tile_tables.h"
#include "tile_tables_ref.inc"

#include <cstdint>
#include <numeric>

static uint64_t fnv(uint64_t h, const void *p, size_t n)
{
  const unsigned char *b = (const unsigned char *)p;
  for (size_t i = 0; i < n; i++) {
    h ^= b[i];
    h *= 1099511628211ull;
  }
  return h;
}
template<typename T> static uint64_t fnv_vec(uint64_t h, const std::vector<T> &v)
{
  const uint64_t n = v.size();
  h = fnv(h, &n, sizeof(n));
  return v.empty() ? h : fnv(h, v.data(), sizeof(T) * v.size());
}

/* tile: unique verts per tile (<= 1024); scramble: permute a leaf's unique verts (tiles that are not spatially compact:
 * more corners staged from other tiles).  r_hashes[12]: one per table + the flags.  Returns the builder's status. */
extern "C" int tt_hashes(const DscMeshDesc *me, const DscPbvhDesc *pd, int tile, int scramble, int mode, uint64_t *r_hashes)
{
  if (tile <= 0 || tile > DSC_TILE || (tile & 31)) return -1; /* like the real layout: a 32-slot group belongs to one tile */
  const int V = me->totvert, T = me->tottri, N = pd->totnode;
  std::vector<int> leaves;
  for (int n = 0; n < N; n++) {
    if (pd->flag[n] & 1) leaves.push_back(n);
  }
  std::sort(leaves.begin(), leaves.end(), [&](int a, int b) { return pd->prim_offset[a] < pd->prim_offset[b]; });
  const int L = (int)leaves.size();
  std::vector<int> slot_of((size_t)V, -1), leaf_ucnt(L), leaf_scnt(L), leaf_pbeg(L), leaf_pcnt(L), leaf_tile0((size_t)L + 1, 0);
  std::vector<DscTileRange> tile_range;
  long long cur = 0;
  uint64_t rng = 0x9e3779b97f4a7c15ull;
  for (int l = 0; l < L; l++) {
    const int n = leaves[l];
    cur = (cur + 31) & ~31ll;
    leaf_ucnt[l] = pd->uniq_verts[n];
    leaf_scnt[l] = pd->face_verts[n];
    leaf_pbeg[l] = pd->prim_offset[n];
    leaf_pcnt[l] = pd->totprim[n];
    const int U = leaf_ucnt[l];
    std::vector<int> ord(pd->vert_indices + pd->vert_offset[n], pd->vert_indices + pd->vert_offset[n] + U);
    if (scramble) {
      for (int i = U - 1; i > 0; i--) {
        rng = rng * 6364136223846793005ull + 1442695040888963407ull;
        std::swap(ord[i], ord[(int)((rng >> 33) % (uint64_t)(i + 1))]);
      }
    }
    for (int i = 0; i < U; i++) slot_of[ord[i]] = (int)cur + i;
    leaf_tile0[l] = (int)tile_range.size();
    if (U == 0) tile_range.push_back(DscTileRange{(int)cur, 0});
    for (int o = 0; o < U; o += tile) tile_range.push_back(DscTileRange{(int)cur + o, std::min(tile, U - o)});
    cur += U;
  }
  leaf_tile0[L] = (int)tile_range.size();
  cur = (cur + 31) & ~31ll;
  for (int v = 0; v < V; v++) {
    if (slot_of[v] < 0) slot_of[v] = (int)cur++;
  }
  const int VP = (int)((cur + 31) & ~31ll) + 32;
  TileTablesIn in;
  in.L = L; in.VP = VP; in.T = T; in.NT = (int)tile_range.size(); in.totpoly = me->totpoly;
  in.leaves = leaves.data();
  in.vert_indices = pd->vert_indices; in.vert_offset = pd->vert_offset; in.prim_indices = pd->prim_indices;
  in.slot_of = slot_of.data();
  in.tile_range = tile_range.data();
  in.leaf_tile0 = leaf_tile0.data();
  in.leaf_ucnt = leaf_ucnt.data(); in.leaf_scnt = leaf_scnt.data(); in.leaf_pbeg = leaf_pbeg.data(); in.leaf_pcnt = leaf_pcnt.data();
  in.tri_vert = me->tri_vert; in.tri_poly = me->tri_poly; in.poly_start = me->poly_loopstart; in.poly_len = me->poly_totloop;
  in.loop_v = me->loop_vert;
  TileTablesOut out;
  int where = -1;
  const int r = mode == 0 ? tile_tables_reference(in, out) : dsc_build_tile_tables(in, out, mode, &where);
  if (r) return r;
  const uint64_t seed = 1469598103934665603ull;
  r_hashes[0] = fnv_vec(seed, out.tri_leaf);
  r_hashes[1] = fnv_vec(fnv_vec(seed, out.vt_off), out.vt_idx);
  r_hashes[2] = fnv_vec(fnv_vec(seed, out.leaf_sslots), out.leaf_sbeg);
  r_hashes[3] = fnv_vec(seed, out.stage);
  r_hashes[4] = fnv_vec(seed, out.e_pv);
  r_hashes[5] = fnv_vec(seed, out.e_halo_leaf);
  r_hashes[6] = fnv_vec(seed, out.tmeta);
  r_hashes[7] = fnv_vec(seed, out.v2_goff);
  r_hashes[8] = fnv_vec(seed, out.v2_idx);
  r_hashes[9] = fnv_vec(seed, out.leaf_fast);
  r_hashes[10] = fnv_vec(seed, out.tile_dims);
  r_hashes[11] = out.any_slow_leaf ? 1 : 0;
  return 0;
}
