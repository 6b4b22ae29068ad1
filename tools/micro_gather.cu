// micro-benchmark: k_gather / k_tag_ancestors latency on a synthetic leaf set (not part of the product)
#include "../dune_sculpt_b200/csrc/dsc_kernels.cuh"
#include <cstdio>
#include <vector>
#include <cmath>
__global__ void k_empty() {}
__global__ void __launch_bounds__(1024) k_empty1024(int *p) { if (threadIdx.x == 2000) *p = 1; }
int main(int argc, char **argv)
{
  int nleaf = argc > 1 ? atoi(argv[1]) : 4096;
  int side = (int)std::sqrt((double)nleaf);
  int tn = 2 * nleaf;
  DevMesh m; memset(&m, 0, sizeof(m));
  std::vector<float> bb(6 * (size_t)tn, 0.f);
  for (int l = 0; l < nleaf; l++) {
    float x0 = -1.f + 2.f * (l % side) / side, y0 = -1.f + 2.f * (l / side) / side, w = 2.f / side;
    bb[0 * tn + l] = x0; bb[1 * tn + l] = y0; bb[2 * tn + l] = -0.05f;
    bb[3 * tn + l] = x0 + w; bb[4 * tn + l] = y0 + w; bb[5 * tn + l] = 0.05f;
  }
  cudaMalloc(&m.bb, bb.size() * 4); cudaMemcpy(m.bb, bb.data(), bb.size() * 4, cudaMemcpyHostToDevice);
  m.obb = m.bb;
  cudaMalloc(&m.node_flag, tn * 4); cudaMemset(m.node_flag, 0, tn * 4);
  cudaMalloc(&m.leaf_state, nleaf * 4); cudaMemset(m.leaf_state, 0, nleaf * 4);
  int *ucnt; cudaMalloc(&ucnt, nleaf * 4); cudaMemset(ucnt, 0, nleaf * 4); m.leaf_ucnt = ucnt;
  cudaMalloc(&m.hit_list, nleaf * 4); cudaMalloc(&m.search_list, nleaf * 4); cudaMalloc(&m.area_list, nleaf * 4);
  cudaMalloc(&m.st, sizeof(DabState)); cudaMemset(m.st, 0, sizeof(DabState));
  m.nleaf = nleaf; m.totnode = tn; m.max_chunks = 4;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int *dummy; cudaMalloc(&dummy, 4);
  float rads[3] = {0.03f, 0.3f, 1.4f};
  for (int it = 0; it < 3; it++) {
    for (int mark = 0; mark < 2; mark++) {
      float r = rads[it];
      for (int w = 0; w < 20; w++) k_gather<<<1, 1024>>>(m, 0.1f, 0.2f, 0.f, r * r, r * r * 0.25f, 0, 1, mark);
      cudaEventRecord(a);
      for (int w = 0; w < 200; w++) k_gather<<<1, 1024>>>(m, 0.1f, 0.2f, 0.f, r * r, r * r * 0.25f, 0, 1, mark);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      DabState st; cudaMemcpy(&st, m.st, sizeof(st), cudaMemcpyDeviceToHost);
      printf("nleaf %d r=%.2f mark=%d: %.2f us per gather (hits %d/%d)\n", nleaf, r, mark, ms * 1000.f / 200, st.hit_count, st.search_count);
    }
  }
  for (int w = 0; w < 20; w++) k_empty<<<1, 32>>>();
  cudaEventRecord(a);
  for (int w = 0; w < 200; w++) k_empty<<<1, 32>>>();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("empty kernel: %.2f us\n", ms * 1000.f / 200);
  cudaEventRecord(a);
  for (int w = 0; w < 200; w++) k_empty1024<<<1, 1024>>>(dummy);
  cudaEventRecord(b); cudaEventSynchronize(b);
  cudaEventElapsedTime(&ms, a, b);
  printf("empty 1024-thread kernel: %.2f us\n", ms * 1000.f / 200);
  cudaEventRecord(a);
  for (int w = 0; w < 200; w++) k_empty1024<<<1184, 256>>>(dummy);
  cudaEventRecord(b); cudaEventSynchronize(b);
  cudaEventElapsedTime(&ms, a, b);
  printf("empty 1184x256 kernel: %.2f us\n", ms * 1000.f / 200);
  return 0;
}
