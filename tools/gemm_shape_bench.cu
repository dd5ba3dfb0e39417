// Which square tile does gemm_dmma_kernel want for the skinny products of the row-distributed Lowdin transform
// (M = n / ranks rows, N = K = n, both spins)?  Times launch_gemm with forced 32 / 64 / 128 tiles and with the heuristic.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/gemm_shape_bench tools/gemm_shape_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../nbed_b200/csrc/gemm.cuh"
using namespace nbd;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s -> %s line %d\n", #x, cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1376;
  const size_t nn = (size_t)n * n;
  double *A, *B, *C;
  CK(cudaMalloc(&A, 2 * nn * 8)); CK(cudaMalloc(&B, 2 * nn * 8)); CK(cudaMalloc(&C, 2 * nn * 8));
  std::vector<double> h(2 * nn);
  for (auto& x : h) x = rand() / (double)RAND_MAX - 0.5;
  CK(cudaMemcpy(A, h.data(), 2 * nn * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(B, h.data(), 2 * nn * 8, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int ranks : {1, 2, 4, 8, 16}) {
    const int M = (n + ranks - 1) / ranks;
    for (int tile : {0, 32, 64, 128}) {
      GemmArgs g{};
      g.M = M; g.N = n; g.K = n;
      g.A = A; g.a_is = n; g.a_ks = 1;          // rows of X (row-major)
      g.B = B; g.b_js = 1; g.b_ks = n;          // F (row-major, [K][N])
      g.C = C; g.ldc = n; g.alpha = 1.0; g.beta = 0.0;
      g.batch = 2; g.strideA = 0; g.strideB = (long)nn; g.strideC = (long)nn;
      float best = 1e30f;
      for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        CK(launch_gemm(0, g, 0, nullptr, 148, tile));
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
      }
      printf("ranks=%d M=%d N=K=%d batch 2, tile %3d%s: %.1f us (%.1f TFLOP/s)\n", ranks, M, n, tile, tile ? "" : " (heuristic)", 1e3 * best,
             2.0 * 2 * M * (double)n * n / (best * 1e-3) / 1e12);
    }
  }
  return 0;
}
