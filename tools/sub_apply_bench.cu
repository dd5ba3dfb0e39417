// Block product out = alpha (A Y - shift Y) - beta Z of the subspace eigensolver: the ring kernel (sub_apply2, 4-CTA
// clusters) against the single-shot bulk-copy kernel (sub_apply3, 8-CTA clusters) - correctness against a plain
// reference kernel and microseconds per product in a dependent chain (the Chebyshev filter's access pattern).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/sub_apply_bench tools/sub_apply_bench.cu
// usage: build/sub_apply_bench [n=1376] [ns=2] [reps=200]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../nbed_b200/csrc/subspace.cuh"
using namespace nbd;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s -> %s line %d\n", #x, cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__global__ void ref_kernel(SubApplyArgs a, int KB, int ns) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long tot = (long)ns * a.n * KB;
  if (idx >= tot) return;
  const int b = idx / ((long)a.n * KB), i = (idx / KB) % a.n, c = idx % KB;
  const double* A = a.A + (long)b * a.n * a.n;
  const double* Y = a.Y + (long)b * a.n * KB;
  double s = 0.0;
  for (int j = 0; j < a.n; ++j) s = fma(A[(long)i * a.n + j], Y[(long)j * KB + c], s);
  double v = a.alpha[b] * (s - a.shift[b] * Y[(long)i * KB + c]);
  if (a.Z) v -= a.beta[b] * a.Z[idx];
  a.out[idx] = v;
}

template <int KB>
static void launch(int variant, const SubApplyArgs& a, int ns, cudaStream_t st) {
  const int n = a.n;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.blockDim = dim3(128);
  cfg.stream = st;
  if (variant >= 2) {
    cfg.gridDim = dim3((n + SUB3_ROWS - 1) / SUB3_ROWS, SUB3_KS, ns);
    cfg.dynamicSmemBytes = sub_apply3_smem_bytes<KB>(n);
    attr[0].val.clusterDim.y = SUB3_KS;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    if (variant == 3) {
      cfg.numAttrs = 2;
      CK(cudaLaunchKernelEx(&cfg, sub_apply3_kernel<KB, 1>, a));
    } else {
      CK(cudaLaunchKernelEx(&cfg, sub_apply3_kernel<KB, 0>, a));
    }
  } else {
    cfg.gridDim = dim3((n + SUB2_ROWS - 1) / SUB2_ROWS, SUB2_KS, ns);
    cfg.dynamicSmemBytes = sub_apply2_smem_bytes<KB>();
    attr[0].val.clusterDim.y = SUB2_KS;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    if (variant == 1) {
      cfg.numAttrs = 2;
      CK(cudaLaunchKernelEx(&cfg, (sub_apply2_kernel<KB, 1>), a));
    } else {
      CK(cudaLaunchKernelEx(&cfg, (sub_apply2_kernel<KB, 0>), a));
    }
  }
}

template <int KB>
static void run(int n, int ns, int reps) {
  const size_t nn = (size_t)n * n, blk = (size_t)n * KB;
  std::vector<double> hA(ns * nn), hY(ns * blk);
  srand(7);
  for (int b = 0; b < ns; ++b)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j <= i; ++j) {
        double v = (rand() / (double)RAND_MAX - 0.5) * 0.02;
        if (i == j) v += -1.0 + 1.9 * i / n;
        hA[b * nn + (size_t)i * n + j] = hA[b * nn + (size_t)j * n + i] = v;
      }
  for (auto& y : hY) y = rand() / (double)RAND_MAX - 0.5;
  double *A, *Y[3], *R;
  CK(cudaMalloc(&A, ns * nn * 8));
  for (auto& y : Y) CK(cudaMalloc(&y, ns * blk * 8));
  CK(cudaMalloc(&R, ns * blk * 8));
  CK(cudaMemcpy(A, hA.data(), ns * nn * 8, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(sub_apply2_kernel<KB, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, sub_apply2_smem_bytes<KB>()));
  CK(cudaFuncSetAttribute(sub_apply2_kernel<KB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sub_apply2_smem_bytes<KB>()));
  CK(cudaFuncSetAttribute(sub_apply3_kernel<KB, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(sub_apply3_kernel<KB, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CK(cudaFuncSetAttribute(sub_apply3_kernel<KB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(sub_apply3_kernel<KB, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  std::vector<double> href(ns * blk), hout(ns * blk);
  for (int variant : {0, 1, 2, 3}) {  // 0 / 1: ring kernel without / with PDL; 2 / 3: single-shot kernel without / with PDL
    if (variant >= 2 && ((n & 1) || sub_apply3_smem_bytes<KB>(n) > 200 * 1024)) continue;
    // correctness: one product with all three terms
    CK(cudaMemcpy(Y[0], hY.data(), ns * blk * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(Y[1], hY.data(), ns * blk * 8, cudaMemcpyHostToDevice));
    SubApplyArgs a{};
    a.A = A; a.Y = Y[0]; a.Z = Y[1]; a.out = R; a.n = n;
    a.alpha[0] = 1.3; a.alpha[1] = 0.7; a.shift[0] = 0.1; a.shift[1] = -0.2; a.beta[0] = 0.4; a.beta[1] = 0.9;
    ref_kernel<<<(unsigned)((ns * blk + 127) / 128), 128, 0, st>>>(a, KB, ns);
    CK(cudaStreamSynchronize(st));
    CK(cudaMemcpy(href.data(), R, ns * blk * 8, cudaMemcpyDeviceToHost));
    a.out = Y[2];
    launch<KB>(variant, a, ns, st);
    CK(cudaStreamSynchronize(st));
    CK(cudaMemcpy(hout.data(), Y[2], ns * blk * 8, cudaMemcpyDeviceToHost));
    double worst = 0.0;
    for (size_t i = 0; i < hout.size(); ++i) worst = std::max(worst, std::fabs(hout[i] - href[i]));
    // timing: dependent chain, scaled so the block stays O(1)
    a.alpha[0] = a.alpha[1] = 0.5; a.shift[0] = a.shift[1] = 0.0; a.beta[0] = a.beta[1] = 0.5;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0, st));
      for (int i = 1; i <= reps; ++i) {
        a.Y = Y[(i - 1) % 3]; a.Z = Y[(i + 1) % 3]; a.out = Y[i % 3];
        launch<KB>(variant, a, ns, st);
      }
      CK(cudaEventRecord(e1, st));
      CK(cudaEventSynchronize(e1));
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = std::min(best, ms);
    }
    printf("n=%d KB=%d ns=%d variant=%d: max|out - ref| = %.3e, %.2f us per product (chain of %d)\n", n, KB, ns, variant, worst,
           1e3 * best / reps, reps);
    fflush(stdout);
  }
  cudaFree(A); for (auto& y : Y) cudaFree(y); cudaFree(R);
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1376, ns = argc > 2 ? atoi(argv[2]) : 2, reps = argc > 3 ? atoi(argv[3]) : 200;
  run<16>(n, ns, reps);
  run<32>(n, ns, reps);
  return 0;
}
