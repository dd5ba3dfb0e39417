// Microbenchmarks that set the roofline denominators MEASURED_PEAKS.json lacks (FP64) and validate the
// primitives the hot kernels are built from (DMMA m8n8k4, bulk-copy + mbarrier pipelines).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/microbench.cu -o build/microbench -lcublas -lcusolver
// Not part of the product path; results are copied into profiles/.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double a0, double b0) {
  double acc[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

// plain streaming read: each thread 2x double2 per iteration, grid-stride
__global__ void k_read(const double2* __restrict__ p, size_t n2, double* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  double s = 0;
  for (; i + 3 * stride < n2; i += 4 * stride) {
    double2 a = p[i], b = p[i + stride], c = p[i + 2 * stride], d = p[i + 3 * stride];
    s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
  }
  for (; i < n2; i += stride) { double2 a = p[i]; s += a.x + a.y; }
  if (s == 123.456) out[0] = s;
}

// ---- bulk-copy (TMA 1-D) + mbarrier ring: one producer lane, all warps consume ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int STAGES, int TILE_BYTES>
__global__ void __launch_bounds__(256, 1) k_bulk(const char* __restrict__ src, size_t ntiles, double* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + STAGES;
  unsigned char* bufs = smem + 1024;
  const int nwarps = blockDim.x / 32;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nwarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // tiles for this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  size_t my = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  double s = 0;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    // prologue
    size_t pre = my < STAGES ? my : STAGES;
    for (size_t k = 0; k < pre; ++k) {
      mbar_expect_tx(&full[k], TILE_BYTES);
      bulk_g2s(bufs + k * TILE_BYTES, src + (blockIdx.x + k * gridDim.x) * (size_t)TILE_BYTES, TILE_BYTES, &full[k]);
    }
  }
  for (size_t k = 0; k < my; ++k) {
    int st = k % STAGES;
    uint32_t ph = (k / STAGES) & 1;
    mbar_wait(&full[st], ph);
    const double2* t = reinterpret_cast<const double2*>(bufs + st * TILE_BYTES);
    for (int i = threadIdx.x; i < TILE_BYTES / 16; i += blockDim.x) { double2 v = t[i]; s += v.x + v.y; }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
    if (warp == 0 && lane == 0 && k + STAGES < my) {
      mbar_wait(&empty[st], ph);
      mbar_expect_tx(&full[st], TILE_BYTES);
      bulk_g2s(bufs + st * TILE_BYTES, src + (blockIdx.x + (k + STAGES) * gridDim.x) * (size_t)TILE_BYTES, TILE_BYTES, &full[st]);
    }
  }
  // reduce to check correctness of the data path
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) atomicAdd(out, s);
}

__global__ void k_fill(double* p, size_t n, double v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs=%d smem/SM=%zu L2=%d MB\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount,
         (size_t)prop.sharedMemPerMultiprocessor, prop.l2CacheSize >> 20);
  int nsm = prop.multiProcessorCount;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double* dout; CK(cudaMalloc(&dout, 64)); CK(cudaMemset(dout, 0, 64));

  // ---- DMMA / DFMA peak ----
  for (int wpc : {4, 8, 16, 32}) {
    int iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      k_dmma<8><<<nsm, wpc * 32>>>(dout, iters, 1.0, 1.0);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    double fl = (double)nsm * wpc * iters * 8 * 512.0;
    printf("DMMA m8n8k4 x8acc  warps/SM=%2d : %.2f TFLOP/s\n", wpc, fl / time_ms(e0, e1) * 1e-9);
  }
  {
    int iters = 20000, wpc = 8;
    CK(cudaEventRecord(e0));
    k_dmma<2><<<nsm, wpc * 32>>>(dout, iters, 1.0, 1.0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    printf("DMMA m8n8k4 x2acc  warps/SM=%2d : %.2f TFLOP/s (dependent-issue latency probe)\n", wpc,
           (double)nsm * wpc * iters * 2 * 512.0 / time_ms(e0, e1) * 1e-9);
    CK(cudaEventRecord(e0));
    k_dmma<1><<<nsm, 32>>>(dout, iters, 1.0, 1.0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    double ms = time_ms(e0, e1);
    printf("DMMA single-warp dependent chain: %.1f ns per DMMA (latency)\n", ms * 1e6 / iters);
  }
  for (int wpc : {8, 16, 32}) {
    int iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      k_dfma<16><<<nsm, wpc * 32>>>(dout, iters, 1.000001, 1e-9);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    double fl = (double)nsm * wpc * 32 * iters * 16 * 2.0;
    printf("DFMA x16acc        warps/SM=%2d : %.2f TFLOP/s\n", wpc, fl / time_ms(e0, e1) * 1e-9);
  }

  // ---- HBM streaming ----
  size_t bytes = (size_t)8 << 30;
  double* buf; CK(cudaMalloc(&buf, bytes));
  k_fill<<<nsm * 8, 256>>>(buf, bytes / 8, 1.0); CK(cudaDeviceSynchronize());
  for (int cps : {2, 4, 8}) {
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      k_read<<<nsm * cps, 512>>>(reinterpret_cast<const double2*>(buf), bytes / 16, dout);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    printf("HBM read (LDG.128) 8 GiB, %d CTAs/SM x512: %.1f GB/s\n", cps, bytes / time_ms(e0, e1) * 1e-6);
  }
  {
    constexpr int ST = 8, TB = 8192;
    size_t ntiles = bytes / TB;
    int smem = 1024 + ST * TB;
    CK(cudaFuncSetAttribute(k_bulk<ST, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaMemset(dout, 0, 8));
      CK(cudaEventRecord(e0));
      k_bulk<ST, TB><<<nsm, 256, smem>>>(reinterpret_cast<const char*>(buf), ntiles, dout);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    double h; CK(cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost));
    printf("HBM read (bulk 8KBx8 stages, 1 CTA/SM): %.1f GB/s  checksum %s\n", bytes / time_ms(e0, e1) * 1e-6,
           h == (double)(bytes / 8) ? "OK" : "BAD");
  }
  {
    constexpr int ST = 6, TB = 32768;
    size_t ntiles = bytes / TB;
    int smem = 1024 + ST * TB;
    CK(cudaFuncSetAttribute(k_bulk<ST, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaMemset(dout, 0, 8));
      CK(cudaEventRecord(e0));
      k_bulk<ST, TB><<<nsm, 256, smem>>>(reinterpret_cast<const char*>(buf), ntiles, dout);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    double h; CK(cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost));
    printf("HBM read (bulk 32KBx6 stages, 1 CTA/SM): %.1f GB/s  checksum %s\n", bytes / time_ms(e0, e1) * 1e-6,
           h == (double)(bytes / 8) ? "OK" : "BAD");
  }
  // ---- L2-resident read (48 MB window re-read) ----
  {
    size_t wbytes = (size_t)48 << 20;
    int reps = 40;
    k_read<<<nsm * 4, 512>>>(reinterpret_cast<const double2*>(buf), wbytes / 16, dout);
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) k_read<<<nsm * 4, 512>>>(reinterpret_cast<const double2*>(buf), wbytes / 16, dout);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    printf("L2 read (48 MiB window x%d): %.1f GB/s\n", reps, (double)wbytes * reps / time_ms(e0, e1) * 1e-6);
  }
  CK(cudaFree(buf));

  // ---- cuBLAS DGEMM peak ----
  {
    cublasHandle_t h; cublasCreate(&h);
    for (int n : {4096, 8192}) {
      double *A, *B, *C; size_t sz = (size_t)n * n * 8;
      CK(cudaMalloc(&A, sz)); CK(cudaMalloc(&B, sz)); CK(cudaMalloc(&C, sz));
      k_fill<<<nsm * 8, 256>>>(A, (size_t)n * n, 0.001); k_fill<<<nsm * 8, 256>>>(B, (size_t)n * n, 0.002);
      double one = 1, zero = 0;
      cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
      float best = 1e30f;
      for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        best = fminf(best, time_ms(e0, e1));
      }
      printf("cuBLAS DGEMM n=%d: %.3f ms  %.2f TFLOP/s (best of 5)\n", n, best, 2.0 * n * n * n / best * 1e-9);
      // sustained: 3 s loop
      if (n == 8192) {
        int loops = (int)(3000.0f / best) + 1;
        CK(cudaEventRecord(e0));
        for (int r = 0; r < loops; ++r) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        printf("cuBLAS DGEMM n=%d sustained (%d loops): %.2f TFLOP/s\n", n, loops, 2.0 * n * n * n * loops / time_ms(e0, e1) * 1e-9);
      }
      // skinny: M=n*?; emulate X=B_P*C: (n x n) x (n x 16)
      CK(cudaFree(A)); CK(cudaFree(B)); CK(cudaFree(C));
    }
    cublasDestroy(h);
  }
  // ---- cuSOLVER dsyevd at the config sizes ----
  {
    cusolverDnHandle_t sh; cusolverDnCreate(&sh);
    for (int n : {174, 688, 1376}) {
      double *A, *W; int* info; size_t sz = (size_t)n * n * 8;
      CK(cudaMalloc(&A, sz)); CK(cudaMalloc(&W, n * 8)); CK(cudaMalloc(&info, 4));
      std::vector<double> hA((size_t)n * n);
      srand(1);
      for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { double v = (rand() / (double)RAND_MAX - 0.5) * 0.1; if (i == j) v += i; hA[(size_t)i * n + j] = v; hA[(size_t)j * n + i] = v; }
      int lwork = 0; cusolverDnDsyevd_bufferSize(sh, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, W, &lwork);
      double* work; CK(cudaMalloc(&work, (size_t)lwork * 8));
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemcpy(A, hA.data(), sz, cudaMemcpyHostToDevice));
        CK(cudaEventRecord(e0));
        cusolverDnDsyevd(sh, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, W, work, lwork, info);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        best = fminf(best, time_ms(e0, e1));
      }
      printf("cuSOLVER dsyevd n=%d: %.3f ms (best of 3)\n", n, best);
      CK(cudaFree(A)); CK(cudaFree(W)); CK(cudaFree(info)); CK(cudaFree(work));
    }
    cusolverDnDestroy(sh);
  }
  printf("done\n");
  return 0;
}
