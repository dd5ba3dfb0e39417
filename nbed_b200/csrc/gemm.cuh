// FP64 tensor-core GEMM used by every GEMM-shaped stage of the hot path (SURVEY.md §2.3 K2 Gram, K4, K6,
// K7, K8):  C[b] = alpha * A[b] * B[b] + beta * C[b]   with arbitrary element strides, a two-level K
// addressing (for the [P][i][mu] half-transformed tensor) and an optional lower-triangle-only tile mask.
// Tiles are staged with cp.async into padded shared memory (conflict-free DMMA fragment loads) through a
// 4-stage ring; accumulators live in registers (tcgen05/TMEM have no f64 kind).
#pragma once
#include "common.cuh"

namespace nbd {

struct GemmArgs {
  int M, N, K;
  const double* A;
  long a_is;        // stride of the row index i of A(i,k)
  long a_ks;        // stride of k inside one K group
  int a_kb;         // K group length (k -> (k / a_kb) * a_kos + (k % a_kb) * a_ks); <=0 : single level
  long a_kos;       // stride between K groups
  const double* B;
  long b_js;        // stride of the column index j of B(k,j)
  long b_ks;
  int b_kb;
  long b_kos;
  double* C;
  long ldc;
  double alpha, beta;
  int batch;
  long strideA, strideB, strideC;
  int lower_only;   // 1: compute only tiles with tile_col <= tile_row (square tiles)
};

__device__ __forceinline__ long koff(int k, long ks, int kb, long kos) {
  if (kb <= 0) return (long)k * ks;
  return (long)(k / kb) * kos + (long)(k % kb) * ks;
}

constexpr int GEMM_BK = 16;
constexpr int GEMM_STAGES = 4;

// A_MC: A tile stored [k][m] (m contiguous in global memory), else [m][k].  B_NC: B tile stored [k][n].
template <int BM, int BN, int WM, int WN, bool A_MC, bool B_NC>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
gemm_dmma_kernel(GemmArgs g) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int BK = GEMM_BK;
  constexpr int A_LD = A_MC ? (BM + 4) : (BK + 4);
  constexpr int B_LD = B_NC ? (BN + 4) : (BK + 4);
  constexpr int A_ELEMS = A_MC ? BK * A_LD : BM * A_LD;
  constexpr int B_ELEMS = B_NC ? BK * B_LD : BN * B_LD;
  constexpr int MI = WM / 8, NI = WN / 8;
  extern __shared__ __align__(16) double gsm[];
  double* As = gsm;
  double* Bs = gsm + GEMM_STAGES * A_ELEMS;

  const int bm = blockIdx.y, bn = blockIdx.x, bz = blockIdx.z;
  if (g.lower_only && bn > bm) return;
  const double* A = g.A + (long)bz * g.strideA;
  const double* B = g.B + (long)bz * g.strideB;
  double* C = g.C + (long)bz * g.strideC;
  const int m0 = bm * BM, n0 = bn * BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int wm = (warp / (BN / WN)) * WM, wn = (warp % (BN / WN)) * WN;

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nk = (g.K + BK - 1) / BK;

  auto load_stage = [&](int kt, int st) {
    const int k0 = kt * BK;
    double* as = As + st * A_ELEMS;
    double* bs = Bs + st * B_ELEMS;
    if (A_MC) {
      for (int e = tid; e < BM * BK; e += NT) {
        const int m = e % BM, k = e / BM;
        const bool ok = (m0 + m < g.M) && (k0 + k < g.K);
        const double* src = ok ? A + (long)(m0 + m) * g.a_is + koff(k0 + k, g.a_ks, g.a_kb, g.a_kos) : A;
        cp_async8(as + k * A_LD + m, src, ok);
      }
    } else {
      for (int e = tid; e < BM * BK; e += NT) {
        const int k = e % BK, m = e / BK;
        const bool ok = (m0 + m < g.M) && (k0 + k < g.K);
        const double* src = ok ? A + (long)(m0 + m) * g.a_is + koff(k0 + k, g.a_ks, g.a_kb, g.a_kos) : A;
        cp_async8(as + m * A_LD + k, src, ok);
      }
    }
    if (B_NC) {
      for (int e = tid; e < BN * BK; e += NT) {
        const int n = e % BN, k = e / BN;
        const bool ok = (n0 + n < g.N) && (k0 + k < g.K);
        const double* src = ok ? B + (long)(n0 + n) * g.b_js + koff(k0 + k, g.b_ks, g.b_kb, g.b_kos) : B;
        cp_async8(bs + k * B_LD + n, src, ok);
      }
    } else {
      for (int e = tid; e < BN * BK; e += NT) {
        const int k = e % BK, n = e / BK;
        const bool ok = (n0 + n < g.N) && (k0 + k < g.K);
        const double* src = ok ? B + (long)(n0 + n) * g.b_js + koff(k0 + k, g.b_ks, g.b_kb, g.b_kos) : B;
        cp_async8(bs + n * B_LD + k, src, ok);
      }
    }
  };

#pragma unroll
  for (int s = 0; s < GEMM_STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<GEMM_STAGES - 2>();
    __syncthreads();
    {
      const int nx = kt + GEMM_STAGES - 1;
      if (nx < nk) load_stage(nx, nx % GEMM_STAGES);
      cp_async_commit();
    }
    const double* as = As + (kt % GEMM_STAGES) * A_ELEMS;
    const double* bs = Bs + (kt % GEMM_STAGES) * B_ELEMS;
#pragma unroll
    for (int k4 = 0; k4 < BK; k4 += 4) {
      double a[MI], b[NI];
#pragma unroll
      for (int i = 0; i < MI; ++i)
        a[i] = A_MC ? as[(k4 + tq) * A_LD + wm + 8 * i + gq] : as[(wm + 8 * i + gq) * A_LD + k4 + tq];
#pragma unroll
      for (int j = 0; j < NI; ++j)
        b[j] = B_NC ? bs[(k4 + tq) * B_LD + wn + 8 * j + gq] : bs[(wn + 8 * j + gq) * B_LD + k4 + tq];
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma(acc[i][j], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int row = m0 + wm + 8 * i + gq;
    if (row >= g.M) continue;
#pragma unroll
    for (int j = 0; j < NI; ++j) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int col = n0 + wn + 8 * j + 2 * tq + r;
        if (col < g.N) {
          double* p = C + (long)row * g.ldc + col;
          double v = g.alpha * acc[i][j][r];
          if (g.beta != 0.0) v += g.beta * (*p);
          *p = v;
        }
      }
    }
  }
}

// Obviously-correct CUDA-core variant (option "gemm_variant" = 1): debugging aid and cross-check.
__global__ void gemm_simple_kernel(GemmArgs g) {
  const int bz = blockIdx.z;
  const double* A = g.A + (long)bz * g.strideA;
  const double* B = g.B + (long)bz * g.strideB;
  double* C = g.C + (long)bz * g.strideC;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y * blockDim.y + threadIdx.y;
  if (row >= g.M || col >= g.N) return;
  if (g.lower_only && (col / 128) > (row / 128)) return;
  double s = 0.0;
  for (int k = 0; k < g.K; ++k)
    s += A[(long)row * g.a_is + koff(k, g.a_ks, g.a_kb, g.a_kos)] * B[(long)col * g.b_js + koff(k, g.b_ks, g.b_kb, g.b_kos)];
  double* p = C + (long)row * g.ldc + col;
  double v = g.alpha * s;
  if (g.beta != 0.0) v += g.beta * (*p);
  *p = v;
}

template <int BM, int BN, int WM, int WN, bool A_MC, bool B_NC>
inline cudaError_t launch_gemm_cfg(cudaStream_t st, const GemmArgs& g) {
  constexpr int BK = GEMM_BK;
  constexpr int A_LD = A_MC ? (BM + 4) : (BK + 4);
  constexpr int B_LD = B_NC ? (BN + 4) : (BK + 4);
  constexpr int A_ELEMS = A_MC ? BK * A_LD : BM * A_LD;
  constexpr int B_ELEMS = B_NC ? BK * B_LD : BN * B_LD;
  constexpr int SMEM = GEMM_STAGES * (A_ELEMS + B_ELEMS) * 8;
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  auto kern = gemm_dmma_kernel<BM, BN, WM, WN, A_MC, B_NC>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.batch > 0 ? g.batch : 1);
  kern<<<grid, NT, SMEM, st>>>(g);
  return cudaGetLastError();
}

// Dispatch on tile size (large problems: 128x128 tiles; small: 32x32) and on operand contiguity.
inline cudaError_t launch_gemm(cudaStream_t st, GemmArgs g, int variant, long* launches) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  if (g.batch <= 0) g.batch = 1;
  if (launches) ++*launches;
  if (variant == 1 || g.K <= 0) {
    dim3 b(32, 8), grid((g.N + 31) / 32, (g.M + 7) / 8, g.batch);
    gemm_simple_kernel<<<grid, b, 0, st>>>(g);
    return cudaGetLastError();
  }
  const bool a_mc = (g.a_is == 1);
  const bool b_nc = (g.b_js == 1);
  // lower_only masks whole tiles above the diagonal; every element with col <= row is still produced for
  // either tile size, which is all the mirror kernel needs.
  const bool small = (g.M <= 256 && g.N <= 256);
#define NBD_GEMM_DISPATCH(BM, BN, WM, WN)                                          \
  do {                                                                             \
    if (a_mc && b_nc) return launch_gemm_cfg<BM, BN, WM, WN, true, true>(st, g);   \
    if (a_mc && !b_nc) return launch_gemm_cfg<BM, BN, WM, WN, true, false>(st, g); \
    if (!a_mc && b_nc) return launch_gemm_cfg<BM, BN, WM, WN, false, true>(st, g); \
    return launch_gemm_cfg<BM, BN, WM, WN, false, false>(st, g);                   \
  } while (0)
  if (small) NBD_GEMM_DISPATCH(32, 32, 16, 16);
  NBD_GEMM_DISPATCH(128, 128, 64, 32);
#undef NBD_GEMM_DISPATCH
}

}  // namespace nbd
