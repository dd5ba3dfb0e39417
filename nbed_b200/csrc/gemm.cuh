// FP64 tensor-core GEMM used by every GEMM-shaped stage of the hot path (SURVEY.md §2.3 K2 Gram, K4, K6,
// K7, K8):  C[b] = alpha * A[b] * B[b] + beta * C[b]   with arbitrary element strides (so that transposes
// and the [P][i][mu] half-transformed tensor need no copies) and an optional lower-triangle-only tile mask.
//
// tcgen05 / TMEM have no f64 kind, so the accumulators live in registers and the MMA is the warp-level
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  Operand tiles are staged with cp.async (LDGSTS) through a 4-stage
// ring of padded shared-memory tiles (conflict-free fragment loads); fragments are double-buffered in
// registers so that the shared-memory latency of step k+1 hides behind the DMMAs of step k.
#pragma once
#include "common.cuh"

namespace nbd {

struct GemmArgs {
  int M, N, K;
  const double* A;
  long a_is;  // stride of the row index i of A(i,k)
  long a_ks;  // stride of k
  const double* B;
  long b_js;  // stride of the column index j of B(k,j)
  long b_ks;
  double* C;
  long ldc;
  double alpha, beta;
  int batch;
  long strideA, strideB, strideC;
  int lower_only;  // 1: compute only tiles with tile_col <= tile_row (square tiles)
  int pdl;         // host only: launch with programmatic stream serialization (may start while the preceding kernel
                   // of the stream still runs, once that kernel has signalled launch_dependents; no data dependency)
};

constexpr int GEMM_BK = 16;
constexpr int GEMM_STAGES = 3;

// A_MC: A tile stored [k][m] (m contiguous in global memory), else [m][k].  B_NC: B tile stored [k][n].
// VEC (only with A_MC && B_NC): operands are 16-byte aligned with even extents, tiles are staged with 16-byte
// cp.async.cg (L2 only - the K Gram co-runs with pass 2, see the caching note in jk.cuh - and half the copy count).
template <int BM, int BN, int WM, int WN, bool A_MC, bool B_NC, bool VEC = false>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
gemm_dmma_kernel(GemmArgs g) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int BK = GEMM_BK;
  constexpr int A_LD = A_MC ? (BM + 4) : (BK + 4);
  constexpr int B_LD = B_NC ? (BN + 4) : (BK + 4);
  constexpr int A_ELEMS = A_MC ? BK * A_LD : BM * A_LD;
  constexpr int B_ELEMS = B_NC ? BK * B_LD : BN * B_LD;
  constexpr int MI = WM / 8, NI = WN / 8;
  constexpr int A_PER_T = BM * BK / NT, B_PER_T = BN * BK / NT;
  static_assert(BM * BK % NT == 0 && BN * BK % NT == 0, "tile must divide evenly over the threads");
  static_assert(NT % BM == 0 || BM % NT == 0, "A_MC mapping");
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

  // ---- per-thread copy plan: element j of this thread is (i_fix + j * i_step, k_fix + j * k_step) ----------
  // A_MC : e = tid + j NT ; m = e % BM , k = e / BM       !A_MC : k = e % BK , m = e / BK
  int a_m, a_k, a_mstep, a_kstep, b_n, b_k, b_nstep, b_kstep;
  if (A_MC) {
    a_m = tid % BM; a_k = tid / BM; a_mstep = (NT % BM == 0) ? 0 : NT; a_kstep = (NT % BM == 0) ? NT / BM : 0;
  } else {
    a_k = tid % BK; a_m = tid / BK; a_mstep = NT / BK; a_kstep = 0;
  }
  if (B_NC) {
    b_n = tid % BN; b_k = tid / BN; b_nstep = (NT % BN == 0) ? 0 : NT; b_kstep = (NT % BN == 0) ? NT / BN : 0;
  } else {
    b_k = tid % BK; b_n = tid / BK; b_nstep = NT / BK; b_kstep = 0;
  }
  const double* a_src = A + (long)(m0 + a_m) * g.a_is + (long)a_k * g.a_ks;
  const double* b_src = B + (long)(n0 + b_n) * g.b_js + (long)b_k * g.b_ks;
  const long a_jstep = (long)a_mstep * g.a_is + (long)a_kstep * g.a_ks;
  const long b_jstep = (long)b_nstep * g.b_js + (long)b_kstep * g.b_ks;
  const int a_dst0 = A_MC ? a_k * A_LD + a_m : a_m * A_LD + a_k;
  const int a_dstep = A_MC ? a_kstep * A_LD + a_mstep : a_mstep * A_LD;
  const int b_dst0 = B_NC ? b_k * B_LD + b_n : b_n * B_LD + b_k;
  const int b_dstep = B_NC ? b_kstep * B_LD + b_nstep : b_nstep * B_LD;

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
    if (VEC) {
      // pairs along the contiguous index: e = tid + j NT ; m = 2 (e % (BM/2)) , k = e / (BM/2)
      constexpr int AV = BM * BK / 2 / NT, BV = BN * BK / 2 / NT;
#pragma unroll
      for (int j = 0; j < AV; ++j) {
        const int e = tid + j * NT, m = 2 * (e % (BM / 2)), k = e / (BM / 2);
        const bool ok = (m0 + m < g.M) && (k0 + k < g.K);  // M even: a pair never straddles the edge
        cp_async16(as + k * A_LD + m, ok ? A + (long)(m0 + m) + (long)(k0 + k) * g.a_ks : A, ok);
      }
#pragma unroll
      for (int j = 0; j < BV; ++j) {
        const int e = tid + j * NT, n = 2 * (e % (BN / 2)), k = e / (BN / 2);
        const bool ok = (n0 + n < g.N) && (k0 + k < g.K);
        cp_async16(bs + k * B_LD + n, ok ? B + (long)(n0 + n) + (long)(k0 + k) * g.b_ks : B, ok);
      }
      return;
    }
    const double* ap = a_src + (long)k0 * g.a_ks;
    const double* bp = b_src + (long)k0 * g.b_ks;
#pragma unroll
    for (int j = 0; j < A_PER_T; ++j) {
      const bool ok = (m0 + a_m + j * a_mstep < g.M) && (k0 + a_k + j * a_kstep < g.K);
      cp_async8(as + a_dst0 + j * a_dstep, ok ? ap + j * a_jstep : A, ok);
    }
#pragma unroll
    for (int j = 0; j < B_PER_T; ++j) {
      const bool ok = (n0 + b_n + j * b_nstep < g.N) && (k0 + b_k + j * b_kstep < g.K);
      cp_async8(bs + b_dst0 + j * b_dstep, ok ? bp + j * b_jstep : B, ok);
    }
  };

#pragma unroll
  for (int s = 0; s < GEMM_STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }

  // fragment offsets inside a stage (k4 added at use)
  int a_off[MI], b_off[NI];
#pragma unroll
  for (int i = 0; i < MI; ++i) a_off[i] = A_MC ? tq * A_LD + wm + 8 * i + gq : (wm + 8 * i + gq) * A_LD + tq;
#pragma unroll
  for (int j = 0; j < NI; ++j) b_off[j] = B_NC ? tq * B_LD + wn + 8 * j + gq : (wn + 8 * j + gq) * B_LD + tq;
  constexpr int A_K4 = A_MC ? 4 * A_LD : 4;
  constexpr int B_K4 = B_NC ? 4 * B_LD : 4;

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
    double a[2][MI], b[2][NI];
#pragma unroll
    for (int i = 0; i < MI; ++i) a[0][i] = as[a_off[i]];
#pragma unroll
    for (int j = 0; j < NI; ++j) b[0][j] = bs[b_off[j]];
#pragma unroll
    for (int s4 = 0; s4 < BK / 4; ++s4) {
      const int cur = s4 & 1, nxt = cur ^ 1;
      if (s4 + 1 < BK / 4) {
#pragma unroll
        for (int i = 0; i < MI; ++i) a[nxt][i] = as[a_off[i] + (s4 + 1) * A_K4];
#pragma unroll
        for (int j = 0; j < NI; ++j) b[nxt][j] = bs[b_off[j] + (s4 + 1) * B_K4];
      }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma(acc[i][j], a[cur][i], b[cur][j]);
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
__global__ void gemm_simple_kernel(GemmArgs g, int tile) {
  const int bz = blockIdx.z;
  const double* A = g.A + (long)bz * g.strideA;
  const double* B = g.B + (long)bz * g.strideB;
  double* C = g.C + (long)bz * g.strideC;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y * blockDim.y + threadIdx.y;
  if (row >= g.M || col >= g.N) return;
  if (g.lower_only && (col / tile) > (row / tile)) return;
  double s = 0.0;
  for (int k = 0; k < g.K; ++k) s += A[(long)row * g.a_is + (long)k * g.a_ks] * B[(long)col * g.b_js + (long)k * g.b_ks];
  double* p = C + (long)row * g.ldc + col;
  double v = g.alpha * s;
  if (g.beta != 0.0) v += g.beta * (*p);
  *p = v;
}

template <int BM, int BN, int WM, int WN, bool A_MC, bool B_NC, bool VEC = false>
inline cudaError_t launch_gemm_cfg(cudaStream_t st, const GemmArgs& g) {
  constexpr int BK = GEMM_BK;
  constexpr int A_LD = A_MC ? (BM + 4) : (BK + 4);
  constexpr int B_LD = B_NC ? (BN + 4) : (BK + 4);
  constexpr int A_ELEMS = A_MC ? BK * A_LD : BM * A_LD;
  constexpr int B_ELEMS = B_NC ? BK * B_LD : BN * B_LD;
  constexpr int SMEM = GEMM_STAGES * (A_ELEMS + B_ELEMS) * 8;
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  auto kern = gemm_dmma_kernel<BM, BN, WM, WN, A_MC, B_NC, VEC>;
  static unsigned long long configured = 0;
  if (first_use_on_current_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return e;
    // GEMMs co-run with other grids (the K Gram next to pass 2): keep the SM at its largest shared-memory carveout
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.batch > 0 ? g.batch : 1);
  if (g.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, g);
  }
  kern<<<grid, NT, SMEM, st>>>(g);
  return cudaGetLastError();
}

// Tile size by problem size: the largest of 128 / 64 / 32 square tiles that still yields about one CTA per SM
// (148 SMs); operand contiguity picks the shared-memory tile orientation.
inline cudaError_t launch_gemm(cudaStream_t st, GemmArgs g, int variant, long* launches, int sm_count = 148,
                               int force_tile = 0) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  if (g.batch <= 0) g.batch = 1;
  if (launches) ++*launches;
  auto ctas = [&](int t) {
    const long tm = (g.M + t - 1) / t, tn = (g.N + t - 1) / t;
    const long per = g.lower_only ? tm * (tm + 1) / 2 : tm * tn;
    return per * g.batch;
  };
  // 128-tiles have the better DMMA : shared-load ratio but quantise badly and monopolise an SM: they are used only
  // when they fill at least three waves.  Measured at n = 1376: the n^3 stages are 17 % faster with 64-tiles, the
  // K Gram equally fast, and the HBM-bound pass 2 that runs next to it loses less.
  int tile = 128;
  if (ctas(128) < 3L * sm_count) tile = 64;
  if (tile == 64 && ctas(64) < (long)sm_count / 2 && g.M <= 256 && g.N <= 256) tile = 32;
  // a thin output (M or N <= 32: the back-transform C = V^T X of the tracked 16 / 32 orbitals, the mu path's gradient
  // blocks) on 64-tiles is 50-90 % padding on less than one wave of CTAs: 32-tiles halve the padding and double the CTAs
  if (tile == 64 && ctas(64) < (long)sm_count && (g.M <= 32 || g.N <= 32)) tile = 32;
  if (force_tile == 32 || force_tile == 64 || force_tile == 128) tile = force_tile;
  if (variant == 1 || g.K <= 0) {
    dim3 b(32, 8), grid((g.N + 31) / 32, (g.M + 7) / 8, g.batch);
    gemm_simple_kernel<<<grid, b, 0, st>>>(g, tile);
    return cudaGetLastError();
  }
  const bool a_mc = (g.a_is == 1);
  const bool b_nc = (g.b_js == 1);
  // lower_only masks whole tiles above the diagonal; every element with col <= row is still produced for
  // any tile size, which is all the mirror kernel needs.
  // 16-byte staging is possible when both operands are contiguous along the tile's leading index, 16-byte aligned,
  // with even extents and even strides everywhere
  auto even = [](long x) { return (x & 1) == 0; };
  const bool vec = a_mc && b_nc && even(g.M) && even(g.N) && even(g.a_ks) && even(g.b_ks) && even(g.strideA) &&
                   even(g.strideB) && ((reinterpret_cast<uintptr_t>(g.A) | reinterpret_cast<uintptr_t>(g.B)) & 15) == 0;
#define NBD_GEMM_DISPATCH(BM, BN, WM, WN)                                          \
  do {                                                                             \
    if (vec) return launch_gemm_cfg<BM, BN, WM, WN, true, true, true>(st, g);      \
    if (a_mc && b_nc) return launch_gemm_cfg<BM, BN, WM, WN, true, true>(st, g);   \
    if (a_mc && !b_nc) return launch_gemm_cfg<BM, BN, WM, WN, true, false>(st, g); \
    if (!a_mc && b_nc) return launch_gemm_cfg<BM, BN, WM, WN, false, true>(st, g); \
    return launch_gemm_cfg<BM, BN, WM, WN, false, false>(st, g);                   \
  } while (0)
  // Narrow outputs (the second ao2mo half-transform L[(P,p)][q] = X C: N = m active orbitals, M = naux * m rows): a
  // 64-column tile would spend 64 / N of the DMMAs on padding (37 % at m = 40), so the column extent of the tile is
  // N rounded up to 8 - one warp per 16 rows, NI = BN / 8 accumulator columns.
  if (force_tile == 0 && !a_mc && !b_nc && g.N > 16 && g.N <= 56 && g.M >= 64 * (long)sm_count) {
    const int bn8 = (g.N + 7) / 8 * 8;
    if (bn8 == 24) return launch_gemm_cfg<64, 24, 16, 24, false, false>(st, g);
    if (bn8 == 32) return launch_gemm_cfg<64, 32, 16, 32, false, false>(st, g);
    if (bn8 == 40) return launch_gemm_cfg<64, 40, 16, 40, false, false>(st, g);
    if (bn8 == 48) return launch_gemm_cfg<64, 48, 16, 48, false, false>(st, g);
    if (bn8 == 56) return launch_gemm_cfg<64, 56, 16, 56, false, false>(st, g);
  }
  if (tile == 32) NBD_GEMM_DISPATCH(32, 32, 16, 16);
  if (tile == 64) NBD_GEMM_DISPATCH(64, 64, 32, 32);
  NBD_GEMM_DISPATCH(128, 128, 64, 32);
#undef NBD_GEMM_DISPATCH
}

}  // namespace nbd
