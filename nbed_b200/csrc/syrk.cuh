// Stream-K symmetric rank-k update for the exchange Gram of the density-fitted K build
// (pyscf/df/df_jk.py get_jk: vk += buf1^T buf1, as reached from nbed/scf/huzinaga_scf.py:156):
//
//   C[b] (lower tiles) = alpha * X[b]^T X[b] + beta * C[b],   X[b] = [K][ld] with the n output indices contiguous.
//
// Why not the generic GEMM of gemm.cuh: with 64 x 64 tiles every CTA re-reads two 64 x K operand panels (10.7 GB of
// L2 -> SM traffic at the C4 shape, competing with the HBM-bound pass 2 that runs next to it), and 128 x 128 tiles
// give only 66 lower tiles per spin - 132 CTAs that each walk the whole contraction index, a bad quantisation on
// 148 SMs.  Here the (tile, k-step) space of all lower tiles of all batch entries is cut into one contiguous range
// per CTA (stream-K): every SM gets the same number of DMMA steps, a tile is finished by at most
// ceil(nk / per) + 1 CTAs, partial tiles go to a workspace and are summed in CTA order (deterministic) by
// syrk_fixup_kernel.  Diagonal tiles stage their single operand panel once.  L2 -> SM traffic: 5.1 GB.
//
// FP64 tensor path as everywhere else: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), operands staged with 16-byte
// cp.async.cg (L2 only: X is rewritten every SCF cycle and the kernel co-runs with pass 2, see jk.cuh).
#pragma once
#include "common.cuh"

namespace nbd {

constexpr int SY_T = 128;  // square output tile
constexpr int SY_BK = 16;
constexpr int SY_STAGES = 3;
constexpr int SY_LD = SY_T + 4;  // padded panel row: conflict-free fragment loads
constexpr int SY_STAGE_ELEMS = SY_BK * SY_LD;
constexpr int SY_THREADS = 256;  // 2 x 4 warps, warp tile 64 x 32
constexpr int SY_SMEM_BYTES = 2 * SY_STAGES * SY_STAGE_ELEMS * 8;

struct SyrkArgs {
  const double* X;
  long ld, strideX;
  double* C;
  long ldc, strideC;
  double* part;  // [grid][maxseg][SY_T * SY_T]
  double alpha, beta;
  int n, K, batch;
  int nt;      // tiles per side
  int ntile;   // lower tiles per batch entry
  int nk;      // k-steps per tile
  int maxseg;  // most tiles one CTA's range can touch
  long total;  // batch * ntile * nk
  long per;    // k-steps per range
  int nranges;            // contiguous (tile, k-step) ranges, handed out through *counter
  unsigned int* counter;  // zeroed before the launch
  int sm_p2, nsm;         // spatial split: sm_p2 of the nsm SMs belong to pass 2 (0 = this kernel runs everywhere)
};

// t = bm (bm + 1) / 2 + bn with 0 <= bn <= bm
__device__ __forceinline__ void syrk_tile_coords(int t, int& bm, int& bn) {
  int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((r + 1) * (r + 2) / 2 <= t) ++r;
  while (r * (r + 1) / 2 > t) --r;
  bm = r;
  bn = t - r * (r + 1) / 2;
}

__global__ void __maxnreg__(192) syrk_streamk_kernel(SyrkArgs g) {
  extern __shared__ __align__(16) double sy_smem[];
  double* As = sy_smem;
  double* Bs = sy_smem + SY_STAGES * SY_STAGE_ELEMS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
  constexpr int MI = 8, NI = 4;
  int a_off[MI], b_off[NI];
#pragma unroll
  for (int i = 0; i < MI; ++i) a_off[i] = tq * SY_LD + wm + 8 * i + gq;
#pragma unroll
  for (int j = 0; j < NI; ++j) b_off[j] = tq * SY_LD + wn + 8 * j + gq;
  constexpr int K4 = 4 * SY_LD;

  // CTAs that land on an SM of the pass-2 group retire at once; the others draw ranges until the queue is dry.  Which
  // CTA computes a range does not matter: partial tiles are stored and summed by range index.
  if (g.sm_p2 > 0 && sm_in_group1(smid(), g.sm_p2, g.nsm)) return;
  __shared__ int s_range;
  for (;;) {
  if (tid == 0) s_range = (int)atomicAdd(g.counter, 1u);
  __syncthreads();
  const int range = s_range;
  __syncthreads();
  if (range >= g.nranges) break;
  long u = (long)range * g.per;
  const long u1 = u + g.per < g.total ? u + g.per : g.total;
  int seg = 0;
  while (u < u1) {
    const long tl = u / g.nk;  // tile index over all batch entries
    const int k0 = (int)(u - tl * g.nk);
    const int k1 = (int)((long)g.nk < k0 + (u1 - u) ? (long)g.nk : k0 + (u1 - u));
    const int bz = (int)(tl / g.ntile), t = (int)(tl - (long)bz * g.ntile);
    int bm, bn;
    syrk_tile_coords(t, bm, bn);
    const bool diag = bm == bn;
    const double* X = g.X + (long)bz * g.strideX;
    const int m0 = bm * SY_T, n0 = bn * SY_T;

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_stage = [&](int kt, int st) {
      const int kb = kt * SY_BK;
      double* as = As + st * SY_STAGE_ELEMS;
      double* bs = Bs + st * SY_STAGE_ELEMS;
#pragma unroll
      for (int j = 0; j < SY_T * SY_BK / 2 / SY_THREADS; ++j) {
        const int e = tid + j * SY_THREADS, m = 2 * (e % (SY_T / 2)), k = e / (SY_T / 2);
        const bool kin = kb + k < g.K;
        const double* row = X + (long)(kb + k) * g.ld;
        const bool oka = kin && (m0 + m < g.n);  // n even: a pair never straddles the edge
        cp_async16(as + k * SY_LD + m, oka ? row + m0 + m : X, oka);
        if (!diag) {
          const bool okb = kin && (n0 + m < g.n);
          cp_async16(bs + k * SY_LD + m, okb ? row + n0 + m : X, okb);
        }
      }
    };

#pragma unroll
    for (int s = 0; s < SY_STAGES - 1; ++s) {
      if (k0 + s < k1) load_stage(k0 + s, s);
      cp_async_commit();
    }
    for (int kt = k0; kt < k1; ++kt) {
      cp_async_wait<SY_STAGES - 2>();
      __syncthreads();
      {
        const int nx = kt + SY_STAGES - 1;
        if (nx < k1) load_stage(nx, (nx - k0) % SY_STAGES);
        cp_async_commit();
      }
      const double* as = As + ((kt - k0) % SY_STAGES) * SY_STAGE_ELEMS;
      const double* bs = diag ? as : Bs + ((kt - k0) % SY_STAGES) * SY_STAGE_ELEMS;
      double a[2][MI], b[2][NI];
#pragma unroll
      for (int i = 0; i < MI; ++i) a[0][i] = as[a_off[i]];
#pragma unroll
      for (int j = 0; j < NI; ++j) b[0][j] = bs[b_off[j]];
#pragma unroll
      for (int s4 = 0; s4 < SY_BK / 4; ++s4) {
        const int cur = s4 & 1, nxt = cur ^ 1;
        if (s4 + 1 < SY_BK / 4) {
#pragma unroll
          for (int i = 0; i < MI; ++i) a[nxt][i] = as[a_off[i] + (s4 + 1) * K4];
#pragma unroll
          for (int j = 0; j < NI; ++j) b[nxt][j] = bs[b_off[j] + (s4 + 1) * K4];
        }
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NI; ++j) dmma(acc[i][j], a[cur][i], b[cur][j]);
      }
    }
    cp_async_wait<0>();
    __syncthreads();  // every warp is done with the ring before the next segment refills it

    if (k0 == 0 && k1 == g.nk) {  // the whole contraction of this tile: write the result
      double* C = g.C + (long)bz * g.strideC;
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int row = m0 + wm + 8 * i + gq;
        if (row >= g.n) continue;
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int col = n0 + wn + 8 * j + 2 * tq + r;
            if (col < g.n) {
              double* p = C + (long)row * g.ldc + col;
              double v = g.alpha * acc[i][j][r];
              if (g.beta != 0.0) v += g.beta * (*p);
              *p = v;
            }
          }
      }
    } else {
      double* P = g.part + ((long)range * g.maxseg + seg) * (SY_T * SY_T);
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          double2 v;
          v.x = acc[i][j][0];
          v.y = acc[i][j][1];
          *reinterpret_cast<double2*>(P + (wm + 8 * i + gq) * SY_T + wn + 8 * j + 2 * tq) = v;
        }
    }
    ++seg;
    u += k1 - k0;
  }
  }
}

// One CTA per tile: sums the partial tiles of the CTAs whose ranges cut this tile, in CTA order.
__global__ void __launch_bounds__(256) syrk_fixup_kernel(SyrkArgs g) {
  const long tl = blockIdx.x;
  const long ua = tl * g.nk, ub = ua + g.nk;
  const int i0 = (int)(ua / g.per), i1 = (int)((ub - 1) / g.per);
  if (i0 == i1) return;  // one CTA held the whole contraction and wrote the result itself
  const int bz = (int)(tl / g.ntile), t = (int)(tl - (long)bz * g.ntile);
  int bm, bn;
  syrk_tile_coords(t, bm, bn);
  double* C = g.C + (long)bz * g.strideC;
  const int m0 = bm * SY_T, n0 = bn * SY_T;
  for (int e = threadIdx.x; e < SY_T * SY_T; e += 256) {
    const int row = m0 + e / SY_T, col = n0 + e % SY_T;
    if (row >= g.n || col > row) continue;  // lower triangle only (the caller mirrors it)
    double s = 0.0;
    for (int i = i0; i <= i1; ++i) {
      const long tfirst = ((long)i * g.per) / g.nk;
      s += __ldcg(g.part + ((long)i * g.maxseg + (tl - tfirst)) * (SY_T * SY_T) + e);
    }
    double* p = C + (long)row * g.ldc + col;
    double v = g.alpha * s;
    if (g.beta != 0.0) v += g.beta * (*p);
    *p = v;
  }
}

inline bool syrk_applicable(const double* X, long ld, long strideX, int n, int K, int batch) {
  return n >= 512 && (n & 1) == 0 && (ld & 1) == 0 && (strideX & 1) == 0 && K >= SY_BK && batch >= 1 &&
         (reinterpret_cast<uintptr_t>(X) & 15) == 0;
}

// Fills the derived fields of `g` for `ranges` work ranges; returns the workspace size in doubles.
inline size_t syrk_plan(SyrkArgs& g, int ranges) {
  g.nt = (g.n + SY_T - 1) / SY_T;
  g.ntile = g.nt * (g.nt + 1) / 2;
  g.nk = (g.K + SY_BK - 1) / SY_BK;
  g.total = (long)g.batch * g.ntile * g.nk;
  g.per = (g.total + ranges - 1) / ranges;
  g.nranges = (int)((g.total + g.per - 1) / g.per);
  g.maxseg = (int)((g.per + g.nk - 1) / g.nk) + 1;
  return (size_t)ranges * g.maxseg * SY_T * SY_T;
}

// `ctas`: CTAs to launch (one per SM; with a spatial split the ones on pass-2 SMs retire at once).
inline cudaError_t launch_syrk(cudaStream_t st, const SyrkArgs& g, int ctas, long* launches) {
  static unsigned long long configured = 0;
  if (first_use_on_current_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(syrk_streamk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SY_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    // the kernel shares its SMs with pass 2 (two 57 KiB CTAs): ask for the largest shared-memory carveout, otherwise an
    // SM configured for this kernel alone cannot take the pass-2 CTAs until the (persistent) Gram CTA has retired
    e = cudaFuncSetAttribute(syrk_streamk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaMemsetAsync(g.counter, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return e;
  syrk_streamk_kernel<<<ctas, SY_THREADS, SY_SMEM_BYTES, st>>>(g);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  syrk_fixup_kernel<<<(unsigned)((long)g.batch * g.ntile), 256, 0, st>>>(g);
  if (launches) *launches += 2;
  return cudaGetLastError();
}

}  // namespace nbd
