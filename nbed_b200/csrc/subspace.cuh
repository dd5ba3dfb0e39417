// Chebyshev-filtered subspace iteration for the occupied block of the Lowdin-orthogonalised Fock matrix.
//
// The reference diagonalises the full n x n matrix every SCF cycle (nbed/scf/huzinaga_scf.py:166-169) although
// the loop only consumes the nelec lowest eigenvectors (get_occ / make_rdm1, :170-174); the full spectrum is
// needed once, for the values returned at the end.  On the GPU the library eigensolver (cuSOLVER dsyevd, 17.6 ms
// per 1376^2 matrix) is the Amdahl term of the iteration, so the cycles in between track the invariant subspace
// of the KB lowest eigenvectors instead: a scaled Chebyshev polynomial of F' damps everything above the block,
// followed by Rayleigh-Ritz, until the residuals of the occupied vectors are below 1e-10 (densities and energies
// then agree with a full diagonalisation far inside the 1e-8 parity tolerance).  The first block of an SCF starts from
// pseudo-random vectors (cold start, amplification-capped filter degree); the filter interval comes from a few
// Lanczos steps and is then widened by ||dF'||_F per cycle.  cuSOLVER still produces the final full (C, eps) and is
// the fallback whenever the filter does not converge or a Ritz value lands above the assumed upper bound.
//
// All kernels are small FP64 CUDA-core kernels over L2-resident data (F' is 15 MB at n = 1376).
#pragma once
#include "common.cuh"

namespace nbd {

constexpr int SUB_ROWS = 16;    // output rows per CTA
constexpr int SUB_JCHUNK = 64;  // contraction chunk per pipeline stage (16 per warp)
constexpr int SUB_STAGES = 4;

struct SubApplyArgs {
  const double* A;  // [batch][n][n] symmetric
  const double* Y;  // [batch][n][KB]
  const double* Z;  // [batch][n][KB] or null
  double* out;      // [batch][n][KB]
  int n;
  double alpha[2], shift[2], beta[2];  // per batch entry: out = alpha * (A Y - shift Y) - beta Z
};

template <int KB>
constexpr int sub_apply_smem_bytes() {
  return SUB_STAGES * (SUB_ROWS * (SUB_JCHUNK + 4) + SUB_JCHUNK * (KB + 4)) * 8;
}

// out = alpha (A Y - shift Y) - beta Z.  FP64 tensor-core (DMMA m8n8k4) skinny product over L2-resident data: a CTA
// owns 16 rows and the whole contraction index, streamed through a 4-stage cp.async ring in chunks of 64; its four
// warps take 16 contraction indices of every chunk each and their partial tiles are summed through shared memory in
// warp order (deterministic).  No inter-CTA reduction, so one step is one short launch.
template <int KB>
__global__ void __launch_bounds__(128) sub_apply_kernel(SubApplyArgs a) {
  constexpr int A_LD = SUB_JCHUNK + 4, Y_LD = KB + 4, NI = KB / 8;
  constexpr int A_ST = SUB_ROWS * A_LD, Y_ST = SUB_JCHUNK * Y_LD;
  extern __shared__ __align__(16) double sub_smem[];
  double* As = sub_smem;
  double* Ys = sub_smem + SUB_STAGES * A_ST;
  const int rb = blockIdx.x, b = blockIdx.z;
  const int n = a.n, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const double* A = a.A + (long)b * n * n;
  const double* Y = a.Y + (long)b * n * KB;
  const int nchunk = (n + SUB_JCHUNK - 1) / SUB_JCHUNK;
  const bool vec = (n & 1) == 0;  // rows of A start 16-byte aligned
  double acc[2][NI][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  auto load = [&](int ch, int buf) {
    const int j0 = ch * SUB_JCHUNK;
    double* as = As + buf * A_ST;
    double* ys = Ys + buf * Y_ST;
    if (vec) {
#pragma unroll
      for (int q = 0; q < SUB_ROWS * SUB_JCHUNK / 2 / 128; ++q) {
        const int e = tid + 128 * q;
        const int r = e / (SUB_JCHUNK / 2), j = 2 * (e % (SUB_JCHUNK / 2));
        const int gi = rb * SUB_ROWS + r, gj = j0 + j;
        const bool ok = gi < n && gj < n;  // n even: a pair never straddles the edge
        cp_async16(as + r * A_LD + j, ok ? A + (long)gi * n + gj : A, ok);
      }
    } else {
#pragma unroll
      for (int q = 0; q < SUB_ROWS * SUB_JCHUNK / 128; ++q) {
        const int e = tid + 128 * q;
        const int r = e / SUB_JCHUNK, j = e % SUB_JCHUNK;
        const int gi = rb * SUB_ROWS + r, gj = j0 + j;
        const bool ok = gi < n && gj < n;
        cp_async8(as + r * A_LD + j, ok ? A + (long)gi * n + gj : A, ok);
      }
    }
#pragma unroll
    for (int q = 0; q < SUB_JCHUNK * KB / 2 / 128; ++q) {
      const int e = tid + 128 * q;
      const int j = e / (KB / 2), c2 = 2 * (e % (KB / 2));
      const bool ok = j0 + j < n;
      cp_async16(ys + j * Y_LD + c2, ok ? Y + (long)(j0 + j) * KB + c2 : Y, ok);
    }
  };
#pragma unroll
  for (int s = 0; s < SUB_STAGES - 1; ++s) {
    if (s < nchunk) load(s, s);
    cp_async_commit();
  }
  for (int ch = 0; ch < nchunk; ++ch) {
    cp_async_wait<SUB_STAGES - 2>();
    __syncthreads();
    {
      const int nx = ch + SUB_STAGES - 1;
      if (nx < nchunk) load(nx, nx % SUB_STAGES);
      cp_async_commit();
    }
    const int buf = ch % SUB_STAGES;
    const double* as = As + buf * A_ST + gq * A_LD + 16 * warp + tq;
    const double* ys = Ys + buf * Y_ST + (16 * warp + tq) * Y_LD + gq;
#pragma unroll
    for (int k4 = 0; k4 < 16; k4 += 4) {
      double af[2], bf[NI];
      af[0] = as[k4];
      af[1] = as[8 * A_LD + k4];
#pragma unroll
      for (int j = 0; j < NI; ++j) bf[j] = ys[k4 * Y_LD + 8 * j];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma(acc[i][j], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  // sum the four warps' partial 16 x KB tiles through shared memory (fixed order), then the epilogue
  double* red = sub_smem;  // [4][16][KB]
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      double* p = red + ((warp * 16) + 8 * i + gq) * KB + 8 * j + 2 * tq;
      p[0] = acc[i][j][0];
      p[1] = acc[i][j][1];
    }
  __syncthreads();
  const double alpha = a.alpha[b], shift = a.shift[b], beta = a.beta[b];
  const double* Z = a.Z ? a.Z + (long)b * n * KB : nullptr;
  double* out = a.out + (long)b * n * KB;
  for (int e = tid; e < SUB_ROWS * KB; e += 128) {
    const int gi = rb * SUB_ROWS + e / KB;
    if (gi >= n) continue;
    const double s = ((red[e] + red[16 * KB + e]) + red[2 * 16 * KB + e]) + red[3 * 16 * KB + e];
    const long o = (long)gi * KB + e % KB;
    double v = alpha * (s - shift * Y[o]);
    if (Z) v -= beta * Z[o];
    out[o] = v;
  }
}

// ---- cluster split-K variant (default) -----------------------------------------------------------------------
// The kernel above runs 86 x 2 CTAs that each walk the whole contraction index: it is latency-bound (17 us per block
// product for 45 MB of L2 reads).  Here a CTA owns 32 rows (half the re-reads of Y) and a QUARTER of the contraction
// index; the four CTAs of a thread-block cluster hold the four partial 32 x KB tiles in their shared memory and the
// cluster's rank-0 CTA sums them over DSMEM in rank order (deterministic) and applies the epilogue.  4x the CTAs in
// flight, a quarter of the serial chunk chain per CTA, no extra launch and no global-memory round trip.
constexpr int SUB2_ROWS = 32;
constexpr int SUB2_JCHUNK = 32;
constexpr int SUB2_STAGES = 4;
constexpr int SUB2_KS = 4;  // cluster size = split of the contraction index

template <int KB>
constexpr int sub_apply2_smem_bytes() {
  // ring of stages, reused for the 4 per-warp partial tiles [4][32][KB]; the CTA's summed tile [32][KB] lives behind it
  constexpr int ring = SUB2_STAGES * (SUB2_ROWS * (SUB2_JCHUNK + 4) + SUB2_JCHUNK * (KB + 4)) * 8;
  constexpr int red = 4 * SUB2_ROWS * KB * 8;
  return (ring > red ? ring : red) + SUB2_ROWS * KB * 8;
}

// PDL = 1 (programmatic stream serialisation, see sub_apply3_kernel): the F' parts of the first ring stages are requested
// before griddepcontrol.wait, everything that touches Y / Z / out behind it.
template <int KB, int PDL = 0>
__global__ void __launch_bounds__(128) sub_apply2_kernel(SubApplyArgs a) {
  constexpr int A_LD = SUB2_JCHUNK + 4, Y_LD = KB + 4, NI = KB / 8, MI = SUB2_ROWS / 8;
  constexpr int A_ST = SUB2_ROWS * A_LD, Y_ST = SUB2_JCHUNK * Y_LD;
  constexpr int RING = SUB2_STAGES * (A_ST + Y_ST), RED = 4 * SUB2_ROWS * KB;
  extern __shared__ __align__(16) double sub_smem[];
  double* As = sub_smem;
  double* Ys = sub_smem + SUB2_STAGES * A_ST;
  double* cta_tile = sub_smem + (RING > RED ? RING : RED);  // [32][KB], read by the cluster's rank-0 CTA
  const int ks = (int)cluster_ctarank();  // which quarter of the contraction index
  const int rb = blockIdx.x, b = blockIdx.z;
  const int n = a.n, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const double* A = a.A + (long)b * n * n;
  const double* Y = a.Y + (long)b * n * KB;
  const int nchunk = (n + SUB2_JCHUNK - 1) / SUB2_JCHUNK;
  const int ch0 = nchunk * ks / SUB2_KS, ch1 = nchunk * (ks + 1) / SUB2_KS;
  const bool vec = (n & 1) == 0;  // rows of A start 16-byte aligned
  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  auto load_a = [&](int ch, int buf) {
    const int j0 = ch * SUB2_JCHUNK;
    double* as = As + buf * A_ST;
    if (vec) {
#pragma unroll
      for (int q = 0; q < SUB2_ROWS * SUB2_JCHUNK / 2 / 128; ++q) {
        const int e = tid + 128 * q;
        const int r = e / (SUB2_JCHUNK / 2), j = 2 * (e % (SUB2_JCHUNK / 2));
        const int gi = rb * SUB2_ROWS + r, gj = j0 + j;
        const bool ok = gi < n && gj < n;
        cp_async16(as + r * A_LD + j, ok ? A + (long)gi * n + gj : A, ok);
      }
    } else {
#pragma unroll
      for (int q = 0; q < SUB2_ROWS * SUB2_JCHUNK / 128; ++q) {
        const int e = tid + 128 * q;
        const int r = e / SUB2_JCHUNK, j = e % SUB2_JCHUNK;
        const int gi = rb * SUB2_ROWS + r, gj = j0 + j;
        const bool ok = gi < n && gj < n;
        cp_async8(as + r * A_LD + j, ok ? A + (long)gi * n + gj : A, ok);
      }
    }
  };
  auto load_y = [&](int ch, int buf) {
    const int j0 = ch * SUB2_JCHUNK;
    double* ys = Ys + buf * Y_ST;
#pragma unroll
    for (int q = 0; q < SUB2_JCHUNK * KB / 2 / 128; ++q) {
      const int e = tid + 128 * q;
      const int j = e / (KB / 2), c2 = 2 * (e % (KB / 2));
      const bool ok = j0 + j < n;
      cp_async16(ys + j * Y_LD + c2, ok ? Y + (long)(j0 + j) * KB + c2 : Y, ok);
    }
  };
  if (PDL) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#pragma unroll
    for (int s = 0; s < SUB2_STAGES - 1; ++s)
      if (ch0 + s < ch1) load_a(ch0 + s, s);
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
#pragma unroll
  for (int s = 0; s < SUB2_STAGES - 1; ++s) {
    if (ch0 + s < ch1) {
      if (!PDL) load_a(ch0 + s, s);
      load_y(ch0 + s, s);
    }
    cp_async_commit();
  }
  for (int ch = ch0; ch < ch1; ++ch) {
    cp_async_wait<SUB2_STAGES - 2>();
    __syncthreads();
    {
      const int nx = ch + SUB2_STAGES - 1;
      if (nx < ch1) {
        load_a(nx, (nx - ch0) % SUB2_STAGES);
        load_y(nx, (nx - ch0) % SUB2_STAGES);
      }
      cp_async_commit();
    }
    const int buf = (ch - ch0) % SUB2_STAGES;
    // warp w takes contraction indices [8 w, 8 w + 8) of the chunk (two k4 steps)
    const double* as = As + buf * A_ST + gq * A_LD + 8 * warp + tq;
    const double* ys = Ys + buf * Y_ST + (8 * warp + tq) * Y_LD + gq;
#pragma unroll
    for (int k4 = 0; k4 < 8; k4 += 4) {
      double af[MI], bf[NI];
#pragma unroll
      for (int i = 0; i < MI; ++i) af[i] = as[8 * i * A_LD + k4];
#pragma unroll
      for (int j = 0; j < NI; ++j) bf[j] = ys[k4 * Y_LD + 8 * j];
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma(acc[i][j], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  // the four warps' partial tiles -> this CTA's tile (fixed order)
  double* red = sub_smem;  // [4][32][KB]
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      double* p = red + ((warp * SUB2_ROWS) + 8 * i + gq) * KB + 8 * j + 2 * tq;
      p[0] = acc[i][j][0];
      p[1] = acc[i][j][1];
    }
  __syncthreads();
  for (int e = tid; e < SUB2_ROWS * KB; e += 128)
    cta_tile[e] = ((red[e] + red[SUB2_ROWS * KB + e]) + red[2 * SUB2_ROWS * KB + e]) + red[3 * SUB2_ROWS * KB + e];
  cluster_sync_all();  // all four partial tiles are in place (release / acquire at cluster scope)
  if (ks == 0) {
    const uint32_t t0 = smem_u32(cta_tile);
    const uint32_t t1 = dsmem_addr(t0, 1), t2 = dsmem_addr(t0, 2), t3 = dsmem_addr(t0, 3);
    const double alpha = b == 0 ? a.alpha[0] : a.alpha[1], shift = b == 0 ? a.shift[0] : a.shift[1],
                 beta = b == 0 ? a.beta[0] : a.beta[1];
    const double* Z = a.Z ? a.Z + (long)b * n * KB : nullptr;
    double* out = a.out + (long)b * n * KB;
    for (int e = tid; e < SUB2_ROWS * KB; e += 128) {
      const int gi = rb * SUB2_ROWS + e / KB;
      if (gi >= n) continue;
      const double s = ((cta_tile[e] + dsmem_ld(t1 + 8u * e)) + dsmem_ld(t2 + 8u * e)) + dsmem_ld(t3 + 8u * e);
      const long o = (long)gi * KB + e % KB;
      double v = alpha * (s - shift * Y[o]);
      if (Z) v -= beta * Z[o];
      out[o] = v;
    }
  }
  cluster_sync_all();  // the peers' shared memory stays alive until rank 0 has read it
}

// ---- single-shot bulk-copy variant (round 2, variant 2) --------------------------------------------------------
// The ring of the kernel above walks 11 chunk steps per CTA, each a cp.async wait + CTA barrier: ~12 us per block
// product although the bytes are worth 3-6 us of L2 bandwidth.  Here the contraction index is split over the EIGHT CTAs of
// a cluster, so that a CTA's whole operand slab - 32 rows x n/8 columns of F' and n/8 rows of Y, 74 KB at n = 1376 - fits
// in shared memory at once: one warp issues all bulk copies (TMA 1-D, one per matrix row / block row) in two halves on
// two mbarriers, every warp owns 8 output rows over the whole slab (no reduction inside the CTA) and starts on the first
// half while the second is still in flight.  The eight partial 32 x KB tiles are summed over DSMEM in rank order
// (deterministic), each CTA finishing 4 rows, with the epilogue's Y / Z operands prefetched at kernel entry.
// Three CTAs per SM: one spin's product (344 CTAs at n = 1376) is a single wave.  Needs even n (16-byte row starts).
constexpr int SUB3_ROWS = 32;
constexpr int SUB3_KS = 8;
__host__ __device__ constexpr int sub_apply3_klen(int n) { return ((n + SUB3_KS - 1) / SUB3_KS + 15) / 16 * 16; }
template <int KB>
__host__ __device__ constexpr int sub_apply3_smem_bytes(int n) {
  // [64 B barriers][A slab 32 x (KL + 4)][Y slab KL x (KB + 4)]; the CTA's partial tile reuses the A slab
  return 64 + (SUB3_ROWS * (sub_apply3_klen(n) + 4) + sub_apply3_klen(n) * (KB + 4)) * 8;
}

// PDL = 1: launched with programmatic stream serialisation.  The CTA lets the next product of the chain start at once
// (its barrier set-up and F' copies do not depend on this one) and itself touches Y / Z / out - the buffers the
// previous product reads or writes - only behind griddepcontrol.wait, i.e. after that grid has completed.
template <int KB, int PDL>
__global__ void __launch_bounds__(128) sub_apply3_kernel(SubApplyArgs a) {
  constexpr int NI = KB / 8, Y_LD = KB + 4;
  extern __shared__ __align__(128) unsigned char s3[];
  const int n = a.n, KL = sub_apply3_klen(n), P = KL + 4;
  uint64_t* bar = reinterpret_cast<uint64_t*>(s3);
  double* As = reinterpret_cast<double*>(s3 + 64);
  double* Ys = As + SUB3_ROWS * P;
  double* tile = As;  // [32][KB] after the product
  const int ks = (int)cluster_ctarank();
  const int rb = blockIdx.x, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const double* A = a.A + (long)b * n * n;
  const double* Y = a.Y + (long)b * n * KB;
  const int k0 = ks * KL;
  const int len = max(0, min(n, k0 + KL) - k0);  // contraction indices of this CTA (even, as n and KL are)
  const int half0 = min(len, KL / 2), half1 = len - half0;
  const int rows_valid = min(SUB3_ROWS, n - rb * SUB3_ROWS);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&bar[0], (uint32_t)((rows_valid * half0 + half0 * KB) * 8));
      mbar_expect_tx(&bar[1], (uint32_t)((rows_valid * half1 + half1 * KB) * 8));
    }
    __syncwarp();
    if (lane < rows_valid) {
      const double* src = A + (long)(rb * SUB3_ROWS + lane) * n + k0;
      if (half0 > 0) bulk_g2s(As + lane * P, src, (uint32_t)(half0 * 8), &bar[0]);
      if (half1 > 0) bulk_g2s(As + lane * P + half0, src + half0, (uint32_t)(half1 * 8), &bar[1]);
    }
  }
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
  // epilogue operands of the 4 rows this CTA finishes (thread e < 4 KB): in flight while the slab arrives
  const int erow = rb * SUB3_ROWS + 4 * ks + tid / KB;
  const bool emine = tid < 4 * KB && erow < n;
  double ey = 0.0, ez = 0.0;
  if (emine) {
    ey = __ldcg(Y + (long)erow * KB + tid % KB);
    if (a.Z) ez = __ldcg(a.Z + (long)b * n * KB + (long)erow * KB + tid % KB);
  }
  if (warp != 0) {
    // rows of the Y slab, padded pitch (conflict-free B fragments): warps 1-3 issue the 128-byte row copies
    for (int j = tid - 32; j < len; j += 96)
      bulk_g2s(Ys + j * Y_LD, Y + (long)(k0 + j) * KB, (uint32_t)(KB * 8), &bar[j < half0 ? 0 : 1]);
  }
  // contraction tail: zero up to the next multiple of 4 (disjoint from the bytes the copies write)
  const int len4 = (len + 3) & ~3;
  for (int e = tid; e < (len4 - len) * SUB3_ROWS; e += 128) As[(e % SUB3_ROWS) * P + len + e / SUB3_ROWS] = 0.0;
  for (int e = tid; e < (len4 - len) * KB; e += 128) Ys[(len + e / KB) * Y_LD + e % KB] = 0.0;
  __syncthreads();
  // warp w: output rows [8 w, 8 w + 8) over the whole slab; two independent accumulator sets (alternate k4 steps)
  double acc[2][NI][2];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[h][j][0] = acc[h][j][1] = 0.0;
  const double* as = As + (8 * warp + gq) * P + tq;
  const double* ys = Ys + tq * Y_LD + gq;
  auto run = [&](int kb, int ke) {  // k4 steps [kb, ke), kb even count from a multiple of 8
    int k = kb;
    for (; k + 8 <= ke; k += 8) {
      const double a0 = as[k], a1 = as[k + 4];
      double b0[NI], b1[NI];
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        b0[j] = ys[k * Y_LD + 8 * j];
        b1[j] = ys[(k + 4) * Y_LD + 8 * j];
      }
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        dmma(acc[0][j], a0, b0[j]);
        dmma(acc[1][j], a1, b1[j]);
      }
    }
    for (; k < ke; k += 4) {
      const double a0 = as[k];
#pragma unroll
      for (int j = 0; j < NI; ++j) dmma(acc[0][j], a0, ys[k * Y_LD + 8 * j]);
    }
  };
  // the first half is KL / 2 long (a multiple of 8) whenever a second half exists
  mbar_wait(&bar[0], 0);
  run(0, half1 > 0 ? half0 : len4);
  if (half1 > 0) {
    mbar_wait(&bar[1], 0);
    run(half0, len4);
  }
  __syncthreads();  // every warp is done with the A slab: it now holds this CTA's partial tile
#pragma unroll
  for (int j = 0; j < NI; ++j) {
    double* t = tile + (8 * warp + gq) * KB + 8 * j + 2 * tq;
    t[0] = acc[0][j][0] + acc[1][j][0];
    t[1] = acc[0][j][1] + acc[1][j][1];
  }
  cluster_sync_all();  // all eight partial tiles are in place (release / acquire at cluster scope)
  if (emine) {
    const uint32_t t0 = smem_u32(tile + (4 * ks) * KB + tid);
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < SUB3_KS; ++q) s += dsmem_ld(dsmem_addr(t0, (uint32_t)q));
    const double alpha = b == 0 ? a.alpha[0] : a.alpha[1], shift = b == 0 ? a.shift[0] : a.shift[1],
                 beta = b == 0 ? a.beta[0] : a.beta[1];
    double v = alpha * (s - shift * ey);
    if (a.Z) v -= beta * ez;
    a.out[(long)b * n * KB + (long)erow * KB + tid % KB] = v;
  }
  cluster_sync_all();  // the peers' shared memory stays alive until every CTA has read it
}

// (A persistent variant that walked a whole Chebyshev filter in one cooperative launch - grid barrier on a monotonic
// counter between the steps - was built and measured in round 2: bit-identical results, half the launches, the same
// time (0.840 against 0.849 ms per cycle at C4).  A block product moves 45 MB through L2 -> SM, ~5.6 us at the measured
// 8 TB/s, so launch latency was never the limiter; removed again, profiles/r02_persistent_filter.md.)
constexpr int SUBF_MAX_STEPS = 26;  // filter steps whose coefficients the host precomputes

// G[b][0] = Y^T Y, G[b][1] = Y^T W  (KB x KB each), rows split over gridDim.x CTAs, deterministic final sum.
template <int KB>
__global__ void __launch_bounds__(256) sub_gram_kernel(const double* __restrict__ Yall, const double* __restrict__ Wall,
                                                       double* __restrict__ part, double* __restrict__ G,
                                                       unsigned int* __restrict__ ticket, int n) {
  __shared__ double Ys[32][KB], Ws[32][KB];
  __shared__ unsigned int is_last;
  const int b = blockIdx.y, tid = threadIdx.x;
  const double* Y = Yall + (long)b * n * KB;
  const double* W = Wall + (long)b * n * KB;
  const int nrb = (n + 31) / 32;
  const int r0 = (int)((long)nrb * blockIdx.x / gridDim.x) * 32, r1 = min(n, (int)((long)nrb * (blockIdx.x + 1) / gridDim.x) * 32);
  constexpr int EPT = (2 * KB * KB + 255) / 256;  // output elements per thread
  double acc[EPT];
#pragma unroll
  for (int q = 0; q < EPT; ++q) acc[q] = 0.0;
  for (int rr = r0; rr < r1; rr += 32) {
    for (int e = tid; e < 32 * KB; e += 256) {
      const int r = e / KB, c = e % KB;
      const bool ok = rr + r < r1;
      Ys[r][c] = ok ? Y[(long)(rr + r) * KB + c] : 0.0;
      Ws[r][c] = ok ? W[(long)(rr + r) * KB + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < EPT; ++q) {
      const int e = tid + 256 * q;
      if (e < 2 * KB * KB) {
        const int which = e / (KB * KB), c = (e / KB) % KB, d = e % KB;
        double s = acc[q];
        if (which == 0) for (int r = 0; r < 32; ++r) s = fma(Ys[r][c], Ys[r][d], s);
        else for (int r = 0; r < 32; ++r) s = fma(Ys[r][c], Ws[r][d], s);
        acc[q] = s;
      }
    }
    __syncthreads();
  }
  double* mypart = part + ((long)b * gridDim.x + blockIdx.x) * 2 * KB * KB;
#pragma unroll
  for (int q = 0; q < EPT; ++q) {
    const int e = tid + 256 * q;
    if (e < 2 * KB * KB) mypart[e] = acc[q];
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int t = atomicAdd(&ticket[b], 1u);
    is_last = (t == gridDim.x - 1) ? 1u : 0u;
    if (is_last) ticket[b] = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int e = tid; e < 2 * KB * KB; e += 256) {
    double s = 0.0;
    for (int k = 0; k < (int)gridDim.x; ++k) s += part[((long)b * gridDim.x + k) * 2 * KB * KB + e];
    G[(long)b * 2 * KB * KB + e] = s;
  }
}

// V = Y M, AV = W M  (M [batch][KB][KB] row-major), and per-column residual sums r_c = sum_i (AV - theta_c V)^2
template <int KB>
__global__ void __launch_bounds__(256) sub_rotate_kernel(const double* __restrict__ Yall, const double* __restrict__ Wall,
                                                         const double* __restrict__ Mall, const double* __restrict__ theta,
                                                         double* __restrict__ Vall, double* __restrict__ AVall,
                                                         double* __restrict__ rpart, int n) {
  __shared__ double Ms[KB][KB + 1];
  __shared__ double red[256];
  const int b = blockIdx.y, tid = threadIdx.x;
  const double* M = Mall + (long)b * KB * KB;
  for (int e = tid; e < KB * KB; e += 256) Ms[e / KB][e % KB] = M[e];
  __syncthreads();
  const double* Y = Yall + (long)b * n * KB;
  const double* W = Wall + (long)b * n * KB;
  double* V = Vall + (long)b * n * KB;
  double* AV = AVall + (long)b * n * KB;
  const int col = tid % KB;
  const double th = theta[b * KB + col];
  double rs = 0.0;
  constexpr int RPB = 256 / KB;
  for (int i = blockIdx.x * RPB + tid / KB; i < n; i += gridDim.x * RPB) {
    double v = 0.0, w = 0.0;
#pragma unroll
    for (int d = 0; d < KB; ++d) {
      const double m = Ms[d][col];
      v = fma(Y[(long)i * KB + d], m, v);
      w = fma(W[(long)i * KB + d], m, w);
    }
    V[(long)i * KB + col] = v;
    AV[(long)i * KB + col] = w;
    const double r = w - th * v;
    rs = fma(r, r, rs);
  }
  red[tid] = rs;
  __syncthreads();
  if (tid < KB) {
    double s = 0.0;
    for (int k = tid; k < 256; k += KB) s += red[k];
    rpart[((long)b * gridDim.x + blockIdx.x) * KB + tid] = s;
  }
}

// Gershgorin-type upper bound of the spectrum: part[b][blk] = max over the block's rows of sum_j |A[b][i][j]|
// (the host takes the max over blocks); one warp per row, 8 rows per CTA.
__global__ void __launch_bounds__(256) sub_gershgorin_kernel(const double* __restrict__ Aall, int n, double* __restrict__ part) {
  __shared__ double red[8];
  const double* A = Aall + (long)blockIdx.y * n * n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  double s = 0.0;
  if (i < n)
    for (int j = lane; j < n; j += 32) s += fabs(A[(long)i * n + j]);
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int w = 0; w < 8; ++w) m = fmax(m, red[w]);
    part[(long)blockIdx.y * gridDim.x + blockIdx.x] = m;
  }
}

// rows [0, KB) of the eigenvector matrix (row-major, row = eigenvector) -> block V [n][KB]
__global__ void sub_gather_block_kernel(const double* __restrict__ Crows, double* __restrict__ V, int n, int kb, long cstride, long vstride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (i < n) V[(long)b * vstride + (long)i * kb + c] = Crows[(long)b * cstride + (long)c * n + i];
}
// block V [n][KB] -> rows [0, KB) of Ct (row = orbital)
__global__ void sub_scatter_block_kernel(const double* __restrict__ V, double* __restrict__ Crows, int n, int kb, long cstride, long vstride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (i < n) Crows[(long)b * cstride + (long)c * n + i] = V[(long)b * vstride + (long)i * kb + c];
}

// ---- spectral bounds: a few Lanczos steps (Zhou & Li), then ||dF||_F updates while the block is tracked ---------
// The recurrence scalars live on the device (coef[b][0..15] = alpha_j, coef[b][16..32] = beta_j, beta_0 = 0), so a whole
// run is one stream of launches and ONE read-back (round 2: ten host round trips per run before).
constexpr int SUB_LZ_COEF = 40;  // doubles per spin in the coefficient array
// One Lanczos step per spin: w = A v - beta_j vprev, part[b][blk] = {sum_i w_i v_i, sum_i w_i^2} over the block's rows
// (one warp per row, 8 rows per CTA; sub_lanczos_scalars_kernel adds the blocks in order).
__global__ void __launch_bounds__(256) sub_lanczos_matvec_kernel(const double* __restrict__ Aall, const double* __restrict__ vall,
                                                                 const double* __restrict__ vprevall, const double* __restrict__ coef,
                                                                 int j, double* __restrict__ wall, double* __restrict__ part, int n) {
  __shared__ double red[8][2];
  const int b = blockIdx.y;
  const double* A = Aall + (long)b * n * n;
  const double* v = vall + (long)b * n;
  const double beta = __ldcg(coef + b * SUB_LZ_COEF + 16 + j);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  double s = 0.0;
  if (i < n)
    for (int jj = lane; jj < n; jj += 32) s = fma(A[(long)i * n + jj], __ldcg(v + jj), s);
  s = warp_sum(s);
  if (lane == 0) {
    double wv = 0.0, ww = 0.0;
    if (i < n) {
      const double vi = __ldcg(v + i);
      s -= beta * __ldcg(vprevall + (long)b * n + i);
      wall[(long)b * n + i] = s;
      wv = s * vi;
      ww = s * s;
    }
    red[warp][0] = wv;
    red[warp][1] = ww;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    part[((long)b * gridDim.x + blockIdx.x) * 2 + threadIdx.x] = t;
  }
}
// alpha_j = sum_blk part[.][0], beta_{j+1} = sqrt(max(0, sum_blk part[.][1] - alpha_j^2)); one CTA of 32 threads per spin,
// blocks added in index order (the order the host used to add them in)
__global__ void sub_lanczos_scalars_kernel(const double* __restrict__ part, int nblk, int j, double* __restrict__ coef) {
  __shared__ double acc[2];
  const int b = blockIdx.x;
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int k = 0; k < nblk; ++k) t += __ldcg(part + ((long)b * nblk + k) * 2 + threadIdx.x);
    acc[threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double a = acc[0];
    coef[b * SUB_LZ_COEF + j] = a;
    coef[b * SUB_LZ_COEF + 16 + j + 1] = sqrt(fmax(0.0, acc[1] - a * a));
  }
}
// vnext = (w - alpha_j v) / beta_{j+1}   (a vanishing beta gives non-finite vectors; the host cuts the run there)
__global__ void sub_lanczos_update_kernel(const double* __restrict__ w, const double* __restrict__ v, double* __restrict__ vnext,
                                          const double* __restrict__ coef, int j, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= n) return;
  const long o = (long)b * n + i;
  const double alpha = __ldcg(coef + b * SUB_LZ_COEF + j), beta = __ldcg(coef + b * SUB_LZ_COEF + 16 + j + 1);
  vnext[o] = (__ldcg(w + o) - alpha * __ldcg(v + o)) * (1.0 / beta);
}
// part[b][blk] = sum over the block's slice of (A - B)^2   (gridDim.x slices per matrix, fixed order)
// and B <- A (the matrix the next cycle compares with) in the same pass
__global__ void __launch_bounds__(256) sub_diffnorm_kernel(const double* __restrict__ A, double* __restrict__ B, long count,
                                                           double* __restrict__ part) {
  __shared__ double red[32];
  const int b = blockIdx.y;
  const long per = (count + gridDim.x - 1) / gridDim.x;
  const long e0 = per * blockIdx.x, e1 = e0 + per < count ? e0 + per : count;
  double s = 0.0;
  for (long e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    const double av = __ldcg(A + (long)b * count + e);
    const double d = av - __ldcg(B + (long)b * count + e);
    B[(long)b * count + e] = av;
    s = fma(d, d, s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) part[(long)b * gridDim.x + blockIdx.x] = s;
}
// deterministic pseudo-random start block in [-1, 1) (SplitMix64 of the element index)
__global__ void sub_random_block_kernel(double* __restrict__ V, long count, unsigned long long seed) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= count) return;
  unsigned long long z = seed * 0xD1B54A32D192ED03ull + (unsigned long long)e + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  V[e] = (double)(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}

}  // namespace nbd
