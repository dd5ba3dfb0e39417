// Density-fitted Coulomb / exchange build over the device-resident 3-centre tensor.
//
//   pass 1  X[P][i][mu] = sum_nu B[P][mu][nu] Ct[i][nu]        (symm_panel_kernel: B streamed ONCE, FP64 DMMA)
//           rho[P]      = sum_{i,mu} w_i Ct[i][mu] X[P][i][mu]  (rho_kernel; = sum_{mu nu} B D, D = sum_i w_i c_i c_i^T)
//   pass 2  J           = sum_P rho[P] B[P]                     (j_pass_kernel: second compulsory pass, HBM bound)
//           K_s         = sum_{P, i in s} w_i X[P][i] X[P][i]^T (gemm.cuh, lower tiles, K-dim = naux*nocc)
//
// The reference reaches the same contraction through pyscf.df.df_jk.get_jk (occupied-orbital branch) from
// nbed/scf/huzinaga_scf.py:156 and nbed/scf/embedded_hcore_funcs.py:34.
#pragma once
#include <vector>
#include "common.cuh"

namespace nbd {

// ---------------------------------------------------------------------------------------------------
// Tile order.  Panels of 32 AOs; tile (I, J), J <= I.  Consumer warp (I % 8) owns the "row" update
// X[I] += B_IJ C_J and warp (J % 8) the "column" update X[J] += B_IJ^T C_I, so inside every run of 8
// consecutive tiles of this skewed order all 8 warps get one task of each kind.
// ---------------------------------------------------------------------------------------------------
inline std::vector<int> build_tile_sequence(int nb) {
  std::vector<int> seq;
  const int nsb = (nb + 7) / 8;
  for (int a = 0; a < nsb; ++a)
    for (int b = 0; b <= a; ++b)
      for (int k = 0; k < 8; ++k)
        for (int j = 0; j < 8; ++j) {
          const int I = 8 * a + (j + k) % 8, J = 8 * b + j;
          if (I >= nb || J >= nb || J > I) continue;
          seq.push_back((I << 16) | J);
        }
  return seq;
}

__host__ __device__ __forceinline__ long packed_index(int mu, int nu) {
  const int hi = mu > nu ? mu : nu, lo = mu > nu ? nu : mu;
  return (long)hi * (hi + 1) / 2 + lo;
}

// ---- layout transforms -----------------------------------------------------------------------------
// packed rows (device staging, [nrows][npair]) -> tiled; one CTA per (tile, row)
__global__ void pack_to_tiled_kernel(const double* __restrict__ packed, double* __restrict__ tiled,
                                     const int* __restrict__ seq, int ntiles, int n, long npair, int row0) {
  const int k = blockIdx.x, p = blockIdx.y;
  const int I = seq[k] >> 16, J = seq[k] & 0xffff;
  const double* src = packed + (long)p * npair;
  double* dst = tiled + ((long)(row0 + p) * ntiles + k) * TILE_ELEMS;
  for (int e = threadIdx.x; e < TILE_ELEMS; e += blockDim.x) {
    const int r = e >> 5, c = e & 31;
    const int mu = 32 * I + r, nu = 32 * J + c;
    double v = 0.0;
    if (mu < n && nu < n) v = src[packed_index(mu, nu)];
    dst[tile_swz(r, c)] = v;
  }
}

__global__ void tiled_to_packed_kernel(const double* __restrict__ tiled, double* __restrict__ packed,
                                       const int* __restrict__ seq, int ntiles, int n, long npair, int row0) {
  const int k = blockIdx.x, p = blockIdx.y;
  const int I = seq[k] >> 16, J = seq[k] & 0xffff;
  double* dst = packed + (long)p * npair;
  const double* src = tiled + ((long)(row0 + p) * ntiles + k) * TILE_ELEMS;
  for (int e = threadIdx.x; e < TILE_ELEMS; e += blockDim.x) {
    const int r = e >> 5, c = e & 31;
    const int mu = 32 * I + r, nu = 32 * J + c;
    if (mu < n && nu < n && nu <= mu) dst[packed_index(mu, nu)] = src[tile_swz(r, c)];
  }
}

// SplitMix64 -> uniform [-1,1): bit-identical to nbed_b200/synthetic.py:hash_uniform
__device__ __forceinline__ double synth_uniform(unsigned long long seed, unsigned long long index) {
  unsigned long long z = (seed << 48) + index + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}

__global__ void synth_tiled_kernel(double* __restrict__ tiled, const int* __restrict__ seq, int ntiles, int n,
                                   long npair, unsigned long long seed, double scale, int global_row0) {
  const int k = blockIdx.x, p = blockIdx.y;
  const int I = seq[k] >> 16, J = seq[k] & 0xffff;
  double* dst = tiled + ((long)p * ntiles + k) * TILE_ELEMS;
  for (int e = threadIdx.x; e < TILE_ELEMS; e += blockDim.x) {
    const int r = e >> 5, c = e & 31;
    const int mu = 32 * I + r, nu = 32 * J + c;
    double v = 0.0;
    if (mu < n && nu < n)
      v = scale * synth_uniform(seed, (unsigned long long)(global_row0 + p) * (unsigned long long)npair +
                                          (unsigned long long)packed_index(mu, nu));
    dst[tile_swz(r, c)] = v;
  }
}

// ---- pass 1: X = B_sym * C, one B tile enters the SM once and is used in both directions -------------
struct XArgs {
  const double* Bt;     // [naux][ntiles][1024]
  const int* seq;       // [ntiles]
  const double* Ct;     // [Ntot][n_ld]
  double* X;            // group-major: column i of aux row P lives at X + xbase[i] + P * xstride[i]  (n_ld doubles)
  const long* xbase;    // [Ntot]
  const long* xstride;  // [Ntot]
  const uint32_t* events;  // per-warp task lists (PanelPlan)
  const int* evbegin;      // [9] offsets of the 8 consumer warps' lists; [9..16] = first tile index of each list
  int naux, ntiles, nb, n_ld, Ntot, nslices, nstages, ncolmax;
  int zrow;  // 1: a zero row follows the slice (stands in for the padded columns of a ragged last slice)
};

// ---------------------------------------------------------------------------------------------------
// Per-warp task lists of the panel kernel.  An event = one tile a consumer warp takes part in:
//   bits 0-7 I, 8-15 J, 16-20 distance (in tiles, cyclic over the row) to this warp's next event,
//   21-25 weight of its arrival on the stage's "empty" barrier, 26 row task, 27 column task.
// A warp only touches tiles it has an event on.  mbarrier parity waits can only tell the current phase from
// the previous one, so two invariants are built into the lists: (1) consecutive events of a warp are at most
// S tiles apart (dummy events - wait + release, no math - are inserted where a warp owns nothing for longer),
// which together with 2S "full" barriers over S data stages guarantees a waiter is never two phases off;
// (2) every tile's participants' weights add up to XK_EMPTY_COUNT, so a stage is refilled only after all of
// them have released it.
// ---------------------------------------------------------------------------------------------------
constexpr int XK_EMPTY_COUNT = 16;
struct PanelPlan {
  int nb = 0, S = 0;
  std::vector<uint32_t> events;
  std::vector<int> begin;  // 9 offsets + 8 first-tile indices
};
inline PanelPlan build_panel_plan(int nb, int S, const std::vector<int>& seq) {
  PanelPlan pl;
  pl.nb = nb;
  pl.S = S;
  const int nt = (int)seq.size();
  std::vector<std::vector<int>> ks(8);          // tile indices with an event, per warp
  std::vector<std::vector<int>> flags(8);       // bit0 row, bit1 col (0 = dummy)
  for (int w = 0; w < 8; ++w) {
    std::vector<int> rk, rf;
    for (int k = 0; k < nt; ++k) {
      const int I = seq[k] >> 16, J = seq[k] & 0xffff;
      const int f = (((I & 7) == w) ? 1 : 0) | ((((J & 7) == w) && I != J) ? 2 : 0);
      if (f) {
        rk.push_back(k);
        rf.push_back(f);
      }
    }
    if (rk.empty()) continue;
    // dummies: first event within the first S tiles, consecutive gaps <= S, cyclic gap <= S
    int prev = -1;
    for (size_t e = 0; e < rk.size(); ++e) {
      while (rk[e] - prev > S) {
        prev += S;
        ks[w].push_back(prev);
        flags[w].push_back(0);
      }
      ks[w].push_back(rk[e]);
      flags[w].push_back(rf[e]);
      prev = rk[e];
    }
    while (nt - prev + ks[w][0] > S) {
      prev = prev + S < nt - 1 ? prev + S : nt - 1;
      ks[w].push_back(prev);
      flags[w].push_back(0);
    }
  }
  std::vector<int> npart(nt, 0), seen(nt, 0);
  for (int w = 0; w < 8; ++w)
    for (int k : ks[w]) ++npart[k];
  pl.begin.assign(17, 0);
  for (int w = 0; w < 8; ++w) {
    pl.begin[w] = (int)pl.events.size();
    pl.begin[9 + w] = ks[w].empty() ? 0 : ks[w][0];
    const int ne = (int)ks[w].size();
    for (int e = 0; e < ne; ++e) {
      const int k = ks[w][e];
      const int gap = e + 1 < ne ? ks[w][e + 1] - k : nt - k + ks[w][0];
      const int weight = seen[k]++ == 0 ? XK_EMPTY_COUNT - (npart[k] - 1) : 1;
      const int I = seq[k] >> 16, J = seq[k] & 0xffff;
      pl.events.push_back((uint32_t)I | ((uint32_t)J << 8) | ((uint32_t)gap << 16) | ((uint32_t)weight << 21) |
                          ((uint32_t)flags[w][e] << 26));
    }
  }
  pl.begin[8] = (int)pl.events.size();
  // self-check of the two invariants (cheap; a violated invariant would be a device-side hang)
  std::vector<int> wsum(nt, 0);
  for (int w = 0; w < 8; ++w) {
    int k = pl.begin[9 + w];
    if (pl.begin[w] < pl.begin[w + 1] && k > S - 1) pl.S = -1;
    for (int e = pl.begin[w]; e < pl.begin[w + 1]; ++e) {
      const uint32_t ev = pl.events[e];
      const int gap = (ev >> 16) & 31;
      if (gap < 1 || gap > S || (int)(ev & 255) != (seq[k] >> 16) || (int)((ev >> 8) & 255) != (seq[k] & 0xffff)) pl.S = -1;
      wsum[k] += (ev >> 21) & 31;
      k = (k + gap) % nt;
    }
    if (pl.begin[w] < pl.begin[w + 1] && k != pl.begin[9 + w]) pl.S = -1;
  }
  for (int k = 0; k < nt; ++k)
    if (wsum[k] != XK_EMPTY_COUNT) pl.S = -1;
  return pl;
}

// 8 consumer warps (warpgroups 0-1) + one producer warpgroup.  Registers are re-balanced with setmaxnreg
// (the accumulators of a 1376-AO row set need ~200 registers per consumer thread; a 9-warp CTA would be
// capped at 168 because the register file is split per SM sub-partition).
constexpr int XK_CONSUMER_WARPS = 8;
constexpr int XK_THREADS = (XK_CONSUMER_WARPS + 4) * 32;
constexpr int XK_CONSUMER_REGS = 240;
constexpr int XK_PRODUCER_REGS = 24;

// NF extra columns (beyond the 8 * NB DMMA columns) ride along on the FP64 FMA pipe: the lane that holds the DMMA
// A-fragment element (row 8 mi + gq, contraction index 4 ks + tq) multiplies it with the orbital value of the same
// contraction index, so accf[mi][f] is this lane's partial sum over the contraction indices == tq (mod 4); the four
// tq lanes are summed once per aux row.  DFMA and DMMA share one pipe at the same FLOP rate (profiles/microbench_r01.md),
// so 8 + NF columns cost (8 + NF) / 8 of a DMMA block instead of 2 blocks.
template <int NB, int NF>
__device__ __forceinline__ void xk_task_row(double (&acc)[4][NB][2], double (&accf)[4][NF > 0 ? NF : 1],
                                            const double* __restrict__ tile, const double* const (&crow)[NB],
                                            const double* __restrict__ cf, int ct_ld, int colbase, int gq, int tq,
                                            const int (&xoff)[4]) {
  // fragments of step ks+1 are loaded before the DMMAs of step ks issue (register double buffering)
  double a[2][4], b[2][NB], f[2][NF > 0 ? NF : 1];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) a[0][mi] = tile[(8 * mi + gq) * 32 + xoff[0]];
#pragma unroll
  for (int ni = 0; ni < NB; ++ni) b[0][ni] = crow[ni][colbase + tq];
#pragma unroll
  for (int fi = 0; fi < NF; ++fi) f[0][fi] = cf[fi * ct_ld + colbase + tq];
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const int cur = ks & 1, nxt = cur ^ 1;
    if (ks + 1 < 8) {
      const int co = (((ks + 1) >> 2) << 4) + xoff[(ks + 1) & 3];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) a[nxt][mi] = tile[(8 * mi + gq) * 32 + co];
#pragma unroll
      for (int ni = 0; ni < NB; ++ni) b[nxt][ni] = crow[ni][colbase + 4 * (ks + 1) + tq];
#pragma unroll
      for (int fi = 0; fi < NF; ++fi) f[nxt][fi] = cf[fi * ct_ld + colbase + 4 * (ks + 1) + tq];
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
      for (int ni = 0; ni < NB; ++ni) dmma(acc[mi][ni], a[cur][mi], b[cur][ni]);
#pragma unroll
      for (int fi = 0; fi < NF; ++fi) accf[mi][fi] = fma(a[cur][mi], f[cur][fi], accf[mi][fi]);
    }
  }
}

template <int NB, int NF>
__device__ __forceinline__ void xk_task_col(double (&acc)[4][NB][2], double (&accf)[4][NF > 0 ? NF : 1],
                                            const double* __restrict__ tile, const double* const (&crow)[NB],
                                            const double* __restrict__ cf, int ct_ld, int colbase, int tq,
                                            const int (&yoff)[4]) {
  double a[2][4], b[2][NB], f[2][NF > 0 ? NF : 1];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) a[0][mi] = tile[tq * 32 + yoff[mi]];
#pragma unroll
  for (int ni = 0; ni < NB; ++ni) b[0][ni] = crow[ni][colbase + tq];
#pragma unroll
  for (int fi = 0; fi < NF; ++fi) f[0][fi] = cf[fi * ct_ld + colbase + tq];
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const int cur = ks & 1, nxt = cur ^ 1;
    if (ks + 1 < 8) {
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) a[nxt][mi] = tile[(4 * (ks + 1) + tq) * 32 + yoff[mi]];
#pragma unroll
      for (int ni = 0; ni < NB; ++ni) b[nxt][ni] = crow[ni][colbase + 4 * (ks + 1) + tq];
#pragma unroll
      for (int fi = 0; fi < NF; ++fi) f[nxt][fi] = cf[fi * ct_ld + colbase + 4 * (ks + 1) + tq];
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
      for (int ni = 0; ni < NB; ++ni) dmma(acc[mi][ni], a[cur][mi], b[cur][ni]);
#pragma unroll
      for (int fi = 0; fi < NF; ++fi) accf[mi][fi] = fma(a[cur][mi], f[cur][fi], accf[mi][fi]);
    }
  }
}

// NSLOT = ceil(nb / 8) panels owned per consumer warp; NB = 8-column DMMA blocks per slice (1 or 2); NF = extra
// FMA-pipe columns per slice (0-2; the host launches NF > 0 only on exactly 8 * NB + NF columns, one slice).
template <int NSLOT, int NB, int NF>
__global__ void __launch_bounds__(XK_THREADS, 1) symm_panel_kernel(XArgs p) {
  extern __shared__ __align__(128) unsigned char xsm[];
  constexpr int NCOL = 8 * NB + NF;
  constexpr int NFA = NF > 0 ? NF : 1;
  const int S = p.nstages;
  uint64_t* full = reinterpret_cast<uint64_t*>(xsm);  // 2S "full" barriers (tile q uses q mod 2S) ...
  uint64_t* empty = full + 32;                        // ... over S data stages / "empty" barriers (q mod S)
  const int ct_ld = p.n_ld + 4;
  const int slice = blockIdx.x % p.nslices;
  const int col0 = slice * NCOL;
  const int ncol = min(NCOL, p.Ntot - col0);
  // layout: [512 B barriers][(ncolmax + zrow) rows of Ct][S stage buffers]   (sizes fixed by the host)
  double* cts = reinterpret_cast<double*>(xsm + 512);
  const size_t ct_bytes = ((size_t)(p.ncolmax + p.zrow) * ct_ld * 8 + 127) & ~(size_t)127;
  double* stages = reinterpret_cast<double*>(xsm + 512 + ct_bytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < 2 * S; ++s) mbar_init(&full[s], 1);
    for (int s = 0; s < S; ++s) mbar_init(&empty[s], XK_EMPTY_COUNT);
    mbar_fence_init();
  }
  __syncthreads();

  const long nitems = (long)p.naux * p.nslices;
  // items of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ... (gridDim.x is a multiple of nslices)

  if (warp >= XK_CONSUMER_WARPS) {
    // ===== producer warpgroup: one lane streams the tiles of every item of this CTA through the ring =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(XK_PRODUCER_REGS));
    if (warp == XK_CONSUMER_WARPS && lane == 0) {
      int st = 0, fb = 0;
      uint32_t ph = 1;  // parity of the "previous round released" phase; the first round needs no wait
      bool first_round = true;
      const uint64_t pol = l2_evict_first_policy();
      for (long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const long P = item / p.nslices;
        const double* src = p.Bt + P * (long)p.ntiles * TILE_ELEMS;
        for (int k = 0; k < p.ntiles; ++k) {
          if (!first_round) mbar_wait(&empty[st], ph);
          mbar_expect_tx(&full[fb], TILE_BYTES);
          bulk_g2s_stream(stages + (size_t)st * TILE_ELEMS, src + (long)k * TILE_ELEMS, TILE_BYTES, &full[fb], pol);
          if (++fb == 2 * S) fb = 0;
          if (++st == S) {
            st = 0;
            ph ^= 1u;
            first_round = false;
          }
        }
      }
    }
    return;
  }

  // ===== consumers =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(XK_CONSUMER_REGS));
  for (int i = 0; i < ncol + p.zrow; ++i) {
    double* dst = cts + (size_t)i * ct_ld;
    if (i < ncol) {
      const double* src = p.Ct + (size_t)(col0 + i) * p.n_ld;
      for (int m = tid; m < p.n_ld; m += XK_CONSUMER_WARPS * 32) dst[m] = src[m];
    } else {
      for (int m = tid; m < p.n_ld; m += XK_CONSUMER_WARPS * 32) dst[m] = 0.0;
    }
  }
  asm volatile("bar.sync 1, %0;" ::"r"(XK_CONSUMER_WARPS * 32) : "memory");

  const int gq = lane >> 2, tq = lane & 3;
  int xoff[4], yoff[4];
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) xoff[q4] = 4 * (q4 ^ (gq & 3)) + tq;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) yoff[mi] = (8 * mi + gq) ^ (4 * tq);
  const double* crow[NB];
#pragma unroll
  for (int ni = 0; ni < NB; ++ni) crow[ni] = cts + (size_t)min(8 * ni + gq, ncol) * ct_ld;
  const double* cf = cts + (size_t)min(8 * NB, ncol) * ct_ld;  // rows of the NF extra columns

  double X[NSLOT][4][NB][2];
  double Xf[NSLOT][4][NFA];
#pragma unroll
  for (int s = 0; s < NSLOT; ++s)
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
      for (int ni = 0; ni < NB; ++ni) X[s][mi][ni][0] = X[s][mi][ni][1] = 0.0;
#pragma unroll
      for (int fi = 0; fi < NFA; ++fi) Xf[s][mi][fi] = 0.0;
    }

  const int e0 = __ldg(p.evbegin + warp), e1 = __ldg(p.evbegin + warp + 1);
  if (e0 == e1) return;  // this warp owns no panel of this matrix size: nothing to compute, wait for or write
  const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
  const uint32_t stages_a = smem_u32(stages);
  (void)stages_a;
  uint32_t fb = (uint32_t)__ldg(p.evbegin + 9 + warp), par = 0;  // first event lies within the first S tiles
  uint32_t ev_next = __ldg(p.events + e0);
  for (long item = blockIdx.x; item < nitems; item += gridDim.x) {
    const long P = item / p.nslices;
    for (int e = e0; e < e1; ++e) {
      const uint32_t ev = ev_next;
      ev_next = __ldg(p.events + (e + 1 < e1 ? e + 1 : e0));
      const int I = ev & 255, J = (ev >> 8) & 255;
      const uint32_t st = fb >= (uint32_t)S ? fb - (uint32_t)S : fb;
      mbar_wait_a(full_a + 8u * fb, par);
      if (ev & (3u << 26)) {
        const double* tile = stages + (size_t)st * TILE_ELEMS;
        if (ev & (1u << 26)) {
          const int slot = I >> 3;
#pragma unroll
          for (int s = 0; s < NSLOT; ++s)
            if (slot == s) xk_task_row<NB, NF>(X[s], Xf[s], tile, crow, cf, ct_ld, 32 * J, gq, tq, xoff);
        }
        if (ev & (2u << 26)) {
          const int slot = J >> 3;
#pragma unroll
          for (int s = 0; s < NSLOT; ++s)
            if (slot == s) xk_task_col<NB, NF>(X[s], Xf[s], tile, crow, cf, ct_ld, 32 * I, tq, yoff);
        }
        __syncwarp();
      }
      if (lane == 0) mbar_arrive_a(empty_a + 8u * st, (ev >> 21) & 31u);
      fb += (ev >> 16) & 31u;
      if (fb >= 2u * (uint32_t)S) {
        fb -= 2u * (uint32_t)S;
        par ^= 1u;
      }
    }
    // write this warp's panels of X[P] (group-major layout) and reset the accumulators
#pragma unroll
    for (int ni = 0; ni < NB; ++ni)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = 8 * ni + 2 * tq + r;
        const bool col_ok = i < ncol;
        double* xo = p.X;
        if (col_ok) xo += __ldcg(p.xbase + col0 + i) + P * __ldcg(p.xstride + col0 + i) + gq;  // tables change per call
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          const int I = 8 * s + warp;
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) {
            if (col_ok && I < p.nb) xo[32 * I + 8 * mi] = X[s][mi][ni][r];
            X[s][mi][ni][r] = 0.0;
          }
        }
      }
    // extra columns: sum the four tq lanes' partials, lane tq == fi writes column 8 * NB + fi
#pragma unroll
    for (int fi = 0; fi < NF; ++fi) {
      double* xo = p.X + __ldcg(p.xbase + col0 + 8 * NB + fi) + P * __ldcg(p.xstride + col0 + 8 * NB + fi) + gq;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int I = 8 * s + warp;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          double v = Xf[s][mi][fi];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          if (tq == fi && I < p.nb) xo[32 * I + 8 * mi] = v;
          Xf[s][mi][fi] = 0.0;
        }
      }
    }
  }
}

// Obviously-correct variant (option "jk_variant" = 1): one CTA per (P, panel), CUDA cores.
__global__ void symm_panel_simple_kernel(const double* __restrict__ Bt, const int* __restrict__ inv,
                                         const double* __restrict__ Ct, double* __restrict__ X,
                                         const long* __restrict__ xbase, const long* __restrict__ xstride, int ntiles,
                                         int nb, int n_ld, int Ntot) {
  const int P = blockIdx.y, I = blockIdx.x;
  const double* bp = Bt + (long)P * ntiles * TILE_ELEMS;
  for (int e = threadIdx.x; e < 32 * Ntot; e += blockDim.x) {
    const int r = e & 31, i = e >> 5;
    const int mu = 32 * I + r;
    double s = 0.0;
    for (int nu = 0; nu < n_ld; ++nu) {
      const int Jn = nu >> 5, c = nu & 31;
      double b;
      if (Jn <= I) b = bp[(long)inv[I * nb + Jn] * TILE_ELEMS + tile_swz(r, c)];
      else b = bp[(long)inv[Jn * nb + I] * TILE_ELEMS + tile_swz(c, r)];
      s += b * Ct[(long)i * n_ld + nu];
    }
    X[xbase[i] + (long)P * xstride[i] + mu] = s;
  }
}

// rho[set][P] = sum_{i in set} sum_mu X[P][i][mu] * Wt[i][mu]   (Wt = sign_i * Ct; X in the group-major layout)
__global__ void rho_kernel(const double* __restrict__ X, const long* __restrict__ xbase,
                           const long* __restrict__ xstride, const double* __restrict__ Wt, double* __restrict__ rho,
                           int naux, int n_ld, int nset, const int* __restrict__ set_begin) {
  __shared__ double red[32];
  const int P = blockIdx.x;
  for (int s = 0; s < nset; ++s) {
    const int i0 = set_begin[s], i1 = set_begin[s + 1];
    double v = 0.0;
    for (int i = i0; i < i1; ++i) {
      const double* x = X + xbase[i] + (long)P * xstride[i];
      const double* w = Wt + (long)i * n_ld;
      for (int e = threadIdx.x; e < n_ld; e += blockDim.x) v = fma(x[e], w[e], v);
    }
    v = block_sum(v, red);
    if (threadIdx.x == 0) rho[(long)s * naux + P] = v;
  }
}

// ---- pass 2: J (tiled layout) = sum_P rho[P] * B[P]; split over P, partials reduced by j_finalize ----
template <int NSET>
__global__ void __launch_bounds__(256) j_pass_kernel(const double2* __restrict__ Bt, const double* __restrict__ rho,
                                                     double2* __restrict__ part, long E2, int naux,
                                                     int rows_per_split) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= E2) return;
  const int split = blockIdx.y;
  const int p0 = split * rows_per_split;
  const int p1 = min(naux, p0 + rows_per_split);
  double2 acc[NSET];
#pragma unroll
  for (int s = 0; s < NSET; ++s) acc[s] = make_double2(0.0, 0.0);
  const double2* b = Bt + idx;
  int p = p0;
  for (; p + 8 <= p1; p += 8) {
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcs(b + (long)(p + u) * E2);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int s = 0; s < NSET; ++s) {
        const double r = __ldcg(rho + (long)s * naux + p + u);
        acc[s].x = fma(r, v[u].x, acc[s].x);
        acc[s].y = fma(r, v[u].y, acc[s].y);
      }
    }
  }
  for (; p < p1; ++p) {
    const double2 v = __ldcs(b + (long)p * E2);
#pragma unroll
    for (int s = 0; s < NSET; ++s) {
      const double r = __ldcg(rho + (long)s * naux + p);
      acc[s].x = fma(r, v.x, acc[s].x);
      acc[s].y = fma(r, v.y, acc[s].y);
    }
  }
#pragma unroll
  for (int s = 0; s < NSET; ++s) part[((long)split * NSET + s) * E2 + idx] = acc[s];
}

// NOTE on caching (a bug found in round 1): rho changes every SCF cycle, and this kernel co-runs with the K Gram on
// the same SMs.  Loaded through the non-coherent path (__ldg / ld.global.nc) it occasionally returned the PREVIOUS
// cycle's values - lines of that cache evidently survive when another grid keeps the SM busy across the launch
// boundary - which perturbed J by ~1e-6 near convergence and made repeated runs differ.  Everything that is rewritten
// between launches and read by a kernel that may share an SM with another grid is therefore loaded with
// ld.global.cg (L2 only): rho here, the partial sums in j_finalize_kernel, the layout tables in the panel kernel.
// tools/determinism.py checks that repeated SCF runs are bit-identical in every mode.
//
// TMA-fed variant of pass 2: persistent CTAs (a few per SM), a producer warp streams tile k of the aux rows of one
// P-range through a ring of 8 KiB bulk copies, four consumer warps accumulate rho[P] * tile in registers.  It keeps
// ~64 KiB in flight per CTA with 160 threads and 17 registers' worth of accumulators, so it reaches HBM speed next
// to the tensor-bound K Gram (whose CTAs leave little room for the register-hungry LDG version above).
constexpr int JP_STAGES = 7;  // 2 CTAs of 57 KiB fit next to one 101 KiB Gram CTA on an SM
constexpr int JP_STAGES_ALONE = 7;  // ring depth on an SM that runs only pass 2 (3 CTAs; 9 stages measured no faster)
constexpr int JP_MAX_STAGES = 12;
constexpr int JP_HEADER = 384;  // full[12] | empty[12] | item_q[16] | pad; the ring starts 128-byte aligned
constexpr int JP_CONSUMERS = 4;    // consumer warps (8 lighter warps were measured slower next to the Gram)
constexpr int JP_THREADS = (JP_CONSUMERS + 1) * 32;
constexpr int JP_Q = 512 / (JP_CONSUMERS * 32);  // double2 per thread per tile
// Work items (P-range, tile) are handed out dynamically (one atomic counter): next to the Gram the CTAs of this
// kernel run at very different speeds depending on what shares their SM, and a static split would wait for the
// slowest.  The producer lane draws the item, publishes it through shared memory ahead of the item's first tile
// (the mbarrier completion orders the two), and releases the consumers with a bare arrival when the queue is dry.
// SAFE: the stage is released only after every lane's shared-memory loads of the tile have RETURNED (a warp vote over a
// value computed from the loaded registers gates the arrival).  Without it the arrival is issued while the last LDS
// of the tile may still be in flight (the FMAs that consume them are scheduled behind it); next to the 8-warp
// stream-K Gram CTA of syrk.cuh that produced wrong tiles (tools/race_probe*.py, profiles/r02_pass2_race.md).
template <int NSET, bool SAFE = true>
__global__ void __launch_bounds__(JP_THREADS) j_pass_tma_kernel(const double* __restrict__ Bt, const double* __restrict__ rho,
                                                                double* __restrict__ part, int ntiles, int naux,
                                                                int nsplit, int rows_per_split,
                                                                unsigned int* __restrict__ next_item, int sm_mod,
                                                                int sm_keep, int nst, int guests = 0, long guest_limit = 0) {
  // spatial split of the pair (pass 2 on some SMs, the K Gram on the others): CTAs that land on an SM of the other
  // group retire at once; the dynamic item queue hands all the work to the CTAs that stay.
  // guests > 0: the first `guests` CTAs that land on a Gram SM stay as well (next_item[1 + smid] counts them).  Next to
  // the Gram's DMMA warps they stream at a fraction of the normal rate, but that is bandwidth the pair would otherwise
  // not use; they stop drawing items at `guest_limit` so that a slow guest never holds the last items of the queue.
  bool guest = false;
  if (sm_mod > 0 && !sm_in_group1(smid(), sm_keep, sm_mod)) {  // sm_keep of the sm_mod SMs run pass 2
    if (guests <= 0) return;
    __shared__ unsigned int s_guest;
    if (threadIdx.x == 0) s_guest = atomicAdd(next_item + 1 + (smid() & 511u), 1u);  // (the host zeroes 512 counters)
    __syncthreads();
    if (s_guest >= (unsigned int)guests) return;
    guest = true;
  }
  extern __shared__ __align__(128) unsigned char jsm[];
  // ring depth nst <= JP_MAX_STAGES is a launch parameter: 7 next to the Gram, 9 when pass 2 has its SMs to itself
  uint64_t* full = reinterpret_cast<uint64_t*>(jsm);
  uint64_t* empty = full + JP_MAX_STAGES;
  volatile long* item_q = reinterpret_cast<volatile long*>(jsm + 192);  // [16] items in flight (>= ring depth + 1), indexed by sequence & 15
  double* stages = reinterpret_cast<double*>(jsm + JP_HEADER);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], JP_CONSUMERS);
    }
    mbar_fence_init();
  }
  __syncthreads();
  // every CTA of this grid is resident from the start: a kernel launched behind it with programmatic stream
  // serialization (the K Gram, no data dependency) may now be scheduled into what is left of the SMs
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long nitems = (long)ntiles * nsplit;  // item = (split, tile): consecutive items are consecutive tiles
  if (warp == JP_CONSUMERS) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 1;
      bool first = true;
      const uint64_t pol = l2_evict_first_policy();
      for (unsigned int seq = 0;; ++seq) {
        long item;
        if (guest && (long)*reinterpret_cast<volatile unsigned int*>(next_item) >= guest_limit) item = -1;
        else item = (long)atomicAdd(next_item, 1u);
        if (item >= nitems) item = -1;
        if (!first) mbar_wait(&empty[st], ph);  // (also guarantees the consumers are done with item_q[seq & 15])
        item_q[seq & 15] = item;
        if (item < 0) {
          mbar_arrive(&full[st]);  // nothing to copy: release the consumers, which then read the sentinel
          break;
        }
        const int sp = (int)(item / ntiles), k = (int)(item % ntiles);
        const int p0 = sp * rows_per_split, p1 = min(naux, p0 + rows_per_split);
        const double* src = Bt + ((long)p0 * ntiles + k) * TILE_ELEMS;
        for (int p = p0; p < p1; ++p, src += (long)ntiles * TILE_ELEMS) {
          if (p > p0 && !first) mbar_wait(&empty[st], ph);
          mbar_expect_tx(&full[st], TILE_BYTES);
          bulk_g2s_stream(stages + (size_t)st * TILE_ELEMS, src, TILE_BYTES, &full[st], pol);
          if (++st == nst) {
            st = 0;
            ph ^= 1u;
            first = false;
          }
        }
      }
    }
    return;
  }
  const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
  uint32_t st = 0, ph = 0;
  for (unsigned int seq = 0;; ++seq) {
    mbar_wait_a(full_a + 8u * st, ph);  // first tile of the item (or the bare arrival of the sentinel)
    const long item = item_q[seq & 15];
    if (item < 0) break;
    const int sp = (int)(item / ntiles), k = (int)(item % ntiles);
    const int p0 = sp * rows_per_split, p1 = min(naux, p0 + rows_per_split);
    double2 acc[NSET][JP_Q];
#pragma unroll
    for (int s = 0; s < NSET; ++s)
#pragma unroll
      for (int q = 0; q < JP_Q; ++q) acc[s][q] = make_double2(0.0, 0.0);
    for (int p = p0; p < p1; ++p) {
      double r[NSET];
#pragma unroll
      for (int s = 0; s < NSET; ++s) r[s] = __ldcg(rho + (long)s * naux + p);
      if (p > p0) mbar_wait_a(full_a + 8u * st, ph);
      const double2* t2 = reinterpret_cast<const double2*>(stages + (size_t)st * TILE_ELEMS);
      double2 v[JP_Q];
#pragma unroll
      for (int q = 0; q < JP_Q; ++q) v[q] = t2[tid + JP_CONSUMERS * 32 * q];
      if (SAFE) {
        // true for every payload a tensor can hold (the pattern is one particular NaN); the vote needs the high word of
        // every loaded double2 in every lane, i.e. all four LDS of the warp have completed when it executes
        bool landed = true;
#pragma unroll
        for (int q = 0; q < JP_Q; ++q) landed = landed && (__double2hiint(v[q].y) != 0x7ff8dead);
        if (__all_sync(0xffffffffu, landed)) {
          if (lane == 0) mbar_arrive_a(empty_a + 8u * st);
        } else {  // same release; the branch only keeps the arrival control-dependent on the vote
          __nanosleep(20);
          if (lane == 0) mbar_arrive_a(empty_a + 8u * st);
        }
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive_a(empty_a + 8u * st);
      }
#pragma unroll
      for (int s = 0; s < NSET; ++s)
#pragma unroll
        for (int q = 0; q < JP_Q; ++q) {
          acc[s][q].x = fma(r[s], v[q].x, acc[s][q].x);
          acc[s][q].y = fma(r[s], v[q].y, acc[s][q].y);
        }
      if (++st == (uint32_t)nst) {
        st = 0;
        ph ^= 1u;
      }
    }
    // part[split][set][tile k][1024]
#pragma unroll
    for (int s = 0; s < NSET; ++s) {
      double2* o = reinterpret_cast<double2*>(part + (((long)sp * NSET + s) * ntiles + k) * TILE_ELEMS);
#pragma unroll
      for (int q = 0; q < JP_Q; ++q) o[tid + JP_CONSUMERS * 32 * q] = acc[s][q];
    }
  }
}

// J[set][mu][nu] (square, symmetric) from the split partials in tiled layout
__global__ void j_finalize_kernel(const double* __restrict__ part, const int* __restrict__ inv, double* __restrict__ J,
                                  int n, int nb, long E, int nsplit, int nset) {
  const int nu = blockIdx.x * blockDim.x + threadIdx.x;
  const int mu = blockIdx.y;
  if (nu >= n) return;
  const int hi = max(mu, nu), lo = min(mu, nu);
  const int I = hi >> 5, Jt = lo >> 5;
  const long off = (long)__ldcg(inv + I * nb + Jt) * TILE_ELEMS + tile_swz(hi & 31, lo & 31);  // (table changes per tensor)
  for (int s = 0; s < nset; ++s) {
    double v = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) v += __ldcg(part + ((long)sp * nset + s) * E + off);  // (L2: see j_pass note)
    J[((long)s * n + mu) * n + nu] = v;
  }
}

// mirror the lower triangle of `batch` square matrices into the upper triangle
__global__ void symmetrize_lower_kernel(double* __restrict__ A, int n, long stride) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  double* a = A + (long)blockIdx.z * stride;
  if (j < n && j > i) a[(long)i * n + j] = a[(long)j * n + i];
}

}  // namespace nbd
