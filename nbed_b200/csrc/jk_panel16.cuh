// EXPERIMENTAL (option "panel_warps" = 16, off by default; compiled and host-checked in round 1, first GPU runs are
// round-2 work - see DESIGN.md section 8): the pass-1 panel kernel with 16 consumer warps.
//
// The 8-warp kernel (jk.cuh) keeps the FP64 pipe of an SM only ~62 % busy on a 10-column slice: with two consumer
// warps per scheduler there is rarely a second warp ready while the first waits for a tile, a shared-memory load or
// a fixed latency.  Four warps per scheduler need the accumulators to shrink to ~60 registers per thread:
//   * a warp owns the panels I == warp (mod 16): 3 slots at n = 1376 instead of 6;
//   * the FMA-pipe columns no longer keep one partial sum per contraction residue (8 doubles per slot): the four
//     `tq` lanes reduce-scatter their partials at the end of every task (6 shuffles), 2 doubles per slot remain.
// The tile order (= storage order of the tensor) uses 16 x 16 super-blocks so that every run of 16 tiles hands each of
// the 16 warps one row task and one column task; it is chosen when the tensor is allocated.
#pragma once
#include "jk.cuh"

namespace nbd {

// Tile order with nw x nw super-blocks (nw = 8 reproduces build_tile_sequence).
inline std::vector<int> build_tile_sequence_w(int nb, int nw) {
  std::vector<int> seq;
  const int nsb = (nb + nw - 1) / nw;
  for (int a = 0; a < nsb; ++a)
    for (int b = 0; b <= a; ++b)
      for (int k = 0; k < nw; ++k)
        for (int j = 0; j < nw; ++j) {
          const int I = nw * a + (j + k) % nw, J = nw * b + j;
          if (I >= nb || J >= nb || J > I) continue;
          seq.push_back((I << 16) | J);
        }
  return seq;
}

// Per-warp event lists for nw consumer warps (same encoding and invariants as build_panel_plan; begin holds
// nw + 1 offsets followed by the nw first-tile indices).
inline PanelPlan build_panel_plan_w(int nb, int S, const std::vector<int>& seq, int nw) {
  PanelPlan pl;
  pl.nb = nb;
  pl.S = S;
  const int nt = (int)seq.size();
  std::vector<std::vector<int>> ks(nw), flags(nw);
  for (int w = 0; w < nw; ++w) {
    std::vector<int> rk, rf;
    for (int k = 0; k < nt; ++k) {
      const int I = seq[k] >> 16, J = seq[k] & 0xffff;
      const int f = ((I % nw == w) ? 1 : 0) | (((J % nw == w) && I != J) ? 2 : 0);
      if (f) {
        rk.push_back(k);
        rf.push_back(f);
      }
    }
    if (rk.empty()) continue;
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
  for (int w = 0; w < nw; ++w)
    for (int k : ks[w]) ++npart[k];
  pl.begin.assign(2 * nw + 1, 0);
  for (int w = 0; w < nw; ++w) {
    pl.begin[w] = (int)pl.events.size();
    pl.begin[nw + 1 + w] = ks[w].empty() ? 0 : ks[w][0];
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
  pl.begin[nw] = (int)pl.events.size();
  // self-check (a violated invariant would be a device-side hang)
  std::vector<int> wsum(nt, 0);
  for (int w = 0; w < nw; ++w) {
    int k = pl.begin[nw + 1 + w];
    if (pl.begin[w] < pl.begin[w + 1] && k > S - 1) pl.S = -1;
    for (int e = pl.begin[w]; e < pl.begin[w + 1]; ++e) {
      const uint32_t ev = pl.events[e];
      const int gap = (ev >> 16) & 31, weight = (ev >> 21) & 31;
      if (gap < 1 || gap > S || weight < 1 || (int)(ev & 255) != (seq[k] >> 16) || (int)((ev >> 8) & 255) != (seq[k] & 0xffff)) pl.S = -1;
      wsum[k] += weight;
      k = (k + gap) % nt;
    }
    if (pl.begin[w] < pl.begin[w + 1] && k != pl.begin[nw + 1 + w]) pl.S = -1;
  }
  for (int k = 0; k < nt; ++k)
    if (wsum[k] != XK_EMPTY_COUNT) pl.S = -1;
  return pl;
}

constexpr int XK16_CONSUMER_WARPS = 16;
constexpr int XK16_THREADS = (XK16_CONSUMER_WARPS + 4) * 32;
constexpr int XK16_CONSUMER_REGS = 112;  // setmaxnreg only moves registers INSIDE the CTA allocation (640 x 96 at launch): the 4 producer warps release 128 x (96 - 24) = 9216, the 16 consumer warps may take 512 x (112 - 96) = 8192; asking for 120 (12288) blocks forever (the hang of the first GPU run, profiles/r02_panel16.md)
constexpr int XK16_PRODUCER_REGS = 24;

// Sum the four tq lanes' partials pf[mi] (mi = 0..3) so that lane tq ends up with the total of mi = 2 (tq & 1) + (tq >> 1).
__device__ __forceinline__ double xk16_reduce_scatter(const double (&pf)[4], int tq) {
  const bool odd = tq & 1, hi = tq & 2;
  double k0 = odd ? pf[2] : pf[0], s0 = odd ? pf[0] : pf[2];
  double k1 = odd ? pf[3] : pf[1], s1 = odd ? pf[1] : pf[3];
  k0 += __shfl_xor_sync(0xffffffffu, s0, 1);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 1);
  const double keep = hi ? k1 : k0, send = hi ? k0 : k1;
  return keep + __shfl_xor_sync(0xffffffffu, send, 2);
}

// One task = one 32 x 32 tile against 8 DMMA columns + NF FMA columns.  ROW: X[I] += B_IJ C_J (A fragment element
// (row 8 mi + gq, contraction 4 ks + tq)); !ROW: X[J] += B_IJ^T C_I (element (contraction 4 ks + tq, row 8 mi + gq)).
template <int NF, bool ROW>
__device__ __forceinline__ void xk16_task(double (&acc)[4][2], double (&accf)[NF > 0 ? NF : 1], const double* __restrict__ tile,
                                          const double* __restrict__ crow, const double* __restrict__ cf, int ct_ld, int colbase,
                                          int gq, int tq, const int (&off)[4]) {
  constexpr int NFA = NF > 0 ? NF : 1;
  double pf[NFA][4];
#pragma unroll
  for (int fi = 0; fi < NFA; ++fi)
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) pf[fi][mi] = 0.0;
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    double a[4], f[NFA];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      if (ROW) a[mi] = tile[(8 * mi + gq) * 32 + ((ks >> 2) << 4) + off[ks & 3]];
      else a[mi] = tile[(4 * ks + tq) * 32 + off[mi]];
    }
    const double b = crow[colbase + 4 * ks + tq];
#pragma unroll
    for (int fi = 0; fi < NF; ++fi) f[fi] = cf[fi * ct_ld + colbase + 4 * ks + tq];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      dmma(acc[mi], a[mi], b);
#pragma unroll
      for (int fi = 0; fi < NF; ++fi) pf[fi][mi] = fma(a[mi], f[fi], pf[fi][mi]);
    }
  }
#pragma unroll
  for (int fi = 0; fi < NF; ++fi) accf[fi] += xk16_reduce_scatter(pf[fi], tq);
}

// NSLOT = ceil(nb / 16) panels per consumer warp; one slice of exactly 8 + NF columns (NF = 1, 2).
template <int NSLOT, int NF>
__global__ void __launch_bounds__(XK16_THREADS, 1) symm_panel16_kernel(XArgs p) {
  extern __shared__ __align__(128) unsigned char xsm[];
  constexpr int NW = XK16_CONSUMER_WARPS;
  constexpr int NCOL = 8 + NF;
  constexpr int NFA = NF > 0 ? NF : 1;
  const int S = p.nstages;
  uint64_t* full = reinterpret_cast<uint64_t*>(xsm);  // 2S "full" barriers over S data stages (see jk.cuh)
  uint64_t* empty = full + 32;
  const int ct_ld = p.n_ld + 4;
  // layout: [512 B barriers][NCOL rows of Ct][S stage buffers]
  double* cts = reinterpret_cast<double*>(xsm + 512);
  const size_t ct_bytes = ((size_t)NCOL * ct_ld * 8 + 127) & ~(size_t)127;
  double* stages = reinterpret_cast<double*>(xsm + 512 + ct_bytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < 2 * S; ++s) mbar_init(&full[s], 1);
    for (int s = 0; s < S; ++s) mbar_init(&empty[s], XK_EMPTY_COUNT);
    mbar_fence_init();
  }
  __syncthreads();
  const long nitems = p.naux;  // one slice: item = aux row

  if (warp >= NW) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(XK16_PRODUCER_REGS));
    if (warp == NW && lane == 0) {
      int st = 0, fb = 0;
      uint32_t ph = 1;
      bool first_round = true;
      const uint64_t pol = l2_evict_first_policy();
      for (long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const double* src = p.Bt + item * (long)p.ntiles * TILE_ELEMS;
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

  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(XK16_CONSUMER_REGS));
  for (int i = 0; i < NCOL; ++i) {
    double* dst = cts + (size_t)i * ct_ld;
    const double* src = p.Ct + (size_t)i * p.n_ld;
    for (int m = tid; m < p.n_ld; m += NW * 32) dst[m] = src[m];
  }
  asm volatile("bar.sync 1, %0;" ::"r"(NW * 32) : "memory");

  const int gq = lane >> 2, tq = lane & 3;
  int xoff[4], yoff[4];
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) xoff[q4] = 4 * (q4 ^ (gq & 3)) + tq;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) yoff[mi] = (8 * mi + gq) ^ (4 * tq);
  const double* crow = cts + (size_t)gq * ct_ld;
  const double* cf = cts + (size_t)8 * ct_ld;

  double X[NSLOT][4][2];
  double Xf[NSLOT][NFA];
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) X[s][mi][0] = X[s][mi][1] = 0.0;
#pragma unroll
    for (int fi = 0; fi < NFA; ++fi) Xf[s][fi] = 0.0;
  }

  const int e0 = __ldg(p.evbegin + warp), e1 = __ldg(p.evbegin + warp + 1);
  if (e0 == e1) return;  // this warp owns no panel of this matrix size
  const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
  uint32_t fb = (uint32_t)__ldg(p.evbegin + NW + 1 + warp), par = 0;
  uint32_t ev_next = __ldg(p.events + e0);
  const int mi_own = ((tq & 1) << 1) | (tq >> 1);  // row block this lane holds of the FMA columns
  for (long item = blockIdx.x; item < nitems; item += gridDim.x) {
    for (int e = e0; e < e1; ++e) {
      const uint32_t ev = ev_next;
      ev_next = __ldg(p.events + (e + 1 < e1 ? e + 1 : e0));
      const int I = ev & 255, J = (ev >> 8) & 255;
      const uint32_t st = fb >= (uint32_t)S ? fb - (uint32_t)S : fb;
      mbar_wait_a(full_a + 8u * fb, par);
      if (ev & (3u << 26)) {
        const double* tile = stages + (size_t)st * TILE_ELEMS;
        if (ev & (1u << 26)) {
          const int slot = I >> 4;
#pragma unroll
          for (int s = 0; s < NSLOT; ++s)
            if (slot == s) xk16_task<NF, true>(X[s], Xf[s], tile, crow, cf, ct_ld, 32 * J, gq, tq, xoff);
        }
        if (ev & (2u << 26)) {
          const int slot = J >> 4;
#pragma unroll
          for (int s = 0; s < NSLOT; ++s)
            if (slot == s) xk16_task<NF, false>(X[s], Xf[s], tile, crow, cf, ct_ld, 32 * I, gq, tq, yoff);
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
    // write this warp's panels of X[item] and reset the accumulators
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = 2 * tq + r;
      double* xo = p.X + __ldcg(p.xbase + i) + item * __ldcg(p.xstride + i) + gq;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int I = NW * s + warp;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          if (I < p.nb) xo[32 * I + 8 * mi] = X[s][mi][r];
          X[s][mi][r] = 0.0;
        }
      }
    }
#pragma unroll
    for (int fi = 0; fi < NF; ++fi) {
      double* xo = p.X + __ldcg(p.xbase + 8 + fi) + item * __ldcg(p.xstride + 8 + fi) + gq + 8 * mi_own;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int I = NW * s + warp;
        if (I < p.nb) xo[32 * I] = Xf[s][fi];
        Xf[s][fi] = 0.0;
      }
    }
  }
}

}  // namespace nbd
