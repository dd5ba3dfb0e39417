// Shared device helpers for the nbed_b200 kernels (sm_100a only).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nbd {

// ---- FP64 tensor-core tile: mma.sync m8n8k4 (SASS DMMA.8x8x4 on sm_100a) -----------------------
// Fragment ownership (g = lane>>2, t = lane&3):  A[g][t]   B[t][g]   C/D[g][2t], C/D[g][2t+1]
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier + bulk async copy (TMA 1-D; SASS UBLKCP / SYNCS) ---------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar, uint32_t count = 1) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// the same on precomputed shared-space addresses (keeps the generic->shared conversion out of inner loops)
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar, uint32_t count = 1) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// streaming variant: the 3-centre tensor is read once per pass and is far larger than L2, so its lines are marked
// evict-first and do not push the reusable operands (X, orbitals, F') out of the cache
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_stream(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}

__device__ __forceinline__ unsigned int smid() {
  unsigned int r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}

// Spatial split of two co-running grids: exactly `p` of the `nsm` SMs, evenly spread over the SM ids (and with them over
// the GPCs), belong to group 1; the rest to group 0.
__device__ __forceinline__ bool sm_in_group1(unsigned int sm, int p, int nsm) {
  return ((sm + 1u) * (unsigned)p) / (unsigned)nsm != (sm * (unsigned)p) / (unsigned)nsm;
}

// ---- thread-block clusters: barrier + distributed shared memory -----------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster; release / acquire at cluster scope (orders shared-memory writes before
// remote reads) - without the GPU-scope fence and L1 invalidate that cooperative_groups' cluster.sync() adds
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ double dsmem_ld(uint32_t addr) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}

// ---- cp.async (LDGSTS) 8-byte with zero fill ----------------------------------------------------
__device__ __forceinline__ void cp_async8(void* dst, const void* src, bool valid) {
  int sz = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- reductions ---------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-wide deterministic sum; result valid in thread 0. `red` = 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    v = l < nw ? red[l] : 0.0;
    v = warp_sum(v);
  }
  return v;
}

// ---- tiled-triangular layout of the 3-centre tensor ---------------------------------------------
// One auxiliary row P = ntiles tiles of 32x32 doubles (8 KiB) in PROCESSING order;
// inside a tile element (r, c) lives at r*32 + (c ^ (4*(r&3))): an XOR swizzle that makes both DMMA
// fragment patterns (row-major A operand and its transpose) bank-conflict-free without padding.
constexpr int TILE = 32;
constexpr int TILE_ELEMS = TILE * TILE;
constexpr int TILE_BYTES = TILE_ELEMS * 8;
__host__ __device__ __forceinline__ int tile_swz(int r, int c) { return r * TILE + (c ^ ((r & 3) << 2)); }

// Function attributes (dynamic shared-memory opt-in) are per device: a call site remembers the devices it has
// configured in a bit mask instead of one process-wide flag, so a process that drives several GPUs stays correct.
inline bool first_use_on_current_device(unsigned long long& mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;
  const unsigned long long bit = 1ull << dev;
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

}  // namespace nbd
