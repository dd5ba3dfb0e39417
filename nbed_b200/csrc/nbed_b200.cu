// nbed_b200 — C-ABI implementation (see include/nbed_b200.h).  sm_100a only.
//
// Everything numerical runs in the hand-written kernels of jk.cuh / gemm.cuh / scf_kernels.cuh; the only
// library routines on the path are the cuSOLVER eigensolvers the north star names (dsyevd / dsygvd, timed
// separately under the "eigh" key) and NCCL's all-reduce for the aux-sharded partial J/K and MO integrals.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <functional>
#include <future>
#include <thread>

#include "common.cuh"
#include "gemm.cuh"
#include "syrk.cuh"
#include "host_util.cuh"
#include "jk.cuh"
#include "scf_kernels.cuh"
#include "integrals.cuh"
#include "subspace.cuh"
#include "small_eigh.cuh"

using namespace nbd;

// ------------------------------------------------------------------------------------------------
// NCCL, bound at run time (dlopen) so that the library loads on hosts without it.
// ------------------------------------------------------------------------------------------------
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string& why) {
    if (lib) return true;
    const char* env = getenv("NBD_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) {
      why = "cannot dlopen libnccl.so.2 (set NBD_NCCL_LIB, or import torch first)";
      return false;
    }
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    Broadcast = (decltype(Broadcast))dlsym(lib, "ncclBroadcast");
    GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
    AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
    if (!AllGather || !GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy || !GetErrorString || !Broadcast || !GroupStart || !GroupEnd) {
      why = "libnccl is missing required symbols";
      return false;
    }
    return true;
  }
};
static NcclApi g_nccl;

#include "host_linalg.h"

// ------------------------------------------------------------------------------------------------
// Context
// ------------------------------------------------------------------------------------------------
struct DiisState {
  // pyscf/lib/diis.py:DIIS bookkeeping (space = 6 for the Huzinaga loop, 8 for CDIIS on the mu path)
  int space = 6;
  int head = 0;
  std::vector<int> bookkeep;
  std::vector<double> H;  // (space+1)^2
  bool have_xprev = false;
  bool err_mode = false;  // true: error vectors are pushed explicitly (CDIIS)
  DBuf<double> x, e, xprev, xnew;
  long len = 0;
  void init(int space_, long len_, bool err_mode_) {
    space = space_;
    len = len_;
    err_mode = err_mode_;
    head = 0;
    bookkeep.clear();
    H.assign((size_t)(space + 1) * (space + 1), 0.0);
    for (int i = 1; i <= space; ++i) H[i] = H[(size_t)i * (space + 1)] = 1.0;
    have_xprev = false;
    x.ensure((size_t)space * len);
    e.ensure((size_t)space * len);
    xprev.ensure(len);
    xnew.ensure(len);
  }
};

void nbd_destroy_blas(struct cublasContext* h);  // defined next to the cuBLAS include (integrals_host.cuh)

struct PlanDev {
  DBuf<uint32_t> events;
  DBuf<int> begin;
};

struct KGroup {
  int set;       // which K matrix
  int c0, c1;    // columns [c0, c1) of the orbital block
  double alpha;  // +1 / -1 (signed eigen-factors of a dense density)
};
// state of a Huzinaga loop between iterations: the staged orbital block of the next J/K
struct HuzLoop {
  int Ntot = 0;
  std::vector<KGroup> groups;
  bool veff_valid = false;  // Kohn-Sham loop: F / vhf already hold get_veff of the current density
};

// exchange-correlation stage (xc.cuh / xc_host.cuh): resident AO values on the caller's grid + workspaces
struct XcState {
  bool ready = false, on = false;
  int code = 0, ng = 0, ng_user = 0;
  double hyb = 0.0;
  double ecoul = 0.0, exc = 0.0;  // of the last Kohn-Sham get_veff (the .ecoul / .exc tags the reference reads)
  DBuf<IntShell> shells;
  DBuf<double> env, c2s, coords, w, ao, rho, grad, sigma, fx, TM, Vpart, V;
};

struct nbd_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cusolverDnHandle_t solver = nullptr;
  struct cublasContext* blas = nullptr;  // only for the one-time triangular solve of the Cholesky decoration (integrals_host.cuh)
  // side stream (+ its own cuSOLVER handle): runs the HBM-bound J pass next to the tensor-bound K Gram, and the
  // second spin's eigensolve next to the first
  cudaStream_t stream2 = nullptr;
  cusolverDnHandle_t solver2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_stage[4] = {nullptr, nullptr, nullptr, nullptr};  // [0,1] lane 0 (main stream), [2,3] lane 1 (copy stream)
  cudaStream_t stream3 = nullptr;  // copy stream of the early result export (nbd_huzinaga_scf)
  cudaEvent_t ev_export = nullptr;
  PinnedBuf pinned2;
  int overlap = 1;
  int small_eigh_warm = 1;  // ... started from the previous SCF cycle's eigenvectors
  DBuf<double> seWarm;
  int se_warm_n = 0, se_warm_batch = 0;
  long se_warm_calls = 0;  // solves since the warm basis was last reset (0 = none valid)
  int small_eigh = 1;   // n <= 32: one-CTA Jacobi eigensolver (small_eigh.cuh) instead of cuSOLVER dsyevd
  int eig_threads = 1;  // second spin's full eigensolve issued from a helper thread on the side stream
  int dist_eig = 1;
  int ks_energy = 0;  // mu / kernel() path with XC on: 1 = pyscf's KS energy_elec (e1 + ecoul + exc), 0 = nbed's patched energy_elec (e1 + tr(vhf D) / 2)
  int dist_orth = 1;  // multi-rank: every rank forms its row block of F' = X F X, one all-gather assembles it
  int gemm_tile = 0;      // tuning: force the GEMM tile size (0 = heuristic)
  int jpass_sm_mod = 0, jpass_sm_keep = 0;  // pass 2 only on SMs with smid % mod < keep (0 = everywhere)
  int jpass_ctas_per_sm = 2;
  int x_cache = 1;        // reuse X = S^-1/2 across nbd_scf_setup calls with a bit-identical overlap matrix
  int x_valid_n = 0;      // nao the cached S / X belong to (0 = none)
  int early_export = 1;   // nbd_huzinaga_scf copies D / Huz to the host on a copy stream while the final eigensolve runs
  int jpass_safe = 1;     // 0: pass 2 releases a ring stage before its loads are known to have returned (experiments)
  int syrk_mode = 1;      // 1: the K Gram runs as the stream-K symmetric rank-k kernel (syrk.cuh) when the shape allows
  int syrk_ctas = 0;      // CTAs of that kernel (0 = one per SM)
  DBuf<double> syrk_ws;   // its partial tiles
  DBuf<unsigned int> d_syrk_counter;
  int pair_guest_reserve = 30;  // guests leave the last 1 / pair_guest_reserve of the item queue alone (3 / 6 / 12 / 30 measured: 5.85 / 5.83 / 5.75 / 5.73 ms)
  int pair_guest = 1;     // pass-2 CTAs that stay on every Gram SM of the split pair (0 / 1 / 2 measured: pass 2 6.04-6.10 / 5.73-5.83 / 5.81 ms, profiles/r02_pair_split.md)
  int pair_split = -1;    // > 0: pass 2 runs on this many SMs (3 CTAs each), the stream-K Gram on the others; 0: both everywhere; -1: automatic
  int jpass_variant = 0;  // 0: TMA-fed persistent pass 2, 1: LDG streaming pass 2
  int panel_stages = 0;  // tuning: cap on the ring depth of the panel kernel (0 = as many as fit)
  int panel_hybrid = 1;  // 9-10 trailing orbital columns: 8 on DMMA + 1-2 on the FMA pipe (0 = pad to 16 DMMA columns)
  std::string err;
  long launches = 0;
  StageTimers timers;
  int jk_variant = 0, gemm_variant = 0;
  long x_budget_bytes = 3L << 30;  // workspace bound for the half-transformed tensor (aux-chunked)
  int sm_count = 148;
  size_t smem_optin = 0;

  // ---- NCCL ----
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;

  // ---- 3-centre tensor ----
  int nao = 0, nb = 0, n_ld = 0, ntiles = 0, naux = 0;
  long npair = 0;
  double* Bt = nullptr;
  std::vector<int> seq, inv;
  DBuf<int> d_seq, d_inv;
  DBuf<double> stage;  // packed-row staging for upload / download
  PinnedBuf pinned;
  PinnedBuf rb;  // small read-backs of the SCF loop (subspace_host.cuh: d2h_small)

  // ---- J/K workspaces ----
  DBuf<double> d_jkpack;  // packed lower triangles of [J | K sets] (multi-GPU all-reduce buffer of the SCF loop)
  int packed_allreduce = 1;
  DBuf<double> d_orb, d_wt, d_X, d_rho, d_jpart, d_jk;  // d_jk = [J sets | K sets] contiguous (all-reduce buffer)
  DBuf<int> d_setbegin;
  std::map<int, PlanDev> plans;  // device copies of the panel kernel's task lists, keyed by ring depth
  DBuf<unsigned int> d_jcounter;
  DBuf<long> d_xtab;  // group-major layout tables of the half-transformed tensor: [xbase | xstride]

  // ---- SCF problem ----
  bool scf_ready = false;
  bool have_virt = false;  // virtual-orbital Huzinaga projector present (gamma_virt S in GSv)
  int nspin = 0, projector = 0;
  int nelec[2] = {0, 0};
  double mu = 0.0;
  DBuf<double> S, Xh, hcore, heff, GS, GSv, F, Huz, vhf, FG, T1, T2, Ct, D, Dold, evals, eigwork, eigwork2, red_part, red_out,
      Corth, Ssave, Ssave2, dm0f, Fpad;
  DBuf<int> devinfo;
  DiisState diis;
  // subspace (Chebyshev-filtered) tracking of the occupied block between full eigensolves
  int eig_mode = 1;  // 0: cuSOLVER every cycle; 1: filtered subspace iteration when eligible (cuSOLVER first / last / fallback)
  bool sub_valid = false;
  int sub_min_nao = 128;  // below this the library eigensolver is cheaper than tracking a 16/32-vector block (C3, n = 174: 538 vs 254 iterations/s, profiles/bench_r02c_C3_*.json)
  bool last_eig_full = true;
  int sub_kb = 0;
  long sub_applies = 0, sub_fallbacks = 0, sub_outer = 0, sub_lanczos = 0, sub_cold_starts = 0;
  double sub_theta[2][32];
  DBuf<double> sV, sY, sZ, sW, sAV, sPart, sG, sGpart, sM, sTheta, sRpart, sBound, sFprev, sLz, sXchg;
  int dist_sub = 1;  // >= 2 ranks, two spins: ranks 0 / 1 track one spin's eigenvector block each, blocks are broadcast
  // spectral bounds of the filter: 0 = Gershgorin every cycle; 1 = Lanczos once, then widened by ||F'_k - F'_{k-1}||_F
  int sub_bound_mode = 1;
  int sub_adaptive = 1;   // filter degree of a tracked block from the observed residual reduction per degree
  double sub_rate = 0.0;  // slowest residual reduction per filter degree seen in the current SCF (0 = none yet)
  int sub_apply_variant = 0;  // 0: automatic (single-shot 8-CTA-cluster kernel when its grid is resident at once, else the 4-CTA-cluster ring kernel); 1: one CTA per 16 rows; 2: single-shot kernel; 3: ring kernel
  int sub_pdl = 1;  // consecutive block products overlap through programmatic dependent launch
  int sub_cold = 1;  // 1: the initial guess starts the block from pseudo-random vectors (no library eigensolve)
  bool sub_bounds_valid = false, sub_is_cold = false;
  double sub_up[2] = {0, 0}, sub_low[2] = {0, 0}, sub_up_ref[2] = {0, 0}, sub_low_ref[2] = {0, 0};
  DBuf<unsigned int> sTicket;
  long sub_rejects = 0;  // SCF runs redone with the library eigensolver because the tracked block failed the final check
  // low-rank form of the occupied environment projector: gamma_s = V_s V_s^T (nbd_scf_set_env_orbitals)
  int env_rank = 0;
  DBuf<double> envV, envZ, envW;  // V rows [ns][r][n];  Z = V^T S [ns][r][n];  W = F V [ns][n][r]
  // bench state
  bool bench_ready = false;
  HuzLoop bench_loop;

  // ---- ao2mo ----
  DBuf<double> mo_c, Lbuf, eri, eri_phys;

  XcState xc;
};

#define LAUNCH_CHECK(ctx)               \
  do {                                  \
    ++(ctx)->launches;                  \
    NBD_CUDA(cudaGetLastError());       \
  } while (0)

static inline dim3 grid1(long cnt, int block) { return dim3((unsigned)((cnt + block - 1) / block)); }

// ------------------------------------------------------------------------------------------------
// GEMM wrappers (row-major views on top of gemm.cuh's stride-generic kernel)
// ------------------------------------------------------------------------------------------------
static void gemm(nbd_ctx* c, int M, int N, int K, const double* A, long a_is, long a_ks, const double* B, long b_js,
                 long b_ks, double* C, long ldc, double alpha, double beta, int batch = 1, long sA = 0, long sB = 0,
                 long sC = 0, int lower = 0, int pdl = 0) {
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.a_is = a_is; g.a_ks = a_ks;
  g.B = B; g.b_js = b_js; g.b_ks = b_ks;
  g.C = C; g.ldc = ldc; g.alpha = alpha; g.beta = beta;
  g.batch = batch; g.strideA = sA; g.strideB = sB; g.strideC = sC; g.lower_only = lower; g.pdl = pdl;
  NBD_CUDA(launch_gemm(c->stream, g, c->gemm_variant, &c->launches, c->sm_count, c->gemm_tile));
}
// C[M][N] = alpha * A[M][K] * B[K][N] + beta * C   (all row-major, leading dimensions given)
static void gemm_nn(nbd_ctx* c, int M, int N, int K, const double* A, long lda, const double* B, long ldb, double* C,
                    long ldc, double alpha = 1.0, double beta = 0.0, int batch = 1, long sA = 0, long sB = 0, long sC = 0) {
  gemm(c, M, N, K, A, lda, 1, B, 1, ldb, C, ldc, alpha, beta, batch, sA, sB, sC);
}
// C = alpha * A^T * B: A stored [K][M]
static void gemm_tn(nbd_ctx* c, int M, int N, int K, const double* A, long lda, const double* B, long ldb, double* C,
                    long ldc, double alpha = 1.0, double beta = 0.0, int batch = 1, long sA = 0, long sB = 0, long sC = 0,
                    int lower = 0) {
  gemm(c, M, N, K, A, 1, lda, B, 1, ldb, C, ldc, alpha, beta, batch, sA, sB, sC, lower);
}
// C = alpha * A * B^T: B stored [N][K]
static void gemm_nt(nbd_ctx* c, int M, int N, int K, const double* A, long lda, const double* B, long ldb, double* C,
                    long ldc, double alpha = 1.0, double beta = 0.0, int batch = 1, long sA = 0, long sB = 0, long sC = 0) {
  gemm(c, M, N, K, A, lda, 1, B, ldb, 1, C, ldc, alpha, beta, batch, sA, sB, sC);
}

static void symmetrize_lower(nbd_ctx* c, double* A, int n, int batch) {
  dim3 g((n + 127) / 128, n, batch);
  symmetrize_lower_kernel<<<g, 128, 0, c->stream>>>(A, n, (long)n * n);
  LAUNCH_CHECK(c);
}

static void all_reduce(nbd_ctx* c, double* buf, size_t count) {
  if (c->world <= 1 || !c->comm) return;
  StageScope ts(c->timers, c->stream, "allreduce");
  ncclResult_t r = g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, c->comm, c->stream);
  if (r != ncclSuccess) fail(NBD_ERR_CUDA, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
}

// ------------------------------------------------------------------------------------------------
// J/K on device-resident orbitals
// ------------------------------------------------------------------------------------------------
template <int NSLOT, int NB, int NF = 0>
static void launch_symm_panel(nbd_ctx* c, const XArgs& a, int grid, size_t smem) {
  auto kern = symm_panel_kernel<NSLOT, NB, NF>;
  NBD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, XK_THREADS, smem, c->stream>>>(a);
  LAUNCH_CHECK(c);
}

// Group-major layout of the half-transformed tensor for a chunk of np aux rows: the columns of group g
// ([c0, c1), width w) form a dense [np * w][n_ld] matrix (row = (P, i)), so the Gram / second half-transform
// GEMMs see a plain strided operand.  Returns the offset of each group; fills the device tables.
struct XLayout {
  std::vector<long> group_base;
  long total = 0;
};
static XLayout make_xlayout(nbd_ctx* c, int np, int Ntot, const std::vector<std::pair<int, int>>& groups) {
  XLayout L;
  std::vector<long> xb(std::max(1, Ntot), 0), xs(std::max(1, Ntot), 0);
  long base = 0;
  int covered = 0;
  for (auto& g : groups) {
    const int w = g.second - g.first;
    NBD_REQUIRE(g.first == covered && w >= 0, NBD_ERR_STATE, "orbital groups must tile the column range");
    L.group_base.push_back(base);
    for (int i = g.first; i < g.second; ++i) {
      xb[i] = base + (long)(i - g.first) * c->n_ld;
      xs[i] = (long)w * c->n_ld;
    }
    base += (long)np * w * c->n_ld;
    covered = g.second;
  }
  NBD_REQUIRE(covered == Ntot, NBD_ERR_STATE, "orbital groups cover %d of %d columns", covered, Ntot);
  L.total = base;
  long* d = c->d_xtab.ensure((size_t)2 * std::max(1, Ntot));
  NBD_CUDA(cudaMemcpyAsync(d, xb.data(), sizeof(long) * Ntot, cudaMemcpyHostToDevice, c->stream));
  NBD_CUDA(cudaMemcpyAsync(d + Ntot, xs.data(), sizeof(long) * Ntot, cudaMemcpyHostToDevice, c->stream));
  return L;
}

// X (group-major, tables in c->d_xtab) for aux rows [p0, p0+np) and all Ntot orbital rows of d_orb ([Ntot][n_ld]).
// One launch of the panel kernel on the orbital columns [col_begin, col_begin + ncols) of d_orb.
static void half_transform_cols(nbd_ctx* c, int p0, int np, const double* d_orb, int Ntot, int col_begin, int ncols,
                                double* d_X, int nb_force, int nf = 0) {
  const long* xbase = c->d_xtab.p + col_begin;
  const long* xstride = c->d_xtab.p + Ntot + col_begin;
  const int nslot = (c->nb + 7) / 8;
  // a ragged last slice reads its padded columns from a zero row behind the slice; exact slices need none
  auto zrow_for = [&](int nbsel) { return (nf > 0 || ncols % (8 * nbsel) == 0) ? 0 : 1; };
  auto smem_for = [&](int ncolmax, int stages, int zrow) {
    const size_t ct = (((size_t)(ncolmax + zrow) * (c->n_ld + 4) * 8) + 127) & ~(size_t)127;
    return (size_t)512 + ct + (size_t)stages * TILE_BYTES;
  };
  int NBsel = nb_force ? nb_force : ((ncols > 8 && nslot <= 6) ? 2 : 1);
  if (NBsel == 2 && smem_for(std::min(16, ncols), 6, zrow_for(2)) > c->smem_optin) NBsel = 1;
  if (nf > 0) NBD_REQUIRE(NBsel == 1 && ncols == 8 + nf && nf <= 2 && nslot <= 6, NBD_ERR_STATE, "bad hybrid panel launch");
  const int ncolmax = nf > 0 ? ncols : std::min(8 * NBsel, ncols);
  const int zrow = zrow_for(NBsel);
  NBD_REQUIRE(smem_for(ncolmax, 2, zrow) <= c->smem_optin, NBD_ERR_UNSUPPORTED, "nao = %d: orbital slice does not fit shared memory", c->nao);
  int stages = c->panel_stages > 0 ? std::min(16, std::max(2, c->panel_stages)) : 16;
  while (stages > 2 && smem_for(ncolmax, stages, zrow) > c->smem_optin) --stages;
  PlanDev& pd = c->plans[stages];  // per-warp task lists for this (matrix size, ring depth); cleared on re-allocation
  if (pd.events.p == nullptr) {
    const PanelPlan plan = build_panel_plan(c->nb, stages, c->seq);
    NBD_REQUIRE(plan.S == stages, NBD_ERR_STATE, "panel task lists failed their self-check (nb = %d, stages = %d)", c->nb, stages);
    uint32_t* de = pd.events.ensure(std::max<size_t>(1, plan.events.size()));
    int* db = pd.begin.ensure(plan.begin.size());
    NBD_CUDA(cudaMemcpyAsync(de, plan.events.data(), sizeof(uint32_t) * plan.events.size(), cudaMemcpyHostToDevice, c->stream));
    NBD_CUDA(cudaMemcpyAsync(db, plan.begin.data(), sizeof(int) * plan.begin.size(), cudaMemcpyHostToDevice, c->stream));
    NBD_CUDA(cudaStreamSynchronize(c->stream));
  }
  XArgs a{};
  a.events = pd.events.p;
  a.evbegin = pd.begin.p;
  a.Bt = c->Bt + (long)p0 * c->ntiles * TILE_ELEMS;
  a.seq = c->d_seq.p;
  a.Ct = d_orb + (long)col_begin * c->n_ld;
  a.X = d_X;
  a.xbase = xbase;
  a.xstride = xstride;
  a.naux = np;
  a.ntiles = c->ntiles;
  a.nb = c->nb;
  a.n_ld = c->n_ld;
  a.Ntot = ncols;
  a.nslices = nf > 0 ? 1 : (ncols + 8 * NBsel - 1) / (8 * NBsel);
  a.nstages = stages;
  a.ncolmax = ncolmax;
  a.zrow = zrow;
  const long nitems = (long)np * a.nslices;
  long grid = std::min<long>(nitems, (long)c->sm_count);
  // gridDim.x must be a multiple of nslices so that a CTA keeps the same orbital slice for all its items
  grid = std::max<long>(a.nslices, grid / a.nslices * a.nslices);
  const size_t smem = smem_for(ncolmax, stages, zrow);
#define NBD_XK(NS, NBB) launch_symm_panel<NS, NBB>(c, a, (int)grid, smem)
#define NBD_XKF(NS, NFF) launch_symm_panel<NS, 1, NFF>(c, a, (int)grid, smem)
  if (nf == 1) {
    if (nslot <= 1) NBD_XKF(1, 1);
    else if (nslot <= 2) NBD_XKF(2, 1);
    else if (nslot <= 4) NBD_XKF(4, 1);
    else NBD_XKF(6, 1);
  } else if (nf == 2) {
    if (nslot <= 1) NBD_XKF(1, 2);
    else if (nslot <= 2) NBD_XKF(2, 2);
    else if (nslot <= 4) NBD_XKF(4, 2);
    else NBD_XKF(6, 2);
  } else if (NBsel == 2) {
    if (nslot <= 1) NBD_XK(1, 2);
    else if (nslot <= 2) NBD_XK(2, 2);
    else if (nslot <= 4) NBD_XK(4, 2);
    else NBD_XK(6, 2);
  } else {
    if (nslot <= 1) NBD_XK(1, 1);
    else if (nslot <= 2) NBD_XK(2, 1);
    else if (nslot <= 4) NBD_XK(4, 1);
    else if (nslot <= 6) NBD_XK(6, 1);
    else if (nslot <= 8) NBD_XK(8, 1);
    else NBD_XK(12, 1);
  }
#undef NBD_XK
#undef NBD_XKF
}

// X (group-major, tables in c->d_xtab) for aux rows [p0, p0+np) and all Ntot orbital rows of d_orb ([Ntot][n_ld]).
static void half_transform(nbd_ctx* c, int p0, int np, const double* d_orb, int Ntot, double* d_X) {
  if (c->jk_variant == 1) {
    dim3 g(c->nb, np);
    symm_panel_simple_kernel<<<g, 256, 0, c->stream>>>(c->Bt + (long)p0 * c->ntiles * TILE_ELEMS, c->d_inv.p, d_orb,
                                                        d_X, c->d_xtab.p, c->d_xtab.p + Ntot, c->ntiles, c->nb, c->n_ld, Ntot);
    LAUNCH_CHECK(c);
    return;
  }
  const int nslot = (c->nb + 7) / 8;
  NBD_REQUIRE(nslot <= 12, NBD_ERR_UNSUPPORTED, "nao = %d exceeds the 3072-AO envelope of the panel kernel", c->nao);
  // 16-column slices are DMMA-bound on padded work, 8-column slices HBM-bound: a trailing remainder of <= 8 columns
  // behind at least one full 16-column slice goes out as its own 8-column launch (40 columns cost 40, not 48)
  const int rem = Ntot % 16;
  // 9 or 10 trailing columns: 8 on the DMMA path + 1-2 on the FMA pipe in the same launch (cost 9-10, not 16)
  if (c->panel_hybrid && nslot <= 6 && (rem == 9 || rem == 10)) {
    if (Ntot > rem) half_transform_cols(c, p0, np, d_orb, Ntot, 0, Ntot - rem, d_X, 2);
    half_transform_cols(c, p0, np, d_orb, Ntot, Ntot - rem, rem, d_X, 1, rem - 8);
    return;
  }
  if (nslot <= 6 && Ntot > 16 && rem > 0 && rem <= 8) {
    half_transform_cols(c, p0, np, d_orb, Ntot, 0, Ntot - rem, d_X, 2);
    half_transform_cols(c, p0, np, d_orb, Ntot, Ntot - rem, rem, d_X, 1);
    return;
  }
  half_transform_cols(c, p0, np, d_orb, Ntot, 0, Ntot, d_X, 0);
}

// d_orb [Ntot][n_ld] (scaled orbitals), d_wt [Ntot][n_ld] (= sign * orbitals).
// J sets: jbegin[0..njset] column boundaries -> d_J [njset][n][n];  K groups -> d_K [nkset][n][n].
// Either output may be null.  Results are all-reduced over the communicator when both live in c->d_jk.
static void jk_device(nbd_ctx* c, const double* d_orb, const double* d_wt, int Ntot, int njset,
                      const std::vector<int>& jbegin, double* d_J, int nkset, const std::vector<KGroup>& kgroups,
                      double* d_K) {
  NBD_REQUIRE(c->Bt != nullptr, NBD_ERR_STATE, "J/K requested before the 3-centre tensor was allocated");
  const int n = c->nao, n_ld = c->n_ld, naux = c->naux;
  const long nn = (long)n * n;
  StageScope ts_total(c->timers, c->stream, "jk_total");
  if (d_K) NBD_CUDA(cudaMemsetAsync(d_K, 0, sizeof(double) * nn * nkset, c->stream));
  if (d_J) NBD_CUDA(cudaMemsetAsync(d_J, 0, sizeof(double) * nn * njset, c->stream));
  if (Ntot == 0 || naux == 0) return;
  // aux chunking bounds the half-transformed tensor
  const long per_row = (long)Ntot * n_ld * 8;
  int chunk = (int)std::max<long>(1, std::min<long>(naux, c->x_budget_bytes / per_row));
  double* X = c->d_X.ensure((size_t)chunk * Ntot * n_ld);
  double* rho = c->d_rho.ensure((size_t)std::max(1, njset) * naux);
  if (d_J) {
    c->d_setbegin.ensure(njset + 1);
    NBD_CUDA(cudaMemcpyAsync(c->d_setbegin.p, jbegin.data(), sizeof(int) * (njset + 1), cudaMemcpyHostToDevice, c->stream));
  }
  std::vector<bool> k_started(nkset, false);
  // column groups of the layout: the K groups when K is wanted (they tile the columns), else one group
  std::vector<std::pair<int, int>> cols;
  if (d_K) for (auto& g : kgroups) cols.push_back({g.c0, g.c1});
  else cols.push_back({0, Ntot});
  // `behind` (optional) is issued directly behind the pass-2 kernel, before the partials are reduced: the K Gram of
  // the programmatic-launch mode below
  // split_pair: the pair [pass 2 || stream-K Gram] of the last chunk runs on disjoint sets of SMs (option "pair_split" =
  // number of pass-2 SMs).  Sharing SMs costs more than it gives: next to the Gram's 8 DMMA warps the pass-2 CTAs of an
  // SM run at a quarter of their speed (6.9 ms for the pair, 4.6 + 3.0 ms alone), while 74 SMs with three pass-2 CTAs
  // each already pull 6.05 TB/s (profiles/r02_pair_split.md).
  bool split_pair = false;
  int pair_p2 = 0;
  auto j_pass = [&](cudaStream_t st, const std::function<void()>& behind = nullptr) {
    StageScope ts(c->timers, st, "jk_j");
    const long E = (long)c->ntiles * TILE_ELEMS, E2 = E / 2;
    dim3 gf((n + 127) / 128, n);
    if (c->jpass_variant == 0) {
      // TMA-fed persistent kernel: items = (P-range, tile), drawn dynamically; ~16 items per CTA keep the tail short
      const int p2 = split_pair ? std::min(pair_p2, c->sm_count - 1) : 0;
      const int sm_mod = p2 > 0 ? c->sm_count : c->jpass_sm_mod, sm_keep = p2 > 0 ? p2 : c->jpass_sm_keep;
      const int ctas = (p2 > 0 ? 3 : std::max(1, c->jpass_ctas_per_sm)) * c->sm_count;
      int nsplit = (int)std::max<long>(1, std::min<long>(std::min(naux, 16), (16L * ctas + c->ntiles - 1) / c->ntiles));
      const int rows_per_split = (naux + nsplit - 1) / nsplit;
      nsplit = (naux + rows_per_split - 1) / rows_per_split;
      unsigned int* counter = c->d_jcounter.ensure(4 + 512);  // [0] item queue, [1 + smid] guest CTAs per SM
      const int nst = p2 > 0 ? JP_STAGES_ALONE : JP_STAGES;
      const int guests = p2 > 0 ? c->pair_guest : 0;
      const size_t smem = JP_HEADER + (size_t)nst * TILE_BYTES, smem_max = JP_HEADER + (size_t)JP_MAX_STAGES * TILE_BYTES;
      static unsigned long long configured = 0;
      if (first_use_on_current_device(configured)) {
        NBD_CUDA(cudaFuncSetAttribute(j_pass_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        NBD_CUDA(cudaFuncSetAttribute(j_pass_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        // pass 2 shares its SMs with the K Gram: both ask for the largest shared-memory carveout
        NBD_CUDA(cudaFuncSetAttribute(j_pass_tma_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        NBD_CUDA(cudaFuncSetAttribute(j_pass_tma_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      }
      for (int s0 = 0; s0 < njset; s0 += 2) {
        const int ns = std::min(2, njset - s0);
        double* part = c->d_jpart.ensure((size_t)nsplit * ns * E);
        const int grid = (int)std::min<long>((long)c->ntiles * nsplit, ctas);
        NBD_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int) * (4 + 512), st));
        // guests leave the last sixth of the queue to the CTAs that own their SM
        const long glimit = (long)c->ntiles * nsplit - (long)c->ntiles * nsplit / std::max(2, c->pair_guest_reserve);
        if (!c->jpass_safe) {  // legacy release order (experiments only)
          NBD_CUDA(cudaFuncSetAttribute(j_pass_tma_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
          NBD_CUDA(cudaFuncSetAttribute(j_pass_tma_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
          if (ns == 2)
            j_pass_tma_kernel<2, false><<<grid, JP_THREADS, smem, st>>>(c->Bt, rho + (long)s0 * naux, part, c->ntiles, naux, nsplit, rows_per_split, counter, sm_mod, sm_keep, nst);
          else
            j_pass_tma_kernel<1, false><<<grid, JP_THREADS, smem, st>>>(c->Bt, rho + (long)s0 * naux, part, c->ntiles, naux, nsplit, rows_per_split, counter, sm_mod, sm_keep, nst);
        } else if (ns == 2)
          j_pass_tma_kernel<2><<<grid, JP_THREADS, smem, st>>>(c->Bt, rho + (long)s0 * naux, part, c->ntiles, naux, nsplit, rows_per_split, counter, sm_mod, sm_keep, nst, guests, glimit);
        else
          j_pass_tma_kernel<1><<<grid, JP_THREADS, smem, st>>>(c->Bt, rho + (long)s0 * naux, part, c->ntiles, naux, nsplit, rows_per_split, counter, sm_mod, sm_keep, nst, guests, glimit);
        LAUNCH_CHECK(c);
        if (behind && s0 == 0) behind();
        j_finalize_kernel<<<gf, 128, 0, st>>>(part, c->d_inv.p, d_J + (long)s0 * nn, n, c->nb, E, nsplit, ns);
        LAUNCH_CHECK(c);
      }
      return;
    }
    const int bx = (int)((E2 + 255) / 256);
    int nsplit = (int)std::min<long>(std::min(naux, 64), std::max<long>(1, ((long)c->sm_count * 16 + bx - 1) / bx));
    const int rows_per_split = (naux + nsplit - 1) / nsplit;
    nsplit = (naux + rows_per_split - 1) / rows_per_split;
    for (int s0 = 0; s0 < njset; s0 += 2) {
      const int ns = std::min(2, njset - s0);
      double* part = c->d_jpart.ensure((size_t)nsplit * ns * E);
      dim3 g(bx, nsplit);
      if (ns == 2)
        j_pass_kernel<2><<<g, 256, 0, st>>>((const double2*)c->Bt, rho + (long)s0 * naux, (double2*)part, E2, naux, rows_per_split);
      else
        j_pass_kernel<1><<<g, 256, 0, st>>>((const double2*)c->Bt, rho + (long)s0 * naux, (double2*)part, E2, naux, rows_per_split);
      LAUNCH_CHECK(c);
      j_finalize_kernel<<<gf, 128, 0, st>>>(part, c->d_inv.p, d_J + (long)s0 * nn, n, c->nb, E, nsplit, ns);
      LAUNCH_CHECK(c);
    }
  };
  bool forked = false, pair_done = false;
  for (int p0 = 0; p0 < naux; p0 += chunk) {
    const int np = std::min(chunk, naux - p0);
    const bool last = p0 + np >= naux;
    const XLayout L = make_xlayout(c, np, Ntot, cols);
    {
      StageScope ts(c->timers, c->stream, "jk_x");
      half_transform(c, p0, np, d_orb, Ntot, X);
    }
    if (d_J) {
      StageScope ts(c->timers, c->stream, "jk_rho");
      rho_kernel<<<np, 256, 0, c->stream>>>(X, c->d_xtab.p, c->d_xtab.p + Ntot, d_wt, rho + p0, naux, n_ld, njset, c->d_setbegin.p);
      LAUNCH_CHECK(c);
    }
    const bool fork = d_J && last && d_K && c->overlap;
    {
      bool all_syrk = c->syrk_mode && c->gemm_variant == 0 && !kgroups.empty();
      for (size_t gi = 0; gi < kgroups.size() && all_syrk; ++gi) {
        const int w = kgroups[gi].c1 - kgroups[gi].c0;
        all_syrk = w > 0 && syrk_applicable(X + L.group_base[gi], n_ld, (long)np * w * n_ld, n, np * w, 1);
      }
      split_pair = fork && (c->overlap == 1 || c->overlap == 2) && c->pair_split != 0 && c->jpass_variant == 0 &&
                   c->jpass_safe && njset <= 2 && all_syrk;
      pair_p2 = c->pair_split;
      if (split_pair && c->pair_split < 0) {
        // Automatic split from the two kernels' own rates (measured on B200, profiles/r02_pair_split.md): an SM that
        // runs only pass 2 streams ~90 GB/s; the Gram runs at ~0.8 of the FP64 tensor peak and loses ~5 % on a
        // subset of the SMs.  p2 balances  bytes / (90 GB/s p2)  against  t_gram nsm / (nsm - p2).  When the Gram
        // dominates (more than 1.5 x the HBM time of pass 2, e.g. 20 occupied orbitals per spin) sharing is better.
        const double bytes = (double)np * c->ntiles * TILE_BYTES;  // one pass over the tiled tensor, whatever njset
        double flops = 0.0;
        const double nt = (n + SY_T - 1) / SY_T;
        for (auto& g : kgroups) flops += nt * (nt + 1) / 2 * SY_T * SY_T * 2.0 * np * (g.c1 - g.c0);
        const double t2 = bytes / 6.9e12, tg = flops / (0.8 * 37.2e12) * 148.0 / c->sm_count;
        if (tg > 1.5 * t2) {
          split_pair = false;
        } else {
          const double a = bytes / 90e9, b = 1.055 * tg * c->sm_count;
          pair_p2 = std::max(c->sm_count / 4, std::min(c->sm_count * 3 / 5, (int)(0.94 * a * c->sm_count / (a + b) + 0.5)));
        }
      }
    }
    if (fork) {
      // pass 2 (HBM-bound, no tensor work) runs on the side stream next to the tensor-bound K Gram of this chunk.
      // Launch order (option "overlap": 1 = pass 2 first, 2 = Gram first) made no measurable difference; with the
      // Gram alone taking 3.3 ms and pass 2 alone 4.5 ms the pair takes 6.3-6.5 ms (7.8 ms back to back).
      NBD_CUDA(cudaEventRecord(c->ev_fork, c->stream));
      if (c->overlap == 1) {
        NBD_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
        j_pass(c->stream2);
        NBD_CUDA(cudaEventRecord(c->ev_join, c->stream2));
        forked = true;
      }
    }
    // K_s (+)= alpha * X_g^T X_g with X_g the dense [np * w][n_ld] matrix of group g; consecutive groups of
    // identical width / sign that feed consecutive sets go out as one batched launch
    auto k_gram = [&](bool pdl_first) {
      size_t gi = 0;
      while (gi < kgroups.size()) {
        const KGroup& g0 = kgroups[gi];
        const int w = g0.c1 - g0.c0;
        size_t gj = gi + 1;
        while (gj < kgroups.size() && kgroups[gj].c1 - kgroups[gj].c0 == w && kgroups[gj].alpha == g0.alpha &&
               kgroups[gj].set == kgroups[gj - 1].set + 1 && k_started[kgroups[gj].set] == k_started[g0.set])
          ++gj;
        const int batch = (int)(gj - gi);
        if (w > 0) {
          const double* Xa = X + L.group_base[gi];
          const long gstride = (long)np * w * n_ld;  // equal-width groups are laid out back to back
          // the stream-K kernel runs alone or on its own SMs; next to pass 2 on shared SMs the 64-tile GEMM is the
          // better neighbour (C4 with 20 occupied orbitals per spin: pair 13.1 ms against 15.3 ms)
          if (c->syrk_mode && !pdl_first && c->gemm_variant == 0 && (!fork || split_pair || c->syrk_mode == 2) &&
              syrk_applicable(Xa, n_ld, gstride, n, np * w, batch)) {
            SyrkArgs sa{};
            sa.X = Xa; sa.ld = n_ld; sa.strideX = gstride;
            sa.C = d_K + (long)g0.set * nn; sa.ldc = n; sa.strideC = nn;
            sa.alpha = g0.alpha; sa.beta = k_started[g0.set] ? 1.0 : 0.0;
            sa.n = n; sa.K = np * w; sa.batch = batch;
            // spatial split of the pair (see j_pass below): the Gram's ranges = its share of the SMs
            const int p2 = split_pair ? std::min(pair_p2, c->sm_count - 1) : 0;
            const int ranges = c->syrk_ctas > 0 ? c->syrk_ctas : c->sm_count - p2;
            sa.part = c->syrk_ws.ensure(syrk_plan(sa, ranges));
            sa.counter = c->d_syrk_counter.ensure(4);
            sa.sm_p2 = p2;
            sa.nsm = c->sm_count;
            NBD_CUDA(launch_syrk(c->stream, sa, p2 > 0 ? c->sm_count : std::min(ranges, 8 * c->sm_count), &c->launches));
          } else {
            gemm(c, n, n, np * w, Xa, 1, n_ld, Xa, 1, n_ld, d_K + (long)g0.set * nn, n, g0.alpha,
                 k_started[g0.set] ? 1.0 : 0.0, batch, gstride, gstride, nn, /*lower=*/1, /*pdl=*/pdl_first ? 1 : 0);
          }
          pdl_first = false;
          for (size_t q = gi; q < gj; ++q) k_started[kgroups[q].set] = true;
        }
        gi = gj;
      }
    };
    if (fork && c->overlap == 3 && c->jpass_variant == 0 && njset <= 2) {
      // Option "overlap" = 3 (experiment, not the default): programmatic dependent launch.  Pass 2 (2 CTAs per SM, all
      // resident at once) signals launch_dependents as soon as its CTAs are up, and the Gram - launched behind it in
      // the SAME stream with programmatic serialization - gets what is left of each SM (one 57 KiB CTA).  Measured on
      // B200 (tools/pdl_probe.cu shows the mechanism itself overlaps two grids): the pair then takes 7.9 ms, the same
      // as back to back - with pass 2 at full speed a lone Gram CTA per SM makes next to no progress, i.e. the two
      // kernels compete for the L2 -> SM path (31.7 GB + 10.7 GB at ~8 TB/s is a 5.2 ms floor for the pair) and the
      // placement only decides who starves.  The two-stream mode (6.5 ms) stays the default.
      // No event may sit between the two launches, so the pair is timed as one stage ("jk_j"; "jk_k" reads 0).
      j_pass(c->stream, [&] { k_gram(true); });
      forked = false;
      pair_done = true;
    } else if (d_K) {
      StageScope ts(c->timers, c->stream, "jk_k");
      k_gram(false);
    }
    if (fork && !forked && !pair_done) {
      NBD_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
      j_pass(c->stream2);
      NBD_CUDA(cudaEventRecord(c->ev_join, c->stream2));
      forked = true;
    }
  }
  if (d_J && !forked && !pair_done) j_pass(c->stream);
  if (forked) NBD_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  if (d_K) symmetrize_lower(c, d_K, n, nkset);
}

// ------------------------------------------------------------------------------------------------
// Reductions
// ------------------------------------------------------------------------------------------------
// out[slot] = reduce(a, b) ; mode as reduce_partial_kernel
static void reduce_to(nbd_ctx* c, const double* a, const double* b, long cnt, int mode, int n, double* d_out) {
  double* part = c->red_part.ensure(REDUCE_BLOCKS);
  reduce_partial_kernel<<<REDUCE_BLOCKS, 256, 0, c->stream>>>(a, b, cnt, mode, n, part);
  LAUNCH_CHECK(c);
  reduce_final_kernel<<<1, 256, 0, c->stream>>>(part, REDUCE_BLOCKS, d_out);
  LAUNCH_CHECK(c);
}

// ------------------------------------------------------------------------------------------------
// Eigensolvers (cuSOLVER; timed under "eigh")
// ------------------------------------------------------------------------------------------------
// A [batch][n][n] symmetric (row-major == column-major); on return rows of A = eigenvectors; w ascending.
// With a communicator of >= 2 ranks and two matrices (the two spins), rank 0 solves the first and rank 1 the
// second, and the eigenvectors / eigenvalues are broadcast: the replicated eigensolve is the Amdahl term of the
// sharded iteration, this halves it (and makes the orbitals bit-identical on all ranks).
static bool eig_distributed(const nbd_ctx* c, int batch) { return c->world >= 2 && c->comm && batch == 2 && c->dist_eig; }
static void eig_exchange(nbd_ctx* c, double* A, double* w, int n) {
  const long nn = (long)n * n;
  StageScope ts(c->timers, c->stream, "eig_bcast");
  ncclResult_t r = g_nccl.GroupStart();
  for (int b = 0; b < 2 && r == ncclSuccess; ++b) {
    r = g_nccl.Broadcast(A + b * nn, A + b * nn, (size_t)nn, ncclDouble, b, c->comm, c->stream);
    if (r == ncclSuccess) r = g_nccl.Broadcast(w + (long)b * n, w + (long)b * n, (size_t)n, ncclDouble, b, c->comm, c->stream);
  }
  if (r == ncclSuccess) r = g_nccl.GroupEnd();
  if (r != ncclSuccess) fail(NBD_ERR_CUDA, "ncclBroadcast (eigenvectors): %s", g_nccl.GetErrorString(r));
}

// warm_slot = 1: the Fock matrix of the SCF loop - the eigenvectors of the previous cycle (kept in c->seWarm while
// c->se_warm_calls > 0) are the starting basis of the small-matrix Jacobi solver; every 16th solve starts from the
// identity again so that the accumulated rotations cannot drift away from orthogonality.
static void eigh_batched(nbd_ctx* c, double* A, double* w, int n, int batch, int warm_slot = 0) {
  if (n <= SE_MAX_N && c->small_eigh) {
    // one launch per batch instead of cuSOLVER's chain of tiny ones (0.1-0.2 ms per matrix at n = 7 / 24)
    int* info = c->devinfo.ensure(8);
    NBD_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * 8, c->stream));
    StageScope ts(c->timers, c->stream, "eigh");
    double* warm = nullptr;
    int use = 0;
    if (warm_slot == 1 && c->small_eigh_warm && batch <= 2) {
      warm = c->seWarm.ensure((size_t)2 * SE_MAX_N * SE_MAX_N);
      use = (c->se_warm_n == n && c->se_warm_batch == batch && c->se_warm_calls > 0 && c->se_warm_calls % 16 != 0) ? 1 : 0;
      if (c->se_warm_n != n || c->se_warm_batch != batch) c->se_warm_calls = 0;
      c->se_warm_n = n;
      c->se_warm_batch = batch;
      ++c->se_warm_calls;
    }
    small_eigh_kernel<<<batch, SE_THREADS, 0, c->stream>>>(A, w, n, warm, use);
    LAUNCH_CHECK(c);
    return;
  }
  int lwork = 0;
  // numpy.linalg.eigh reads the lower triangle of the row-major matrix == the UPPER triangle column-major
  NBD_SOLVER(cusolverDnDsyevd_bufferSize(c->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, A, n, w, &lwork));
  double* work = c->eigwork.ensure((size_t)lwork);
  int* info = c->devinfo.ensure(8);
  NBD_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * 8, c->stream));
  const bool dist = eig_distributed(c, batch);
  // single rank, two spins: the second solve is issued from a helper thread on the side stream.  dsyevd blocks its
  // calling thread on internal synchronisations, so only two issuing threads let the two solves overlap
  // (35.2 -> 30.8 ms at n = 1376, profiles/eig_bench_r01.log)
  const bool two_threads = !dist && batch == 2 && c->overlap && c->eig_threads && n >= 512;
  {
    StageScope ts(c->timers, c->stream, "eigh");
    std::future<cusolverStatus_t> side;
    if (two_threads) {
      double* work2 = c->eigwork2.ensure((size_t)lwork);
      NBD_CUDA(cudaEventRecord(c->ev_fork, c->stream));
      NBD_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
      const int dev_id = c->device;
      cusolverDnHandle_t h2 = c->solver2;
      double* A2 = A + (long)n * n;
      double* w2 = w + n;
      side = std::async(std::launch::async, [=] {
        cudaSetDevice(dev_id);
        return cusolverDnDsyevd(h2, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, A2, n, w2, work2, lwork, info + 1);
      });
    }
    for (int b = 0; b < batch; ++b) {
      if ((dist && b != c->rank) || (two_threads && b == 1)) continue;
      NBD_SOLVER(cusolverDnDsyevd(c->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, A + (long)b * n * n, n,
                                  w + (long)b * n, work, lwork, info + b));
    }
    if (two_threads) {
      const cusolverStatus_t st2 = side.get();
      NBD_CUDA(cudaEventRecord(c->ev_join, c->stream2));
      NBD_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
      if (st2 != CUSOLVER_STATUS_SUCCESS) fail(NBD_ERR_CUDA, "cusolverDnDsyevd (side stream): status %d", (int)st2);
    }
  }
  if (dist) eig_exchange(c, A, w, n);
}
static void check_devinfo(nbd_ctx* c, int count, const char* what) {
  int h[8] = {0};
  NBD_CUDA(cudaMemcpyAsync(h, c->devinfo.p, sizeof(int) * count, cudaMemcpyDeviceToHost, c->stream));
  NBD_CUDA(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < count; ++i)
    if (h[i] != 0) fail(NBD_ERR_CUDA, "%s: cuSOLVER devInfo = %d", what, h[i]);
}

// ------------------------------------------------------------------------------------------------
// ABI helpers
// ------------------------------------------------------------------------------------------------
template <class F>
static int guarded(nbd_ctx* ctx, F&& f) {
  if (!ctx) return NBD_ERR_ARG;
  try {
    NBD_CUDA(cudaSetDevice(ctx->device));
    f();
    return NBD_OK;
  } catch (const Error& e) {
    ctx->err = e.msg;
    ctx->timers.resolve();
    return e.code;
  } catch (const std::exception& e) {
    ctx->err = e.what();
    return NBD_ERR_CUDA;
  }
}
static void finish_call(nbd_ctx* c) {
  NBD_CUDA(cudaStreamSynchronize(c->stream));
  c->timers.resolve();
}
// Large copies to / from PAGEABLE host memory (numpy arrays of the caller) are pipelined through two page-locked
// staging buffers: the DMA of chunk i overlaps the (multi-threaded) host memcpy of chunk i-1.  cudaMemcpyAsync on
// pageable memory reaches ~5 GB/s on these hosts; page-locking the caller's buffer per call costs more than that.
constexpr size_t STAGE_CHUNK = 16u << 20;
static bool host_is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}
// Copy threads (option "copy_threads"; default: half the host's hardware threads, 2..12).  The destination of a result
// copy is usually a fresh numpy array: most of the time goes into first-touch page faults, which only parallelise
// over threads.
static int g_copy_threads = 0;
static void parallel_memcpy(char* dst, const char* src, size_t n) {
  constexpr int TMAX = 12;
  if (g_copy_threads <= 0) g_copy_threads = std::max(2, std::min(TMAX, (int)std::thread::hardware_concurrency() / 2));
  const int T = std::min(TMAX, g_copy_threads);
  if (n < (4u << 20) || T < 2) {
    memcpy(dst, src, n);
    return;
  }
  std::future<void> f[TMAX - 1];
  const size_t part = (n / T + 63) & ~(size_t)63;
  for (int t = 1; t < T; ++t) {
    const size_t o = std::min(n, part * t), e = std::min(n, part * (t + 1));
    f[t - 1] = std::async(std::launch::async, [=] { memcpy(dst + o, src + o, e - o); });
  }
  memcpy(dst, src, std::min(n, part));
  for (int t = 1; t < T; ++t) f[t - 1].get();
}
// lane 0: the library's stream and staging buffers; lane 1: the copy stream (a second pair of staging buffers), used by
// the early export of D / Huz that runs next to the final eigensolve.
static void staged_copy(nbd_ctx* c, char* host, char* dev, size_t bytes, bool to_host, int lane = 0) {
  char* st = (char*)(lane ? c->pinned2 : c->pinned).ensure(2 * STAGE_CHUNK);
  cudaEvent_t* evs = c->ev_stage + 2 * lane;
  cudaStream_t stream = lane ? c->stream3 : c->stream;
  if (!evs[0]) {
    NBD_CUDA(cudaEventCreateWithFlags(&evs[0], cudaEventDisableTiming));
    NBD_CUDA(cudaEventCreateWithFlags(&evs[1], cudaEventDisableTiming));
  }
  const size_t nch = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
  if (to_host) {
    std::future<void> fut[2];
    for (size_t i = 0; i < nch; ++i) {
      const int b = (int)(i & 1);
      const size_t off = i * STAGE_CHUNK, n = std::min(STAGE_CHUNK, bytes - off);
      if (fut[b].valid()) fut[b].get();  // staging buffer b is free again
      NBD_CUDA(cudaMemcpyAsync(st + b * STAGE_CHUNK, dev + off, n, cudaMemcpyDeviceToHost, stream));
      NBD_CUDA(cudaEventRecord(evs[b], stream));
      cudaEvent_t ev = evs[b];
      char* src = st + b * STAGE_CHUNK;
      char* dst = host + off;
      const int dev_id = c->device;
      fut[b] = std::async(std::launch::async, [=] {
        cudaSetDevice(dev_id);
        cudaEventSynchronize(ev);
        parallel_memcpy(dst, src, n);
      });
    }
    for (auto& f : fut)
      if (f.valid()) f.get();
  } else {
    for (size_t i = 0; i < nch; ++i) {
      const int b = (int)(i & 1);
      const size_t off = i * STAGE_CHUNK, n = std::min(STAGE_CHUNK, bytes - off);
      if (i >= 2) NBD_CUDA(cudaEventSynchronize(evs[b]));  // DMA out of staging buffer b has finished
      parallel_memcpy(st + b * STAGE_CHUNK, host + off, n);
      NBD_CUDA(cudaMemcpyAsync(dev + off, st + b * STAGE_CHUNK, n, cudaMemcpyHostToDevice, stream));
      NBD_CUDA(cudaEventRecord(evs[b], stream));
    }
    NBD_CUDA(cudaEventSynchronize(evs[0]));
    NBD_CUDA(cudaEventSynchronize(evs[1]));
  }
}
static void h2d(nbd_ctx* c, double* dst, const double* src, size_t count) {
  const size_t bytes = count * sizeof(double);
  if (bytes >= (8u << 20) && host_is_pageable(src)) {
    staged_copy(c, (char*)src, (char*)dst, bytes, false);
    return;
  }
  NBD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
}
static void d2h(nbd_ctx* c, double* dst, const double* src, size_t count) {
  const size_t bytes = count * sizeof(double);
  if (bytes >= (8u << 20) && host_is_pageable(dst)) {
    staged_copy(c, (char*)dst, (char*)src, bytes, true);
    return;
  }
  NBD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
}
// Device -> host copy on the copy stream (lane 1), blocking: for a helper thread that exports results which are already
// final while the main thread keeps the library's stream busy.  The caller has made stream3 wait for the producers.
static void d2h_lane1(nbd_ctx* c, double* dst, const double* src, size_t count) {
  const size_t bytes = count * sizeof(double);
  if (bytes >= (8u << 20) && host_is_pageable(dst)) {
    staged_copy(c, (char*)dst, (char*)src, bytes, true, 1);
    return;
  }
  NBD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream3));
  NBD_CUDA(cudaStreamSynchronize(c->stream3));
}

extern "C" {

int nbd_version(void) { return 100; }

int nbd_create(nbd_ctx** out, int device) {
  if (!out) return NBD_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return NBD_ERR_CUDA;
  nbd_ctx* c = new nbd_ctx();
  c->device = device;
  try {
    NBD_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    NBD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) fail(NBD_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    NBD_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    NBD_SOLVER(cusolverDnCreate(&c->solver));
    NBD_SOLVER(cusolverDnSetStream(c->solver, c->stream));
    NBD_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    NBD_SOLVER(cusolverDnCreate(&c->solver2));
    NBD_SOLVER(cusolverDnSetStream(c->solver2, c->stream2));
    NBD_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    NBD_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  } catch (const Error& e) {
    fprintf(stderr, "nbd_create: %s\n", e.msg.c_str());
    int code = e.code;
    delete c;
    return code;
  }
  *out = c;
  return NBD_OK;
}

int nbd_destroy(nbd_ctx* c) {
  if (!c) return NBD_ERR_ARG;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  if (c->Bt) cudaFree(c->Bt);
  if (c->solver) cusolverDnDestroy(c->solver);
  if (c->blas) nbd_destroy_blas(c->blas);
  if (c->solver2) cusolverDnDestroy(c->solver2);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  for (auto e : c->ev_stage)
    if (e) cudaEventDestroy(e);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->stream3) cudaStreamDestroy(c->stream3);
  if (c->ev_export) cudaEventDestroy(c->ev_export);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return NBD_OK;
}

const char* nbd_last_error(nbd_ctx* c) { return c ? c->err.c_str() : "null context"; }

int nbd_set_option(nbd_ctx* c, const char* key, long value) {
  if (!c || !key) return NBD_ERR_ARG;
  std::string k(key);
  if (k == "jk_variant") c->jk_variant = (int)value;
  else if (k == "gemm_variant") c->gemm_variant = (int)value;
  else if (k == "x_budget_mb") c->x_budget_bytes = value << 20;
  else if (k == "timers") c->timers.enabled = value != 0;
  else if (k == "overlap") c->overlap = (int)value;
  else if (k == "eig_threads") c->eig_threads = (int)value;
  else if (k == "small_eigh") c->small_eigh = (int)value;
  else if (k == "small_eigh_warm") { c->small_eigh_warm = (int)value; c->se_warm_calls = 0; }
  else if (k == "dist_eig") c->dist_eig = (int)value;
  else if (k == "dist_orth") c->dist_orth = (int)value;
  else if (k == "ks_energy") c->ks_energy = (int)value;
  else if (k == "panel_stages") c->panel_stages = (int)value;
  else if (k == "panel_hybrid") c->panel_hybrid = (int)value;
  else if (k == "jpass_variant") c->jpass_variant = (int)value;
  else if (k == "gemm_tile") c->gemm_tile = (int)value;
  else if (k == "syrk") c->syrk_mode = (int)value;
  else if (k == "pair_split") c->pair_split = (int)value;
  else if (k == "pair_guest") c->pair_guest = (int)value;
  else if (k == "pair_guest_reserve") c->pair_guest_reserve = (int)value;
  else if (k == "jpass_safe") c->jpass_safe = (int)value;
  else if (k == "copy_threads") g_copy_threads = (int)value;
  else if (k == "x_cache") { if (c->x_cache != (int)value) c->x_valid_n = 0; c->x_cache = (int)value; }
  else if (k == "early_export") c->early_export = (int)value;
  else if (k == "packed_allreduce") c->packed_allreduce = (int)value;
  else if (k == "jpass_sm_mod") c->jpass_sm_mod = (int)value;
  else if (k == "jpass_sm_keep") c->jpass_sm_keep = (int)value;
  else if (k == "jpass_ctas_per_sm") c->jpass_ctas_per_sm = (int)value;
  else if (k == "syrk_ctas") c->syrk_ctas = (int)value;
  else if (k == "eig_mode") { c->eig_mode = (int)value; c->sub_valid = false; }
  else if (k == "sub_bound") { c->sub_bound_mode = (int)value; c->sub_bounds_valid = false; }
  else if (k == "sub_cold") c->sub_cold = (int)value;
  else if (k == "sub_apply_variant") c->sub_apply_variant = (int)value;
  else if (k == "sub_pdl") c->sub_pdl = (int)value;
  else if (k == "sub_adaptive") c->sub_adaptive = (int)value;
  else if (k == "dist_sub") c->dist_sub = (int)value;
  else if (k == "sub_min_nao") { c->sub_min_nao = (int)value; c->sub_valid = false; }
  else return NBD_ERR_ARG;
  return NBD_OK;
}

double nbd_timer_ms(nbd_ctx* c, const char* key) {
  if (!c || !key) return -1.0;
  // cumulative counters of the subspace eigensolver (since context creation)
  if (!strcmp(key, "count:sub_applies")) return (double)c->sub_applies;
  if (!strcmp(key, "count:sub_outer")) return (double)c->sub_outer;
  if (!strcmp(key, "count:sub_fallbacks")) return (double)c->sub_fallbacks;
  if (!strcmp(key, "count:sub_lanczos")) return (double)c->sub_lanczos;
  if (!strcmp(key, "count:sub_cold_starts")) return (double)c->sub_cold_starts;
  if (!strcmp(key, "count:sub_rejects")) return (double)c->sub_rejects;
  auto it = c->timers.ms.find(key);
  return it == c->timers.ms.end() ? 0.0 : it->second;
}

long nbd_launch_count(nbd_ctx* c) { return c ? c->launches : -1; }

void* nbd_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void nbd_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// ---- NCCL ---------------------------------------------------------------------------------------
int nbd_comm_unique_id(void* unique_id_128) {
  std::string why;
  if (!unique_id_128 || !g_nccl.load(why)) return NBD_ERR_CUDA;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return NBD_ERR_CUDA;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(unique_id_128, &id, 128);
  return NBD_OK;
}

int nbd_comm_init(nbd_ctx* c, const void* unique_id_128, int rank, int world) {
  return guarded(c, [&] {
    NBD_REQUIRE(unique_id_128 && world >= 1 && rank >= 0 && rank < world, NBD_ERR_ARG, "bad communicator arguments");
    std::string why;
    if (!g_nccl.load(why)) fail(NBD_ERR_CUDA, "%s", why.c_str());
    ncclUniqueId id;
    memcpy(&id, unique_id_128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) fail(NBD_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    c->rank = rank;
    c->world = world;
  });
}

// ---- 3-centre tensor ------------------------------------------------------------------------------
int nbd_cderi_alloc(nbd_ctx* c, int nao, int naux_local) {
  return guarded(c, [&] {
    NBD_REQUIRE(nao > 0 && naux_local >= 0, NBD_ERR_ARG, "nao = %d, naux = %d", nao, naux_local);
    NBD_REQUIRE(naux_local <= 65535, NBD_ERR_UNSUPPORTED, "naux_local = %d > 65535", naux_local);
    if (c->Bt) {
      NBD_CUDA(cudaFree(c->Bt));
      c->Bt = nullptr;
    }
    c->nao = nao;
    c->nb = (nao + TILE - 1) / TILE;
    c->n_ld = c->nb * TILE;
    c->ntiles = c->nb * (c->nb + 1) / 2;
    c->npair = (long)nao * (nao + 1) / 2;
    c->naux = naux_local;
    c->seq = build_tile_sequence(c->nb);
    c->plans.clear();
    NBD_REQUIRE((int)c->seq.size() == c->ntiles, NBD_ERR_STATE, "tile sequence has %zu entries, expected %d", c->seq.size(), c->ntiles);
    c->inv.assign((size_t)c->nb * c->nb, -1);
    for (int k = 0; k < c->ntiles; ++k) c->inv[(size_t)(c->seq[k] >> 16) * c->nb + (c->seq[k] & 0xffff)] = k;
    c->d_seq.ensure(c->ntiles);
    c->d_inv.ensure((size_t)c->nb * c->nb);
    NBD_CUDA(cudaMemcpyAsync(c->d_seq.p, c->seq.data(), sizeof(int) * c->ntiles, cudaMemcpyHostToDevice, c->stream));
    NBD_CUDA(cudaMemcpyAsync(c->d_inv.p, c->inv.data(), sizeof(int) * c->nb * c->nb, cudaMemcpyHostToDevice, c->stream));
    const size_t bytes = (size_t)std::max(1, naux_local) * c->ntiles * TILE_BYTES;
    NBD_CUDA(cudaMalloc(&c->Bt, bytes));
    c->scf_ready = false;
    c->bench_ready = false;
    finish_call(c);
  });
}

static int stage_rows(nbd_ctx* c) {
  const long budget = 256L << 20;
  return (int)std::max<long>(1, std::min<long>(4096, budget / (c->npair * 8)));
}

int nbd_cderi_upload(nbd_ctx* c, const double* rows, int row0, int nrows) {
  return guarded(c, [&] {
    NBD_REQUIRE(c->Bt, NBD_ERR_STATE, "nbd_cderi_alloc first");
    NBD_REQUIRE(rows && row0 >= 0 && nrows >= 0 && row0 + nrows <= c->naux, NBD_ERR_ARG, "rows [%d, %d) outside [0, %d)", row0, row0 + nrows, c->naux);
    const int step = stage_rows(c);
    double* st = c->stage.ensure((size_t)step * c->npair);
    for (int r = 0; r < nrows; r += step) {
      const int k = std::min(step, nrows - r);
      h2d(c, st, rows + (long)r * c->npair, (size_t)k * c->npair);
      dim3 g(c->ntiles, k);
      pack_to_tiled_kernel<<<g, 256, 0, c->stream>>>(st, c->Bt, c->d_seq.p, c->ntiles, c->nao, c->npair, row0 + r);
      LAUNCH_CHECK(c);
      NBD_CUDA(cudaStreamSynchronize(c->stream));
    }
    finish_call(c);
  });
}

int nbd_cderi_synth(nbd_ctx* c, unsigned long long seed, double scale, int global_row0) {
  return guarded(c, [&] {
    NBD_REQUIRE(c->Bt, NBD_ERR_STATE, "nbd_cderi_alloc first");
    if (c->naux > 0) {
      dim3 g(c->ntiles, c->naux);
      synth_tiled_kernel<<<g, 256, 0, c->stream>>>(c->Bt, c->d_seq.p, c->ntiles, c->nao, c->npair, seed, scale, global_row0);
      LAUNCH_CHECK(c);
    }
    finish_call(c);
  });
}

int nbd_cderi_download(nbd_ctx* c, double* rows, int row0, int nrows) {
  return guarded(c, [&] {
    NBD_REQUIRE(c->Bt, NBD_ERR_STATE, "nbd_cderi_alloc first");
    NBD_REQUIRE(rows && row0 >= 0 && nrows >= 0 && row0 + nrows <= c->naux, NBD_ERR_ARG, "rows [%d, %d) outside [0, %d)", row0, row0 + nrows, c->naux);
    const int step = stage_rows(c);
    double* st = c->stage.ensure((size_t)step * c->npair);
    for (int r = 0; r < nrows; r += step) {
      const int k = std::min(step, nrows - r);
      dim3 g(c->ntiles, k);
      tiled_to_packed_kernel<<<g, 256, 0, c->stream>>>(c->Bt, st, c->d_seq.p, c->ntiles, c->nao, c->npair, row0 + r);
      LAUNCH_CHECK(c);
      d2h(c, rows + (long)r * c->npair, st, (size_t)k * c->npair);
      NBD_CUDA(cudaStreamSynchronize(c->stream));
    }
    finish_call(c);
  });
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Host-orbital J/K (nbd_jk) and dense-density J/K (nbd_jk_dm)
// ------------------------------------------------------------------------------------------------
// d_src: device [ncols][src_ld] orbital rows (row = orbital); writes rows [row0, row0+ncols) of c->d_orb / c->d_wt.
static void stage_orbitals(nbd_ctx* c, const double* d_src, long src_ld, int ncols, int row0, double fixed_scale,
                           const double* d_rowscale, double wt_sign) {
  if (ncols <= 0) return;
  dim3 g((c->n_ld + 127) / 128, ncols);
  pad_rows_kernel<<<g, 128, 0, c->stream>>>(d_src, src_ld, c->d_orb.p + (long)row0 * c->n_ld, c->n_ld, c->nao, d_rowscale, fixed_scale);
  LAUNCH_CHECK(c);
  pad_rows_kernel<<<g, 128, 0, c->stream>>>(d_src, src_ld, c->d_wt.p + (long)row0 * c->n_ld, c->n_ld, c->nao, d_rowscale, fixed_scale * wt_sign);
  LAUNCH_CHECK(c);
}

extern "C" int nbd_jk(nbd_ctx* c, int nset, const int* ncol, const double* orb, const double* sign, double* vj, double* vk) {
  return guarded(c, [&] {
    c->timers.reset();
    NBD_REQUIRE(c->Bt, NBD_ERR_STATE, "nbd_cderi_alloc first");
    NBD_REQUIRE(nset >= 1 && nset <= 64 && ncol && orb, NBD_ERR_ARG, "bad nset / pointers");
    const int n = c->nao;
    const long nn = (long)n * n;
    int Ntot = 0;
    for (int s = 0; s < nset; ++s) {
      NBD_REQUIRE(ncol[s] >= 0, NBD_ERR_ARG, "negative column count");
      Ntot += ncol[s];
    }
    // host -> device with the columns of each set reordered: positive-sign columns first, then negative ones
    c->d_orb.ensure((size_t)std::max(1, Ntot) * c->n_ld);
    c->d_wt.ensure((size_t)std::max(1, Ntot) * c->n_ld);
    std::vector<int> jbegin(nset + 1, 0);
    std::vector<KGroup> groups;
    std::vector<double> hrows((size_t)std::max(1, Ntot) * n), hsign(std::max(1, Ntot), 1.0);
    {
      long off = 0;  // offset into orb (per set: [n][ncol[s]] row-major)
      int row = 0, sidx = 0;
      for (int s = 0; s < nset; ++s) {
        const int w = ncol[s];
        const int start = row;
        for (int pass = 0; pass < 2; ++pass) {
          const int r0 = row;
          for (int i = 0; i < w; ++i) {
            const double sg = sign ? sign[sidx + i] : 1.0;
            if ((pass == 0) != (sg >= 0.0)) continue;
            for (int m = 0; m < n; ++m) hrows[(size_t)row * n + m] = orb[off + (long)m * w + i];
            hsign[row] = sg >= 0.0 ? 1.0 : -1.0;
            ++row;
          }
          if (row > r0) groups.push_back(KGroup{s, r0, row, pass == 0 ? 1.0 : -1.0});
        }
        (void)start;
        jbegin[s + 1] = row;
        off += (long)n * w;
        sidx += w;
      }
    }
    double* d_stage = c->stage.ensure((size_t)std::max(1, Ntot) * n + std::max(1, Ntot));
    if (Ntot > 0) {
      h2d(c, d_stage, hrows.data(), (size_t)Ntot * n);
      h2d(c, d_stage + (size_t)Ntot * n, hsign.data(), Ntot);
      dim3 g((c->n_ld + 127) / 128, Ntot);
      pad_rows_kernel<<<g, 128, 0, c->stream>>>(d_stage, n, c->d_orb.p, c->n_ld, n, nullptr, 1.0);
      LAUNCH_CHECK(c);
      pad_rows_kernel<<<g, 128, 0, c->stream>>>(d_stage, n, c->d_wt.p, c->n_ld, n, d_stage + (size_t)Ntot * n, 1.0);
      LAUNCH_CHECK(c);
    }
    const int nj = vj ? nset : 0, nk = vk ? nset : 0;
    double* buf = c->d_jk.ensure((size_t)std::max(1, nj + nk) * nn);
    jk_device(c, c->d_orb.p, c->d_wt.p, Ntot, nj, jbegin, vj ? buf : nullptr, nk, groups, vk ? buf + (size_t)nj * nn : nullptr);
    all_reduce(c, buf, (size_t)(nj + nk) * nn);
    if (vj) d2h(c, vj, buf, (size_t)nj * nn);
    if (vk) d2h(c, vk, buf + (size_t)nj * nn, (size_t)nk * nn);
    finish_call(c);
  });
}

// Factor symmetric dense densities dm[s] = sum_i sign_i c_i c_i^T (eigen-decomposition) into c->d_orb / c->d_wt.
// d_dm: device [nset][n][n] (destroyed).  Returns Ntot and fills jbegin / groups.  tol drops |w| <= tol * max|w|.
static int factor_densities(nbd_ctx* c, double* d_dm, int nset, std::vector<int>& jbegin, std::vector<KGroup>& groups) {
  const int n = c->nao;
  double* w = c->evals.ensure((size_t)nset * n);
  eigh_batched(c, d_dm, w, n, nset);
  std::vector<double> hw((size_t)nset * n);
  d2h(c, hw.data(), w, (size_t)nset * n);
  check_devinfo(c, std::min(nset, 8), "density factorisation");
  // rows of d_dm are now eigenvectors; scale by sqrt|w| and keep the significant ones (ascending order:
  // negatives first, positives last)
  jbegin.assign(nset + 1, 0);
  groups.clear();
  int total = 0;
  std::vector<std::array<int, 4>> plan;  // set, first row, count, sign
  for (int s = 0; s < nset; ++s) {
    double wmax = 0.0;
    for (int i = 0; i < n; ++i) wmax = std::max(wmax, std::fabs(hw[(size_t)s * n + i]));
    const double thr = wmax * 1e-14;
    int nneg = 0, npos = 0;
    for (int i = 0; i < n; ++i) {
      const double v = hw[(size_t)s * n + i];
      if (v < -thr) ++nneg;
      else if (v > thr) ++npos;
    }
    // eigenvalues ascend: the first nneg rows are the negative ones, the last npos the positive ones
    plan.push_back({s, n - npos, npos, +1});
    plan.push_back({s, 0, nneg, -1});
    total += npos + nneg;
  }
  c->d_orb.ensure((size_t)std::max(1, total) * c->n_ld);
  c->d_wt.ensure((size_t)std::max(1, total) * c->n_ld);
  int row = 0;
  for (auto& pl : plan) {
    const int s = pl[0], r0 = pl[1], cnt = pl[2], sg = pl[3];
    if (cnt > 0) {
      double* src = d_dm + (long)s * n * n + (long)r0 * n;
      dim3 g((n + 127) / 128, cnt);
      scale_rows_kernel<<<g, 128, 0, c->stream>>>(src, w + (long)s * n + r0, n, 1);
      LAUNCH_CHECK(c);
      stage_orbitals(c, src, n, cnt, row, 1.0, nullptr, (double)sg);
      groups.push_back(KGroup{s, row, row + cnt, (double)sg});
      row += cnt;
    }
    if (sg < 0) jbegin[s + 1] = row;
  }
  return total;
}

extern "C" int nbd_jk_dm(nbd_ctx* c, int nset, const double* dm, double* vj, double* vk) {
  return guarded(c, [&] {
    c->timers.reset();
    NBD_REQUIRE(c->Bt, NBD_ERR_STATE, "nbd_cderi_alloc first");
    NBD_REQUIRE(nset >= 1 && nset <= 8 && dm, NBD_ERR_ARG, "bad nset / pointers");
    const int n = c->nao;
    const long nn = (long)n * n;
    double* d_dm = c->dm0f.ensure((size_t)nset * nn);
    h2d(c, d_dm, dm, (size_t)nset * nn);
    std::vector<int> jbegin;
    std::vector<KGroup> groups;
    const int Ntot = factor_densities(c, d_dm, nset, jbegin, groups);
    const int nj = vj ? nset : 0, nk = vk ? nset : 0;
    double* buf = c->d_jk.ensure((size_t)std::max(1, nj + nk) * nn);
    jk_device(c, c->d_orb.p, c->d_wt.p, Ntot, nj, jbegin, vj ? buf : nullptr, nk, groups, vk ? buf + (size_t)nj * nn : nullptr);
    all_reduce(c, buf, (size_t)(nj + nk) * nn);
    if (vj) d2h(c, vj, buf, (size_t)nj * nn);
    if (vk) d2h(c, vk, buf + (size_t)nj * nn, (size_t)nk * nn);
    finish_call(c);
  });
}

#include "subspace_host.cuh"
#include "integrals_host.cuh"
#include "xc_host.cuh"
#include "scf_host.cuh"
#include "ao2mo_host.cuh"
