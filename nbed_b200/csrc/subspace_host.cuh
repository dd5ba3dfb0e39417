// Host driver of the Chebyshev-filtered subspace iteration (subspace.cuh).  Included by nbed_b200.cu before
// scf_host.cuh.
#pragma once

static int sub_block_size(const nbd_ctx* c) {
  int omax = 0;
  for (int s = 0; s < c->nspin; ++s) omax = std::max(omax, c->nspin == 2 ? c->nelec[s] : (c->nelec[0] + c->nelec[1]) / 2);
  if (c->eig_mode != 1 || c->projector != NBD_HUZINAGA || c->nao < c->sub_min_nao || omax < 1) return 0;
  // (a 32-vector block for <= 10 occupied orbitals was measured in round 2: 3 % fewer block products at C4, each twice
  // as expensive - profiles/r02x_block32.md)
  if (omax <= 10) return 16;
  if (omax <= 24) return 32;
  return 0;
}

static void sub_alloc(nbd_ctx* c, int kb) {
  const size_t blk = (size_t)2 * c->nao * kb;
  for (DBuf<double>* b : {&c->sV, &c->sY, &c->sZ, &c->sW, &c->sAV}) b->ensure(blk);
  c->sG.ensure((size_t)2 * 2 * kb * kb);
  c->sGpart.ensure((size_t)2 * 64 * 2 * kb * kb);
  c->sM.ensure((size_t)2 * kb * kb);
  c->sTheta.ensure((size_t)2 * kb);
  c->sRpart.ensure((size_t)2 * 64 * kb);
  c->sBound.ensure(8);
  if (c->sTicket.cap < 4096) {
    c->sTicket.ensure(4096);
    NBD_CUDA(cudaMemsetAsync(c->sTicket.p, 0, sizeof(unsigned int) * 4096, c->stream));
  }
}

// after a full eigensolve: rows of `eigrows` ([nspin][n][n], row = eigenvector, ascending) -> tracked block
static void sub_init_from_full(nbd_ctx* c, const double* eigrows, const double* evals_dev) {
  const int kb = sub_block_size(c);
  c->sub_valid = false;
  if (!kb) return;
  const int n = c->nao;
  sub_alloc(c, kb);
  dim3 g((n + 127) / 128, kb, c->nspin);
  sub_gather_block_kernel<<<g, 128, 0, c->stream>>>(eigrows, c->sV.p, n, kb, (long)n * n, (long)n * kb);
  LAUNCH_CHECK(c);
  std::vector<double> w((size_t)c->nspin * n);
  d2h(c, w.data(), evals_dev, (size_t)c->nspin * n);
  NBD_CUDA(cudaStreamSynchronize(c->stream));
  for (int s = 0; s < c->nspin; ++s)
    for (int k = 0; k < kb; ++k) c->sub_theta[s][k] = w[(size_t)s * n + k];
  c->sub_kb = kb;
  c->sub_valid = true;
}

// Small device -> host read-backs go through a pinned scratch area: a pageable destination makes the driver stage the
// copy, which costs several microseconds more per round trip - and the SCF loop makes ~10 of them per cycle.
// `slot` separates read-backs that are in flight at the same time (slots of 4096 doubles).
static double* rb_slot(nbd_ctx* c, int slot) { return (double*)c->rb.ensure((size_t)8 * 4096 * sizeof(double)) + (size_t)slot * 4096; }
static double* d2h_small(nbd_ctx* c, int slot, const double* src, size_t count) {
  NBD_REQUIRE(count <= 4096, NBD_ERR_STATE, "read-back of %zu doubles does not fit a slot", count);
  double* h = rb_slot(c, slot);
  NBD_CUDA(cudaMemcpyAsync(h, src, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  return h;
}

// k-step Lanczos per spin (batched): up = theta_max + beta_k (+ 2 % of the Ritz spread), low = theta_min - beta_k.
// Spins [s0, s0 + ns) of the problem; Fp, up, low are indexed by the LOCAL spin 0 .. ns-1 (the caller offsets them).
// The recurrence scalars stay on the device; the host reads all of them back once and cuts the run at a breakdown.
static void sub_lanczos_bounds(nbd_ctx* c, const double* Fp, double* up, double* low, int s0, int ns) {
  const int n = c->nao;
  const int steps = std::min(10, n);
  const int nblk = (n + 7) / 8;
  double* buf = c->sLz.ensure((size_t)3 * ns * n + (size_t)2 * ns * nblk + (size_t)2 * SUB_LZ_COEF);
  double *v = buf, *vprev = buf + (size_t)ns * n, *w = buf + (size_t)2 * ns * n, *part = buf + (size_t)3 * ns * n,
         *coef = part + (size_t)2 * ns * nblk;
  // deterministic start vector (same hash as the start block), normalised on the host
  std::vector<double> v0((size_t)ns * n);
  for (int s = 0; s < ns; ++s) {
    double nrm = 0.0;
    for (int i = 0; i < n; ++i) {
      unsigned long long z = (unsigned long long)((s0 + s) * n + i) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z = z ^ (z >> 31);
      const double x = (double)(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
      v0[(size_t)s * n + i] = x;
      nrm += x * x;
    }
    nrm = 1.0 / std::sqrt(nrm);
    for (int i = 0; i < n; ++i) v0[(size_t)s * n + i] *= nrm;
  }
  h2d(c, v, v0.data(), v0.size());
  NBD_CUDA(cudaMemsetAsync(vprev, 0, sizeof(double) * ns * n, c->stream));
  NBD_CUDA(cudaMemsetAsync(coef, 0, sizeof(double) * 2 * SUB_LZ_COEF, c->stream));
  for (int j = 0; j < steps; ++j) {
    sub_lanczos_matvec_kernel<<<dim3(nblk, ns), 256, 0, c->stream>>>(Fp, v, vprev, coef, j, w, part, n);
    LAUNCH_CHECK(c);
    sub_lanczos_scalars_kernel<<<ns, 32, 0, c->stream>>>(part, nblk, j, coef);
    LAUNCH_CHECK(c);
    if (j + 1 == steps) break;
    sub_lanczos_update_kernel<<<dim3((n + 255) / 256, ns), 256, 0, c->stream>>>(w, v, vprev, coef, j, n);
    LAUNCH_CHECK(c);
    std::swap(v, vprev);
  }
  const double* hc = d2h_small(c, 2, coef, (size_t)2 * SUB_LZ_COEF);
  NBD_CUDA(cudaStreamSynchronize(c->stream));
  double alpha[2][16] = {}, beta[2][17] = {};
  int k = 0;
  for (int j = 0; j < steps; ++j) {
    bool breakdown = false;
    for (int s = 0; s < ns; ++s) {
      alpha[s][j] = hc[s * SUB_LZ_COEF + j];
      beta[s][j + 1] = hc[s * SUB_LZ_COEF + 16 + j + 1];
      if (!(beta[s][j + 1] > 1e-10 * (std::fabs(alpha[s][j]) + 1e-300))) breakdown = true;
    }
    k = j + 1;
    if (breakdown) break;
  }
  for (int s = 0; s < ns; ++s) {
    double lo, hi;
    lanczos_ritz_bounds(k, alpha[s], beta[s], &lo, &hi);
    const double spread = (hi - lo) * 0.02;
    up[s] = hi + spread;
    low[s] = lo - spread;
  }
  ++c->sub_lanczos;
}

// Interval [low, up] containing the spectrum of Fp (per spin) for the Chebyshev filter.  Mode 1: the bounds of the
// last matrix widened by the Frobenius norm of the change (rigorous: |lambda(A + E) - lambda(A)| <= ||E||_2 <= ||E||_F),
// refreshed by a new Lanczos run whenever they have drifted by a quarter of the width.  Mode 0: row-sum bound.
// Two halves, so that the read-back of ||dF'||_F rides on the first Rayleigh-Ritz step's host round trip:
// sub_bounds_begin launches the norm (and the copy of F' the next cycle compares with) and returns the host address the
// partial sums will arrive at; sub_bounds_finish - after a stream synchronisation - turns them into bounds.
static const double* sub_bounds_begin(nbd_ctx* c, const double* Fp, int s0, int ns) {
  const long nn = (long)c->nao * c->nao;
  if (c->sub_bound_mode == 0) return nullptr;
  constexpr int NBLK = 64;
  double* prev = c->sFprev.ensure((size_t)c->nspin * nn) + (long)s0 * nn;
  if (!c->sub_bounds_valid) {
    NBD_CUDA(cudaMemcpyAsync(prev, Fp, sizeof(double) * ns * nn, cudaMemcpyDeviceToDevice, c->stream));
    return nullptr;
  }
  double* bp = c->sBound.ensure((size_t)2 * NBLK);
  sub_diffnorm_kernel<<<dim3(NBLK, ns), 256, 0, c->stream>>>(Fp, prev, nn, bp);
  LAUNCH_CHECK(c);
  return d2h_small(c, 1, bp, (size_t)ns * NBLK);
}
static void sub_bounds_finish(nbd_ctx* c, const double* Fp, const double* hb, double* up, double* low, int s0, int ns) {
  const int n = c->nao;
  if (c->sub_bound_mode == 0) {
    const int nblk = (n + 7) / 8;
    double* bp = c->sBound.ensure((size_t)2 * nblk);
    sub_gershgorin_kernel<<<dim3(nblk, ns), 256, 0, c->stream>>>(Fp, n, bp);
    LAUNCH_CHECK(c);
    std::vector<double> g((size_t)ns * nblk);
    d2h(c, g.data(), bp, g.size());
    NBD_CUDA(cudaStreamSynchronize(c->stream));
    for (int s = 0; s < ns; ++s) {
      up[s] = 0.0;
      for (int k = 0; k < nblk; ++k) up[s] = std::max(up[s], g[(size_t)s * nblk + k]);
      low[s] = -up[s];
    }
    return;
  }
  constexpr int NBLK = 64;
  bool refresh = !c->sub_bounds_valid || !hb;
  if (!refresh) {
    for (int s = 0; s < ns; ++s) {
      double d2 = 0.0;
      for (int k = 0; k < NBLK; ++k) d2 += hb[(size_t)s * NBLK + k];
      const double d = std::sqrt(d2) * (1.0 + 1e-12);
      up[s] = c->sub_up[s0 + s] + d;
      low[s] = c->sub_low[s0 + s] - d;
      if (!(up[s] - c->sub_up_ref[s0 + s] <= 0.25 * (c->sub_up_ref[s0 + s] - c->sub_low_ref[s0 + s]))) refresh = true;
    }
  }
  if (refresh) {
    sub_lanczos_bounds(c, Fp, up, low, s0, ns);
    for (int s = 0; s < ns; ++s) {
      c->sub_up_ref[s0 + s] = up[s];
      c->sub_low_ref[s0 + s] = low[s];
    }
  }
  for (int s = 0; s < ns; ++s) {
    c->sub_up[s0 + s] = up[s];
    c->sub_low[s0 + s] = low[s];
  }
  c->sub_bounds_valid = true;
}

// Start block of a cold solve: pseudo-random vectors; the first Rayleigh-Ritz step of sub_solve_t orthonormalises them.
static void sub_init_cold(nbd_ctx* c) {
  const int kb = sub_block_size(c);
  c->sub_valid = false;
  if (!kb || c->sub_bound_mode != 1) return;
  sub_alloc(c, kb);
  const long count = (long)c->nspin * c->nao * kb;
  sub_random_block_kernel<<<grid1(count, 256), 256, 0, c->stream>>>(c->sV.p, count, 20240ull);
  LAUNCH_CHECK(c);
  c->sub_kb = kb;
  c->sub_valid = true;
  c->sub_is_cold = true;
  ++c->sub_cold_starts;
}

// pdl: this product directly follows another product of the same chain in the stream (filter steps 2 .. degree and the
// Rayleigh-Ritz product behind a filter): it may start while that one is still running (sub_apply3_kernel<KB, 1>).
template <int KB>
static void sub_apply(nbd_ctx* c, const double* A, const double* Y, const double* Z, double* out, const double* alpha,
                      const double* shift, const double* beta, int ns, bool pdl = false) {
  const int n = c->nao;
  pdl = pdl && c->sub_pdl;
  SubApplyArgs a{};
  a.A = A; a.Y = Y; a.Z = Z; a.out = out;
  a.n = n;
  const int nrb = (n + SUB_ROWS - 1) / SUB_ROWS;
  for (int b = 0; b < 2; ++b) {
    const int q = b < ns ? b : 0;
    a.alpha[b] = alpha[q]; a.shift[b] = shift[q]; a.beta[b] = beta[q];
  }
  // Kernel choice (tools/sub_apply_bench.cu on B200, profiles/r02y_sub_apply_bench.md): the single-shot kernel wins when
  // its whole grid is resident at once - one spin at n = 1376 (7.6 against 9.8 us per product), both spins at n = 688 -
  // the ring kernel otherwise (both spins at n = 1376: 10.2 against 14.0 us) and on small matrices.
  const int smem3 = sub_apply3_smem_bytes<KB>(n);
  const long ctas3 = (long)((n + SUB3_ROWS - 1) / SUB3_ROWS) * SUB3_KS * ns;
  const bool fits3 = (n & 1) == 0 && n >= 512 && smem3 <= 200 * 1024 &&
                     ctas3 <= (long)c->sm_count * ((227 * 1024) / (smem3 + 1024));
  if (c->sub_apply_variant == 2 || (c->sub_apply_variant == 0 && fits3)) {
    NBD_REQUIRE((n & 1) == 0 && smem3 <= 200 * 1024, NBD_ERR_ARG, "sub_apply_variant = 2 needs an even nao that fits shared memory");
    // single-shot bulk-copy kernel: 8-CTA clusters split the contraction index, whole operand slab resident
    static unsigned long long configured3 = 0;
    if (first_use_on_current_device(configured3)) {
      NBD_CUDA(cudaFuncSetAttribute(sub_apply3_kernel<KB, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      NBD_CUDA(cudaFuncSetAttribute(sub_apply3_kernel<KB, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      NBD_CUDA(cudaFuncSetAttribute(sub_apply3_kernel<KB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      NBD_CUDA(cudaFuncSetAttribute(sub_apply3_kernel<KB, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((n + SUB3_ROWS - 1) / SUB3_ROWS, SUB3_KS, ns);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem3;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = SUB3_KS;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 2 : 1;
    if (pdl) NBD_CUDA(cudaLaunchKernelEx(&cfg, sub_apply3_kernel<KB, 1>, a));
    else NBD_CUDA(cudaLaunchKernelEx(&cfg, sub_apply3_kernel<KB, 0>, a));
    LAUNCH_CHECK(c);
    ++c->sub_applies;
    return;
  }
  if (c->sub_apply_variant != 1) {  // 0 (shape does not suit the single-shot kernel) or 3 (forced)
    // cluster split-K kernel: 32 rows per CTA, 4 CTAs of a cluster share the contraction index, DSMEM reduction
    constexpr int smem2 = sub_apply2_smem_bytes<KB>();
    static unsigned long long configured2 = 0;
    if (first_use_on_current_device(configured2)) {
      NBD_CUDA(cudaFuncSetAttribute(sub_apply2_kernel<KB, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
      NBD_CUDA(cudaFuncSetAttribute(sub_apply2_kernel<KB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((n + SUB2_ROWS - 1) / SUB2_ROWS, SUB2_KS, ns);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem2;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = SUB2_KS;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 2 : 1;
    if (pdl) NBD_CUDA(cudaLaunchKernelEx(&cfg, (sub_apply2_kernel<KB, 1>), a));
    else NBD_CUDA(cudaLaunchKernelEx(&cfg, (sub_apply2_kernel<KB, 0>), a));
    LAUNCH_CHECK(c);
    ++c->sub_applies;
    return;
  }
  dim3 g(nrb, 1, ns);
  constexpr int smem = sub_apply_smem_bytes<KB>();
  static unsigned long long configured = 0;
  if (first_use_on_current_device(configured)) NBD_CUDA(cudaFuncSetAttribute(sub_apply_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  sub_apply_kernel<KB><<<g, 128, smem, c->stream>>>(a);
  LAUNCH_CHECK(c);
  ++c->sub_applies;
}

// Tracks the KB lowest eigenvectors of spins [s0, s0 + ns) of Fp ([nspin][n][n], Lowdin basis) starting from c->sV.
// On success c->sV holds the Ritz vectors, c->sub_theta the Ritz values; returns false when it did not converge.
// Host round trips per Rayleigh-Ritz step: two (Gram matrices out / rotation in, residuals out), both through the pinned
// read-back area; the norm ||F'_k - F'_{k-1}||_F that widens the filter interval arrives with the first of them.
template <int KB>
static bool sub_solve_t(nbd_ctx* c, const double* Fp_all, int s0, int ns) {
  const int n = c->nao;
  const long blk = (long)n * KB, nn = (long)n * n;
  constexpr int NBLK = 64;
  const bool cold = c->sub_is_cold;
  c->sub_is_cold = false;
  const int max_degree = 24, max_outer = cold ? 24 : 14;
  const double tol = 1e-10;
  // device views of the spins handled here (sV / sY are re-read from the context: the rotation swaps them)
  const double* Fp = Fp_all + (long)s0 * nn;
  double *sZ = c->sZ.p + s0 * blk, *sW = c->sW.p + s0 * blk, *sAV = c->sAV.p + s0 * blk;
  double *sG = c->sG.p + (long)s0 * 2 * KB * KB, *sGpart = c->sGpart.p + (long)s0 * NBLK * 2 * KB * KB,
         *sM = c->sM.p + (long)s0 * KB * KB, *sTheta = c->sTheta.p + (long)s0 * KB, *sRpart = c->sRpart.p + (long)s0 * NBLK * KB;
  unsigned int* ticket = c->sTicket.p + 2048 + s0;
  double bound[2] = {0, 0}, lowb[2] = {0, 0}, bu_used[2] = {0, 0};
  const double* hb = sub_bounds_begin(c, Fp, s0, ns);
  const double one[2] = {1.0, 1.0}, zero[2] = {0.0, 0.0};
  double* cur = c->sV.p + s0 * blk;  // block to Rayleigh-Ritz next
  double worst_prev = -1.0;
  int deg_prev = 0;
  std::vector<double> M((size_t)ns * KB * KB), th((size_t)ns * KB);
  for (int outer = 0; outer < max_outer; ++outer) {
    double *sV = c->sV.p + s0 * blk, *sY = c->sY.p + s0 * blk;
    if (outer > 0) {
      // scaled Chebyshev filter of degree `degree` damping [a, bound] (Zhou & Saad), per spin
      double e[2], cc[2] = {0, 0}, sig[2], sig1[2], al[2];
      int degree = max_degree;
      // A tracked block late in the SCF starts close to converged: the residual reduction per filter degree observed
      // earlier in this SCF (c->sub_rate, the slowest seen) tells how many degrees the remaining gap to `tol` needs.
      if (!cold && c->sub_adaptive && c->sub_rate > 0.0 && c->sub_rate < 1.0 && worst_prev > 0.0) {
        const double need = std::log(0.1 * tol / worst_prev) / std::log(c->sub_rate);
        if (need < max_degree) degree = std::max(6, (int)std::ceil(need) + 2);
      }
      for (int s = 0; s < ns; ++s) {
        const double a = c->sub_theta[s0 + s][KB - 1], a0 = c->sub_theta[s0 + s][0];
        const double bu = std::max(bound[s], a + 1e-3) + 1e-7 * std::fabs(bound[s]) + 1e-9;
        bu_used[s] = bu;
        // mode 0 has no estimate of the lowest eigenvalue: tracked blocks (theta_0 converged) take the full degree
        const double lmin = c->sub_bound_mode == 1 ? std::min(lowb[s], a0) : a0;
        degree = std::min(degree, chebyshev_degree(a, bu, lmin, 1e10, max_degree));
        deg_prev = degree;
        e[s] = 0.5 * (bu - a);
        cc[s] = 0.5 * (bu + a);
        sig1[s] = e[s] / (a0 - cc[s]);
        sig[s] = sig1[s];
        al[s] = sig1[s] / e[s];
      }
      // coefficients of all steps (Zhou & Saad's scaled three-term recurrence)
      double alv[2][SUBF_MAX_STEPS] = {}, bev[2][SUBF_MAX_STEPS] = {};
      for (int s = 0; s < ns; ++s) alv[s][1] = al[s];
      for (int i = 2; i <= degree; ++i)
        for (int s = 0; s < ns; ++s) {
          const double sig2 = 1.0 / (2.0 / sig1[s] - sig[s]);
          alv[s][i] = 2.0 * sig2 / e[s];
          bev[s][i] = sig[s] * sig2;
          sig[s] = sig2;
        }
      double* const bufs[3] = {sV, sY, sZ};
      cur = bufs[degree % 3];
      for (int i = 1; i <= degree; ++i) {
        double ai[2] = {alv[0][i], alv[1][i]}, bi[2] = {bev[0][i], bev[1][i]};
        sub_apply<KB>(c, Fp, bufs[(i - 1) % 3], i >= 2 ? bufs[(i - 2) % 3] : nullptr, bufs[i % 3], ai, cc, i >= 2 ? bi : zero, ns, i >= 2);
      }
    }
    // W = F' Y ; G = Y^T Y, H = Y^T W ; host Rayleigh-Ritz ; V = Y M, AV = W M, residuals
    sub_apply<KB>(c, Fp, cur, nullptr, sW, one, zero, zero, ns, outer > 0);
    sub_gram_kernel<KB><<<dim3(NBLK, ns), 256, 0, c->stream>>>(cur, sW, sGpart, sG, ticket, n);
    LAUNCH_CHECK(c);
    const double* G = d2h_small(c, 0, sG, (size_t)ns * 2 * KB * KB);
    NBD_CUDA(cudaStreamSynchronize(c->stream));
    if (outer == 0) sub_bounds_finish(c, Fp, hb, bound, lowb, s0, ns);  // G was copied before any Lanczos run reuses slot 2
    for (int s = 0; s < ns; ++s)
      if (!sub_rayleigh_ritz(KB, G + (size_t)s * 2 * KB * KB, G + (size_t)s * 2 * KB * KB + KB * KB,
                             M.data() + (size_t)s * KB * KB, th.data() + (size_t)s * KB))
        return false;
    // a Ritz value above the assumed upper end of the spectrum: the filter interval was wrong, let the library decide
    for (int s = 0; s < ns; ++s)
      if (outer > 0 && th[(size_t)s * KB + KB - 1] > bu_used[s]) {
        c->sub_bounds_valid = false;
        return false;
      }
    h2d(c, sM, M.data(), M.size());
    h2d(c, sTheta, th.data(), th.size());
    // rotate into a buffer that is not `cur` (cur may alias sV / sY / sZ); if that buffer is sY, sV and sY trade places
    double* vout = (cur == sV) ? sY : sV;
    sub_rotate_kernel<KB><<<dim3(NBLK, ns), 256, 0, c->stream>>>(cur, sW, sM, sTheta, vout, sAV, sRpart, n);
    LAUNCH_CHECK(c);
    if (vout != sV) {
      std::swap(c->sV.p, c->sY.p);
      std::swap(c->sV.cap, c->sY.cap);
    }
    const double* rp = d2h_small(c, 0, sRpart, (size_t)ns * NBLK * KB);
    NBD_CUDA(cudaStreamSynchronize(c->stream));
    double worst = 0.0;
    for (int s = 0; s < ns; ++s) {
      const int o = c->nspin == 2 ? c->nelec[s0 + s] : (c->nelec[0] + c->nelec[1]) / 2;
      for (int k = 0; k < KB; ++k) c->sub_theta[s0 + s][k] = th[(size_t)s * KB + k];
      for (int k = 0; k < o; ++k) {
        double r2 = 0.0;
        for (int bl = 0; bl < NBLK; ++bl) r2 += rp[((size_t)s * NBLK + bl) * KB + k];
        worst = std::max(worst, std::sqrt(r2));
      }
    }
    ++c->sub_outer;
    if (getenv("NBD_SUB_DEBUG")) fprintf(stderr, "[sub] s0=%d ns=%d outer=%d cold=%d degree=%d worst=%.3e rate=%.4f\n", s0, ns, outer, (int)cold, outer > 0 ? deg_prev : 0, worst, c->sub_rate);
    cur = c->sV.p + s0 * blk;
    if (worst < tol) return true;
    if (!cold && outer > 0 && deg_prev > 0 && worst_prev > 0.0 && worst < worst_prev)
      c->sub_rate = std::max(c->sub_rate, std::pow(worst / worst_prev, 1.0 / deg_prev));
    worst_prev = worst;
  }
  return false;
}

// With a communicator of >= 2 ranks and two spins, rank 0 tracks the alpha block and rank 1 the beta block - each block
// product then moves half the data of the joint one - and the converged blocks (+ Ritz values and a status word) are
// broadcast, so every rank continues with bit-identical orbitals.  The collectives are issued on every rank whatever
// the local outcome; a failure on either owner sends all ranks to the library eigensolver together.
template <int KB>
static bool sub_solve_split(nbd_ctx* c, const double* Fp) {
  const int n = c->nao;
  const long blk = (long)n * KB;
  bool ok_local = true;
  if (c->rank < 2) ok_local = sub_solve_t<KB>(c, Fp, c->rank, 1);
  else c->sub_is_cold = false;
  StageScope ts(c->timers, c->stream, "eig_bcast");
  double* x = c->sXchg.ensure((size_t)2 * (KB + 2));
  if (c->rank < 2) {
    double hx[KB + 2];
    for (int k = 0; k < KB; ++k) hx[k] = c->sub_theta[c->rank][k];
    hx[KB] = ok_local ? 1.0 : 0.0;
    hx[KB + 1] = c->sub_bounds_valid ? 1.0 : 0.0;
    h2d(c, x + (long)c->rank * (KB + 2), hx, KB + 2);
  }
  ncclResult_t r = g_nccl.GroupStart();
  for (int s = 0; s < 2 && r == ncclSuccess; ++s) {
    r = g_nccl.Broadcast(c->sV.p + s * blk, c->sV.p + s * blk, (size_t)blk, ncclDouble, s, c->comm, c->stream);
    if (r == ncclSuccess) r = g_nccl.Broadcast(x + (long)s * (KB + 2), x + (long)s * (KB + 2), (size_t)(KB + 2), ncclDouble, s, c->comm, c->stream);
  }
  if (r == ncclSuccess) r = g_nccl.GroupEnd();
  if (r != ncclSuccess) fail(NBD_ERR_CUDA, "ncclBroadcast (eigenvector blocks): %s", g_nccl.GetErrorString(r));
  double hx[2 * (KB + 2)];
  d2h(c, hx, x, 2 * (KB + 2));
  NBD_CUDA(cudaStreamSynchronize(c->stream));
  bool ok = true;
  for (int s = 0; s < 2; ++s) {
    for (int k = 0; k < KB; ++k) c->sub_theta[s][k] = hx[s * (KB + 2) + k];
    ok = ok && hx[s * (KB + 2) + KB] == 1.0;
  }
  return ok;
}

static bool sub_solve(nbd_ctx* c, const double* Fp) {
  StageScope ts(c->timers, c->stream, "eig_sub");
  const bool split = c->world >= 2 && c->comm && c->nspin == 2 && c->dist_sub;
  if (c->sub_kb == 16) return split ? sub_solve_split<16>(c, Fp) : sub_solve_t<16>(c, Fp, 0, c->nspin);
  if (c->sub_kb == 32) return split ? sub_solve_split<32>(c, Fp) : sub_solve_t<32>(c, Fp, 0, c->nspin);
  return false;
}
