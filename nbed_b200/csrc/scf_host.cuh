// Embedded-SCF drivers behind the C-ABI: static set-up, the Huzinaga loop (nbed/scf/huzinaga_scf.py:93-206),
// the mu-shift loop (pyscf.scf.hf.kernel as driven by nbed/driver.py:500-538) and the benchmark stepper.
// Included by nbed_b200.cu (single translation unit).
#pragma once

// ---- helpers ------------------------------------------------------------------------------------
static int scf_nocc(const nbd_ctx* c, int s) { return c->nspin == 2 ? c->nelec[s] : (c->nelec[0] + c->nelec[1]) / 2; }
static double scf_occ(const nbd_ctx* c) { return c->nspin == 2 ? 1.0 : 2.0; }

// occupied rows of Ct -> scaled orbital block for J/K;  returns Ntot
static int scf_stage_occupied(nbd_ctx* c) {
  const int n = c->nao;
  int tot = 0;
  for (int s = 0; s < c->nspin; ++s) tot += scf_nocc(c, s);
  c->d_orb.ensure((size_t)std::max(1, tot) * c->n_ld);
  c->d_wt.ensure((size_t)std::max(1, tot) * c->n_ld);
  int row = 0;
  for (int s = 0; s < c->nspin; ++s) {
    const int o = scf_nocc(c, s);
    stage_orbitals(c, c->Ct.p + (long)s * n * n, n, o, row, std::sqrt(scf_occ(c)), nullptr, 1.0);
    row += o;
  }
  return tot;
}

// J/K of the staged orbital block -> vhf_s = J - kscale K_s and F_s = heff_s + vhf_s.
// Kohn-Sham objects (c->xc.on; pyscf/dft/uks.py:get_veff as reached from huzinaga_scf.py:55,156): kscale = the hybrid
// fraction, vhf_s += V_xc,s[D] for the density in c->D (the one the staged orbitals belong to), and the scalars the
// reference reads from the result's tags: ecoul = tr((D_a + D_b) J) / 2, exc = int f - hyb sum_s tr(D_s K_s) / 2.
static void scf_build_fock(nbd_ctx* c, int Ntot, const std::vector<KGroup>& groups) {
  const int n = c->nao;
  const long nn = (long)n * n;
  const int ns = c->nspin;
  const bool ks = c->xc.on;
  double* buf = c->d_jk.ensure((size_t)(1 + ns) * nn);
  std::vector<int> jbegin = {0, Ntot};
  jk_device(c, c->d_orb.p, c->d_wt.p, Ntot, 1, jbegin, buf, ns, groups, buf + nn);
  // UHF: J - K_s; RHF: J - K / 2; UKS: J - hyb K_s; RKS: J - hyb K / 2   (pyscf/dft/{uks,rks}.py:get_veff)
  const double kscale = (ks ? c->xc.hyb : 1.0) * (ns == 2 ? 1.0 : 0.5);
  // (option packed_allreduce = 2 takes this route on a single rank as well, where the all-reduce is a no-op: test hook)
  if (((c->world > 1 && c->comm && c->packed_allreduce) || c->packed_allreduce == 2) && !ks && n >= 256) {
    // symmetric partial sums: the ranks exchange the lower triangles only (half the all-reduce bytes), and the Fock
    // assembly reads the packed sums directly.  (Kohn-Sham objects keep the square buffers: their traces read J and K.)
    const long npack = (long)n * (n + 1) / 2;
    double* pk = c->d_jkpack.ensure((size_t)(1 + ns) * npack);
    {
      StageScope ts(c->timers, c->stream, "allreduce");
      pack_lower_kernel<<<dim3((n + 127) / 128, n, 1 + ns), 128, 0, c->stream>>>(buf, pk, n, npack);
      LAUNCH_CHECK(c);
    }
    all_reduce(c, pk, (size_t)(1 + ns) * npack);
    StageScope ts(c->timers, c->stream, "fock");
    fock_from_heff_packed_kernel<<<dim3((n + 127) / 128, n), 128, 0, c->stream>>>(c->heff.p, pk, kscale, c->F.p, c->vhf.p, n, npack, ns);
    LAUNCH_CHECK(c);
    return;
  }
  all_reduce(c, buf, (size_t)(1 + ns) * nn);
  {
    StageScope ts(c->timers, c->stream, "fock");
    fock_from_heff_kernel<<<grid1(nn, 256), 256, 0, c->stream>>>(c->heff.p, buf, buf + nn, kscale, c->F.p, c->vhf.p, nn, ns);
    LAUNCH_CHECK(c);
  }
  if (!ks) return;
  double in3[3];
  double* out = c->red_out.ensure(64);
  if (ns == 1) {
    // restricted Kohn-Sham object (pyscf/dft/rks.py:get_veff; the object type of the reference's tests/test_scf.py:19-40):
    // nr_rks of the total density = nr_uks of the spin-unpolarised pair (D / 2, D / 2) - same energy, V_xc of either spin
    double* pair = c->dm0f.ensure((size_t)2 * nn);
    halve_to_pair_kernel<<<grid1(nn, 256), 256, 0, c->stream>>>(c->D.p, pair, nn);
    LAUNCH_CHECK(c);
    xc_eval_device(c, pair, in3);
    StageScope ts(c->timers, c->stream, "fock");
    xc_add_potential_kernel<<<grid1(nn, 256), 256, 0, c->stream>>>(c->xc.V.p, c->F.p, c->vhf.p, nn);
    LAUNCH_CHECK(c);
    reduce_to(c, c->D.p, buf, nn, 0, n, out + 52);       // tr(D J)
    reduce_to(c, c->D.p, buf + nn, nn, 0, n, out + 53);  // tr(D K)
    double h[2];
    d2h(c, h, out + 52, 2);
    NBD_CUDA(cudaStreamSynchronize(c->stream));
    c->xc.ecoul = 0.5 * h[0];
    c->xc.exc = in3[0] - 0.25 * c->xc.hyb * h[1];
    return;
  }
  xc_eval_device(c, c->D.p, in3);
  StageScope ts(c->timers, c->stream, "fock");
  xc_add_potential_kernel<<<grid1(2 * nn, 256), 256, 0, c->stream>>>(c->xc.V.p, c->F.p, c->vhf.p, 2 * nn);
  LAUNCH_CHECK(c);
  reduce_to(c, c->D.p, buf, nn, 0, n, out + 52);             // tr(D_a J)
  reduce_to(c, c->D.p + nn, buf, nn, 0, n, out + 53);        // tr(D_b J)
  reduce_to(c, c->D.p, buf + nn, 2 * nn, 0, n, out + 54);    // sum_s tr(D_s K_s)
  double h[3];
  d2h(c, h, out + 52, 3);
  NBD_CUDA(cudaStreamSynchronize(c->stream));
  c->xc.ecoul = 0.5 * (h[0] + h[1]);
  c->xc.exc = in3[0] - 0.5 * c->xc.hyb * h[2];
}
static std::vector<KGroup> scf_occ_groups(const nbd_ctx* c) {
  std::vector<KGroup> g;
  int row = 0;
  for (int s = 0; s < c->nspin; ++s) {
    const int o = scf_nocc(c, s);
    g.push_back(KGroup{s, row, row + o, 1.0});
    row += o;
  }
  return g;
}

// Huz_s = -cfac (F_s gammaS_s + (F_s gammaS_s)^T);  F_s += Huz_s      (huzinaga_scf.py:65-90,159-160)
static void scf_apply_huzinaga(nbd_ctx* c) {
  StageScope ts(c->timers, c->stream, "fock");
  const int n = c->nao;
  const long nn = (long)n * n;
  if (c->env_rank > 0) {
    // gamma = V V^T has rank r << n: F gamma S = (F V)(V^T S), 4 n^2 r flop instead of 2 n^3  (r = 155 of 1376 at C4)
    const int r = c->env_rank;
    gemm_nt(c, n, r, n, c->F.p, n, c->envV.p, n, c->envW.p, r, 1.0, 0.0, c->nspin, nn, (long)r * n, (long)n * r);
    gemm_nn(c, n, n, r, c->envW.p, r, c->envZ.p, n, c->FG.p, n, 1.0, 0.0, c->nspin, (long)n * r, (long)r * n, nn);
  } else {
    gemm_nn(c, n, n, n, c->F.p, n, c->GS.p, n, c->FG.p, n, 1.0, 0.0, c->nspin, nn, nn, nn);
  }
  const double* FGv = nullptr;
  const double* W = nullptr;
  if (c->have_virt) {  // virtual-orbital projector (huzinaga_scf.py:82-88): FGv = F GSv, W = GSv^T FGv
    gemm_nn(c, n, n, n, c->F.p, n, c->GSv.p, n, c->T1.p, n, 1.0, 0.0, c->nspin, nn, nn, nn);
    gemm_tn(c, n, n, n, c->GSv.p, n, c->T1.p, n, c->T2.p, n, 1.0, 0.0, c->nspin, nn, nn, nn);
    FGv = c->T1.p;
    W = c->T2.p;
  }
  dim3 g((n + 31) / 32, (n + 31) / 32, c->nspin), b(32, 8);
  huzinaga_apply_kernel<<<g, b, 0, c->stream>>>(c->FG.p, FGv, W, c->nspin == 2 ? 1.0 : 0.5, c->Huz.p, c->F.p, n);
  LAUNCH_CHECK(c);
}

// F' = X F X ; eigh ; Ct = V X (rows = MOs)       (huzinaga_scf.py:166-169)
// allow_subspace: between full cuSOLVER solves the occupied block is tracked by filtered subspace iteration
// (subspace.cuh); c->last_eig_full tells the caller whether Ct / evals hold the complete spectrum.
// allow_cold: with no tracked block yet, start one from pseudo-random vectors instead of a library eigensolve.
static void scf_diagonalise_lowdin(nbd_ctx* c, bool allow_subspace = false, bool allow_cold = false) {
  const int n = c->nao;
  const long nn = (long)n * n;
  if (c->world > 1 && c->comm && c->dist_orth && n >= 256) {
    // Row-distributed: rank r forms rows R_r of F' = (X[R_r, :] F) X - two skinny GEMMs, no exchange in between -
    // and ONE in-place all-gather per spin assembles F' on every rank (bit-identical everywhere: every row block is
    // computed by exactly one rank).  Replaces 4 n^3 replicated flop per spin by 4 n^3 / N + 8 n^2 bytes of NVLink.
    const int rpr = (n + c->world - 1) / c->world, r0 = std::min(n, c->rank * rpr), rows = std::min(n, r0 + rpr) - r0;
    const long pad = (long)c->world * rpr * n;
    // when the row blocks tile the matrix exactly the all-gather assembles F' in place, otherwise through a padded copy
    const bool in_place = pad == nn;
    double* Fp = in_place ? c->T2.p : c->Fpad.ensure((size_t)c->nspin * pad);
    {
      StageScope ts(c->timers, c->stream, "orth");
      if (rows > 0) {
        gemm_nn(c, rows, n, n, c->Xh.p + (long)r0 * n, n, c->F.p, n, c->T1.p, n, 1.0, 0.0, c->nspin, 0, nn, nn);
        gemm_nn(c, rows, n, n, c->T1.p, n, c->Xh.p, n, Fp + (long)r0 * n, n, 1.0, 0.0, c->nspin, nn, 0, pad);
      }
    }
    {
      StageScope ts(c->timers, c->stream, "orth_gather");
      ncclResult_t r = g_nccl.GroupStart();
      for (int s = 0; s < c->nspin && r == ncclSuccess; ++s)
        r = g_nccl.AllGather(Fp + s * pad + (long)c->rank * rpr * n, Fp + s * pad, (size_t)rpr * n, ncclDouble, c->comm, c->stream);
      if (r == ncclSuccess) r = g_nccl.GroupEnd();
      if (r != ncclSuccess) fail(NBD_ERR_CUDA, "ncclAllGather (F'): %s", g_nccl.GetErrorString(r));
      for (int s = 0; s < c->nspin && !in_place; ++s)
        NBD_CUDA(cudaMemcpyAsync(c->T2.p + s * nn, Fp + s * pad, sizeof(double) * nn, cudaMemcpyDeviceToDevice, c->stream));
    }
  } else {
    StageScope ts(c->timers, c->stream, "orth");
    gemm_nn(c, n, n, n, c->F.p, n, c->Xh.p, n, c->T1.p, n, 1.0, 0.0, c->nspin, nn, 0, nn);
    // F' is symmetric: only the lower tiles of X (F X) are computed, then mirrored
    gemm(c, n, n, n, c->Xh.p, n, 1, c->T1.p, 1, n, c->T2.p, n, 1.0, 0.0, c->nspin, 0, nn, nn, /*lower=*/1);
    symmetrize_lower(c, c->T2.p, n, c->nspin);
  }
  if (allow_cold && c->sub_cold && !c->sub_valid) {
    sub_init_cold(c);
    allow_subspace = c->sub_valid;
  }
  if (allow_subspace && c->sub_valid && sub_block_size(c) == c->sub_kb) {
    if (sub_solve(c, c->T2.p)) {
      StageScope ts(c->timers, c->stream, "orth");
      const int kb = c->sub_kb;
      // Ct[0..kb) = V^T X  (rows = the kb lowest MOs), eps[0..kb) = Ritz values
      gemm(c, kb, n, n, c->sV.p, 1, kb, c->Xh.p, 1, n, c->Ct.p, n, 1.0, 0.0, c->nspin, (long)n * kb, 0, nn);
      for (int s = 0; s < c->nspin; ++s) h2d(c, c->evals.p + (long)s * n, c->sub_theta[s], kb);
      c->last_eig_full = false;
      return;
    }
    ++c->sub_fallbacks;  // not converged: fall through to the library eigensolver on the same F'
  }
  eigh_batched(c, c->T2.p, c->evals.p, n, c->nspin, /*warm_slot=*/1);
  {
    StageScope ts(c->timers, c->stream, "orth");
    gemm_nn(c, n, n, n, c->T2.p, n, c->Xh.p, n, c->Ct.p, n, 1.0, 0.0, c->nspin, nn, 0, nn);
  }
  c->last_eig_full = true;
  sub_init_from_full(c, c->T2.p, c->evals.p);
}

// complete (C, eps) of the last Fock matrix when the last cycle only tracked the occupied block
static void scf_complete_spectrum(nbd_ctx* c) {
  if (c->last_eig_full) return;
  scf_diagonalise_lowdin(c, false);
}

// D_s = occ * sum_{i < o_s} c_i c_i^T  (old D kept in Dold)     (huzinaga_scf.py:170-174)
static void scf_make_density(nbd_ctx* c) {
  StageScope ts(c->timers, c->stream, "density");
  const int n = c->nao;
  const long nn = (long)n * n;
  std::swap(c->D.p, c->Dold.p);
  std::swap(c->D.cap, c->Dold.cap);
  for (int s = 0; s < c->nspin; ++s) {
    const int o = scf_nocc(c, s);
    if (o == 0) {
      NBD_CUDA(cudaMemsetAsync(c->D.p + s * nn, 0, sizeof(double) * nn, c->stream));
      continue;
    }
    gemm_tn(c, n, n, o, c->Ct.p + s * nn, n, c->Ct.p + s * nn, n, c->D.p + s * nn, n, scf_occ(c), 0.0, 1, 0, 0, 0, /*lower=*/1);
  }
  symmetrize_lower(c, c->D.p, n, c->nspin);
}

// out8 (host) = per spin [a.D, b.D, c.D, |D - Dold|^2]
static void scf_traces(nbd_ctx* c, const double* a, const double* b, const double* cc, bool with_dd, double* out8) {
  StageScope ts(c->timers, c->stream, "energy");
  const long nn = (long)c->nao * c->nao;
  double* part = c->red_part.ensure((size_t)REDUCE_BLOCKS * 9);
  double* out = c->red_out.ensure(64);
  scf_traces_partial_kernel<<<REDUCE_BLOCKS, 256, 0, c->stream>>>(a, b, cc, c->D.p, with_dd ? c->Dold.p : nullptr, nn, c->nspin, part);
  LAUNCH_CHECK(c);
  scf_traces_final_kernel<<<1, 256, 0, c->stream>>>(part, REDUCE_BLOCKS, 4 * c->nspin, out);
  LAUNCH_CHECK(c);
  const double* h = d2h_small(c, 3, out, 4 * c->nspin);
  NBD_CUDA(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 4 * c->nspin; ++i) out8[i] = h[i];
}

// ---- DIIS (pyscf/lib/diis.py:DIIS.update; CDIIS pushes its own error vector) ------------------------
// x: device vector of d.len doubles, overwritten with the extrapolated vector when one is produced.
static void diis_update(nbd_ctx* c, DiisState& d, double* x, const double* errvec) {
  StageScope ts(c->timers, c->stream, "diis");
  const long len = d.len;
  const size_t bytes = sizeof(double) * len;
  if (errvec) {  // push_err_vec
    if (d.head >= d.space) d.head = 0;
    NBD_CUDA(cudaMemcpyAsync(d.e.p + (long)d.head * len, errvec, bytes, cudaMemcpyDeviceToDevice, c->stream));
  }
  // push_vec
  while ((int)d.bookkeep.size() >= d.space) d.bookkeep.erase(d.bookkeep.begin());
  if (errvec) {
    d.bookkeep.push_back(d.head);
    NBD_CUDA(cudaMemcpyAsync(d.x.p + (long)d.head * len, x, bytes, cudaMemcpyDeviceToDevice, c->stream));
    d.head += 1;
  } else if (!d.have_xprev) {
    NBD_CUDA(cudaMemcpyAsync(d.xprev.p, x, bytes, cudaMemcpyDeviceToDevice, c->stream));
    d.have_xprev = true;
  } else {
    if (d.head >= d.space) d.head = 0;
    d.bookkeep.push_back(d.head);
    NBD_CUDA(cudaMemcpyAsync(d.x.p + (long)d.head * len, x, bytes, cudaMemcpyDeviceToDevice, c->stream));
    sub_kernel<<<grid1(len, 256), 256, 0, c->stream>>>(x, d.xprev.p, d.e.p + (long)d.head * len, len);
    LAUNCH_CHECK(c);
    d.head += 1;
  }
  const int nd = (int)d.bookkeep.size();
  if (nd < 1) return;  // min_space = 1
  // Gram row of the newest error vector against all stored ones
  MultiDotArgs ma{};
  ma.nd = nd;
  for (int i = 0; i < nd; ++i) ma.ys[i] = d.e.p + (long)i * len;
  double* part = c->red_part.ensure((size_t)REDUCE_BLOCKS * 9);
  double* out = c->red_out.ensure(64);
  multidot_partial_kernel<<<REDUCE_BLOCKS, 256, 0, c->stream>>>(d.e.p + (long)(d.head - 1) * len, ma, len, part);
  LAUNCH_CHECK(c);
  multidot_final_kernel<<<1, 256, 0, c->stream>>>(part, REDUCE_BLOCKS, nd, out + 16);
  LAUNCH_CHECK(c);
  const double* hrow = d2h_small(c, 4, out + 16, nd);
  NBD_CUDA(cudaStreamSynchronize(c->stream));
  const int ldh = d.space + 1;
  for (int i = 0; i < nd; ++i) {
    d.H[(size_t)d.head * ldh + i + 1] = hrow[i];
    d.H[(size_t)(i + 1) * ldh + d.head] = hrow[i];
  }
  std::vector<double> coef = diis_coefficients(d.H, ldh, nd);
  LinCombArgs la{};
  la.nd = nd;
  for (int i = 0; i < nd; ++i) {
    la.xs[i] = d.x.p + (long)i * len;
    la.coef[i] = coef[i + 1];
  }
  lincomb_kernel<<<grid1(len, 256), 256, 0, c->stream>>>(la, x, len);
  LAUNCH_CHECK(c);
  if (d.have_xprev) NBD_CUDA(cudaMemcpyAsync(d.xprev.p, x, bytes, cudaMemcpyDeviceToDevice, c->stream));
}

// ---- static set-up -----------------------------------------------------------------------------------
extern "C" int nbd_scf_setup(nbd_ctx* c, int nspin, const int* nelec, const double* ovlp, const double* hcore,
                             const double* v_emb, const double* dm_env, int projector, double mu) {
  return guarded(c, [&] {
    c->timers.reset();
    NBD_REQUIRE(c->Bt, NBD_ERR_STATE, "nbd_cderi_alloc first (nao is taken from the 3-centre tensor)");
    NBD_REQUIRE((nspin == 1 || nspin == 2) && nelec && ovlp && hcore && v_emb && dm_env, NBD_ERR_ARG, "bad nspin / pointers");
    NBD_REQUIRE(projector == NBD_HUZINAGA || projector == NBD_MU_SHIFT, NBD_ERR_ARG, "unknown projector %d", projector);
    const int n = c->nao;
    const long nn = (long)n * n;
    NBD_REQUIRE(nelec[0] >= 0 && nelec[1] >= 0 && nelec[0] <= n && nelec[1] <= n, NBD_ERR_ARG, "nelec out of range");
    c->nspin = nspin;
    c->nelec[0] = nelec[0];
    c->nelec[1] = nelec[1];
    c->projector = projector;
    c->mu = mu;
    c->S.ensure(nn); c->Xh.ensure(nn); c->hcore.ensure(nn);
    for (DBuf<double>* b : {&c->heff, &c->GS, &c->F, &c->Huz, &c->vhf, &c->FG, &c->T1, &c->T2, &c->Ct, &c->D, &c->Dold, &c->Corth})
      b->ensure((size_t)2 * nn);
    c->evals.ensure((size_t)8 * n);
    // The same molecule is embedded several times (DFT-in-DFT and HF-in-DFT, both projectors: nbed/driver.py:1138-1231):
    // when the uploaded overlap matrix is bit-identical to the one X = S^-1/2 was last built from, X is reused and the
    // eigendecomposition (17.6 ms at n = 1376) is skipped.  Option "x_cache" = 0 switches the reuse off.
    bool reuse_x = false;
    if (c->x_cache && c->x_valid_n == n) {
      h2d(c, c->FG.p, ovlp, nn);
      double* out = c->red_out.ensure(64);
      reduce_to(c, c->FG.p, c->S.p, nn, 1, n, out + 48);
      double diff = 1.0;
      d2h(c, &diff, out + 48, 1);
      NBD_CUDA(cudaStreamSynchronize(c->stream));
      reuse_x = diff == 0.0;
    }
    c->x_valid_n = 0;
    if (!reuse_x) h2d(c, c->S.p, ovlp, nn);
    h2d(c, c->hcore.p, hcore, nn);
    h2d(c, c->T1.p, v_emb, (size_t)nspin * nn);   // V_emb
    h2d(c, c->T2.p, dm_env, (size_t)nspin * nn);  // gamma_env
    if (!reuse_x) {
      // X = S^-1/2 = (w^-1/4 V)^T (w^-1/4 V)     (huzinaga_scf.py:128)
      NBD_CUDA(cudaMemcpyAsync(c->FG.p, c->S.p, sizeof(double) * nn, cudaMemcpyDeviceToDevice, c->stream));
      eigh_batched(c, c->FG.p, c->evals.p, n, 1);
      {
        std::vector<double> w(n);
        d2h(c, w.data(), c->evals.p, n);
        check_devinfo(c, 1, "overlap eigendecomposition");
        NBD_REQUIRE(w[0] > 0.0, NBD_ERR_ARG, "overlap matrix is not positive definite (lowest eigenvalue %g)", w[0]);
      }
      {
        dim3 g((n + 127) / 128, n);
        scale_rows_kernel<<<g, 128, 0, c->stream>>>(c->FG.p, c->evals.p, n, 2);
        LAUNCH_CHECK(c);
      }
      gemm_tn(c, n, n, n, c->FG.p, n, c->FG.p, n, c->Xh.p, n, 1.0, 0.0, 1, 0, 0, 0, /*lower=*/1);
      symmetrize_lower(c, c->Xh.p, n, 1);
    }
    // gamma S (huzinaga_scf.py:132)
    gemm_nn(c, n, n, n, c->T2.p, n, c->S.p, n, c->GS.p, n, 1.0, 0.0, nspin, nn, 0, nn);
    if (projector == NBD_MU_SHIFT) {
      // P_s = S gamma_s S (driver.py:433-449); heff = h + V + mu P (driver.py:518,529)
      gemm_nn(c, n, n, n, c->S.p, n, c->GS.p, n, c->FG.p, n, 1.0, 0.0, nspin, 0, nn, nn);
      heff_kernel<<<grid1(nn, 256), 256, 0, c->stream>>>(c->hcore.p, c->T1.p, c->FG.p, mu, c->heff.p, nn, nspin);
    } else {
      heff_kernel<<<grid1(nn, 256), 256, 0, c->stream>>>(c->hcore.p, c->T1.p, nullptr, 0.0, c->heff.p, nn, nspin);
    }
    LAUNCH_CHECK(c);
    c->scf_ready = true;
    c->se_warm_calls = 0;  // the small-matrix eigensolver's warm basis never survives a change of problem
    c->bench_ready = false;
    c->have_virt = false;
    c->env_rank = 0;
    c->sub_valid = false;  // a tracked eigenvector block never survives a change of problem
    c->sub_bounds_valid = false;
    c->last_eig_full = true;
    finish_call(c);
    c->x_valid_n = n;  // S and X of this problem are complete on the device
  });
}

// Optional virtual-orbital environment projector of huzinaga_scf (dm_environment_virtual, huzinaga_scf.py:133-136,
// 82-88): dm_env_virt [nspin][nao][nao]; null or all-zero switches it off.  Call after nbd_scf_setup.
extern "C" int nbd_scf_set_virtual_projector(nbd_ctx* c, const double* dm_env_virt) {
  return guarded(c, [&] {
    NBD_REQUIRE(c->scf_ready && c->projector == NBD_HUZINAGA, NBD_ERR_STATE, "nbd_scf_setup(projector = NBD_HUZINAGA) first");
    const int n = c->nao;
    const long nn = (long)n * n;
    c->have_virt = false;
    if (dm_env_virt) {
      c->GSv.ensure((size_t)2 * nn);
      h2d(c, c->T2.p, dm_env_virt, (size_t)c->nspin * nn);
      gemm_nn(c, n, n, n, c->T2.p, n, c->S.p, n, c->GSv.p, n, 1.0, 0.0, c->nspin, nn, 0, nn);  // gamma_virt S  (:134)
      c->have_virt = true;
    }
    c->bench_ready = false;
    finish_call(c);
  });
}

// Optional low-rank factor of the occupied environment density: dm_env_s = c_env_s c_env_s^T with c_env [nspin][nao][r]
// (the localizer's c_enviro, nbed/localizers/system.py:33: dm_enviro = c_enviro c_enviro^T).  The projector product
// F gamma S of get_huzinaga_operator (huzinaga_scf.py:77) then runs as two rank-r GEMMs.  The factor is checked against
// the gamma S of nbd_scf_setup on the device; a factor that does not reproduce it is refused (NBD_ERR_ARG) and the
// dense product stays in use.  Call after nbd_scf_setup.
extern "C" int nbd_scf_set_env_orbitals(nbd_ctx* c, int r, const double* c_env) {
  return guarded(c, [&] {
    NBD_REQUIRE(c->scf_ready && c->projector == NBD_HUZINAGA, NBD_ERR_STATE, "nbd_scf_setup(projector = NBD_HUZINAGA) first");
    NBD_REQUIRE(r >= 1 && c_env, NBD_ERR_ARG, "bad rank / pointer");
    const int n = c->nao, ns = c->nspin;
    const long nn = (long)n * n;
    c->env_rank = 0;
    if (2 * r > n) {  // no saving: keep the dense product
      finish_call(c);
      return;
    }
    c->envV.ensure((size_t)ns * r * n);
    c->envZ.ensure((size_t)ns * r * n);
    c->envW.ensure((size_t)ns * n * r);
    h2d(c, c->envW.p, c_env, (size_t)ns * n * r);  // [ns][n][r]
    {
      dim3 g((r + 31) / 32, (n + 31) / 32, ns), b(32, 8);
      transpose_kernel<<<g, b, 0, c->stream>>>(c->envW.p, c->envV.p, n, r);  // -> [ns][r][n]
      LAUNCH_CHECK(c);
    }
    gemm_nn(c, r, n, n, c->envV.p, n, c->S.p, n, c->envZ.p, n, 1.0, 0.0, ns, (long)r * n, 0, (long)r * n);
    // check: V (V^T S) == gamma S
    gemm_tn(c, n, n, r, c->envV.p, n, c->envZ.p, n, c->T1.p, n, 1.0, 0.0, ns, (long)r * n, (long)r * n, nn);
    double* out = c->red_out.ensure(64);
    reduce_to(c, c->T1.p, c->GS.p, ns * nn, 1, n, out + 40);
    reduce_to(c, c->GS.p, c->GS.p, ns * nn, 0, n, out + 41);
    double h[2];
    d2h(c, h, out + 40, 2);
    NBD_CUDA(cudaStreamSynchronize(c->stream));
    NBD_REQUIRE(h[0] <= 1e-20 * std::max(h[1], 1e-300), NBD_ERR_ARG,
                "c_env does not factor dm_env: |c c^T S - dm S|_F / |dm S|_F = %.3e", std::sqrt(h[0] / std::max(h[1], 1e-300)));
    c->env_rank = r;
    c->bench_ready = false;
    finish_call(c);
  });
}

// ---- Huzinaga loop -----------------------------------------------------------------------------------

// initial guess of huzinaga_scf.py:139-148 (dm0 == null) or a dense user density
static void huz_initial(nbd_ctx* c, const double* dm0, HuzLoop& L) {
  const int n = c->nao;
  const long nn = (long)n * n;
  c->sub_valid = false;  // a new SCF never inherits the eigenvector block or the spectral bounds of the last one
  c->sub_bounds_valid = false;
  c->sub_rate = 0.0;
  c->se_warm_calls = 0;  // ... nor the warm basis of the small-matrix eigensolver (repeated runs stay bit-identical)
  if (!dm0) {
    NBD_CUDA(cudaMemcpyAsync(c->F.p, c->heff.p, sizeof(double) * nn * c->nspin, cudaMemcpyDeviceToDevice, c->stream));
    scf_apply_huzinaga(c);
    scf_diagonalise_lowdin(c, false, true);
    scf_make_density(c);
    L.Ntot = scf_stage_occupied(c);
    L.groups = scf_occ_groups(c);
  } else {
    h2d(c, c->D.p, dm0, (size_t)c->nspin * nn);
    double* tmp = c->dm0f.ensure((size_t)c->nspin * nn);
    NBD_CUDA(cudaMemcpyAsync(tmp, c->D.p, sizeof(double) * nn * c->nspin, cudaMemcpyDeviceToDevice, c->stream));
    std::vector<int> jb;
    L.Ntot = factor_densities(c, tmp, c->nspin, jb, L.groups);
  }
}

// one pass of the loop body (huzinaga_scf.py:154-201); out: per-spin energies and max_s ||dD_s||_F
// Kohn-Sham objects (:176-180, calculate_ks_energy :36-62): the energy needs get_veff of the NEW density; the reference
// then builds the same get_veff again at the top of the next cycle - here it is built once and reused (same numbers).
static void huz_iteration(nbd_ctx* c, int iter, bool use_diis, HuzLoop& L, double* energy, double* norm_ddm) {
  const bool ks = c->xc.on;
  if (!(ks && L.veff_valid)) scf_build_fock(c, L.Ntot, L.groups);     // :156-157
  scf_apply_huzinaga(c);                                              // :159-160
  if (use_diis && iter > 1) diis_update(c, c->diis, c->F.p, nullptr);  // :162-164
  scf_diagonalise_lowdin(c, true, true);                              // :166-169
  scf_make_density(c);                                                // :170-174
  L.Ntot = scf_stage_occupied(c);
  L.groups = scf_occ_groups(c);
  if (ks) {
    scf_build_fock(c, L.Ntot, L.groups);                              // :55  vhf_updated = get_veff(dm = new density)
    L.veff_valid = true;
  }
  double t[8];
  scf_traces(c, c->heff.p, c->vhf.p, c->Huz.p, true, t);              // :182-194
  double nd = 0.0;
  for (int s = 0; s < c->nspin; ++s) {
    energy[s] = ks ? (c->xc.ecoul + c->xc.exc) + t[4 * s + 0] + t[4 * s + 2]   // :56-61
                   : t[4 * s + 0] + 0.5 * t[4 * s + 1] + t[4 * s + 2];
    nd = std::max(nd, std::sqrt(t[4 * s + 3]));
  }
  *norm_ddm = nd;
}

static void scf_export(nbd_ctx* c, double* mo_coeff, double* mo_energy, double* dm, double* extra, const double* d_extra) {
  const int n = c->nao;
  const long nn = (long)n * n;
  if (mo_coeff) {
    dim3 g((n + 31) / 32, (n + 31) / 32, c->nspin), b(32, 8);
    transpose_kernel<<<g, b, 0, c->stream>>>(c->Ct.p, c->T1.p, n, n);
    LAUNCH_CHECK(c);
    d2h(c, mo_coeff, c->T1.p, (size_t)c->nspin * nn);
  }
  if (mo_energy) d2h(c, mo_energy, c->evals.p, (size_t)c->nspin * n);
  if (dm) d2h(c, dm, c->D.p, (size_t)c->nspin * nn);
  if (extra) d2h(c, extra, d_extra, (size_t)c->nspin * nn);
}

extern "C" int nbd_huzinaga_scf(nbd_ctx* c, int max_cycle, double conv_tol, double dm_conv_tol, int use_diis,
                                const double* dm0, double* mo_coeff, double* mo_energy, double* dm, double* huz,
                                double* trace, nbd_scf_result* result) {
  return guarded(c, [&] {
    c->timers.reset();
    NBD_REQUIRE(c->scf_ready && c->projector == NBD_HUZINAGA, NBD_ERR_STATE, "nbd_scf_setup(projector = NBD_HUZINAGA) first");
    NBD_REQUIRE(max_cycle >= 1, NBD_ERR_ARG, "max_cycle = %d", max_cycle);
    const long nn = (long)c->nao * c->nao;
    {
    StageScope ts_all(c->timers, c->stream, "scf_total");
    const int eig_mode_in = c->eig_mode;
    double eprev[2] = {0.0, 0.0}, e[2] = {0.0, 0.0}, nd = 0.0;
    int conv = 0, cycles = 0;
    bool exported = false;  // D / Huz of the accepted attempt already sit in the caller's buffers
    for (int attempt = 0; attempt < 2; ++attempt) {
    HuzLoop L;
    c->sub_valid = false;  // every SCF run starts from scratch (cold block / full diagonalisation / the caller's density)
    c->sub_bounds_valid = false;
    c->last_eig_full = true;
    huz_initial(c, dm0, L);
    c->diis.init(6, c->nspin * nn, false);
    eprev[0] = eprev[1] = e[0] = e[1] = nd = 0.0;
    conv = cycles = 0;
    for (int i = 0; i < max_cycle; ++i) {
      StageScope ts(c->timers, c->stream, "iter_total");
      huz_iteration(c, i, use_diis != 0, L, e, &nd);
      ++cycles;
      double run_diff = 0.0;
      for (int s = 0; s < c->nspin; ++s) run_diff = std::max(run_diff, std::fabs(e[s] - eprev[s]));
      if (trace) {
        trace[3 * i + 0] = e[0];
        trace[3 * i + 1] = c->nspin == 2 ? e[1] : e[0];
        trace[3 * i + 2] = nd;
      }
      if (run_diff < conv_tol && nd < dm_conv_tol) {  // :196
        conv = 1;
        break;
      }
      eprev[0] = e[0];
      eprev[1] = e[1];
    }
    // Full (C, eps) of the last Fock matrix for the returned values.  It also audits the tracked block: a small
    // residual proves an invariant subspace, not that it is the LOWEST one, so the Ritz values of the occupied
    // orbitals must equal the nocc lowest eigenvalues of the complete spectrum.  If they do not (a missed or crossed
    // level), the whole SCF is redone with the library eigensolver in every cycle.
    const bool tracked = !c->last_eig_full;
    double theta[2][32];
    memcpy(theta, c->sub_theta, sizeof theta);
    // D and Huz are final here; only (C, eps) still need the full-spectrum solve.  Their copies to the caller's buffers
    // (2 x 30 MB at n = 1376, through the staging pipeline when the destination is pageable) run from a helper thread on
    // the copy stream while this thread drives the eigensolve.
    std::future<void> early;
    bool exported_early = false;
    if (tracked && c->early_export && (dm || huz)) {
      if (!c->stream3) NBD_CUDA(cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking));
      if (!c->ev_export) NBD_CUDA(cudaEventCreateWithFlags(&c->ev_export, cudaEventDisableTiming));
      NBD_CUDA(cudaEventRecord(c->ev_export, c->stream));
      NBD_CUDA(cudaStreamWaitEvent(c->stream3, c->ev_export, 0));
      const int dev_id = c->device;
      const size_t cnt = (size_t)c->nspin * nn;
      const double *dD = c->D.p, *dH = c->Huz.p;
      early = std::async(std::launch::async, [=] {
        NBD_CUDA(cudaSetDevice(dev_id));
        if (dm) d2h_lane1(c, dm, dD, cnt);
        if (huz) d2h_lane1(c, huz, dH, cnt);
      });
      exported_early = true;
    }
    scf_complete_spectrum(c);
    check_devinfo(c, c->nspin, "Fock eigendecomposition");
    if (early.valid()) early.get();
    if (tracked) {
      double worst = 0.0;
      std::vector<double> w(32);
      for (int s = 0; s < c->nspin; ++s) {
        const int o = std::min(scf_nocc(c, s), 32);
        d2h(c, w.data(), c->evals.p + (long)s * c->nao, o);
        NBD_CUDA(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < o; ++k) worst = std::max(worst, std::fabs(w[k] - theta[s][k]) / std::max(1.0, std::fabs(w[k])));
      }
      if (worst > 1e-8) {
        ++c->sub_rejects;
        c->eig_mode = 0;
        continue;
      }
    }
    exported = exported_early;
    break;
    }  // attempt
    c->eig_mode = eig_mode_in;
    scf_export(c, mo_coeff, mo_energy, exported ? nullptr : dm, exported ? nullptr : huz, c->Huz.p);
    if (result) {
      result->converged = conv;
      result->cycles = cycles;
      result->energy[0] = e[0];
      result->energy[1] = c->nspin == 2 ? e[1] : e[0];
      result->e_tot = c->nspin == 2 ? e[0] + e[1] : e[0];
      result->norm_ddm = nd;
      result->norm_grad = 0.0;
    }
    }  // stop events of the scopes above are recorded before the final synchronize
    finish_call(c);
  });
}

// ---- benchmark stepper: the same loop body, one call per iteration --------------------------------------
extern "C" int nbd_scf_bench_init(nbd_ctx* c) {
  return guarded(c, [&] {
    c->timers.reset();
    NBD_REQUIRE(c->scf_ready && c->projector == NBD_HUZINAGA, NBD_ERR_STATE, "nbd_scf_setup(projector = NBD_HUZINAGA) first");
    const long nn = (long)c->nao * c->nao;
    c->bench_loop = HuzLoop();
    {
      StageScope ts(c->timers, c->stream, "iter_total");
      huz_initial(c, nullptr, c->bench_loop);
      c->diis.init(6, c->nspin * nn, false);
    }
    c->bench_ready = true;
    finish_call(c);
  });
}

extern "C" int nbd_scf_bench_iteration(nbd_ctx* c, int iter, double* energy2, double* norm_ddm) {
  int rc = guarded(c, [&] {
    c->timers.reset();
    NBD_REQUIRE(c->bench_ready, NBD_ERR_STATE, "nbd_scf_bench_init first");
    double e[2] = {0, 0}, nd = 0;
    {
      StageScope ts(c->timers, c->stream, "iter_total");
      huz_iteration(c, iter, true, c->bench_loop, e, &nd);
    }
    if (energy2) {
      energy2[0] = e[0];
      energy2[1] = c->nspin == 2 ? e[1] : e[0];
    }
    if (norm_ddm) *norm_ddm = nd;
  });
  if (rc == NBD_OK) rc = guarded(c, [&] { finish_call(c); });
  return rc;
}

// ---- mu-shift loop ---------------------------------------------------------------------------------------
// e_tot = e1 + e_coul + e_nuc with the spin-resolved core Hamiltonian (embedded_hcore_funcs.py:11-46)
static double mu_energy(nbd_ctx* c, double e_nuc) {
  double t[8];
  scf_traces(c, c->heff.p, c->vhf.p, nullptr, false, t);
  double e1 = 0.0, ec = 0.0;
  for (int s = 0; s < c->nspin; ++s) {
    e1 += t[4 * s + 0];
    ec += t[4 * s + 1];
  }
  if (c->xc.on && c->ks_energy) return e1 + c->xc.ecoul + c->xc.exc + e_nuc;  // pyscf/dft/rks.py:energy_elec
  return e1 + 0.5 * ec + e_nuc;
}

// generalised eigensolve per spin of Fsrc (copied), rows of Crows = MOs (spins distributed over ranks 0 / 1
// when a communicator exists, see eigh_batched)
static void mu_eig(nbd_ctx* c, const double* Fsrc, double* Crows) {
  const int n = c->nao;
  const long nn = (long)n * n;
  int lwork = 0;
  double* sv = c->Ssave.ensure(nn);
  NBD_SOLVER(cusolverDnDsygvd_bufferSize(c->solver, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, Crows, n, sv, n, c->evals.p, &lwork));
  double* work = c->eigwork.ensure((size_t)lwork);
  int* info = c->devinfo.ensure(8);
  NBD_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * 8, c->stream));
  NBD_CUDA(cudaMemcpyAsync(Crows, Fsrc, sizeof(double) * nn * c->nspin, cudaMemcpyDeviceToDevice, c->stream));
  const bool dist = eig_distributed(c, c->nspin);
  // single rank, two spins: the second solve is issued from a helper thread on the side stream (as in eigh_batched:
  // the library call blocks its calling thread on internal synchronisations)
  const bool two_threads = !dist && c->nspin == 2 && c->overlap && c->eig_threads && n >= 512;
  {
    StageScope ts(c->timers, c->stream, "eigh");
    std::future<cusolverStatus_t> side;
    if (two_threads) {
      double* work2 = c->eigwork2.ensure((size_t)lwork);
      double* sv2 = c->Ssave2.ensure(nn);
      NBD_CUDA(cudaEventRecord(c->ev_fork, c->stream));
      NBD_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
      NBD_CUDA(cudaMemcpyAsync(sv2, c->S.p, sizeof(double) * nn, cudaMemcpyDeviceToDevice, c->stream2));
      const int dev_id = c->device;
      cusolverDnHandle_t h2 = c->solver2;
      double* A2 = Crows + nn;
      double* w2 = c->evals.p + n;
      side = std::async(std::launch::async, [=] {
        cudaSetDevice(dev_id);
        return cusolverDnDsygvd(h2, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, A2, n, sv2, n, w2, work2,
                                lwork, info + 1);
      });
    }
    for (int s = 0; s < c->nspin; ++s) {
      if ((dist && s != c->rank) || (two_threads && s == 1)) continue;
      // scipy.linalg.eigh(f, s): dsygvd overwrites the overlap with its Cholesky factor -> work on a copy
      NBD_CUDA(cudaMemcpyAsync(sv, c->S.p, sizeof(double) * nn, cudaMemcpyDeviceToDevice, c->stream));
      NBD_SOLVER(cusolverDnDsygvd(c->solver, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n,
                                  Crows + s * nn, n, sv, n, c->evals.p + (long)s * n, work, lwork, info + s));
    }
    if (two_threads) {
      const cusolverStatus_t st2 = side.get();
      NBD_CUDA(cudaEventRecord(c->ev_join, c->stream2));
      NBD_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
      if (st2 != CUSOLVER_STATUS_SUCCESS) fail(NBD_ERR_CUDA, "cusolverDnDsygvd (side stream): status %d", (int)st2);
    }
  }
  if (dist) eig_exchange(c, Crows, c->evals.p, n);
}

// ||g||, g_s = C_vir^T F_s C_occ   (pyscf/scf/uhf.py:get_grad)
static double mu_grad_norm(nbd_ctx* c) {
  StageScope ts(c->timers, c->stream, "energy");
  const int n = c->nao;
  const long nn = (long)n * n;
  double* out = c->red_out.ensure(64);
  double tot = 0.0;
  double h[2] = {0, 0};
  for (int s = 0; s < c->nspin; ++s) {
    const int o = scf_nocc(c, s), v = n - o;
    if (o == 0 || v == 0) continue;
    // T[i][mu] = sum_nu Ct[i][nu] F[nu][mu] ; G[a][i] = sum_mu Ct[o+a][mu] T[i][mu]
    gemm_nn(c, o, n, n, c->Ct.p + s * nn, n, c->F.p + s * nn, n, c->T1.p, n);
    gemm_nt(c, v, o, n, c->Ct.p + s * nn + (long)o * n, n, c->T1.p, n, c->T2.p, o);
    reduce_to(c, c->T2.p, c->T2.p, (long)v * o, 0, n, out + 32 + s);
  }
  d2h(c, h, out + 32, 2);
  NBD_CUDA(cudaStreamSynchronize(c->stream));
  for (int s = 0; s < c->nspin; ++s) {
    const int o = scf_nocc(c, s);
    if (o == 0 || o == n) continue;
    tot += h[s] * (c->nspin == 1 ? 4.0 : 1.0);  // RHF gradient carries a factor 2
  }
  return std::sqrt(tot);
}

// CDIIS error vector per spin: Corth^T (S D F) Corth, antisymmetrised (pyscf/scf/diis.py:get_err_vec_orth)
static void mu_cdiis_errvec(nbd_ctx* c, double* err) {
  StageScope ts(c->timers, c->stream, "diis");
  const int n = c->nao;
  const long nn = (long)n * n;
  const int ns = c->nspin;
  gemm_nn(c, n, n, n, c->S.p, n, c->D.p, n, c->T1.p, n, 1.0, 0.0, ns, 0, nn, nn);      // S D
  gemm_nn(c, n, n, n, c->T1.p, n, c->F.p, n, c->T2.p, n, 1.0, 0.0, ns, nn, nn, nn);    // (S D) F
  gemm_nn(c, n, n, n, c->Corth.p, n, c->T2.p, n, c->T1.p, n, 1.0, 0.0, ns, nn, nn, nn);  // Corth^T (.)   (rows = MOs)
  gemm_nt(c, n, n, n, c->T1.p, n, c->Corth.p, n, c->T2.p, n, 1.0, 0.0, ns, nn, nn, nn);  // (.) Corth
  dim3 g((n + 127) / 128, n, ns);
  antisym_kernel<<<g, 128, 0, c->stream>>>(c->T2.p, err, n);
  LAUNCH_CHECK(c);
}

extern "C" int nbd_mu_scf(nbd_ctx* c, int max_cycle, double conv_tol, double e_nuc, const double* dm0,
                          double* mo_coeff, double* mo_energy, double* mo_occ, double* dm, double* vhf_out,
                          double* trace, nbd_scf_result* result) {
  return guarded(c, [&] {
    c->timers.reset();
    NBD_REQUIRE(c->scf_ready && c->projector == NBD_MU_SHIFT, NBD_ERR_STATE, "nbd_scf_setup(projector = NBD_MU_SHIFT) first");
    NBD_REQUIRE(dm0 && max_cycle >= 1, NBD_ERR_ARG, "dm0 is required (PySCF's minao guess needs basis data)");
    // the reference's mu path is spin-resolved only: _env_projector indexes dm_enviro[0] (nbed/driver.py:439) and
    // the driver always builds UHF/UKS objects, so a rank-2 call has no reference behaviour to reproduce
    NBD_REQUIRE(c->nspin == 2, NBD_ERR_UNSUPPORTED, "mu-shift embedding is spin-resolved (rank-3 inputs) in the reference");
    const int n = c->nao, ns = c->nspin;
    const long nn = (long)n * n;
    const double conv_tol_grad = std::sqrt(conv_tol);
    c->Ssave.ensure(nn);
    {
    StageScope ts_all(c->timers, c->stream, "scf_total");
    // dm = dm0 ; vhf = get_veff(dm) ; e_tot           (pyscf/scf/hf.py:kernel prologue)
    h2d(c, c->D.p, dm0, (size_t)ns * nn);
    {
      double* tmp = c->dm0f.ensure((size_t)ns * nn);
      NBD_CUDA(cudaMemcpyAsync(tmp, c->D.p, sizeof(double) * nn * ns, cudaMemcpyDeviceToDevice, c->stream));
      std::vector<int> jb;
      std::vector<KGroup> groups;
      const int Ntot = factor_densities(c, tmp, ns, jb, groups);
      scf_build_fock(c, Ntot, groups);
    }
    double e_tot = mu_energy(c, e_nuc);
    // CDIIS with Corth from eig(F0, S)
    c->diis.init(8, ns * nn, true);
    mu_eig(c, c->F.p, c->Corth.p);
    double* err = c->FG.p;
    int conv = 0, cycles = 0;
    double norm_g = 0.0, norm_dd = 0.0;
    int tr = 0;
    auto new_density = [&]() {  // eig -> occ -> dm -> vhf -> e_tot -> F = h + vhf
      mu_eig(c, c->F.p, c->Ct.p);
      scf_make_density(c);
      const int Ntot = scf_stage_occupied(c);
      scf_build_fock(c, Ntot, scf_occ_groups(c));
    };
    auto scalars = [&](double& e_new) {
      e_new = mu_energy(c, e_nuc);
      norm_g = mu_grad_norm(c);
      double t[8];
      scf_traces(c, nullptr, nullptr, nullptr, true, t);
      double dd = 0.0;
      for (int s = 0; s < ns; ++s) dd += t[4 * s + 3];
      norm_dd = std::sqrt(dd);
      if (trace) {
        trace[3 * tr + 0] = e_new;
        trace[3 * tr + 1] = norm_g;
        trace[3 * tr + 2] = norm_dd;
      }
      ++tr;
    };
    for (int cycle = 0; cycle < max_cycle; ++cycle) {
      StageScope ts(c->timers, c->stream, "iter_total");
      const double last_e = e_tot;
      if (cycle >= 1) {  // get_fock(..., cycle, diis): CDIIS from cycle 1
        mu_cdiis_errvec(c, err);
        diis_update(c, c->diis, c->F.p, err);
      }
      new_density();
      scalars(e_tot);
      ++cycles;
      if (std::fabs(e_tot - last_e) < conv_tol && norm_g < conv_tol_grad) {
        conv = 1;
        break;
      }
    }
    if (conv) {  // extra cycle (conv_check)
      const double last_e = e_tot;
      new_density();
      scalars(e_tot);
      conv = (std::fabs(e_tot - last_e) < conv_tol * 10 || norm_g < conv_tol_grad * 3) ? 1 : 0;
    }
    check_devinfo(c, c->nspin, "generalised eigensolve");
    scf_export(c, mo_coeff, mo_energy, dm, vhf_out, c->vhf.p);
    if (mo_occ)
      for (int s = 0; s < ns; ++s)
        for (int i = 0; i < n; ++i) mo_occ[(long)s * n + i] = i < scf_nocc(c, s) ? scf_occ(c) : 0.0;
    if (result) {
      result->converged = conv;
      result->cycles = cycles;
      result->e_tot = e_tot;
      result->energy[0] = result->energy[1] = 0.0;
      result->norm_ddm = norm_dd;
      result->norm_grad = norm_g;
    }
    }
    finish_call(c);
  });
}
