// Active-space AO->MO transform behind the C-ABI (nbed/ham_builder.py:53-156,158-216).
// Included by nbed_b200.cu (single translation unit).
#pragma once

// Upload host C [n][m] (columns = MOs) as rows [row0, row0+m) of the padded orbital block c->d_orb.
static void ao2mo_stage_mos(nbd_ctx* c, const double* C, int m, int row0) {
  const int n = c->nao;
  std::vector<double> rows((size_t)m * n);
  for (int p = 0; p < m; ++p)
    for (int mu = 0; mu < n; ++mu) rows[(size_t)p * n + mu] = C[(size_t)mu * m + p];
  double* st = c->stage.ensure((size_t)m * n);
  // the staging buffer is reused by the next call: serialise
  NBD_CUDA(cudaStreamSynchronize(c->stream));
  h2d(c, st, rows.data(), (size_t)m * n);
  dim3 g((c->n_ld + 127) / 128, m);
  pad_rows_kernel<<<g, 128, 0, c->stream>>>(st, n, c->d_orb.p + (long)row0 * c->n_ld, c->n_ld, n, nullptr, 1.0);
  LAUNCH_CHECK(c);
  NBD_CUDA(cudaStreamSynchronize(c->stream));
}

// Device part of the transform: leaves out[blk][p][r][s][q] = (pq|rs) (4 blocks) in c->eri_phys.
static void ao2mo_device(nbd_ctx* c, int m, const double* ca, const double* cb) {
    const int n_ld = c->n_ld, naux = c->naux;
    const bool restricted = (cb == nullptr || cb == ca);
    const int nsp = restricted ? 1 : 2;
    const int Ntot = nsp * m;
    const long m2 = (long)m * m, m4 = m2 * m2;
    {
      StageScope ts_all(c->timers, c->stream, "ao2mo_total");
      c->d_orb.ensure((size_t)Ntot * n_ld);
      ao2mo_stage_mos(c, ca, m, 0);
      if (!restricted) ao2mo_stage_mos(c, cb, m, m);
      double* L = c->Lbuf.ensure((size_t)nsp * std::max(1, naux) * m2);
      const int nblk = restricted ? 1 : 3;
      double* eri = c->eri.ensure((size_t)nblk * m4);
      const long per_row = (long)Ntot * n_ld * 8;
      const int chunk = (int)std::max<long>(1, std::min<long>(std::max(1, naux), c->x_budget_bytes / per_row));
      double* X = c->d_X.ensure((size_t)chunk * Ntot * n_ld);
      std::vector<std::pair<int, int>> cols;
      for (int s = 0; s < nsp; ++s) cols.push_back({s * m, (s + 1) * m});
      for (int p0 = 0; p0 < naux; p0 += chunk) {
        const int np = std::min(chunk, naux - p0);
        const XLayout XL = make_xlayout(c, np, Ntot, cols);
        {
          StageScope ts(c->timers, c->stream, "ao2mo_half");  // X_s[(P,p)][mu] = sum_nu B[P][mu][nu] C_s[nu][p]
          half_transform(c, p0, np, c->d_orb.p, Ntot, X);
        }
        StageScope ts(c->timers, c->stream, "ao2mo_l");  // L_s[(P,p)][q] = sum_mu X_s[(P,p)][mu] C_s[mu][q]
        for (int s = 0; s < nsp; ++s)
          gemm(c, np * m, m, c->nao, X + XL.group_base[s], n_ld, 1, c->d_orb.p + (long)s * m * n_ld, n_ld, 1,
               L + ((long)s * naux + p0) * m2, m, 1.0, 0.0);
      }
      {
        StageScope ts(c->timers, c->stream, "ao2mo_eri");  // (pq|rs) = sum_P L[P][pq] L'[P][rs]
        if (naux == 0) {
          NBD_CUDA(cudaMemsetAsync(eri, 0, sizeof(double) * nblk * m4, c->stream));
        } else {
          // same-spin blocks are symmetric under (pq) <-> (rs): lower tiles + mirror (half the flops of two of the three)
          gemm_tn(c, (int)m2, (int)m2, naux, L, m2, L, m2, eri, m2, 1.0, 0.0, 1, 0, 0, 0, /*lower=*/1);
          if (!restricted) {
            double* Lb = L + (long)naux * m2;
            gemm_tn(c, (int)m2, (int)m2, naux, Lb, m2, Lb, m2, eri + m4, m2, 1.0, 0.0, 1, 0, 0, 0, /*lower=*/1);
            gemm_tn(c, (int)m2, (int)m2, naux, L, m2, Lb, m2, eri + 2 * m4, m2);
          }
          symmetrize_lower(c, eri, (int)m2, restricted ? 1 : 2);
        }
      }
      all_reduce(c, eri, (size_t)nblk * m4);
      {
        StageScope ts(c->timers, c->stream, "ao2mo_perm");  // restore(1) + transpose(0,2,3,1): out[p][r][s][q] = (pq|rs)
        double* phys = c->eri_phys.ensure((size_t)4 * m4);
        const dim3 g = grid1(m4, 256);
        for (int blk = 0; blk < 4; ++blk) {
          const double* src = restricted ? eri : (blk == 0 ? eri : blk == 1 ? eri + m4 : eri + 2 * m4);
          chem_to_phys_kernel<<<g, 256, 0, c->stream>>>(src, phys + (long)blk * m4, m, (!restricted && blk == 3) ? 1 : 0);
          LAUNCH_CHECK(c);
        }
      }
    }
}

extern "C" int nbd_ao2mo(nbd_ctx* c, int m, const double* ca, const double* cb, double* out) {
  return guarded(c, [&] {
    c->timers.reset();
    NBD_REQUIRE(c->Bt, NBD_ERR_STATE, "nbd_cderi_alloc first");
    NBD_REQUIRE(m >= 1 && ca && out, NBD_ERR_ARG, "bad m / pointers");
    ao2mo_device(c, m, ca, cb);
    d2h(c, out, c->eri_phys.p, (size_t)4 * m * m * m * m);
    finish_call(c);
  });
}

// out (device, c->red_out) [2][m][m] = C_s^T h_s C_s
static double* one_body_device(nbd_ctx* c, int m, int nspin_h, const double* hcore, const double* ca, const double* cb) {
  const int n = c->nao;
  const long nn = (long)n * n, nm = (long)n * m, m2 = (long)m * m;
  if (!cb) cb = ca;
  double* h = c->T1.ensure((size_t)2 * nn);
  double* C = c->mo_c.ensure((size_t)2 * nm);
  double* T = c->T2.ensure((size_t)std::max(2 * nn, 2 * nm));
  double* o = c->red_out.ensure((size_t)std::max<long>(64, 2 * m2));
  h2d(c, h, hcore, (size_t)nspin_h * nn);
  h2d(c, C, ca, nm);
  h2d(c, C + nm, cb, nm);
  for (int s = 0; s < 2; ++s) {
    const double* hs = h + (nspin_h == 2 ? s * nn : 0);
    gemm_nn(c, n, m, n, hs, n, C + s * nm, m, T + s * nm, m);
    gemm_tn(c, m, m, n, C + s * nm, m, T + s * nm, m, o + s * m2, m);
  }
  return o;
}

extern "C" int nbd_one_body(nbd_ctx* c, int m, int nspin_h, const double* hcore, const double* ca, const double* cb,
                            double* out) {
  return guarded(c, [&] {
    c->timers.reset();
    NBD_REQUIRE(c->nao > 0, NBD_ERR_STATE, "nbd_cderi_alloc first (nao)");
    NBD_REQUIRE(m >= 1 && (nspin_h == 1 || nspin_h == 2) && hcore && ca && out, NBD_ERR_ARG, "bad arguments");
    const double* o = one_body_device(c, m, nspin_h, hcore, ca, cb);
    d2h(c, out, o, (size_t)2 * m * m);
    finish_call(c);
  });
}

// Replaces HamiltonianBuilder.build() end to end (nbed/ham_builder.py:218-254): one-body transform, the four
// two-body blocks, the spin-orbital scatter with the EQ_TOLERANCE truncation and the 0.5 factor - all on the
// device; only h1 [2m][2m] and h2 [2m]^4 cross PCIe.
extern "C" int nbd_build_hamiltonian(nbd_ctx* c, int m, int nspin_h, const double* hcore, const double* ca,
                                     const double* cb, double eq_tol, double two_body_scale, double* h1, double* h2) {
  return guarded(c, [&] {
    c->timers.reset();
    NBD_REQUIRE(c->Bt, NBD_ERR_STATE, "nbd_cderi_alloc first");
    NBD_REQUIRE(m >= 1 && (nspin_h == 1 || nspin_h == 2) && hcore && ca && h1 && h2, NBD_ERR_ARG, "bad arguments");
    const long m2 = (long)m * m, m4 = m2 * m2, q2 = 4 * m2, q4 = 16 * m4;
    {
      StageScope ts_all(c->timers, c->stream, "build_total");
      const double* one = one_body_device(c, m, nspin_h, hcore, ca, cb);
      double* d_h1 = c->red_part.ensure((size_t)std::max<long>(q2, (long)REDUCE_BLOCKS * 9));
      spinorb_one_kernel<<<grid1(q2, 256), 256, 0, c->stream>>>(one, d_h1, m, eq_tol);
      LAUNCH_CHECK(c);
      ao2mo_device(c, m, ca, cb);  // -> c->eri_phys (uses T1/T2-free workspaces only)
      double* d_h2 = c->Lbuf.ensure((size_t)q4);  // L is dead after the (pq|rs) GEMMs
      {
        StageScope ts(c->timers, c->stream, "spinorb");
        spinorb_two_kernel<<<grid1(q4, 256), 256, 0, c->stream>>>(c->eri_phys.p, d_h2, m, eq_tol, two_body_scale);
        LAUNCH_CHECK(c);
      }
      d2h(c, h1, d_h1, (size_t)q2);
      d2h(c, h2, d_h2, (size_t)q4);
    }
    finish_call(c);
  });
}

extern "C" int nbd_spinorb_from_spatial(nbd_ctx* c, int m, const double* one, const double* two, double eq_tol,
                                        double two_body_scale, double* h1, double* h2) {
  return guarded(c, [&] {
    c->timers.reset();
    NBD_REQUIRE(m >= 1 && one && two && h1 && h2, NBD_ERR_ARG, "bad arguments");
    const long m2 = (long)m * m, m4 = m2 * m2;
    const long q2 = 4 * m2, q4 = 16 * m4;
    double* d_two = c->eri_phys.ensure((size_t)4 * m4);
    double* d_one = c->mo_c.ensure((size_t)2 * m2 + q2);
    double* d_h2 = c->eri.ensure((size_t)q4);
    h2d(c, d_two, two, (size_t)4 * m4);
    h2d(c, d_one, one, (size_t)2 * m2);
    {
      StageScope ts(c->timers, c->stream, "spinorb");
      spinorb_two_kernel<<<grid1(q4, 256), 256, 0, c->stream>>>(d_two, d_h2, m, eq_tol, two_body_scale);
      LAUNCH_CHECK(c);
      spinorb_one_kernel<<<grid1(q2, 256), 256, 0, c->stream>>>(d_one, d_one + 2 * m2, m, eq_tol);
      LAUNCH_CHECK(c);
    }
    d2h(c, h2, d_h2, (size_t)q4);
    d2h(c, h1, d_one + 2 * m2, (size_t)q2);
    finish_call(c);
  });
}
