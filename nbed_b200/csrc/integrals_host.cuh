// Host driver of the device integral generator (integrals.cuh) behind the C-ABI: libcint-format basis in, the
// density-fitting tensor out.  Included by nbed_b200.cu (single translation unit).
#pragma once
#include <cublas_v2.h>

#include "integrals.cuh"

#define NBD_CUBLAS(call)                                                                            \
  do {                                                                                              \
    cublasStatus_t s_ = (call);                                                                     \
    if (s_ != CUBLAS_STATUS_SUCCESS) ::nbd::fail(NBD_ERR_CUDA, "%s:%d %s: cublas status %d", __FILE__, __LINE__, #call, (int)s_); \
  } while (0)

void nbd_destroy_blas(struct cublasContext* h) { cublasDestroy(h); }

// libcint slots (cint.h)
constexpr int CINT_ATM_SLOTS = 6, CINT_BAS_SLOTS = 8, CINT_PTR_COORD = 1;
constexpr int CINT_ATOM_OF = 0, CINT_ANG_OF = 1, CINT_NPRIM_OF = 2, CINT_NCTR_OF = 3, CINT_PTR_EXP = 5, CINT_PTR_COEFF = 6;

// Unit-normalised real solid harmonics in Cartesian monomials (Helgaker, Jorgensen, Olsen eq. 6.4.47-50), rows
// m = -l..l, columns in libcint's Cartesian order; l = 0, 1 are libcint's CINTcommon_fac_sp constants (p: x, y, z).
static std::vector<double> cart2sph_matrix(int l) {
  const int nc = ncart(l);
  std::vector<double> out((size_t)(2 * l + 1) * nc, 0.0);
  if (l == 0) {
    out[0] = 0.282094791773878143;
    return out;
  }
  if (l == 1) {
    for (int i = 0; i < 3; ++i) out[i * 3 + i] = 0.488602511902919921;
    return out;
  }
  auto fact = [](int k) { double f = 1.0; for (int i = 2; i <= k; ++i) f *= i; return f; };
  auto binom = [&](int a, int b) { return (b < 0 || b > a) ? 0.0 : fact(a) / (fact(b) * fact(a - b)); };
  auto cidx = [&](int ex, int ey) {  // position of (ex, ey, l - ex - ey) in libcint's order
    int k = 0;
    for (int lx = l; lx >= 0; --lx)
      for (int ly = l - lx; ly >= 0; --ly, ++k)
        if (lx == ex && ly == ey) return k;
    return -1;
  };
  for (int m = -l; m <= l; ++m) {
    const int am = std::abs(m);
    const double nlm = 1.0 / (std::pow(2.0, am) * fact(l)) * std::sqrt(2.0 * fact(l + am) * fact(l - am) / (m == 0 ? 2.0 : 1.0));
    const int vm2 = m < 0 ? 1 : 0;
    for (int t = 0; t <= (l - am) / 2; ++t)
      for (int u = 0; u <= t; ++u)
        for (int v2 = vm2; v2 <= am; v2 += 2) {
          const double cf = (((t + (v2 - vm2) / 2) & 1) ? -1.0 : 1.0) * std::pow(0.25, t) * binom(l, t) * binom(l - t, am + t) *
                            binom(t, u) * binom(am, v2);
          const int ex = 2 * t + am - 2 * u - v2, ey = 2 * u + v2;
          out[(size_t)(m + l) * nc + cidx(ex, ey)] += nlm * cf;
        }
    const double sph = std::sqrt((2.0 * l + 1.0) / (4.0 * 3.14159265358979323846));
    for (int k = 0; k < nc; ++k) out[(size_t)(m + l) * nc + k] *= sph;
  }
  return out;
}

struct DfBasis {
  std::vector<IntShell> ao, aux;
  int nao = 0, naux = 0, lmax_ao = 0, lmax_aux = 0, npp_max = 1;
};

// libcint (atm, bas, env) of the concatenated mol + auxmol -> segmented shell tables (one entry per contraction)
static DfBasis parse_basis(const int* atm, int natm, const int* bas, int nbas, const double* env, int nenv, int nbas_ao) {
  DfBasis B;
  NBD_REQUIRE(atm && bas && env && natm > 0 && nbas > 0 && nbas_ao > 0 && nbas_ao <= nbas, NBD_ERR_ARG, "bad basis arrays");
  for (int ib = 0; ib < nbas; ++ib) {
    const int* b = bas + (size_t)ib * CINT_BAS_SLOTS;
    const int ia = b[CINT_ATOM_OF], l = b[CINT_ANG_OF], np = b[CINT_NPRIM_OF], nc = b[CINT_NCTR_OF];
    NBD_REQUIRE(ia >= 0 && ia < natm && np >= 1 && nc >= 1, NBD_ERR_ARG, "shell %d: atom %d, nprim %d, nctr %d", ib, ia, np, nc);
    const int pc = atm[(size_t)ia * CINT_ATM_SLOTS + CINT_PTR_COORD];
    NBD_REQUIRE(pc >= 0 && pc + 3 <= nenv && b[CINT_PTR_EXP] >= 0 && b[CINT_PTR_EXP] + np <= nenv && b[CINT_PTR_COEFF] >= 0 &&
                    b[CINT_PTR_COEFF] + np * nc <= nenv, NBD_ERR_ARG, "shell %d points outside env", ib);
    const bool is_ao = ib < nbas_ao;
    NBD_REQUIRE(l >= 0 && l <= (is_ao ? INT_LMAX_AO : INT_LMAX_AUX), NBD_ERR_UNSUPPORTED,
                "shell %d has l = %d (implemented: orbital l <= %d, auxiliary l <= %d)", ib, l, INT_LMAX_AO, INT_LMAX_AUX);
    for (int ic = 0; ic < nc; ++ic) {  // libcint orders the functions of a shell contraction-major
      IntShell s{};
      s.x = env[pc]; s.y = env[pc + 1]; s.z = env[pc + 2];
      s.l = l; s.nprim = np;
      s.ptr_exp = b[CINT_PTR_EXP];
      s.ptr_coef = b[CINT_PTR_COEFF] + ic * np;
      if (is_ao) {
        s.ao_off = B.nao; B.nao += 2 * l + 1; B.lmax_ao = std::max(B.lmax_ao, l); B.ao.push_back(s);
      } else {
        s.ao_off = B.naux; B.naux += 2 * l + 1; B.lmax_aux = std::max(B.lmax_aux, l); B.aux.push_back(s);
      }
    }
  }
  for (auto& s : B.ao) B.npp_max = std::max(B.npp_max, s.nprim * s.nprim);
  return B;
}

struct DfDevice {
  DBuf<IntShell> ao, aux;
  DBuf<double> env, c2s, j3c, j2c;
  DBuf<int> pairs;
};

// (P | mu >= nu) [naux][npair] and (P|Q) [naux][naux] on the device
static void df_integrals_device(nbd_ctx* c, const DfBasis& B, const double* env, int nenv, DfDevice& D) {
  NBD_REQUIRE(!B.aux.empty(), NBD_ERR_ARG, "no auxiliary shells: nbas_ao must be smaller than nbas");
  const long npair = (long)B.nao * (B.nao + 1) / 2;
  D.ao.ensure(B.ao.size());
  D.aux.ensure(B.aux.size());
  D.env.ensure(nenv);
  NBD_CUDA(cudaMemcpyAsync(D.ao.p, B.ao.data(), sizeof(IntShell) * B.ao.size(), cudaMemcpyHostToDevice, c->stream));
  NBD_CUDA(cudaMemcpyAsync(D.aux.p, B.aux.data(), sizeof(IntShell) * B.aux.size(), cudaMemcpyHostToDevice, c->stream));
  NBD_CUDA(cudaMemcpyAsync(D.env.p, env, sizeof(double) * nenv, cudaMemcpyHostToDevice, c->stream));
  std::vector<double> t;
  for (int l = 0; l <= INT_LMAX_AUX; ++l) {
    const auto m = cart2sph_matrix(l);
    t.insert(t.end(), m.begin(), m.end());
  }
  D.c2s.ensure(t.size());
  NBD_CUDA(cudaMemcpyAsync(D.c2s.p, t.data(), sizeof(double) * t.size(), cudaMemcpyHostToDevice, c->stream));
  std::vector<int> pl;
  for (int i = 0; i < (int)B.ao.size(); ++i)
    for (int j = 0; j <= i; ++j) {
      pl.push_back(i);
      pl.push_back(j);
    }
  D.pairs.ensure(pl.size());
  NBD_CUDA(cudaMemcpyAsync(D.pairs.p, pl.data(), sizeof(int) * pl.size(), cudaMemcpyHostToDevice, c->stream));
  D.j3c.ensure((size_t)B.naux * npair);
  D.j2c.ensure((size_t)B.naux * B.naux);
  IntArgs a{};
  a.ao = D.ao.p; a.aux = D.aux.p; a.env = D.env.p; a.c2s = D.c2s.p;
  a.nsh_ao = (int)B.ao.size(); a.nsh_aux = (int)B.aux.size();
  a.npair = npair; a.naux = B.naux;
  // shared memory: as many primitive-pair records (p, P, coefficient, E^x, E^y, E^z) as fit 48 KB
  const int lm = B.lmax_ao, rec_max = 5 + 3 * (lm + 1) * (lm + 1) * (2 * lm + 1);
  const int pp_cap = std::max(1, std::min(B.npp_max, 6000 / rec_max));
  const size_t smem = (size_t)pp_cap * rec_max * sizeof(double);
  const int npairs = (int)(pl.size() / 2);
  a.out = D.j3c.p;
  StageScope ts(c->timers, c->stream, "int3c2e");
#define NBD_INT3C(LAB, LC) int3c2e_kernel<LAB, LC><<<npairs, INT_THREADS, smem, c->stream>>>(a, D.pairs.p, pp_cap)
  const bool big_aux = B.lmax_aux > 2;
  if (lm <= 1) { if (big_aux) NBD_INT3C(2, 4); else NBD_INT3C(2, 2); }
  else if (lm == 2) { if (big_aux) NBD_INT3C(4, 4); else NBD_INT3C(4, 2); }
  else { if (big_aux) NBD_INT3C(6, 4); else NBD_INT3C(6, 2); }
#undef NBD_INT3C
  LAUNCH_CHECK(c);
  a.out = D.j2c.p;
  const long nap = (long)a.nsh_aux * (a.nsh_aux + 1) / 2;
  int2c2e_kernel<<<(unsigned)((nap + INT_THREADS - 1) / INT_THREADS), INT_THREADS, 0, c->stream>>>(a);
  LAUNCH_CHECK(c);
}

extern "C" int nbd_int3c2e(nbd_ctx* c, const int* atm, int natm, const int* bas, int nbas, const double* env, int nenv,
                           int nbas_ao, double* j3c, double* j2c) {
  return guarded(c, [&] {
    c->timers.reset();
    const DfBasis B = parse_basis(atm, natm, bas, nbas, env, nenv, nbas_ao);
    DfDevice D;
    df_integrals_device(c, B, env, nenv, D);
    const long npair = (long)B.nao * (B.nao + 1) / 2;
    if (j3c) d2h(c, j3c, D.j3c.p, (size_t)B.naux * npair);
    if (j2c) d2h(c, j2c, D.j2c.p, (size_t)B.naux * B.naux);
    finish_call(c);
  });
}

extern "C" int nbd_basis_dims(const int* bas, int nbas, int nbas_ao, int* nao, int* naux) {
  if (!bas || !nao || !naux || nbas_ao < 0 || nbas_ao > nbas) return NBD_ERR_ARG;
  *nao = *naux = 0;
  for (int ib = 0; ib < nbas; ++ib) {
    const int* b = bas + (size_t)ib * CINT_BAS_SLOTS;
    (ib < nbas_ao ? *nao : *naux) += (2 * b[CINT_ANG_OF] + 1) * b[CINT_NCTR_OF];
  }
  return NBD_OK;
}

extern "C" int nbd_cderi_from_basis(nbd_ctx* c, const int* atm, int natm, const int* bas, int nbas, const double* env,
                                    int nenv, int nbas_ao, int global_row0, int naux_local) {
  if (!c) return NBD_ERR_ARG;
  int nao = 0, naux = 0;
  if (nbd_basis_dims(bas, nbas, nbas_ao, &nao, &naux) != NBD_OK || nao <= 0 || naux <= 0) {
    c->err = "bad basis arrays";
    return NBD_ERR_ARG;
  }
  if (naux_local < 0) naux_local = naux - global_row0;
  if (global_row0 < 0 || naux_local < 0 || global_row0 + naux_local > naux) {
    c->err = "aux row range outside the auxiliary basis";
    return NBD_ERR_ARG;
  }
  const int rc = nbd_cderi_alloc(c, nao, naux_local);
  if (rc != NBD_OK) return rc;
  return guarded(c, [&] {
    c->timers.reset();
    const DfBasis B = parse_basis(atm, natm, bas, nbas, env, nenv, nbas_ao);
    DfDevice D;
    df_integrals_device(c, B, env, nenv, D);
    const long npair = (long)B.nao * (B.nao + 1) / 2;
    {
      // pyscf.df.incore.cholesky_eri: (P|Q) = L L^T (dpotrf), cderi = L^-1 (P|mu nu) (triangular solve, in place).
      // Row-major [naux][npair] is the column-major npair x naux matrix B: cderi = B L^-T.
      StageScope ts(c->timers, c->stream, "cholesky");
      int lwork = 0;
      NBD_SOLVER(cusolverDnDpotrf_bufferSize(c->solver, CUBLAS_FILL_MODE_LOWER, B.naux, D.j2c.p, B.naux, &lwork));
      double* work = c->eigwork.ensure((size_t)lwork);
      int* info = c->devinfo.ensure(8);
      NBD_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * 8, c->stream));
      NBD_SOLVER(cusolverDnDpotrf(c->solver, CUBLAS_FILL_MODE_LOWER, B.naux, D.j2c.p, B.naux, work, lwork, info));
      int h = 0;
      NBD_CUDA(cudaMemcpyAsync(&h, info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      NBD_CUDA(cudaStreamSynchronize(c->stream));
      NBD_REQUIRE(h == 0, NBD_ERR_ARG, "auxiliary Coulomb metric is not positive definite (dpotrf info = %d): linearly dependent "
                  "auxiliary basis; PySCF's eigen-decomposition fallback is not implemented", h);
      if (!c->blas) {
        NBD_CUBLAS(cublasCreate(&c->blas));
        NBD_CUBLAS(cublasSetStream(c->blas, c->stream));
      }
      const double one = 1.0;
      NBD_CUBLAS(cublasDtrsm(c->blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, (int)npair, B.naux,
                             &one, D.j2c.p, B.naux, D.j3c.p, (int)npair));
    }
    for (int r = 0; r < naux_local; r += 32768) {
      const int k = std::min(32768, naux_local - r);
      dim3 g(c->ntiles, k);
      pack_to_tiled_kernel<<<g, 256, 0, c->stream>>>(D.j3c.p + (long)(global_row0 + r) * npair, c->Bt, c->d_seq.p, c->ntiles, c->nao,
                                                     c->npair, r);
      LAUNCH_CHECK(c);
    }
    finish_call(c);
  });
}
