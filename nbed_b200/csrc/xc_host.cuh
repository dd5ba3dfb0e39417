// Host driver of the exchange-correlation stage (xc.cuh) and its hook into the embedded-SCF loops.
// Included by nbed_b200.cu after integrals_host.cuh (shares the libcint-format basis parser).
#pragma once
#include "xc.cuh"

constexpr int XC_CHUNK = 2048;  // grid points per split-K slab of the V = Phi^T M contraction

static double xc_hybrid_fraction(int code) { return code == NBD_XC_B3LYP ? 0.2 : 0.0; }

// Replaces: the grid-side set-up of a pyscf KS object - numint.eval_ao(mol, grids.coords, deriv=1) - for the grid the
// caller passes (PySCF's mf.grids.coords / weights).  AO values and gradients stay resident: [4][ng_pad][nao].
extern "C" int nbd_xc_setup(nbd_ctx* c, int xc_code, const int* atm, int natm, const int* bas, int nbas, const double* env,
                            int nenv, int ngrid, const double* coords, const double* weights) {
  return guarded(c, [&] {
    c->timers.reset();
    NBD_REQUIRE(xc_code == NBD_XC_B3LYP || xc_code == NBD_XC_LDA, NBD_ERR_UNSUPPORTED, "xc functional code %d (implemented: 1 = b3lyp, 2 = lda: slater + vwn_rpa)", xc_code);
    NBD_REQUIRE(ngrid >= 1 && coords && weights, NBD_ERR_ARG, "bad grid");
    NBD_REQUIRE(c->Bt, NBD_ERR_STATE, "nbd_cderi_alloc / nbd_cderi_from_basis first (nao is taken from the 3-centre tensor)");
    const DfBasis B = parse_basis(atm, natm, bas, nbas, env, nenv, nbas);
    NBD_REQUIRE(B.nao == c->nao, NBD_ERR_ARG, "basis has %d spherical functions, the 3-centre tensor %d", B.nao, c->nao);
    XcState& x = c->xc;
    x.ready = false;
    const int ng = (ngrid + XC_CHUNK - 1) / XC_CHUNK * XC_CHUNK;  // padded with zero-weight copies of the last point
    std::vector<double> hc((size_t)3 * ng), hw((size_t)ng, 0.0);
    memcpy(hc.data(), coords, sizeof(double) * 3 * ngrid);
    memcpy(hw.data(), weights, sizeof(double) * ngrid);
    for (int g = ngrid; g < ng; ++g)
      for (int k = 0; k < 3; ++k) hc[(size_t)3 * g + k] = coords[(size_t)3 * (ngrid - 1) + k];
    x.coords.ensure((size_t)3 * ng);
    x.w.ensure(ng);
    h2d(c, x.coords.p, hc.data(), (size_t)3 * ng);
    h2d(c, x.w.p, hw.data(), ng);
    x.shells.ensure(B.ao.size());
    x.env.ensure(nenv);
    NBD_CUDA(cudaMemcpyAsync(x.shells.p, B.ao.data(), sizeof(IntShell) * B.ao.size(), cudaMemcpyHostToDevice, c->stream));
    NBD_CUDA(cudaMemcpyAsync(x.env.p, env, sizeof(double) * nenv, cudaMemcpyHostToDevice, c->stream));
    std::vector<double> t;
    for (int l = 0; l <= INT_LMAX_AUX; ++l) {
      const auto m = cart2sph_matrix(l);
      t.insert(t.end(), m.begin(), m.end());
    }
    x.c2s.ensure(t.size());
    NBD_CUDA(cudaMemcpyAsync(x.c2s.p, t.data(), sizeof(double) * t.size(), cudaMemcpyHostToDevice, c->stream));
    const int n = c->nao;
    x.ao.ensure((size_t)4 * ng * n);
    NBD_CUDA(cudaStreamSynchronize(c->stream));  // hc / hw / t go out of scope
    {
      StageScope ts(c->timers, c->stream, "xc_ao");
      xc_eval_ao_kernel<<<(ng + 127) / 128, 128, 0, c->stream>>>(x.shells.p, (int)B.ao.size(), x.env.p, x.c2s.p, x.coords.p, ng, n, x.ao.p);
      LAUNCH_CHECK(c);
    }
    x.rho.ensure((size_t)2 * ng);
    x.grad.ensure((size_t)6 * ng);
    x.sigma.ensure((size_t)3 * ng);
    x.fx.ensure((size_t)6 * ng);
    x.TM.ensure((size_t)2 * ng * n);
    x.Vpart.ensure((size_t)(ng / XC_CHUNK) * n * n);
    x.V.ensure((size_t)2 * n * n);
    x.ng = ng;
    x.ng_user = ngrid;
    x.code = xc_code;
    x.hyb = xc_hybrid_fraction(xc_code);
    x.ready = true;
    finish_call(c);
  });
}

// numint.nr_uks: V_s (device, c->xc.V) for the device densities d_dm [2][n][n]; integrals[3] = (int f, N_alpha, N_beta)
static void xc_eval_device(nbd_ctx* c, const double* d_dm, double* integrals3) {
  XcState& x = c->xc;
  NBD_REQUIRE(x.ready, NBD_ERR_STATE, "nbd_xc_setup first");
  StageScope ts(c->timers, c->stream, "xc");
  const int n = c->nao, ng = x.ng;
  const long nn = (long)n * n, plane = (long)ng * n;
  // T_s = Phi D_s  (ng x n x n)
  gemm_nn(c, ng, n, n, x.ao.p, n, d_dm, n, x.TM.p, n, 1.0, 0.0, 2, 0, nn, plane);
  xc_density_kernel<<<(unsigned)(((long)ng * 32 + 255) / 256), 256, 0, c->stream>>>(x.TM.p, x.ao.p, plane, 0, ng, ng, n, x.rho.p, x.grad.p);
  LAUNCH_CHECK(c);
  xc_sigma_kernel<<<(ng + 255) / 256, 256, 0, c->stream>>>(x.grad.p, ng, x.sigma.p);
  LAUNCH_CHECK(c);
  xc_functional_kernel<<<(ng + 127) / 128, 128, 0, c->stream>>>(x.code, x.rho.p, x.sigma.p, ng, x.fx.p);
  LAUNCH_CHECK(c);
  double* part = c->red_part.ensure((size_t)REDUCE_BLOCKS * 9);
  double* out = c->red_out.ensure(64);
  xc_integrate_partial_kernel<<<REDUCE_BLOCKS, 256, 0, c->stream>>>(x.w.p, x.fx.p, x.rho.p, ng, part);
  LAUNCH_CHECK(c);
  xc_integrate_final_kernel<<<1, 32, 0, c->stream>>>(part, REDUCE_BLOCKS, out + 48);
  LAUNCH_CHECK(c);
  // M_s = w (1/2 vrho Phi + g . grad Phi), then V_s = Phi^T M_s + (Phi^T M_s)^T with the grid index split over slabs
  for (int g0 = 0; g0 < ng; g0 += 65535) {  // gridDim.y is limited to 65535
    dim3 gg((n + 127) / 128, std::min(65535, ng - g0), 2);
    xc_potential_rows_kernel<<<gg, 128, 0, c->stream>>>(x.ao.p, plane, x.w.p, x.fx.p, x.grad.p, g0, ng, n, x.TM.p);
    LAUNCH_CHECK(c);
  }
  const int nslab = ng / XC_CHUNK;
  for (int s = 0; s < 2; ++s) {
    gemm_tn(c, n, n, XC_CHUNK, x.ao.p, n, x.TM.p + s * plane, n, x.Vpart.p, n, 1.0, 0.0, nslab, (long)XC_CHUNK * n, (long)XC_CHUNK * n, nn);
    sum_partials_kernel<<<grid1(nn, 256), 256, 0, c->stream>>>(x.Vpart.p, nslab, nn, x.V.p + s * nn);
    LAUNCH_CHECK(c);
  }
  {
    dim3 g((n + 127) / 128, n, 2);
    add_transpose_kernel<<<g, 128, 0, c->stream>>>(x.V.p, n);
    LAUNCH_CHECK(c);
  }
  d2h(c, integrals3, out + 48, 3);
  NBD_CUDA(cudaStreamSynchronize(c->stream));
}

// Replaces: pyscf.dft.numint.NumInt.nr_uks(mol, grids, xc, dms) (+ libxc): dm host [2][nao][nao];
// outputs nelec[2], exc (the grid integral of the energy density; no exact-exchange part), vxc host [2][nao][nao].
extern "C" int nbd_xc_nr_uks(nbd_ctx* c, const double* dm, double* nelec, double* exc, double* vxc) {
  return guarded(c, [&] {
    c->timers.reset();
    NBD_REQUIRE(dm, NBD_ERR_ARG, "null density");
    NBD_REQUIRE(c->xc.ready, NBD_ERR_STATE, "nbd_xc_setup first");
    const long nn = (long)c->nao * c->nao;
    double* d_dm = c->dm0f.ensure((size_t)2 * nn);
    h2d(c, d_dm, dm, (size_t)2 * nn);
    double in3[3];
    xc_eval_device(c, d_dm, in3);
    if (exc) *exc = in3[0];
    if (nelec) {
      nelec[0] = in3[1];
      nelec[1] = in3[2];
    }
    if (vxc) d2h(c, vxc, c->xc.V.p, (size_t)2 * nn);
    finish_call(c);
  });
}

// Switches the Kohn-Sham branch of the SCF drivers on or off (needs nbd_xc_setup): get_veff = J - hyb K + V_xc with
// the .ecoul / .exc the reference reads (huzinaga_scf.py:55-56), calculate_ks_energy for the Huzinaga loop (:36-62).
extern "C" int nbd_scf_set_xc(nbd_ctx* c, int on) {
  if (!c) return NBD_ERR_ARG;
  if (on && !c->xc.ready) {
    c->err = "nbd_xc_setup first";
    return NBD_ERR_STATE;
  }
  c->xc.on = on != 0;
  c->bench_ready = false;
  return NBD_OK;
}
