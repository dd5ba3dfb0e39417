/* nbed_b200 — C-ABI of the B200-native Nbed hot path (sm_100a).
 *
 * Drop-in boundary.  The reference (UCL-CCS/Nbed) is pure Python on top of PySCF; the
 * arithmetic of its hot path is reached through PySCF's ctypes bindings.  Every entry point below states
 * the reference interface it replaces (paths relative to the reference tree).  Conventions mirror
 * PySCF's own ctypes style: C-contiguous float64 host arrays passed as plain pointers, sizes by value,
 * caller allocates outputs, calls block until the result is in the caller's buffer, no exceptions cross
 * the ABI: every function returns 0 on success or a negative nbd_status, and nbd_last_error() gives text.
 *
 * One context per process and per GPU (one process per GPU; multi-GPU = aux-index shards + NCCL
 * all-reduce, see nbd_comm_init).  A context is not re-entrant.
 */
#ifndef NBED_B200_H
#define NBED_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nbd_ctx nbd_ctx;

enum nbd_status {
  NBD_OK = 0,
  NBD_ERR_CUDA = -1,       /* a CUDA / cuSOLVER / NCCL call failed                         */
  NBD_ERR_ARG = -2,        /* bad argument (shape, null pointer, state)                      */
  NBD_ERR_STATE = -3,      /* call order violated (e.g. J/K before the 3-centre tensor)      */
  NBD_ERR_UNSUPPORTED = -4 /* outside the implemented envelope (e.g. nao > 3072)             */
};

enum nbd_projector { NBD_HUZINAGA = 0, NBD_MU_SHIFT = 1 };

/* ---- context ------------------------------------------------------------------------------- */
int nbd_version(void);
int nbd_create(nbd_ctx** out, int device);
int nbd_destroy(nbd_ctx* ctx);
const char* nbd_last_error(nbd_ctx* ctx);
/* Tuning / debugging knobs (listed with nbd_timer_ms below).  Returns NBD_ERR_ARG for an unknown key. */
int nbd_set_option(nbd_ctx* ctx, const char* key, long value);
/* Per-stage device timings (CUDA events) of the last call, in milliseconds.
 * keys: "jk_x" (pass 1), "jk_rho", "jk_j" (pass 2), "jk_k" (Gram), "jk_total", "allreduce", "fock", "diis", "orth",
 *       "eigh" (cuSOLVER), "eig_sub" (filtered subspace iteration), "eig_bcast", "density", "energy", "iter_total",
 *       "scf_total", "ao2mo_half", "ao2mo_l", "ao2mo_eri", "ao2mo_perm", "ao2mo_total", "spinorb", "build_total", "int3c2e", "cholesky", "xc_ao", "xc".
 * Cumulative counters since context creation: "count:sub_applies", "count:sub_outer", "count:sub_fallbacks",
 * "count:sub_cold_starts", "count:sub_lanczos", "count:sub_rejects" (SCF runs redone with cuSOLVER because the tracked
 * block's Ritz values disagreed with the final full spectrum).
 * Options of nbd_set_option: "jk_variant", "gemm_variant" (1 = simple reference kernels), "jpass_variant"
 * (0 = TMA-fed, 1 = LDG streaming), "eig_mode" (0 = cuSOLVER every cycle, 1 = subspace tracking), "sub_min_nao",
 * "overlap" (0 = pass 2 behind the K Gram, 1 = side stream, 3 = same stream with programmatic dependent launch),
 * "sub_bound" (1 = Lanczos + ||dF||_F spectral bounds, 0 = row sums), "sub_cold" (1 = initial guess by cold-start
 * subspace iteration, 0 = cuSOLVER), "dist_eig", "panel_stages", "panel_hybrid" (1 = 9-10 column slices run 8 columns on DMMA + 1-2 on the
 * FMA pipe, 0 = padded to 16 DMMA columns), "x_budget_mb", "timers"; round 2: "sub_apply_variant" (block product of the
 * subspace eigensolver: 0 = automatic, 1 = one CTA per 16 rows, 2 = single-shot 8-CTA-cluster kernel, 3 = 4-CTA-cluster
 * ring kernel), "sub_pdl" (1 = consecutive block products overlap through programmatic dependent launch), "pair_split"
 * (pass-2 SMs of the pair [pass 2 || K Gram]; -1 = automatic), "pair_guest" (pass-2 guest CTAs per Gram SM),
 * "pair_guest_reserve", "small_eigh" (1 = one-CTA Jacobi eigensolver for nao <= 32, 0 = cuSOLVER), "early_export"
 * (1 = nbd_huzinaga_scf copies D / Huz out while the final eigensolve runs), "x_cache", "copy_threads". */
double nbd_timer_ms(nbd_ctx* ctx, const char* key);
/* Number of kernels launched by this library since the context was created (bench "gpu_launches"). */
long nbd_launch_count(nbd_ctx* ctx);
/* Page-locked host buffers for results (optional): device->host copies into them run at full PCIe speed instead
 * of being staged through the driver (numpy arrays of the Python layer are allocated with these). */
void* nbd_host_alloc(size_t bytes);
void nbd_host_free(void* p);

/* ---- multi-GPU plumbing -------------------------------------------------------------------- */
/* NCCL communicator over the ranks that share one 3-centre tensor by auxiliary index.
 * unique_id: the 128 bytes of ncclUniqueId made by rank 0 (nbd_comm_unique_id) and broadcast by the host
 * layer (torch.distributed).  After this call J/K and ao2mo results are all-reduced across ranks. */
int nbd_comm_unique_id(void* unique_id_128);
int nbd_comm_init(nbd_ctx* ctx, const void* unique_id_128, int rank, int world);

/* ---- 3-centre tensor (P|mu nu): uploaded once per geometry, NOT in the timed loop ----------- */
/* Replaces: pyscf `mf.with_df._cderi` (what `mf.density_fit()` builds through libcint; reference J/K call
 * sites nbed/scf/huzinaga_scf.py:156, nbed/scf/embedded_hcore_funcs.py:34, nbed/driver.py:533).
 * naux_local rows of this rank's shard are stored in the tiled-triangular device layout. */
int nbd_cderi_alloc(nbd_ctx* ctx, int nao, int naux_local);
/* rows [row0, row0+nrows) of the LOCAL shard, host layout = PySCF packed-lower [nrows, nao(nao+1)/2]. */
int nbd_cderi_upload(nbd_ctx* ctx, const double* cderi_rows, int row0, int nrows);
/* Synthetic tensor filled on the device (bench / large-size tests): value of global row
 * (global_row0 + local row) as defined by nbed_b200/synthetic.py:hash_uniform. */
int nbd_cderi_synth(nbd_ctx* ctx, unsigned long long seed, double scale, int global_row0);
/* Read back local rows in packed-lower layout (round-trip test of the layout transform). */
int nbd_cderi_download(nbd_ctx* ctx, double* cderi_rows, int row0, int nrows);

/* Density-fitting integrals generated ON THE DEVICE from a libcint-format basis.
 * Replaces: pyscf.df.incore.cholesky_eri(mol, auxbasis) - aux_e2(mol, auxmol, 'int3c2e', aosym='s2ij'),
 * auxmol.intor('int2c2e'), Cholesky decoration - which mf.density_fit() runs through libcint once per geometry
 * (J/K call sites: nbed/scf/huzinaga_scf.py:156, nbed/driver.py:533).  Not part of the timed loop.
 * atm [natm][6], bas [nbas][8], env [nenv]: PySCF's mol._atm / _bas / _env of the concatenated mol + auxmol
 * (gto.mole.conc_mol, as aux_e2 builds it): orbital shells [0, nbas_ao), auxiliary shells [nbas_ao, nbas); real
 * spherical functions; coefficients as PySCF stores them (primitive and contraction normalisation folded in);
 * orbital l <= 3, auxiliary l <= 4, general contractions (nctr > 1) accepted.
 * nbd_int3c2e: undecorated (P|mu>=nu) [naux][nao(nao+1)/2] and (P|Q) [naux][naux] to host buffers (either may be NULL).
 * nbd_cderi_from_basis: allocates the device tensor (nao from the basis) and fills aux rows
 * [global_row0, global_row0 + naux_local) of cderi = L^-1 (P|mu nu), (P|Q) = L L^T; naux_local = -1: all rows.
 * nbd_basis_dims: nao / naux of such a basis (no context needed). */
int nbd_basis_dims(const int* bas, int nbas, int nbas_ao, int* nao, int* naux);
int nbd_int3c2e(nbd_ctx* ctx, const int* atm, int natm, const int* bas, int nbas, const double* env, int nenv,
                int nbas_ao, double* j3c, double* j2c);
int nbd_cderi_from_basis(nbd_ctx* ctx, const int* atm, int natm, const int* bas, int nbas, const double* env, int nenv,
                         int nbas_ao, int global_row0, int naux_local);

/* ---- exchange-correlation (Kohn-Sham objects) ------------------------------------------------------ */
/* Replaces, for UKS objects, what scf_method.get_veff reaches beyond J/K (nbed/scf/huzinaga_scf.py:55,156 with the
 * objects of nbed/driver.py:289-313; calculate_ks_energy :36-62): pyscf.dft.numint.eval_ao + NumInt.nr_uks + libxc.
 * The quadrature is an INPUT: coords [ngrid][3] (Bohr) and weights [ngrid] = PySCF's mf.grids.coords / weights.
 * atm / bas / env: libcint arrays of the ORBITAL basis (all nbas shells), l <= 3.  xc_code: 1 = 'b3lyp' (libxc
 * HYB_GGA_XC_B3LYP: 0.08 Slater + 0.72 B88 + 0.19 VWN_RPA + 0.81 LYP, 20 % exact exchange), 2 = Slater + VWN_RPA (LDA).
 * AO values and gradients on the grid stay resident on the device.  After the 3-centre tensor exists. */
int nbd_xc_setup(nbd_ctx* ctx, int xc_code, const int* atm, int natm, const int* bas, int nbas, const double* env,
                 int nenv, int ngrid, const double* coords, const double* weights);
/* NumInt.nr_uks: dm host [2][nao][nao] -> nelec[2], exc (grid integral of the energy density, without the
 * exact-exchange part), vxc host [2][nao][nao].  Any output may be NULL. */
int nbd_xc_nr_uks(nbd_ctx* ctx, const double* dm, double* nelec, double* exc, double* vxc);
/* on = 1: nbd_huzinaga_scf / nbd_mu_scf treat the SCF object as UKS: get_veff = J - hyb K + V_xc, the Huzinaga loop
 * uses calculate_ks_energy (ecoul + exc + tr[D (h + Huz + V)]); nbd_mu_scf keeps nbed's patched energy_elec unless
 * option "ks_energy" = 1 (plain pyscf UKS.energy_elec: e1 + ecoul + exc).  nbd_huzinaga_scf also takes restricted
 * Kohn-Sham objects (nspin = 1, pyscf/dft/rks.py: J - hyb / 2 K + V_xc[D / 2, D / 2], scalar energy); nbd_mu_scf is
 * spin-resolved like the reference's mu path. */
int nbd_scf_set_xc(nbd_ctx* ctx, int on);

/* ---- J/K --------------------------------------------------------------------------------------- */
/* Replaces: pyscf.df.df_jk.get_jk occupied-orbital branch (reached from scf_method.get_veff / get_jk /
 * get_j: nbed/scf/huzinaga_scf.py:55,156; nbed/scf/embedded_hcore_funcs.py:34; nbed/driver.py:344-345,
 * 391,533,627,847,849).
 * orb: host [nset][nao][ncol[s]] concatenated per set, C-contiguous: scaled occupied orbitals
 *      C_occ*sqrt(occ) (or signed eigen-factors of a dense density, sign[] = +1/-1 per column, may be NULL).
 * vj : host [nset][nao][nao] (NULL to skip J)   vk: host [nset][nao][nao] (NULL to skip K).
 * Results are summed over all ranks of the communicator. */
int nbd_jk(nbd_ctx* ctx, int nset, const int* ncol, const double* orb, const double* sign,
           double* vj, double* vk);
/* Dense-density entry (pyscf get_jk without mo_coeff tags): dm host [nset][nao][nao], symmetric. */
int nbd_jk_dm(nbd_ctx* ctx, int nset, const double* dm, double* vj, double* vk);

/* ---- embedded SCF ------------------------------------------------------------------------------ */
/* Static matrices of one embedded SCF problem.
 * Replaces the set-up of nbed/scf/huzinaga_scf.py:126-136 (S, S^-1/2, gamma*S) and of
 * nbed/driver.py:433-449,518 (mu * S gamma S + V folded into the core Hamiltonian).
 * ovlp [nao][nao]; hcore [nao][nao]; v_emb [nspin][nao][nao]; dm_env [nspin][nao][nao] (nspin = 1: RHF
 * rank-2 convention with the doubled density and the -1/2 factor of huzinaga_scf.py:80). */
int nbd_scf_setup(nbd_ctx* ctx, int nspin, const int* nelec, const double* ovlp, const double* hcore,
                  const double* v_emb, const double* dm_env, int projector, double mu);

/* Optional virtual-orbital environment projector of huzinaga_scf (argument dm_environment_virtual,
 * nbed/scf/huzinaga_scf.py:96,133-136 and the second term of get_huzinaga_operator :82-88; built by the PAO
 * localizer in nbed/driver.py:566-575).  dm_env_virt [nspin][nao][nao], NULL switches it off.  After nbd_scf_setup. */
int nbd_scf_set_virtual_projector(nbd_ctx* ctx, const double* dm_env_virt);

/* Optional low-rank factor of the occupied environment density: dm_env[s] = c_env[s] c_env[s]^T, c_env host
 * [nspin][nao][r] - the localizer's c_enviro (nbed/localizers/system.py:33 builds dm_enviro from exactly this product).
 * With it the projector product F gamma S of get_huzinaga_operator (nbed/scf/huzinaga_scf.py:77) runs as two rank-r
 * GEMMs (4 n^2 r flop instead of 2 n^3).  The factor is verified on the device against the dm_env of nbd_scf_setup;
 * NBD_ERR_ARG (and the dense product stays in use) when it does not reproduce it.  After nbd_scf_setup. */
int nbd_scf_set_env_orbitals(nbd_ctx* ctx, int r, const double* c_env);

typedef struct nbd_scf_result {
  int converged;      /* conv flag                                              */
  int cycles;         /* number of Fock builds executed in the loop             */
  double e_tot;       /* mu path: PySCF e_tot (incl. e_nuc passed in); Huzinaga: sum of last scf_energy */
  double energy[2];   /* Huzinaga: per-spin scf_energy of the last cycle (huzinaga_scf.py:185) */
  double norm_ddm;    /* last density change                                    */
  double norm_grad;   /* mu path: last orbital-gradient norm                    */
} nbd_scf_result;

/* Replaces: nbed.scf.huzinaga_scf (nbed/scf/huzinaga_scf.py:93-206), HF objects.
 * dm0: optional initial guess [nspin][nao][nao] (NULL = core guess of :139-148).
 * Outputs (caller-allocated, may be NULL): mo_coeff [nspin][nao][nao] (columns = MOs),
 * mo_energy [nspin][nao], dm [nspin][nao][nao], huz [nspin][nao][nao] (pre-DIIS operator of the last
 * cycle), trace [max_cycle][3] = per-cycle (E_alpha, E_beta, max ||dD||_F). */
int nbd_huzinaga_scf(nbd_ctx* ctx, int max_cycle, double conv_tol, double dm_conv_tol, int use_diis,
                     const double* dm0, double* mo_coeff, double* mo_energy, double* dm, double* huz,
                     double* trace, nbd_scf_result* result);

/* Replaces: pyscf.scf.hf.kernel as driven by NbedDriver._mu_embed (nbed/driver.py:500-538) with the
 * patched get_hcore (:529) and energy_elec (nbed/scf/embedded_hcore_funcs.py:11-46); CDIIS error vector
 * F D S - S D F; generalised eigensolve = cuSOLVER dsygvd.  dm0 [nspin][nao][nao] is required.
 * trace [max_cycle+1][3] = per-cycle (e_tot, ||g||, ||dD||). */
int nbd_mu_scf(nbd_ctx* ctx, int max_cycle, double conv_tol, double e_nuc, const double* dm0,
               double* mo_coeff, double* mo_energy, double* mo_occ, double* dm, double* vhf,
               double* trace, nbd_scf_result* result);

/* One Fock-build iteration on device-resident state, for benchmarks (K1-K6 of SURVEY.md §2.3):
 * J/K from the current occupied orbitals, Fock + projector, DIIS push/extrapolate, orthogonalise,
 * eigensolve, back-transform, density, energy and convergence scalars.  Requires nbd_scf_setup and a
 * prior nbd_scf_bench_init (core guess).  with_eigh = 0 skips nothing: the eigensolve is always executed;
 * its time is reported separately under the "eigh" timer key. */
int nbd_scf_bench_init(nbd_ctx* ctx);
int nbd_scf_bench_iteration(nbd_ctx* ctx, int iter, double* energy2, double* norm_ddm);

/* ---- active-space AO->MO transform --------------------------------------------------------------- */
/* Replaces: HamiltonianBuilder._two_body_integrals (nbed/ham_builder.py:98-156): the four
 * pyscf.ao2mo.kernel + ao2mo.restore(1) + transpose(0,2,3,1) blocks.
 * ca, cb: host [nao][m] (pass cb = ca or NULL for the restricted case).
 * out: host [4][m][m][m][m], out[blk][p][r][s][q] = (p q | r s), blk = aaaa, bbbb, aabb, bbaa. */
int nbd_ao2mo(nbd_ctx* ctx, int m, const double* ca, const double* cb, double* out);
/* Replaces: HamiltonianBuilder._one_body_integrals (nbed/ham_builder.py:53-96): C^T h C per spin.
 * hcore [nspin_h][nao][nao] (nspin_h = 1 or 2), out [2][m][m]. */
int nbd_one_body(nbd_ctx* ctx, int m, int nspin_h, const double* hcore, const double* ca, const double* cb,
                 double* out);
/* Replaces: HamiltonianBuilder.build() end to end (nbed/ham_builder.py:218-254) = _one_body_integrals +
 * _two_body_integrals + _spinorb_from_spatial + the 0.5 factor, with the MO integrals kept on the device
 * (only h1 [2m][2m] and h2 [2m]^4 are copied back).  Arguments as nbd_one_body / nbd_ao2mo /
 * nbd_spinorb_from_spatial. */
int nbd_build_hamiltonian(nbd_ctx* ctx, int m, int nspin_h, const double* hcore, const double* ca, const double* cb,
                          double eq_tol, double two_body_scale, double* h1, double* h2);
/* Replaces: HamiltonianBuilder._spinorb_from_spatial + the 0.5 factor of build()
 * (nbed/ham_builder.py:158-216,254).  one [2][m][m], two [4][m][m][m][m] (host);
 * h1 [2m][2m], h2 [2m][2m][2m][2m] (host); |x| < eq_tol -> 0; h2 is scaled by two_body_scale (0.5). */
int nbd_spinorb_from_spatial(nbd_ctx* ctx, int m, const double* one, const double* two, double eq_tol,
                             double two_body_scale, double* h1, double* h2);

#ifdef __cplusplus
}
#endif
#endif /* NBED_B200_H */
