"""Thin object wrapper over the C-ABI context: one per process and GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import NBD_HUZINAGA, NBD_MU_SHIFT, NbdError, ScfResult, f64, pinned_empty, ptr, result_empty

TIMER_KEYS = (
    "jk_x", "jk_rho", "jk_j", "jk_k", "jk_total", "allreduce", "fock", "diis", "orth", "eigh", "density", "energy",
    "eig_sub", "eig_bcast", "orth_gather", "iter_total", "scf_total", "ao2mo_half", "ao2mo_l", "ao2mo_eri", "ao2mo_perm", "ao2mo_total", "spinorb",
)


class B200Context:
    """Owns the device-resident 3-centre tensor shard and all workspaces of one GPU."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.nbd_create(C.byref(self._h), int(device))
        if rc != 0:
            raise NbdError(rc, f"nbd_create(device={device}) failed: no usable sm_100a GPU (no CPU fallback exists)")
        self.device = device
        self.nao = 0
        self.naux_local = 0
        self.rank, self.world = 0, 1

    # -- plumbing ------------------------------------------------------------------------------
    def _ck(self, rc: int):
        if rc != 0:
            raise NbdError(rc, self._lib.nbd_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.nbd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int):
        self._ck(self._lib.nbd_set_option(self._h, key.encode(), int(value)))

    def timer_ms(self, key: str) -> float:
        return float(self._lib.nbd_timer_ms(self._h, key.encode()))

    def timers(self) -> dict:
        return {k: self.timer_ms(k) for k in TIMER_KEYS if self.timer_ms(k) > 0.0}

    @property
    def launch_count(self) -> int:
        return int(self._lib.nbd_launch_count(self._h))

    # -- multi-GPU -----------------------------------------------------------------------------
    def comm_init_from_torch(self):
        """Join the NCCL communicator of the torch.distributed world (one process per GPU)."""
        import torch
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            rc = self._lib.nbd_comm_unique_id(C.cast(buf, C.c_void_p))
            if rc != 0:
                raise NbdError(rc, "nbd_comm_unique_id failed (libnccl not loadable)")
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            uid = uid.cuda()
        dist.broadcast(uid, src=0)
        raw = bytes(uid.cpu().tolist())
        buf = (C.c_ubyte * 128).from_buffer_copy(raw)
        self._ck(self._lib.nbd_comm_init(self._h, C.cast(buf, C.c_void_p), rank, world))
        self.rank, self.world = rank, world

    # -- 3-centre tensor -----------------------------------------------------------------------
    def cderi_alloc(self, nao: int, naux_local: int):
        self._ck(self._lib.nbd_cderi_alloc(self._h, int(nao), int(naux_local)))
        self.nao, self.naux_local = int(nao), int(naux_local)

    def cderi_upload(self, rows: np.ndarray, row0: int = 0):
        rows = f64(rows)
        assert rows.ndim == 2 and rows.shape[1] == self.nao * (self.nao + 1) // 2
        self._ck(self._lib.nbd_cderi_upload(self._h, ptr(rows), int(row0), rows.shape[0]))

    def cderi_synth(self, seed: int, scale: float, global_row0: int = 0):
        self._ck(self._lib.nbd_cderi_synth(self._h, int(seed), float(scale), int(global_row0)))

    def cderi_download(self, row0: int, nrows: int) -> np.ndarray:
        out = np.empty((nrows, self.nao * (self.nao + 1) // 2))
        self._ck(self._lib.nbd_cderi_download(self._h, ptr(out), int(row0), int(nrows)))
        return out

    def load_cderi(self, cderi: np.ndarray):
        """Allocate and upload a complete (already local) packed-lower tensor [naux, nao(nao+1)/2]."""
        cderi = f64(cderi)
        npair = cderi.shape[1]
        nao = int((np.sqrt(8 * npair + 1) - 1) // 2)
        assert nao * (nao + 1) // 2 == npair
        self.cderi_alloc(nao, cderi.shape[0])
        if cderi.shape[0]:
            self.cderi_upload(cderi, 0)

    # -- density-fitting integrals on the device (libcint-format basis of mol + auxmol) ---------------------
    @staticmethod
    def _basis_arrays(atm, bas, env):
        atm = np.ascontiguousarray(atm, dtype=np.int32).reshape(-1, 6)
        bas = np.ascontiguousarray(bas, dtype=np.int32).reshape(-1, 8)
        return atm, bas, f64(env)

    def basis_dims(self, bas, nbas_ao: int):
        bas = np.ascontiguousarray(bas, dtype=np.int32).reshape(-1, 8)
        nao, naux = C.c_int(), C.c_int()
        rc = self._lib.nbd_basis_dims(ptr(bas), bas.shape[0], int(nbas_ao), C.cast(C.byref(nao), C.c_void_p),
                                      C.cast(C.byref(naux), C.c_void_p))
        if rc != 0:
            raise NbdError(rc, "bad basis arrays")
        return nao.value, naux.value

    def int3c2e(self, atm, bas, env, nbas_ao: int):
        """Undecorated (P|mu>=nu) [naux, nao(nao+1)/2] and (P|Q) [naux, naux] (aux_e2 'int3c2e' s2ij, 'int2c2e')."""
        atm, bas, env = self._basis_arrays(atm, bas, env)
        nao, naux = self.basis_dims(bas, nbas_ao)
        j3c, j2c = np.empty((naux, nao * (nao + 1) // 2)), np.empty((naux, naux))
        self._ck(self._lib.nbd_int3c2e(self._h, ptr(atm), atm.shape[0], ptr(bas), bas.shape[0], ptr(env), env.size,
                                       int(nbas_ao), ptr(j3c), ptr(j2c)))
        return j3c, j2c

    def cderi_from_basis(self, atm, bas, env, nbas_ao: int, global_row0: int = 0, naux_local: int = -1):
        """``mf.density_fit()``: Cholesky-decorated tensor generated on the device, rows [global_row0, +naux_local)."""
        atm, bas, env = self._basis_arrays(atm, bas, env)
        nao, naux = self.basis_dims(bas, nbas_ao)
        self._ck(self._lib.nbd_cderi_from_basis(self._h, ptr(atm), atm.shape[0], ptr(bas), bas.shape[0], ptr(env), env.size,
                                                int(nbas_ao), int(global_row0), int(naux_local)))
        self.nao = nao
        self.naux_local = naux - global_row0 if naux_local < 0 else int(naux_local)
        return nao, naux

    # -- exchange-correlation (UKS objects) ------------------------------------------------------------------
    XC_CODES = {"b3lyp": 1, "lda": 2, "lda,vwn_rpa": 2}

    def xc_setup(self, xc: str, atm, bas, env, coords, weights):
        """AO values / gradients of the orbital basis (libcint arrays) on the caller's grid, resident on the device."""
        code = self.XC_CODES.get(str(xc).lower())
        if code is None:
            raise NbdError(-4, f"xc functional {xc!r} is not implemented on the device (available: {sorted(self.XC_CODES)})")
        atm, bas, env = self._basis_arrays(atm, bas, env)
        coords, weights = f64(coords), f64(weights)
        if coords.ndim != 2 or coords.shape[1] != 3 or weights.shape != (coords.shape[0],):
            raise ValueError("coords must be (ngrid, 3) and weights (ngrid,)")
        self._ck(self._lib.nbd_xc_setup(self._h, code, ptr(atm), atm.shape[0], ptr(bas), bas.shape[0], ptr(env), env.size,
                                        coords.shape[0], ptr(coords), ptr(weights)))

    def xc_nr_uks(self, dm):
        """(nelec [2], exc, vxc [2, nao, nao]) = NumInt.nr_uks for the spin densities dm [2, nao, nao]."""
        dm = f64(np.asarray(dm).reshape(2, self.nao, self.nao))
        nelec, exc, vxc = np.zeros(2), C.c_double(), np.empty((2, self.nao, self.nao))
        self._ck(self._lib.nbd_xc_nr_uks(self._h, ptr(dm), ptr(nelec), C.cast(C.byref(exc), C.c_void_p), ptr(vxc)))
        return nelec, exc.value, vxc

    def scf_set_xc(self, on: bool):
        self._ck(self._lib.nbd_scf_set_xc(self._h, int(bool(on))))

    # -- J/K -------------------------------------------------------------------------------------
    def jk_orbitals(self, orbs, signs=None, with_j=True, with_k=True):
        """orbs: list of (nao, ncol) scaled occupied-orbital blocks; returns (vj, vk) of shape (nset, nao, nao)."""
        n = self.nao
        nset = len(orbs)
        ncol = np.array([o.shape[1] for o in orbs], dtype=np.int32)
        flat = np.concatenate([f64(o).ravel() for o in orbs]) if ncol.sum() else np.zeros(1)
        sg = None
        if signs is not None:
            sg = np.concatenate([f64(s).ravel() for s in signs]) if ncol.sum() else np.zeros(1)
        vj = result_empty((nset, n, n)) if with_j else None
        vk = result_empty((nset, n, n)) if with_k else None
        self._ck(self._lib.nbd_jk(self._h, nset, ptr(ncol), ptr(flat), ptr(sg), ptr(vj), ptr(vk)))
        return vj, vk

    def jk_dm(self, dm, with_j=True, with_k=True):
        dm = f64(dm)
        shape = dm.shape
        dms = dm.reshape(-1, self.nao, self.nao)
        vj = result_empty(dms.shape) if with_j else None
        vk = result_empty(dms.shape) if with_k else None
        self._ck(self._lib.nbd_jk_dm(self._h, dms.shape[0], ptr(dms), ptr(vj), ptr(vk)))
        return (vj.reshape(shape) if with_j else None), (vk.reshape(shape) if with_k else None)

    # -- embedded SCF ---------------------------------------------------------------------------
    def scf_setup(self, nelec, ovlp, hcore, v_emb, dm_env, projector: int, mu: float = 0.0):
        """``hcore`` may be (n, n) or - once ``get_hcore`` has been patched by a previous embedding
        (nbed/driver.py:529,597) - spin-resolved (nspin, n, n); the reference simply broadcasts
        ``get_hcore() + embedding_potential`` (huzinaga_scf.py:140,157), so the spin-resolved part is folded into the
        potential with the same host addition and the C-ABI receives a zero (n, n) core Hamiltonian."""
        v_emb, dm_env, ovlp, hcore = f64(v_emb), f64(dm_env), f64(ovlp), f64(hcore)
        n = self.nao
        if v_emb.ndim != dm_env.ndim or v_emb.ndim not in (2, 3):
            raise ValueError("v_emb / dm_env must both be (n, n) or both (2, n, n)")
        nspin = 1 if v_emb.ndim == 2 else 2
        want = (n, n) if nspin == 1 else (2, n, n)
        if ovlp.shape != (n, n) or v_emb.shape != want or dm_env.shape != want:
            raise ValueError(f"shape mismatch: ovlp {ovlp.shape}, v_emb {v_emb.shape}, dm_env {dm_env.shape}; the device "
                             f"tensor has nao = {n}")
        if hcore.shape == want and nspin == 2:
            v_emb, hcore = f64(hcore + v_emb), np.zeros((n, n))
        elif hcore.shape != (n, n):
            raise ValueError(f"hcore has shape {hcore.shape}; expected ({n}, {n}) or {want}")
        ne = np.array(list(nelec), dtype=np.int32)
        if ne.shape != (2,):
            raise ValueError("nelec must be (n_alpha, n_beta)")
        self._nspin = nspin
        self._ck(self._lib.nbd_scf_setup(self._h, nspin, ptr(ne), ptr(ovlp), ptr(hcore), ptr(v_emb),
                                         ptr(dm_env), int(projector), float(mu)))

    def scf_set_env_orbitals(self, c_env):
        """Low-rank factor of the environment density (dm_env = c_env c_env^T, shape (n, r) or (2, n, r))."""
        n, ns = self.nao, self._nspin
        c_env = f64(c_env)
        r = c_env.shape[-1]
        if c_env.shape != ((n, r) if ns == 1 else (2, n, r)):
            raise ValueError(f"c_env has shape {c_env.shape}")
        self._ck(self._lib.nbd_scf_set_env_orbitals(self._h, int(r), ptr(c_env)))

    def scf_set_virtual_projector(self, dm_env_virt):
        n, ns = self.nao, self._nspin
        shape = (n, n) if ns == 1 else (2, n, n)
        d = None if dm_env_virt is None else f64(np.asarray(dm_env_virt).reshape(shape))
        self._ck(self._lib.nbd_scf_set_virtual_projector(self._h, ptr(d)))

    def huzinaga_scf(self, max_cycle, conv_tol, dm_conv_tol=1e-6, use_diis=True, dm0=None):
        n, ns = self.nao, self._nspin
        shape = (n, n) if ns == 1 else (2, n, n)
        c, dm, huz = result_empty(shape), result_empty(shape), result_empty(shape)
        e = np.empty(shape[:-1])
        trace = np.zeros((max_cycle, 3))
        res = ScfResult()
        d0 = None if dm0 is None else f64(np.asarray(dm0).reshape(shape))
        self._ck(self._lib.nbd_huzinaga_scf(self._h, int(max_cycle), float(conv_tol), float(dm_conv_tol),
                                            int(bool(use_diis)), ptr(d0), ptr(c), ptr(e), ptr(dm), ptr(huz),
                                            ptr(trace), C.byref(res)))
        info = dict(converged=bool(res.converged), cycles=res.cycles, energy=np.array(res.energy[:ns]),
                    norm_ddm=res.norm_ddm, trace=trace[: res.cycles].copy())
        return c, e, dm, huz, info

    def mu_scf(self, max_cycle, conv_tol, e_nuc, dm0):
        n, ns = self.nao, self._nspin
        shape = (n, n) if ns == 1 else (2, n, n)
        c, dm, vhf = result_empty(shape), result_empty(shape), result_empty(shape)
        e, occ = np.empty(shape[:-1]), np.empty(shape[:-1])
        trace = np.zeros((max_cycle + 1, 3))
        res = ScfResult()
        d0 = f64(np.asarray(dm0).reshape(shape))
        self._ck(self._lib.nbd_mu_scf(self._h, int(max_cycle), float(conv_tol), float(e_nuc), ptr(d0), ptr(c), ptr(e),
                                      ptr(occ), ptr(dm), ptr(vhf), ptr(trace), C.byref(res)))
        ntr = res.cycles + (1 if res.converged or trace[res.cycles].any() else 0)
        info = dict(converged=bool(res.converged), cycles=res.cycles, e_tot=res.e_tot, norm_ddm=res.norm_ddm,
                    norm_grad=res.norm_grad, trace=trace[:ntr].copy())
        return c, e, occ, dm, vhf, info

    def scf_bench_init(self):
        self._ck(self._lib.nbd_scf_bench_init(self._h))

    def scf_bench_iteration(self, it: int):
        e = np.zeros(2)
        nd = C.c_double()
        self._ck(self._lib.nbd_scf_bench_iteration(self._h, int(it), ptr(e), C.cast(C.byref(nd), C.c_void_p)))
        return e, nd.value

    # -- active-space integrals -------------------------------------------------------------------
    def ao2mo(self, ca, cb=None):
        ca = f64(ca)
        m = ca.shape[1]
        cbp = None if cb is None else f64(cb)
        out = result_empty((4, m, m, m, m))
        self._ck(self._lib.nbd_ao2mo(self._h, m, ptr(ca), ptr(cbp), ptr(out)))
        return out

    def one_body(self, hcore, ca, cb=None):
        hcore, ca = f64(hcore), f64(ca)
        m = ca.shape[1]
        nspin_h = 1 if hcore.ndim == 2 else 2
        cbp = None if cb is None else f64(cb)
        out = np.empty((2, m, m))
        self._ck(self._lib.nbd_one_body(self._h, m, nspin_h, ptr(hcore), ptr(ca), ptr(cbp), ptr(out)))
        return out

    def spinorb_from_spatial(self, one, two, eq_tol=1e-8, two_body_scale=1.0):
        one, two = f64(one), f64(two)
        m = one.shape[-1]
        h1 = np.empty((2 * m, 2 * m))
        h2 = result_empty((2 * m,) * 4)
        self._ck(self._lib.nbd_spinorb_from_spatial(self._h, m, ptr(one), ptr(two), float(eq_tol),
                                                    float(two_body_scale), ptr(h1), ptr(h2)))
        return h1, h2


    def build_hamiltonian(self, hcore, ca, cb=None, eq_tol=1e-8, two_body_scale=0.5):
        """(h1 [2m, 2m], h2 [2m]^4) of HamiltonianBuilder.build() with the MO integrals kept on the device."""
        hcore, ca = f64(hcore), f64(ca)
        m = ca.shape[1]
        nspin_h = 1 if hcore.ndim == 2 else 2
        cbp = None if cb is None else f64(cb)
        h1 = np.empty((2 * m, 2 * m))
        h2 = result_empty((2 * m,) * 4)
        self._ck(self._lib.nbd_build_hamiltonian(self._h, m, nspin_h, ptr(hcore), ptr(ca), ptr(cbp), float(eq_tol),
                                                 float(two_body_scale), ptr(h1), ptr(h2)))
        return h1, h2

    

__all__ = ["B200Context", "NBD_HUZINAGA", "NBD_MU_SHIFT", "NbdError"]
