"""CPU oracle for the Nbed hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the checker or
as the reported CPU baseline.  The product path (``nbed_b200``) never imports this package and fails
loudly when its CUDA library is missing.

Contents
--------
``pyscf_restatement``  NumPy FP64 restatement of the PySCF 2.9.0 routines the reference reaches on the
                       hot path (pyscf is pinned in /root/reference/uv.lock:2571-2572 and is NOT
                       installed here): DF ``get_jk``, ``get_veff``, ``lib.diis.DIIS``, ``scf.diis.CDIIS``,
                       ``get_occ``, ``make_rdm1``, ``scf.hf.kernel``, DF ``ao2mo`` and ``ao2mo.restore``.
``nbed_restatement``   NumPy restatement of the reference's own hot-path functions, each citing the
                       reference file:line it follows.
``stubs``              ~stub ``pyscf`` / ``openfermion`` modules so the reference's own pure-NumPy code
                       (``huzinaga_scf``, ``get_huzinaga_operator``, ``energy_elec``,
                       ``_spinorb_from_spatial``) can be imported UNMODIFIED from /root/reference in the
                       dev container to validate the restatement and to generate ``tests/golden``.
``gaussian_integrals`` minimal s/p Gaussian integral generator (libcint is absent) used to rebuild the
                       water/STO-3G system of the reference's tests, so that the oracle can be pinned on
                       the reference's own golden energies (tests/test_driver.py:56-57,76).
``c_kernels.c``        plain-C (OpenMP) restatement of the DF-J/K contraction, built by ``oracle/Makefile`` into
                       ``oracle/_build`` and bound by ``c_binding``; cross-checked against the NumPy restatement.
``fock_space``         Fock-space matrix and Jordan-Wigner Pauli terms of a second-quantised Hamiltonian (stands in
                       for openfermion in the builder tests).
(The shape-synthetic problem generator of SURVEY.md §8(d) lives in ``nbed_b200/synthetic.py``: it is shared by the
product's benchmark and the tests and contains no reference arithmetic.)

Parity status: PINNED for the pieces listed in DESIGN.md §oracle (reference-run fixtures + the
water/STO-3G HF/FCI goldens); the DFT-derived known answers of tests/test_scf.py need libxc and remain
unpinned (stated in DESIGN.md).
"""
