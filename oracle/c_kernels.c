/* Plain-C restatement of the hot contraction of the Nbed embedded-SCF Fock build.  TEST INFRASTRUCTURE ONLY
 * (checker and reported CPU baseline; never linked or called by the product).
 *
 * Restates pyscf.df.df_jk.get_jk (pyscf 2.9.0, reached from nbed/scf/huzinaga_scf.py:156 and
 * nbed/scf/embedded_hcore_funcs.py:34 of the reference) for densities given by their occupied orbitals:
 *     rho_P   = sum_{mu nu} B[P,mu nu] D[mu nu]          (dmtril @ eri1.T)
 *     vj      = sum_P rho_P B[P]                          (rho @ eri1, unpack_tril)
 *     buf1    = B[P] * orbo  ;  vk += buf1^T buf1         (occupied-orbital K)
 * cderi is PySCF's packed-lower layout [naux][nao(nao+1)/2].  OpenMP over the auxiliary index like libcvhf.
 */
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* pyscf.lib.unpack_tril (NPdunpack_tril): packed rows -> full symmetric matrices */
void oracle_unpack_tril(const double* packed, double* out, long nrows, int n) {
  const long npair = (long)n * (n + 1) / 2, nn = (long)n * n;
#pragma omp parallel for schedule(static)
  for (long p = 0; p < nrows; ++p) {
    const double* src = packed + p * npair;
    double* dst = out + p * nn;
    for (int i = 0; i < n; ++i) {
      const double* row = src + (long)i * (i + 1) / 2;
      for (int j = 0; j <= i; ++j) {
        dst[(long)i * n + j] = row[j];
        dst[(long)j * n + i] = row[j];
      }
    }
  }
}

/* vj [nset][n][n], vk [nset][n][n] from scaled occupied orbitals orb[s] = [n][ncol[s]] (concatenated). */
void oracle_df_jk_occ(const double* cderi, long naux, int n, int nset, const int* ncol, const double* orb,
                      double* vj, double* vk) {
  const long npair = (long)n * (n + 1) / 2, nn = (long)n * n;
  int ntot = 0;
  for (int s = 0; s < nset; ++s) ntot += ncol[s];
  memset(vj, 0, sizeof(double) * nset * nn);
  memset(vk, 0, sizeof(double) * nset * nn);
  double* rho = (double*)calloc((size_t)naux * nset, sizeof(double));
#pragma omp parallel
  {
    double* x = (double*)malloc(sizeof(double) * (size_t)n * (ntot > 0 ? ntot : 1)); /* buf1[mu][i] */
    double* kloc = (double*)calloc((size_t)nset * nn, sizeof(double));
#pragma omp for schedule(dynamic, 4)
    for (long p = 0; p < naux; ++p) {
      const double* b = cderi + p * npair;
      /* buf1 = B_P * orbo for all sets (symmetric packed matrix times skinny block) */
      memset(x, 0, sizeof(double) * (size_t)n * ntot);
      long off = 0;
      int c0 = 0;
      for (int s = 0; s < nset; ++s) {
        const int w = ncol[s];
        const double* c = orb + off;
        for (int i = 0; i < n; ++i) {
          const double* row = b + (long)i * (i + 1) / 2;
          double* xi = x + (long)i * ntot + c0;
          const double* ci = c + (long)i * w;
          for (int j = 0; j < i; ++j) {
            const double v = row[j];
            double* xj = x + (long)j * ntot + c0;
            const double* cj = c + (long)j * w;
            for (int k = 0; k < w; ++k) {
              xi[k] += v * cj[k];
              xj[k] += v * ci[k];
            }
          }
          for (int k = 0; k < w; ++k) xi[k] += row[i] * ci[k];
        }
        /* rho_P = sum_i c_i^T B c_i */
        double r = 0.0;
        for (int i = 0; i < n; ++i)
          for (int k = 0; k < w; ++k) r += c[(long)i * w + k] * x[(long)i * ntot + c0 + k];
        rho[p * nset + s] = r;
        /* vk_s += buf1 buf1^T (lower triangle) */
        double* ks = kloc + (long)s * nn;
        for (int i = 0; i < n; ++i) {
          const double* xi2 = x + (long)i * ntot + c0;
          for (int j = 0; j <= i; ++j) {
            const double* xj2 = x + (long)j * ntot + c0;
            double acc = 0.0;
            for (int k = 0; k < w; ++k) acc += xi2[k] * xj2[k];
            ks[(long)i * n + j] += acc;
          }
        }
        off += (long)n * w;
        c0 += w;
      }
    }
#pragma omp critical
    for (long e = 0; e < (long)nset * nn; ++e) vk[e] += kloc[e];
    free(kloc);
    free(x);
  }
  /* vj_s = sum_P rho_P B_P */
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    for (int s = 0; s < nset; ++s) {
      double* jrow = vj + (long)s * nn + (long)i * n;
      for (long p = 0; p < naux; ++p) {
        const double r = rho[p * nset + s];
        const double* row = cderi + p * npair + (long)i * (i + 1) / 2;
        for (int j = 0; j <= i; ++j) jrow[j] += r * row[j];
      }
    }
  }
  for (int s = 0; s < nset; ++s)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j) {
        vj[(long)s * nn + (long)j * n + i] = vj[(long)s * nn + (long)i * n + j];
        vk[(long)s * nn + (long)j * n + i] = vk[(long)s * nn + (long)i * n + j];
      }
  free(rho);
}

/* Counter-based synthetic 3-centre tensor of nbed_b200/synthetic.py (SplitMix64 of the packed index), bit-identical to
 * synthetic.hash_uniform / the device generator: rows [row0, row0 + nrows) of B[P][mu >= nu], packed-lower.  Lets the
 * oracle stream the 31 GB tensor of BASELINE config 4 block by block, the way pyscf's with_df.loop() feeds
 * df_jk.get_jk from disk. */
static inline unsigned long long oracle_splitmix64(unsigned long long x) {
  unsigned long long z = x + 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
void oracle_synth_rows(unsigned long long seed, int n, double scale, long row0, long nrows, double* out) {
  const unsigned long long npair = (unsigned long long)n * (n + 1) / 2;
#pragma omp parallel for schedule(static)
  for (long r = 0; r < nrows; ++r) {
    const unsigned long long base = (seed << 48) + (unsigned long long)(row0 + r) * npair;
    double* dst = out + (unsigned long long)r * npair;
    for (unsigned long long k = 0; k < npair; ++k) {
      const unsigned long long z = oracle_splitmix64(base + k);
      const double u = (double)(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
      dst[k] = u * scale;
    }
  }
}
