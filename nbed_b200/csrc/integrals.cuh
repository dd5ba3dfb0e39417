// Three-centre / two-centre Coulomb integrals over contracted real-spherical Gaussians on the device
// (SURVEY.md section 8(f) rank 4): the tensor the reference obtains once per geometry from libcint through
//   mf.density_fit()  ->  pyscf.df.incore.cholesky_eri  ->  aux_e2(mol, auxmol, 'int3c2e', aosym='s2ij'),
//                                                            auxmol.intor('int2c2e')
// (call sites of the resulting J/K: nbed/scf/huzinaga_scf.py:156, nbed/driver.py:533).  Not part of the timed SCF
// loop; a plain McMurchie-Davidson generator: Hermite expansion of every primitive pair, Hermite Coulomb integrals
// R_tuv from the Boys function, Cartesian -> real-spherical transformation with libcint's conventions
// (CINTcommon_fac_sp for s and p, unit-normalised real solid harmonics for l >= 2, m = -l..l, Cartesian components in
// the order xx, xy, xz, yy, yz, zz).  Coefficients arrive already normalised the way PySCF stores them in mol._env.
//
// Decomposition: one CTA per AO shell pair (i >= j).  The CTA expands the pair's primitive products once into shared
// memory (p, P, coefficient, E^x, E^y, E^z tables); each thread then takes auxiliary shells k = tid, tid + 64, ...,
// builds R_tuv per primitive triple in local memory and accumulates the Cartesian block, transforms it and writes its
// rows of the packed-lower result [naux][nao (nao + 1) / 2] - disjoint writes, no atomics, deterministic.
#pragma once
#include "common.cuh"

namespace nbd {

constexpr int INT_LMAX_AO = 3;   // s, p, d, f orbital shells
constexpr int INT_LMAX_AUX = 4;  // up to g auxiliary shells
constexpr int INT_THREADS = 64;

struct IntShell {
  double x, y, z;
  int l, nprim;
  int ptr_exp, ptr_coef;  // into env
  int ao_off;             // first spherical function of the shell
};

struct IntArgs {
  const IntShell* ao;   // [nsh_ao]
  const IntShell* aux;  // [nsh_aux]
  const double* env;
  const double* c2s;  // cart -> spherical matrices, l = 0 .. 4 back to back (c2s_off)
  int nsh_ao, nsh_aux;
  long npair;   // nao (nao + 1) / 2
  double* out;  // 3c: [naux][npair];  2c: [naux][naux]
  int naux;
};

__host__ __device__ constexpr int ncart(int l) { return (l + 1) * (l + 2) / 2; }
__host__ __device__ constexpr int tri3(int L) { return (L + 1) * (L + 2) * (L + 3) / 6; }  // #(t,u,v): t+u+v <= L
__device__ __forceinline__ int c2s_off(int l) {  // sum_{l' < l} (2 l' + 1) ncart(l')
  const int o[5] = {0, 1, 10, 40, 110};
  return o[l];
}
// compact index of (t, u, v), t + u + v <= L
__device__ __forceinline__ int tuv_idx(int L, int t, int u, int v) {
  // entries with first index < t: sum_{s < t} (L - s + 1)(L - s + 2) / 2
  const int M = L - t;
  const int before_t = tri3(L) - tri3(M);
  return before_t + u * (M + 1) - u * (u - 1) / 2 + v;
}
// Cartesian component c of shell l -> (lx, ly, lz), libcint order
__device__ __forceinline__ void cart_lmn(int l, int c, int& lx, int& ly, int& lz) {
  int k = 0;
  for (lx = l; lx >= 0; --lx)
    for (ly = l - lx; ly >= 0; --ly, ++k)
      if (k == c) {
        lz = l - lx - ly;
        return;
      }
}

// Boys function F_0 .. F_mmax at T, relative accuracy ~1e-15: ascending series for F_mmax + downward recursion
// (T < 35), erf-based F_0 + upward recursion otherwise (stable there: (2m + 1) / 2T < 1).
__device__ inline void boys(int mmax, double T, double* F) {
  if (T < 35.0) {
    const double et = exp(-T);
    double term = 1.0 / (2.0 * mmax + 1.0), sum = term;
    for (int k = 1; k < 400; ++k) {
      term *= 2.0 * T / (2.0 * mmax + 2.0 * k + 1.0);
      sum += term;
      if (term < 1e-17 * sum) break;
    }
    F[mmax] = et * sum;
    for (int m = mmax; m > 0; --m) F[m - 1] = (2.0 * T * F[m] + et) / (2.0 * m - 1.0);
  } else {
    const double et = exp(-T);
    F[0] = 0.5 * sqrt(3.14159265358979323846 / T) * erf(sqrt(T));
    for (int m = 0; m < mmax; ++m) F[m + 1] = ((2.0 * m + 1.0) * F[m] - et) / (2.0 * T);
  }
}

// E[i][j][t], i <= la, j <= lb, t <= la + lb, one Cartesian direction; leading dimensions (lb + 1), (la + lb + 1)
__device__ inline void hermite_e(int la, int lb, double a, double b, double ab, double* E) {
  const double p = a + b, mu = a * b / p;
  const double xpa = -b / p * ab, xpb = a / p * ab, h = 0.5 / p;
  const int nt = la + lb + 1, nj = lb + 1;
  for (int e = 0; e < (la + 1) * nj * nt; ++e) E[e] = 0.0;
  E[0] = exp(-mu * ab * ab);
  for (int i = 0; i <= la; ++i)
    for (int j = 0; j <= lb; ++j) {
      if (i == 0 && j == 0) continue;
      double* cur = E + (i * nj + j) * nt;
      const double* prev = j == 0 ? E + ((i - 1) * nj + j) * nt : E + (i * nj + j - 1) * nt;
      const double xq = j == 0 ? xpa : xpb;
      for (int t = 0; t <= i + j; ++t) {
        double v = xq * prev[t];
        if (t + 1 < nt) v += (t + 1) * prev[t + 1];
        if (t > 0) v += h * prev[t - 1];
        cur[t] = v;
      }
    }
}

// R[tuv_idx(L, t, u, v)] = R^0_{tuv}(alpha, pq); scratch `tmp` of the same size (tri3(L))
__device__ inline void hermite_r(int L, double alpha, double px, double py, double pz, double* R, double* tmp) {
  double F[INT_LMAX_AO * 2 + INT_LMAX_AUX + 1];
  boys(L, alpha * (px * px + py * py + pz * pz), F);
  // n runs from L down to 0; after the step for n, `cur` holds R^n_{tuv} for t + u + v <= L - n
  double* cur = (L & 1) ? tmp : R;  // L + 1 steps: the last one (n = 0) must land in R
  double* prev = (L & 1) ? R : tmp;
  double m2a = 1.0;
  for (int n = 0; n < L; ++n) m2a *= -2.0 * alpha;
  for (int n = L; n >= 0; --n) {
    const int top = L - n;
    cur[0] = m2a * F[n];
    for (int t = 0; t <= top; ++t)
      for (int u = 0; u <= top - t; ++u)
        for (int v = 0; v <= top - t - u; ++v) {
          if (t + u + v == 0) continue;
          double val;
          if (t > 0) {
            val = px * prev[tuv_idx(L, t - 1, u, v)];
            if (t > 1) val += (t - 1) * prev[tuv_idx(L, t - 2, u, v)];
          } else if (u > 0) {
            val = py * prev[tuv_idx(L, t, u - 1, v)];
            if (u > 1) val += (u - 1) * prev[tuv_idx(L, t, u - 2, v)];
          } else {
            val = pz * prev[tuv_idx(L, t, u, v - 1)];
            if (v > 1) val += (v - 1) * prev[tuv_idx(L, t, u, v - 2)];
          }
          cur[tuv_idx(L, t, u, v)] = val;
        }
    double* s = cur;
    cur = prev;
    prev = s;
    if (n > 0) m2a /= -2.0 * alpha;
  }
}

// ---- three-centre integrals ------------------------------------------------------------------------
// shared-memory record of one primitive pair: [p, Px, Py, Pz, coef] + E^x, E^y, E^z
template <int LAB_MAX, int LC_MAX>
__global__ void __launch_bounds__(INT_THREADS) int3c2e_kernel(IntArgs a, const int* __restrict__ pair_list, int pp_cap) {
  extern __shared__ __align__(16) double ism[];
  constexpr int LTOT = LAB_MAX + LC_MAX;
  const int pi = pair_list[2 * blockIdx.x], pj = pair_list[2 * blockIdx.x + 1];
  const IntShell A = a.ao[pi], B = a.ao[pj];
  const int la = A.l, lb = B.l, lab = la + lb;
  const int esz = (la + 1) * (lb + 1) * (lab + 1);
  const int rec = 5 + 3 * esz;
  const int npp = A.nprim * B.nprim;
  const int nca = ncart(la), ncb = ncart(lb);
  const int nsa = 2 * la + 1, nsb = 2 * lb + 1;
  const double abx = A.x - B.x, aby = A.y - B.y, abz = A.z - B.z;

  // per-thread state (local memory): Cartesian block accumulator, R tables, aux Hermite coefficients, G
  double cart[ncart(LAB_MAX > INT_LMAX_AO ? INT_LMAX_AO : LAB_MAX) * ncart(LAB_MAX > INT_LMAX_AO ? INT_LMAX_AO : LAB_MAX) * ncart(LC_MAX)];
  double R[tri3(LTOT)], Rtmp[tri3(LTOT)];
  double Ec[(LC_MAX + 1) * (LC_MAX + 1)];
  double G[tri3(LAB_MAX)];

  for (int k0 = 0; k0 < a.nsh_aux; k0 += INT_THREADS) {
    const int k = k0 + threadIdx.x;
    const bool active = k < a.nsh_aux;
    IntShell C = a.aux[active ? k : 0];
    const int lc = C.l, ncc = ncart(lc), L = lab + lc;
    if (active)
      for (int e = 0; e < nca * ncb * ncc; ++e) cart[e] = 0.0;
    // primitive pairs in batches that fit the shared-memory budget
    for (int pp0 = 0; pp0 < npp; pp0 += pp_cap) {
      const int nb = min(pp_cap, npp - pp0);
      __syncthreads();
      for (int q = threadIdx.x; q < nb; q += INT_THREADS) {
        const int ia = (pp0 + q) / B.nprim, ib = (pp0 + q) % B.nprim;
        const double ea = a.env[A.ptr_exp + ia], eb = a.env[B.ptr_exp + ib];
        const double p = ea + eb;
        double* r = ism + (size_t)q * rec;
        r[0] = p;
        r[1] = (ea * A.x + eb * B.x) / p;
        r[2] = (ea * A.y + eb * B.y) / p;
        r[3] = (ea * A.z + eb * B.z) / p;
        r[4] = a.env[A.ptr_coef + ia] * a.env[B.ptr_coef + ib];
        hermite_e(la, lb, ea, eb, abx, r + 5);
        hermite_e(la, lb, ea, eb, aby, r + 5 + esz);
        hermite_e(la, lb, ea, eb, abz, r + 5 + 2 * esz);
      }
      __syncthreads();
      if (!active) continue;
      for (int ic = 0; ic < C.nprim; ++ic) {
        const double g = a.env[C.ptr_exp + ic], wc = a.env[C.ptr_coef + ic];
        // single-Gaussian Hermite coefficients E^{i}_t (b -> 0 limit), identical for x, y, z
        hermite_e(lc, 0, g, 0.0, 0.0, Ec);
        for (int q = 0; q < nb; ++q) {
          const double* r = ism + (size_t)q * rec;
          const double p = r[0];
          const double alpha = p * g / (p + g);
          hermite_r(L, alpha, r[1] - C.x, r[2] - C.y, r[3] - C.z, R, Rtmp);
          const double pref = r[4] * wc * 34.986836655249725 / (p * g * sqrt(p + g));  // 2 pi^(5/2)
          const double *Ex = r + 5, *Ey = Ex + esz, *Ez = Ey + esz;
          for (int cc = 0; cc < ncc; ++cc) {
            int cx, cy, cz;
            cart_lmn(lc, cc, cx, cy, cz);
            // G[tuv] = sum_{t'u'v'} (-1)^(t'+u'+v') Ec[cx][t'] Ec[cy][u'] Ec[cz][v'] R[t+t', u+u', v+v']
            for (int t = 0; t <= lab; ++t)
              for (int u = 0; u <= lab - t; ++u)
                for (int v = 0; v <= lab - t - u; ++v) {
                  double s = 0.0;
                  for (int tp = 0; tp <= cx; ++tp) {
                    const double ex = Ec[cx * (lc + 1) + tp];
                    if (ex == 0.0) continue;
                    for (int up = 0; up <= cy; ++up) {
                      const double exy = ex * Ec[cy * (lc + 1) + up];
                      if (exy == 0.0) continue;
                      for (int vp = 0; vp <= cz; ++vp) {
                        const double e3 = exy * Ec[cz * (lc + 1) + vp];
                        if (e3 == 0.0) continue;
                        const double rv = R[tuv_idx(L, t + tp, u + up, v + vp)];
                        s += ((tp + up + vp) & 1) ? -e3 * rv : e3 * rv;
                      }
                    }
                  }
                  G[tuv_idx(lab, t, u, v)] = s;
                }
            for (int ca = 0; ca < nca; ++ca) {
              int ax, ay, az;
              cart_lmn(la, ca, ax, ay, az);
              for (int cb = 0; cb < ncb; ++cb) {
                int bx, by, bz;
                cart_lmn(lb, cb, bx, by, bz);
                const double* ex = Ex + (ax * (lb + 1) + bx) * (lab + 1);
                const double* ey = Ey + (ay * (lb + 1) + by) * (lab + 1);
                const double* ez = Ez + (az * (lb + 1) + bz) * (lab + 1);
                double s = 0.0;
                for (int t = 0; t <= ax + bx; ++t)
                  for (int u = 0; u <= ay + by; ++u) {
                    const double etu = ex[t] * ey[u];
                    for (int v = 0; v <= az + bz; ++v) s += etu * ez[v] * G[tuv_idx(lab, t, u, v)];
                  }
                cart[(ca * ncb + cb) * ncc + cc] += pref * s;
              }
            }
          }
        }
      }
    }
    if (!active) continue;
    // Cartesian -> spherical on all three indices, then the packed-lower rows of the result
    const double* Ta = a.c2s + c2s_off(la);
    const double* Tb = a.c2s + c2s_off(lb);
    const double* Tc = a.c2s + c2s_off(lc);
    const int nsc = 2 * lc + 1;
    for (int mc = 0; mc < nsc; ++mc)
      for (int ma = 0; ma < nsa; ++ma)
        for (int mb = 0; mb < nsb; ++mb) {
          const int mu = A.ao_off + ma, nu = B.ao_off + mb;
          if (pi == pj && nu > mu) continue;
          double s = 0.0;
          for (int ca = 0; ca < nca; ++ca) {
            const double ta = Ta[ma * nca + ca];
            if (ta == 0.0) continue;
            for (int cb = 0; cb < ncb; ++cb) {
              const double tab = ta * Tb[mb * ncb + cb];
              if (tab == 0.0) continue;
              for (int cc = 0; cc < ncc; ++cc) s += tab * Tc[mc * ncc + cc] * cart[(ca * ncb + cb) * ncc + cc];
            }
          }
          const long hi = mu > nu ? mu : nu, lo = mu > nu ? nu : mu;
          a.out[(long)(C.ao_off + mc) * a.npair + hi * (hi + 1) / 2 + lo] = s;
        }
  }
}

// ---- two-centre metric (P|Q): one thread per auxiliary shell pair (i >= k) -----------------------------
__global__ void __launch_bounds__(INT_THREADS) int2c2e_kernel(IntArgs a) {
  constexpr int LM = INT_LMAX_AUX;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long npairs = (long)a.nsh_aux * (a.nsh_aux + 1) / 2;
  if (idx >= npairs) return;
  int i = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
  while ((long)i * (i + 1) / 2 > idx) --i;
  while ((long)(i + 1) * (i + 2) / 2 <= idx) ++i;
  const int k = (int)(idx - (long)i * (i + 1) / 2);
  const IntShell A = a.aux[i], C = a.aux[k];
  const int la = A.l, lc = C.l, L = la + lc;
  const int nca = ncart(la), ncc = ncart(lc);
  double cart[ncart(LM) * ncart(LM)];
  double R[tri3(2 * LM)], Rtmp[tri3(2 * LM)];
  double Ea[(LM + 1) * (LM + 1)], Ec[(LM + 1) * (LM + 1)];
  for (int e = 0; e < nca * ncc; ++e) cart[e] = 0.0;
  for (int ia = 0; ia < A.nprim; ++ia) {
    const double p = a.env[A.ptr_exp + ia], wa = a.env[A.ptr_coef + ia];
    hermite_e(la, 0, p, 0.0, 0.0, Ea);
    for (int ic = 0; ic < C.nprim; ++ic) {
      const double g = a.env[C.ptr_exp + ic], wc = a.env[C.ptr_coef + ic];
      hermite_e(lc, 0, g, 0.0, 0.0, Ec);
      hermite_r(L, p * g / (p + g), A.x - C.x, A.y - C.y, A.z - C.z, R, Rtmp);
      const double pref = wa * wc * 34.986836655249725 / (p * g * sqrt(p + g));
      for (int ca = 0; ca < nca; ++ca) {
        int ax, ay, az;
        cart_lmn(la, ca, ax, ay, az);
        for (int cc = 0; cc < ncc; ++cc) {
          int cx, cy, cz;
          cart_lmn(lc, cc, cx, cy, cz);
          double s = 0.0;
          for (int t = 0; t <= ax; ++t)
            for (int u = 0; u <= ay; ++u)
              for (int v = 0; v <= az; ++v) {
                const double e1 = Ea[ax * (la + 1) + t] * Ea[ay * (la + 1) + u] * Ea[az * (la + 1) + v];
                if (e1 == 0.0) continue;
                for (int tp = 0; tp <= cx; ++tp)
                  for (int up = 0; up <= cy; ++up)
                    for (int vp = 0; vp <= cz; ++vp) {
                      const double e2 = Ec[cx * (lc + 1) + tp] * Ec[cy * (lc + 1) + up] * Ec[cz * (lc + 1) + vp];
                      if (e2 == 0.0) continue;
                      const double rv = R[tuv_idx(L, t + tp, u + up, v + vp)];
                      s += ((tp + up + vp) & 1) ? -e1 * e2 * rv : e1 * e2 * rv;
                    }
              }
          cart[ca * ncc + cc] += pref * s;
        }
      }
    }
  }
  const double* Ta = a.c2s + c2s_off(la);
  const double* Tc = a.c2s + c2s_off(lc);
  for (int ma = 0; ma < 2 * la + 1; ++ma)
    for (int mc = 0; mc < 2 * lc + 1; ++mc) {
      double s = 0.0;
      for (int ca = 0; ca < nca; ++ca)
        for (int cc = 0; cc < ncc; ++cc) s += Ta[ma * nca + ca] * Tc[mc * ncc + cc] * cart[ca * ncc + cc];
      const long r = A.ao_off + ma, c = C.ao_off + mc;
      a.out[r * a.naux + c] = s;
      a.out[c * a.naux + r] = s;
    }
}

// zero the strictly-upper triangle (row-major) of an n x n matrix
__global__ void zero_upper_kernel(double* __restrict__ A, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j < n && j > i) A[(long)i * n + j] = 0.0;
}

}  // namespace nbd
