// Small dense host linear algebra used by the SCF drivers (pure C++, no CUDA): the <= 9 x 9 DIIS equations and the
// 16 / 32-dimensional Rayleigh-Ritz step of the subspace eigensolver.  Unit-tested on the CPU by tests/cpp/linalg_test.cpp.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <vector>

// ------------------------------------------------------------------------------------------------
// Small host linear algebra for the DIIS equations (<= 9 x 9): restates what pyscf/lib/diis.py asks of
// scipy.linalg.eigh / numpy.linalg.solve (reference call site nbed/scf/huzinaga_scf.py:164).
// ------------------------------------------------------------------------------------------------
inline void jacobi_eigh(int n, std::vector<double> a, std::vector<double>& w, std::vector<double>& v) {
  v.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) v[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int p = 0; p < n; ++p) {
      diag += a[(size_t)p * n + p] * a[(size_t)p * n + p];
      for (int q = p + 1; q < n; ++q) off += a[(size_t)p * n + q] * a[(size_t)p * n + q];
    }
    if (off <= 1e-34 * (diag + 2.0 * off) || off < 1e-300) break;  // off-diagonal norm below 1e-17 of the matrix norm
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = a[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double app = a[(size_t)p * n + p], aqq = a[(size_t)q * n + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
          a[(size_t)k * n + p] = c * akp - s * akq;
          a[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
          a[(size_t)p * n + k] = c * apk - s * aqk;
          a[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = v[(size_t)k * n + p], vkq = v[(size_t)k * n + q];
          v[(size_t)k * n + p] = c * vkp - s * vkq;
          v[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  w.resize(n);
  for (int i = 0; i < n; ++i) w[i] = a[(size_t)i * n + i];
}

inline bool lu_solve(int n, std::vector<double> a, std::vector<double> b, std::vector<double>& x) {
  for (int k = 0; k < n; ++k) {
    int piv = k;
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(a[(size_t)i * n + k]) > std::fabs(a[(size_t)piv * n + k])) piv = i;
    if (a[(size_t)piv * n + k] == 0.0) return false;
    if (piv != k) {
      for (int j = 0; j < n; ++j) std::swap(a[(size_t)k * n + j], a[(size_t)piv * n + j]);
      std::swap(b[k], b[piv]);
    }
    for (int i = k + 1; i < n; ++i) {
      const double f = a[(size_t)i * n + k] / a[(size_t)k * n + k];
      if (f == 0.0) continue;
      for (int j = k; j < n; ++j) a[(size_t)i * n + j] -= f * a[(size_t)k * n + j];
      b[i] -= f * b[k];
    }
  }
  x.assign(n, 0.0);
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < n; ++j) s -= a[(size_t)i * n + j] * x[j];
    x[i] = s / a[(size_t)i * n + i];
  }
  return true;
}

// c = solution of H c = (1,0,...,0) with the pseudo-inverse fallback of pyscf/lib/diis.py:extrapolate
inline std::vector<double> diis_coefficients(const std::vector<double>& Hfull, int ldh, int nd) {
  const int m = nd + 1;
  std::vector<double> h((size_t)m * m), g(m, 0.0), w, v, c;
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) h[(size_t)i * m + j] = Hfull[(size_t)i * ldh + j];
  g[0] = 1.0;
  jacobi_eigh(m, h, w, v);
  bool singular = false;
  for (int i = 0; i < m; ++i)
    if (std::fabs(w[i]) < 1e-14) singular = true;
  if (!singular && lu_solve(m, h, g, c)) return c;
  c.assign(m, 0.0);
  for (int k = 0; k < m; ++k) {
    if (std::fabs(w[k]) <= 1e-14) continue;
    double proj = 0.0;
    for (int i = 0; i < m; ++i) proj += v[(size_t)i * m + k] * g[i];
    for (int i = 0; i < m; ++i) c[i] += v[(size_t)i * m + k] * proj / w[k];
  }
  return c;
}

// Symmetric eigensolver for the 16 / 32-dimensional Rayleigh-Ritz matrices: Householder reduction to tridiagonal
// form followed by the implicit-shift QL iteration with accumulated transformations (about 20x fewer operations than
// cyclic Jacobi at n = 32, where the host step used to cost more than the device work of a whole filter pass).
// a: symmetric n x n (row-major, both triangles); w: eigenvalues (unsorted); v: eigenvectors in columns.
// Returns false if an eigenvalue needed more than 60 QL sweeps (the caller then falls back to jacobi_eigh).
inline bool householder_ql_eigh(int n, std::vector<double> a, std::vector<double>& w, std::vector<double>& v) {
  // the accumulated transformation is kept transposed (row i = vector i) so that both kinds of update below run
  // over contiguous memory
  std::vector<double> vt((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) vt[(size_t)i * n + i] = 1.0;
  std::vector<double> d(n, 0.0), e(n, 0.0), u(n), p(n), tr(n);
  // Householder reflections H_k = I - beta u u^T on the trailing indices k+1.. zero column k below the subdiagonal
  for (int k = 0; k + 2 < n; ++k) {
    double scale = 0.0;
    for (int i = k + 1; i < n; ++i) scale = std::max(scale, std::fabs(a[(size_t)i * n + k]));
    double below = 0.0;
    for (int i = k + 2; i < n; ++i) below = std::max(below, std::fabs(a[(size_t)i * n + k]));
    if (below == 0.0) continue;  // already tridiagonal in this column
    double nrm = 0.0;
    for (int i = k + 1; i < n; ++i) {
      u[i] = a[(size_t)i * n + k] / scale;
      nrm += u[i] * u[i];
    }
    nrm = std::sqrt(nrm);
    const double alpha = u[k + 1] >= 0.0 ? -nrm : nrm;
    u[k + 1] -= alpha;
    double uu = 0.0;
    for (int i = k + 1; i < n; ++i) uu += u[i] * u[i];
    const double beta = 2.0 / uu;
    // p = beta * A u ; K = beta/2 * u.p ; q = p - K u ; A -= u q^T + q u^T   (trailing block)
    double up = 0.0;
    for (int i = k + 1; i < n; ++i) {
      double t = 0.0;
      for (int j = k + 1; j < n; ++j) t += a[(size_t)i * n + j] * u[j];
      p[i] = beta * t;
      up += u[i] * p[i];
    }
    const double K = 0.5 * beta * up;
    for (int i = k + 1; i < n; ++i) p[i] -= K * u[i];
    for (int i = k + 1; i < n; ++i)
      for (int j = k + 1; j < n; ++j) a[(size_t)i * n + j] -= u[i] * p[j] + p[i] * u[j];
    a[(size_t)(k + 1) * n + k] = a[(size_t)k * n + k + 1] = alpha * scale;
    for (int i = k + 2; i < n; ++i) a[(size_t)i * n + k] = a[(size_t)k * n + i] = 0.0;
    // V <- V H_k, i.e. V^T <- H_k V^T
    for (int r = 0; r < n; ++r) tr[r] = 0.0;
    for (int j = k + 1; j < n; ++j)
      for (int r = 0; r < n; ++r) tr[r] += vt[(size_t)j * n + r] * u[j];
    for (int j = k + 1; j < n; ++j) {
      const double bu = beta * u[j];
      for (int r = 0; r < n; ++r) vt[(size_t)j * n + r] -= bu * tr[r];
    }
  }
  for (int i = 0; i < n; ++i) {
    d[i] = a[(size_t)i * n + i];
    e[i] = i + 1 < n ? a[(size_t)(i + 1) * n + i] : 0.0;  // e[i] couples i and i + 1
  }
  // implicit QL with Wilkinson-type shift
  for (int l = 0; l < n; ++l) {
    for (int iter = 0;; ++iter) {
      int m = l;
      for (; m + 1 < n; ++m) {
        const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
      }
      if (m == l) break;
      if (iter == 60) return false;
      double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
      double r = std::sqrt(g * g + 1.0);
      g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? r : -r));
      double sn = 1.0, cs = 1.0, pp = 0.0;
      int i = m - 1;
      for (; i >= l; --i) {
        double f = sn * e[i];
        const double b = cs * e[i];
        r = std::sqrt(f * f + g * g);  // (the Rayleigh-Ritz matrices are scaled to unit diagonal Gram: no overflow)
        e[i + 1] = r;
        if (r == 0.0) {
          d[i + 1] -= pp;
          e[m] = 0.0;
          break;
        }
        sn = f / r;
        cs = g / r;
        g = d[i + 1] - pp;
        r = (d[i] - g) * sn + 2.0 * cs * b;
        pp = sn * r;
        d[i + 1] = g + pp;
        g = cs * r - b;
        double* v0 = vt.data() + (size_t)i * n;
        double* v1 = v0 + n;
        for (int k = 0; k < n; ++k) {
          const double x0 = v0[k], x1 = v1[k];
          v1[k] = sn * x0 + cs * x1;
          v0[k] = cs * x0 - sn * x1;
        }
      }
      if (r == 0.0 && i >= l) continue;
      d[l] -= pp;
      e[l] = g;
      e[m] = 0.0;
    }
  }
  w = d;
  v.resize((size_t)n * n);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < n; ++k) v[(size_t)k * n + i] = vt[(size_t)i * n + k];
  return true;
}

// host: orthonormalise + Rayleigh-Ritz.  G = Y^T Y, H = Y^T A Y  ->  M (Y M orthonormal Ritz vectors), theta ascending
inline bool sub_rayleigh_ritz(int kb, const double* G, const double* H, double* M, double* theta) {
  std::vector<double> d(kb), gs((size_t)kb * kb), hs((size_t)kb * kb), L((size_t)kb * kb, 0.0), Li((size_t)kb * kb, 0.0);
  for (int i = 0; i < kb; ++i) {
    if (!(G[(size_t)i * kb + i] > 0.0)) return false;
    d[i] = 1.0 / std::sqrt(G[(size_t)i * kb + i]);
  }
  for (int i = 0; i < kb; ++i)
    for (int j = 0; j < kb; ++j) {
      gs[(size_t)i * kb + j] = 0.5 * (G[(size_t)i * kb + j] + G[(size_t)j * kb + i]) * d[i] * d[j];
      hs[(size_t)i * kb + j] = 0.5 * (H[(size_t)i * kb + j] + H[(size_t)j * kb + i]) * d[i] * d[j];
    }
  for (int j = 0; j < kb; ++j) {  // Cholesky gs = L L^T
    double s = gs[(size_t)j * kb + j];
    for (int k = 0; k < j; ++k) s -= L[(size_t)j * kb + k] * L[(size_t)j * kb + k];
    if (!(s > 1e-12)) return false;  // (columns are normalised: a tiny pivot means a numerically dependent block)
    L[(size_t)j * kb + j] = std::sqrt(s);
    for (int i = j + 1; i < kb; ++i) {
      double t = gs[(size_t)i * kb + j];
      for (int k = 0; k < j; ++k) t -= L[(size_t)i * kb + k] * L[(size_t)j * kb + k];
      L[(size_t)i * kb + j] = t / L[(size_t)j * kb + j];
    }
  }
  for (int j = 0; j < kb; ++j) {  // Li = L^-1 (lower)
    Li[(size_t)j * kb + j] = 1.0 / L[(size_t)j * kb + j];
    for (int i = j + 1; i < kb; ++i) {
      double t = 0.0;
      for (int k = j; k < i; ++k) t -= L[(size_t)i * kb + k] * Li[(size_t)k * kb + j];
      Li[(size_t)i * kb + j] = t / L[(size_t)i * kb + i];
    }
  }
  std::vector<double> t1((size_t)kb * kb, 0.0), ht((size_t)kb * kb, 0.0), w, q;
  for (int i = 0; i < kb; ++i)  // t1 = Li hs
    for (int j = 0; j < kb; ++j) {
      double t = 0.0;
      for (int k = 0; k <= i; ++k) t += Li[(size_t)i * kb + k] * hs[(size_t)k * kb + j];
      t1[(size_t)i * kb + j] = t;
    }
  for (int i = 0; i < kb; ++i)  // ht = t1 Li^T
    for (int j = 0; j < kb; ++j) {
      double t = 0.0;
      for (int k = 0; k <= j; ++k) t += t1[(size_t)i * kb + k] * Li[(size_t)j * kb + k];
      ht[(size_t)i * kb + j] = t;
    }
  for (int i = 0; i < kb; ++i)
    for (int j = 0; j < i; ++j) ht[(size_t)i * kb + j] = ht[(size_t)j * kb + i] = 0.5 * (ht[(size_t)i * kb + j] + ht[(size_t)j * kb + i]);
  if (!householder_ql_eigh(kb, ht, w, q)) jacobi_eigh(kb, ht, w, q);  // columns of q = eigenvectors
  std::vector<int> order(kb);
  for (int i = 0; i < kb; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return w[a] < w[b]; });
  for (int c = 0; c < kb; ++c) {
    theta[c] = w[order[c]];
    for (int i = 0; i < kb; ++i) {  // M[i][c] = d_i * sum_k Li[k][i] q[k][order c]
      double t = 0.0;
      for (int k = i; k < kb; ++k) t += Li[(size_t)k * kb + i] * q[(size_t)k * kb + order[c]];
      M[(size_t)i * kb + c] = d[i] * t;
    }
  }
  return true;
}


// Extreme Ritz values of a k-step Lanczos run: alpha[0..k) diagonal, beta[1..k) off-diagonal of the tridiagonal
// matrix, beta[k] the norm of the last residual.  theta_max + beta[k] bounds the spectrum from above (Zhou & Li,
// "Bounding the spectrum of large Hermitian matrices"); theta_min - beta[k] is the matching estimate from below.
inline void lanczos_ritz_bounds(int k, const double* alpha, const double* beta, double* low, double* up) {
  std::vector<double> t((size_t)k * k, 0.0), w, v;
  for (int i = 0; i < k; ++i) {
    t[(size_t)i * k + i] = alpha[i];
    if (i + 1 < k) t[(size_t)i * k + i + 1] = t[(size_t)(i + 1) * k + i] = beta[i + 1];
  }
  if (!householder_ql_eigh(k, t, w, v)) jacobi_eigh(k, t, w, v);
  const double lo = *std::min_element(w.begin(), w.end()), hi = *std::max_element(w.begin(), w.end());
  *low = lo - beta[k];
  *up = hi + beta[k];
}

// Degree of the Chebyshev filter on [a, bu] such that the lowest eigenvalue (estimate lmin <= a) is amplified by at
// most `cap` relative to the interval: keeps the Gram matrix of a filtered block that is still far from the invariant
// subspace well inside double precision.  Converged blocks (a - lmin << bu - a) always get max_degree.
inline int chebyshev_degree(double a, double bu, double lmin, double cap, int max_degree) {
  const double x0 = 1.0 + 2.0 * std::max(0.0, a - lmin) / std::max(1e-300, bu - a);
  if (x0 <= 1.0 + 1e-14) return max_degree;
  const double m = std::acosh(cap) / std::acosh(x0);
  return (int)std::max(2.0, std::min((double)max_degree, std::floor(m)));
}
