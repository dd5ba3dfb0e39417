// CPU unit test of nbed_b200/csrc/host_linalg.h (g++ -std=c++17).
#include <cstdio>
#include <random>
#include "../../nbed_b200/csrc/host_linalg.h"

static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { ++fails; printf("FAIL %s:%d ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } while (0)

int main() {
  std::mt19937 rng(7);
  std::normal_distribution<double> N(0.0, 1.0);
  // Jacobi eigensolver: A v = w v, V orthonormal, for several sizes (incl. the 16 / 32 of the Rayleigh-Ritz step)
  for (int n : {1, 2, 7, 9, 16, 32}) {
    std::vector<double> a((size_t)n * n), w, v;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j <= i; ++j) a[(size_t)i * n + j] = a[(size_t)j * n + i] = N(rng) * (i == j ? 5.0 : 1.0);
    jacobi_eigh(n, a, w, v);
    double res = 0, orth = 0;
    for (int k = 0; k < n; ++k) {
      for (int i = 0; i < n; ++i) {
        double s = 0;
        for (int j = 0; j < n; ++j) s += a[(size_t)i * n + j] * v[(size_t)j * n + k];
        res = std::max(res, std::fabs(s - w[k] * v[(size_t)i * n + k]));
      }
      for (int l = 0; l < n; ++l) {
        double s = 0;
        for (int i = 0; i < n; ++i) s += v[(size_t)i * n + k] * v[(size_t)i * n + l];
        orth = std::max(orth, std::fabs(s - (k == l ? 1.0 : 0.0)));
      }
    }
    CHECK(res < 1e-12 && orth < 1e-13, "jacobi n=%d residual %.2e orth %.2e", n, res, orth);
  }
  // Householder + QL eigensolver: same checks, plus agreement with Jacobi, degenerate and already-diagonal inputs
  for (int n : {1, 2, 3, 7, 10, 16, 32}) {
    for (int kind = 0; kind < 3; ++kind) {
      std::vector<double> a((size_t)n * n, 0.0), w, v, wj, vj;
      if (kind == 0) {
        for (int i = 0; i < n; ++i)
          for (int j = 0; j <= i; ++j) a[(size_t)i * n + j] = a[(size_t)j * n + i] = N(rng) * (i == j ? 5.0 : 1.0);
      } else if (kind == 1) {  // diagonal with repeated entries
        for (int i = 0; i < n; ++i) a[(size_t)i * n + i] = (double)(i / 2);
      } else {  // rank-2 plus identity: (n - 2)-fold degenerate eigenvalue
        std::vector<double> x(n), y(n);
        for (int i = 0; i < n; ++i) { x[i] = N(rng); y[i] = N(rng); }
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j) a[(size_t)i * n + j] = (i == j ? 1.0 : 0.0) + x[i] * x[j] - 0.5 * y[i] * y[j];
      }
      CHECK(householder_ql_eigh(n, a, w, v), "QL did not converge n=%d kind=%d", n, kind);
      jacobi_eigh(n, a, wj, vj);
      double res = 0, orth = 0, scale = 1.0;
      for (int i = 0; i < n * n; ++i) scale = std::max(scale, std::fabs(a[i]));
      for (int k = 0; k < n; ++k) {
        for (int i = 0; i < n; ++i) {
          double s = 0;
          for (int j = 0; j < n; ++j) s += a[(size_t)i * n + j] * v[(size_t)j * n + k];
          res = std::max(res, std::fabs(s - w[k] * v[(size_t)i * n + k]));
        }
        for (int l = 0; l < n; ++l) {
          double s = 0;
          for (int i = 0; i < n; ++i) s += v[(size_t)i * n + k] * v[(size_t)i * n + l];
          orth = std::max(orth, std::fabs(s - (k == l ? 1.0 : 0.0)));
        }
      }
      std::sort(w.begin(), w.end());
      std::sort(wj.begin(), wj.end());
      double dw = 0;
      for (int k = 0; k < n; ++k) dw = std::max(dw, std::fabs(w[k] - wj[k]));
      CHECK(res < 2e-13 * scale * n && orth < 1e-13 * n && dw < 1e-12 * scale, "QL n=%d kind=%d residual %.2e orth %.2e vs jacobi %.2e", n, kind, res, orth, dw);
    }
  }
  // LU solve against a manufactured solution, with a zero leading pivot (the bordered DIIS matrix has H[0][0] = 0)
  {
    const int n = 5;
    std::vector<double> a((size_t)n * n), x0(n), b(n, 0.0), x;
    for (auto& e : a) e = N(rng);
    a[0] = 0.0;
    for (int i = 0; i < n; ++i) x0[i] = N(rng);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) b[i] += a[(size_t)i * n + j] * x0[j];
    CHECK(lu_solve(n, a, b, x), "lu_solve reported a singular matrix");
    double err = 0;
    for (int i = 0; i < n; ++i) err = std::max(err, std::fabs(x[i] - x0[i]));
    CHECK(err < 1e-11, "lu_solve error %.2e", err);
  }
  // DIIS coefficients (pyscf/lib/diis.py:extrapolate): H c = (1, 0, ...); two orthogonal unit error vectors -> 1/2, 1/2;
  // and the pseudo-inverse branch with an exactly duplicated error vector
  {
    const int ld = 7;
    std::vector<double> H((size_t)ld * ld, 0.0);
    for (int i = 1; i < ld; ++i) H[i] = H[(size_t)i * ld] = 1.0;
    H[1 * ld + 1] = 1.0; H[2 * ld + 2] = 1.0;
    auto c = diis_coefficients(H, ld, 2);
    CHECK(std::fabs(c[1] - 0.5) < 1e-14 && std::fabs(c[2] - 0.5) < 1e-14 && std::fabs(c[0] + 0.5) < 1e-14, "diis 2-vector case %.3g %.3g %.3g", c[0], c[1], c[2]);
    H[1 * ld + 2] = H[2 * ld + 1] = 1.0;  // e1 == e2: singular, pseudo-inverse keeps the sum rule
    c = diis_coefficients(H, ld, 2);
    CHECK(std::fabs(c[1] + c[2] - 1.0) < 1e-12 && std::fabs(c[1] - c[2]) < 1e-12, "diis singular case %.3g %.3g", c[1], c[2]);
  }
  // Rayleigh-Ritz: for a badly scaled, non-orthogonal block Y (Gram G, projected H), M must give
  // M^T G M = I and M^T H M = diag(theta) with theta ascending
  for (int kb : {16, 32}) {
    const int n = 80;
    std::vector<double> Y((size_t)n * kb), A((size_t)n * n), G((size_t)kb * kb, 0.0), Hm((size_t)kb * kb, 0.0), M((size_t)kb * kb), th(kb);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j <= i; ++j) A[(size_t)i * n + j] = A[(size_t)j * n + i] = N(rng);
    for (int i = 0; i < n; ++i)
      for (int c = 0; c < kb; ++c) Y[(size_t)i * kb + c] = N(rng) * std::pow(10.0, -0.4 * c);  // columns spanning 6-12 decades
    std::vector<double> AY((size_t)n * kb, 0.0);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        for (int c = 0; c < kb; ++c) AY[(size_t)i * kb + c] += A[(size_t)i * n + j] * Y[(size_t)j * kb + c];
    for (int c = 0; c < kb; ++c)
      for (int d = 0; d < kb; ++d)
        for (int i = 0; i < n; ++i) {
          G[(size_t)c * kb + d] += Y[(size_t)i * kb + c] * Y[(size_t)i * kb + d];
          Hm[(size_t)c * kb + d] += Y[(size_t)i * kb + c] * AY[(size_t)i * kb + d];
        }
    CHECK(sub_rayleigh_ritz(kb, G.data(), Hm.data(), M.data(), th.data()), "rayleigh-ritz failed kb=%d", kb);
    double e1 = 0, e2 = 0;
    for (int c = 0; c < kb; ++c)
      for (int d = 0; d < kb; ++d) {
        double g = 0, h = 0;
        for (int p = 0; p < kb; ++p)
          for (int q = 0; q < kb; ++q) {
            g += M[(size_t)p * kb + c] * G[(size_t)p * kb + q] * M[(size_t)q * kb + d];
            h += M[(size_t)p * kb + c] * Hm[(size_t)p * kb + q] * M[(size_t)q * kb + d];
          }
        e1 = std::max(e1, std::fabs(g - (c == d ? 1.0 : 0.0)));
        e2 = std::max(e2, std::fabs(h - (c == d ? th[c] : 0.0)));
      }
    bool sorted = true;
    for (int c = 1; c < kb; ++c) sorted = sorted && th[c] >= th[c - 1];
    CHECK(e1 < 1e-9 && e2 < 1e-8 && sorted, "rayleigh-ritz kb=%d |M^T G M - I| %.2e |M^T H M - theta| %.2e sorted %d", kb, e1, e2, (int)sorted);
  }
  // a rank-deficient block is refused (the caller falls back to the library eigensolver)
  {
    const int kb = 16;
    std::vector<double> G((size_t)kb * kb, 1.0), Hm((size_t)kb * kb, 1.0), M((size_t)kb * kb), th(kb);
    CHECK(!sub_rayleigh_ritz(kb, G.data(), Hm.data(), M.data(), th.data()), "rank-1 block accepted");
  }
  // Lanczos bounds: 10 steps on a 200 x 200 symmetric matrix with a known spectrum enclose it from both sides
  {
    const int n = 200, k = 10;
    std::vector<double> A((size_t)n * n, 0.0), d(n);
    for (int i = 0; i < n; ++i) d[i] = -11.0 + 10.5 * i / (n - 1.0);
    // A = Q diag(d) Q^T with Q a product of Givens rotations (exact spectrum, dense matrix)
    for (int i = 0; i < n; ++i) A[(size_t)i * n + i] = d[i];
    for (int sweep = 0; sweep < 3; ++sweep)
      for (int i = 0; i + 1 < n; ++i) {
        const int j = (i * 7 + sweep * 13 + 1) % n;
        if (j == i) continue;
        const double c = std::cos(0.7 + i), s = std::sin(0.7 + i);
        for (int r = 0; r < n; ++r) {
          const double x = A[(size_t)r * n + i], y = A[(size_t)r * n + j];
          A[(size_t)r * n + i] = c * x - s * y;
          A[(size_t)r * n + j] = s * x + c * y;
        }
        for (int r = 0; r < n; ++r) {
          const double x = A[(size_t)i * n + r], y = A[(size_t)j * n + r];
          A[(size_t)i * n + r] = c * x - s * y;
          A[(size_t)j * n + r] = s * x + c * y;
        }
      }
    std::vector<double> v(n), vp(n, 0.0), w(n);
    double nrm = 0;
    for (int i = 0; i < n; ++i) { v[i] = N(rng); nrm += v[i] * v[i]; }
    for (auto& e : v) e /= std::sqrt(nrm);
    double alpha[16] = {}, beta[17] = {};
    for (int j = 0; j < k; ++j) {
      double a = 0, ww = 0;
      for (int i = 0; i < n; ++i) {
        double t = 0;
        for (int c = 0; c < n; ++c) t += A[(size_t)i * n + c] * v[c];
        w[i] = t - beta[j] * vp[i];
        a += w[i] * v[i];
        ww += w[i] * w[i];
      }
      alpha[j] = a;
      beta[j + 1] = std::sqrt(std::max(0.0, ww - a * a));
      for (int i = 0; i < n; ++i) { vp[i] = v[i]; v[i] = (w[i] - a * vp[i]) / beta[j + 1]; }
    }
    double lo, up;
    lanczos_ritz_bounds(k, alpha, beta, &lo, &up);
    CHECK(up >= -0.5 && up < 4.0 && lo <= -11.0 && lo > -16.0, "lanczos bounds [%.3f, %.3f] for a spectrum [-11, -0.5]", lo, up);
  }
  // filter degree: full degree for a converged block, reduced while the block is far from the invariant subspace
  {
    CHECK(chebyshev_degree(-10.8, 2.0, -11.2, 1e10, 24) == 24, "tracked block should use the full degree");
    const int m = chebyshev_degree(-4.8, 2.0, -13.8, 1e10, 24);
    CHECK(m >= 8 && m <= 16, "cold block degree %d", m);
    CHECK(chebyshev_degree(0.0, 1.0, -1e6, 1e10, 24) == 2, "degree floor");
  }
  printf("linalg_test fails=%d\n", fails);
  return fails != 0;
}
