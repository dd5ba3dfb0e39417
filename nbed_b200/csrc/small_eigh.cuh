// Symmetric eigensolver for SMALL matrices (n <= 32), one CTA per matrix, everything in shared memory.
//
// The reference diagonalises the Lowdin-orthogonalised Fock matrix with numpy.linalg.eigh in every SCF cycle
// (nbed/scf/huzinaga_scf.py:145,168) and S with fractional_matrix_power (:128).  For the small configurations
// (H2O / STO-3G n = 7, H2O / cc-pVDZ n = 24) cuSOLVER's dsyevd costs 0.1-0.2 ms per matrix - two thirds of the whole SCF
// cycle (profiles/bench_r02ah_C2.json) - because it is a chain of tiny launches.  This kernel is ONE launch: two-sided
// cyclic Jacobi with the round-robin (tournament) ordering, all n/2 rotations of a round computed from the same matrix
// and applied in parallel (columns of A and V, then rows of A), until the off-diagonal norm is below 1e-15 |A|_F.
// Jacobi is at least as accurate as the QR / divide-and-conquer path (eigenvalues to a few ulp of |A|).  Output in the
// library call's convention: eigenvalues ascending, eigenvectors as ROWS of the row-major output (= LAPACK's columns).
#pragma once
#include "common.cuh"

namespace nbd {

constexpr int SE_MAX_N = 32;
constexpr int SE_THREADS = 256;  // 8 warps: warp w owns the pairs w and w + 8 of a round, lane = row / column index

// warm (optional): [batch][n][n] eigenvector rows of the PREVIOUS solve of a slowly changing matrix (the Fock matrix of
// consecutive SCF cycles).  With use_warm the kernel first rotates A into that basis (A' = E A E^T, nearly diagonal), so
// the sweeps start a few orders of magnitude closer to convergence, and accumulates its rotations on E^T instead of the
// identity; the result is the same decomposition of A.  The new eigenvector rows are always written back to `warm`.
__global__ void __launch_bounds__(SE_THREADS) small_eigh_kernel(double* __restrict__ Aall, double* __restrict__ wall, int n,
                                                                double* __restrict__ warm = nullptr, int use_warm = 0) {
  __shared__ double A[SE_MAX_N][SE_MAX_N + 1];
  __shared__ double V[SE_MAX_N][SE_MAX_N + 1];
  __shared__ double T[SE_MAX_N][SE_MAX_N + 1];
  __shared__ double red[SE_THREADS / 32][2];
  __shared__ double norms[2];
  __shared__ int rank_of[SE_MAX_N];
  double* Ag = Aall + (long)blockIdx.x * n * n;
  double* w = wall + (long)blockIdx.x * n;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // numpy.linalg.eigh reads the lower triangle
  for (int e = tid; e < SE_MAX_N * SE_MAX_N; e += SE_THREADS) {
    const int i = e >> 5, j = e & 31;
    A[i][j] = (i < n && j < n) ? Ag[i >= j ? (long)i * n + j : (long)j * n + i] : 0.0;
    V[i][j] = i == j ? 1.0 : 0.0;
  }
  __syncthreads();
  if (warm != nullptr && use_warm) {
    const double* E = warm + (long)blockIdx.x * n * n;
    for (int e = tid; e < SE_MAX_N * SE_MAX_N; e += SE_THREADS) {
      const int i = e >> 5, k = e & 31;
      V[i][k] = (i < n && k < n) ? E[(long)k * n + i] : (i == k ? 1.0 : 0.0);  // columns of V = previous eigenvectors
    }
    __syncthreads();
    for (int e = tid; e < SE_MAX_N * SE_MAX_N; e += SE_THREADS) {  // T = A V
      const int i = e >> 5, k = e & 31;
      double t = 0.0;
      for (int j = 0; j < n; ++j) t = fma(A[i][j], V[j][k], t);
      T[i][k] = t;
    }
    __syncthreads();
    for (int e = tid; e < SE_MAX_N * SE_MAX_N; e += SE_THREADS) {  // A' = V^T T, symmetrised
      const int k = e >> 5, l = e & 31;
      if (k < n && l < n && l <= k) {
        double t = 0.0, u = 0.0;
        for (int i = 0; i < n; ++i) {
          t = fma(V[i][k], T[i][l], t);
          u = fma(V[i][l], T[i][k], u);
        }
        A[k][l] = A[l][k] = 0.5 * (t + u);
      }
    }
    __syncthreads();
  }
  const int m = (n + 1) & ~1;  // players of the tournament (index n = a bye when n is odd)
  const int npair = m / 2;
  for (int sweep = 0; sweep < 40; ++sweep) {
    // convergence: off-diagonal against the whole Frobenius norm (rows / columns >= n are zero)
    double off = 0.0, tot = 0.0;
    for (int e = tid; e < SE_MAX_N * SE_MAX_N; e += SE_THREADS) {
      const int i = e >> 5, j = e & 31;
      const double v = A[i][j] * A[i][j];
      tot += v;
      if (i != j) off += v;
    }
    off = warp_sum(off);
    tot = warp_sum(tot);
    if (lane == 0) {
      red[warp][0] = off;
      red[warp][1] = tot;
    }
    __syncthreads();
    if (tid == 0) {
      double o = 0.0, t = 0.0;
      for (int k = 0; k < SE_THREADS / 32; ++k) {
        o += red[k][0];
        t += red[k][1];
      }
      norms[0] = o;
      norms[1] = t;
    }
    __syncthreads();
    if (norms[0] <= 1e-30 * norms[1] || norms[0] == 0.0) break;
    for (int round = 0; round < m - 1; ++round) {
      // this warp's (up to two) pairs and their rotations, computed by every lane from the matrix before the round
      int pp[2], qq[2];
      double cc[2], ss[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = warp + 8 * h;
        int p = 0, q = 0;
        double c = 1.0, s = 0.0;
        if (k < npair) {
          if (k == 0) {
            p = m - 1;
            q = round;
          } else {
            p = round + k;
            if (p >= m - 1) p -= m - 1;
            q = round - k;
            if (q < 0) q += m - 1;
          }
          if (p > q) {
            const int t = p;
            p = q;
            q = t;
          }
          if (q < n) {
            const double apq = A[p][q];
            if (apq != 0.0) {
              // t = 2 a_pq / (d + sign(d) sqrt(d^2 + 4 a_pq^2)), d = a_qq - a_pp: the smaller root, |t| <= 1
              const double d = A[q][q] - A[p][p];
              const double r = sqrt(fma(d, d, 4.0 * apq * apq));
              const double t = 2.0 * apq / (d >= 0.0 ? d + r : d - r);
              c = rsqrt(fma(t, t, 1.0));
              s = t * c;
            }
          } else {
            q = p;  // bye
          }
        }
        pp[h] = p;
        qq[h] = q;
        cc[h] = c;
        ss[h] = s;
      }
      __syncthreads();
      // columns of A and V (lane = row): (x_p, x_q) <- (c x_p - s x_q, s x_p + c x_q)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int p = pp[h], q = qq[h];
        if (p == q || ss[h] == 0.0) continue;
        const double c = cc[h], s = ss[h];
        const double ap = A[lane][p], aq = A[lane][q];
        A[lane][p] = c * ap - s * aq;
        A[lane][q] = s * ap + c * aq;
        const double vp = V[lane][p], vq = V[lane][q];
        V[lane][p] = c * vp - s * vq;
        V[lane][q] = s * vp + c * vq;
      }
      __syncthreads();
      // rows of A (lane = column); the rotated pair element is zero by construction: set it exactly
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int p = pp[h], q = qq[h];
        if (p == q || ss[h] == 0.0) continue;
        const double c = cc[h], s = ss[h];
        const double ap = A[p][lane], aq = A[q][lane];
        A[p][lane] = lane == q ? 0.0 : c * ap - s * aq;
        A[q][lane] = lane == p ? 0.0 : s * ap + c * aq;
      }
      __syncthreads();
    }
  }
  // ascending order (ties by index), eigenvectors out as rows
  if (tid < n) {
    const double li = A[tid][tid];
    int r = 0;
    for (int j = 0; j < n; ++j) {
      const double lj = A[j][j];
      r += (lj < li || (lj == li && j < tid)) ? 1 : 0;
    }
    rank_of[tid] = r;
    w[r] = li;
  }
  __syncthreads();
  for (int e = tid; e < n * n; e += SE_THREADS) {
    const int k = e / n, i = e % n;  // eigenvector k (column k of V), component i
    Ag[(long)rank_of[k] * n + i] = V[i][k];
    if (warm != nullptr) warm[(long)blockIdx.x * n * n + (long)rank_of[k] * n + i] = V[i][k];
  }
}

}  // namespace nbd
