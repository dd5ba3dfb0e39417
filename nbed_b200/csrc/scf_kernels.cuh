// Element-wise / reduction stages of the embedded-SCF iteration: all HBM-bound,
// one pass per matrix, warp-shuffle + fixed-order block reductions (deterministic, no atomics).
#pragma once
#include "common.cuh"

namespace nbd {

// F_s = h + V_s + J - kscale*K_s ; vhf_s = J - kscale*K_s      (nbed/scf/huzinaga_scf.py:156-157)
__global__ void fock_assemble_kernel(const double* __restrict__ h, const double* __restrict__ V,
                                     const double* __restrict__ J, const double* __restrict__ K, double kscale,
                                     double* __restrict__ F, double* __restrict__ vhf, long nn, int nspin) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const double hj = h[i], j = J[i];
  for (int s = 0; s < nspin; ++s) {
    const double v = j - kscale * K[(long)s * nn + i];
    vhf[(long)s * nn + i] = v;
    F[(long)s * nn + i] = hj + V[(long)s * nn + i] + v;
  }
}

// h_eff_s = h + V_s (+ mu * P_s)   (nbed/driver.py:518,529)
__global__ void heff_kernel(const double* __restrict__ h, const double* __restrict__ V, const double* __restrict__ P,
                            double mu, double* __restrict__ out, long nn, int nspin) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  for (int s = 0; s < nspin; ++s) {
    double v = h[i] + V[(long)s * nn + i];
    if (P) v += mu * P[(long)s * nn + i];
    out[(long)s * nn + i] = v;
  }
}

// out = a + b (batched)
__global__ void add_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out,
                           long cnt) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cnt) out[i] = a[i] + b[i];
}

// Huz_s = -c (FG_s + FG_s^T) [ -c (FGv_s + FGv_s^T - 2 W_s) ];  F_s += Huz_s      (nbed/scf/huzinaga_scf.py:78-90,160)
// FG = F (gamma_occ S); optional virtual-projector part: FGv = F (gamma_virt S), W = (gamma_virt S)^T FGv.
__global__ void huzinaga_apply_kernel(const double* __restrict__ FG, const double* __restrict__ FGv,
                                      const double* __restrict__ W, double c, double* __restrict__ huz,
                                      double* __restrict__ F, int n) {
  __shared__ double tile[32][33];
  const long base = (long)blockIdx.z * n * n;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = bx + r, j = by + threadIdx.x;  // transposed block: rows from the x-block
    double v = 0.0;
    if (i < n && j < n) {
      v = FG[base + (long)i * n + j];
      if (FGv) v += FGv[base + (long)i * n + j];
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = by + r, j = bx + threadIdx.x;
    if (i < n && j < n) {
      const long e = base + (long)i * n + j;
      double s = FG[e] + tile[threadIdx.x][r];
      if (FGv) s += FGv[e] - 2.0 * W[e];
      const double v = -c * s;
      huz[e] = v;
      F[e] += v;
    }
  }
}

// out[b][j][i] = in[b][i][j]
__global__ void transpose_kernel(const double* __restrict__ in, double* __restrict__ out, int rows, int cols) {
  __shared__ double tile[32][33];
  const long base = (long)blockIdx.z * rows * cols;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = by + r, j = bx + threadIdx.x;
    tile[r][threadIdx.x] = (i < rows && j < cols) ? in[base + (long)i * cols + j] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = bx + r, j = by + threadIdx.x;  // out is cols x rows
    if (i < cols && j < rows) out[base + (long)i * rows + j] = tile[threadIdx.x][r];
  }
}

// dst[r][0..ld) = scale[r] * src[r][0..n) zero-padded; rows selected by index list (or identity)
__global__ void pad_rows_kernel(const double* __restrict__ src, long src_ld, double* __restrict__ dst, int ld, int n,
                                const double* __restrict__ scale, double fixed_scale) {
  const int r = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ld) return;
  const double s = (scale ? scale[r] : 1.0) * fixed_scale;
  dst[(long)r * ld + c] = c < n ? s * src[(long)r * src_ld + c] : 0.0;
}

// rows scaled: A[r][:] *= f(w[r])   mode 0: w^-1/2  mode 1: sqrt(|w|)  mode 2: w^-1/4
__global__ void scale_rows_kernel(double* __restrict__ A, const double* __restrict__ w, int n, int mode) {
  const int r = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const double f = mode == 0 ? 1.0 / sqrt(w[r]) : (mode == 1 ? sqrt(fabs(w[r])) : 1.0 / sqrt(sqrt(w[r])));
  A[(long)r * n + c] *= f;
}

// Fixed-shape two-stage reductions: REDUCE_BLOCKS partials, then one block sums them in index order.
constexpr int REDUCE_BLOCKS = 296;

// mode 0: sum a[i]*b[i]    mode 1: sum (a[i]-b[i])^2    mode 2: sum a[i][j]*b[j][i] (n x n matrices)
__global__ void reduce_partial_kernel(const double* __restrict__ a, const double* __restrict__ b, long cnt, int mode,
                                      int n, double* __restrict__ part) {
  __shared__ double red[32];
  double v = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (long)gridDim.x * blockDim.x) {
    if (mode == 0) v += a[i] * b[i];
    else if (mode == 1) { const double d = a[i] - b[i]; v += d * d; }
    else { const long r = i / n, c = i % n; v += a[i] * b[c * n + r]; }
  }
  v = block_sum(v, red);
  if (threadIdx.x == 0) part[blockIdx.x] = v;
}
__global__ void reduce_final_kernel(const double* __restrict__ part, int nparts, double* __restrict__ out) {
  __shared__ double red[32];
  double v = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) v += part[i];
  v = block_sum(v, red);
  if (threadIdx.x == 0) *out = v;
}

// out = sum_k coef[k] * xs[k]   (DIIS extrapolation, pyscf/lib/diis.py:extrapolate)
struct LinCombArgs {
  const double* xs[9];
  double coef[9];
  int nd;
};
__global__ void lincomb_kernel(LinCombArgs a, double* __restrict__ out, long cnt) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  double v = 0.0;
  for (int k = 0; k < a.nd; ++k) v += a.xs[k][i] * a.coef[k];
  out[i] = v;
}
__global__ void sub_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out,
                           long cnt) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cnt) out[i] = a[i] - b[i];
}
// out = a^T - a   per n x n matrix (CDIIS error vector  (SDF)^T - SDF, pyscf/scf/diis.py)
__global__ void antisym_kernel(const double* __restrict__ a, double* __restrict__ out, int n) {
  const long base = (long)blockIdx.z * n * n;
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j < n) out[base + (long)i * n + j] = a[base + (long)j * n + i] - a[base + (long)i * n + j];
}
// sum of squares of rows [row0, n) of a (n x ncols) matrix: orbital-gradient norm (virtual rows)
__global__ void rows_sqsum_kernel(const double* __restrict__ a, int row0, int nrows, int ncols, long ld,
                                  double* __restrict__ out) {
  __shared__ double red[32];
  double v = 0.0;
  const long cnt = (long)(nrows - row0) * ncols;
  for (long e = threadIdx.x; e < cnt; e += blockDim.x) {
    const long r = row0 + e / ncols, c = e % ncols;
    const double x = a[r * ld + c];
    v += x * x;
  }
  v = block_sum(v, red);
  if (threadIdx.x == 0) *out = v;
}

// F_s = heff_s + J - kscale*K_s ; vhf_s = J - kscale*K_s   (heff = hcore + V_emb [+ mu*P]; huzinaga_scf.py:156-157)
__global__ void fock_from_heff_kernel(const double* __restrict__ heff, const double* __restrict__ J,
                                      const double* __restrict__ K, double kscale, double* __restrict__ F,
                                      double* __restrict__ vhf, long nn, int nspin) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const double j = J[i];
  for (int s = 0; s < nspin; ++s) {
    const double v = j - kscale * K[(long)s * nn + i];
    vhf[(long)s * nn + i] = v;
    F[(long)s * nn + i] = heff[(long)s * nn + i] + v;
  }
}

// Multi-GPU form of the same step.  J and K are symmetric, so the ranks all-reduce only the lower triangles: the
// (1 + nspin) matrices are packed row by row ([m][i (i + 1) / 2 + j], j <= i) - 22.7 MB instead of 45.5 MB per Fock
// build at n = 1376 - and the Fock assembly reads both triangles from the packed sums (no unpacking pass).
__global__ void pack_lower_kernel(const double* __restrict__ A, double* __restrict__ P, int n, long npack) {
  const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.z;
  if (j <= i) P[(long)m * npack + (long)i * (i + 1) / 2 + j] = A[(long)m * n * n + (long)i * n + j];
}
__global__ void fock_from_heff_packed_kernel(const double* __restrict__ heff, const double* __restrict__ JKp, double kscale,
                                             double* __restrict__ F, double* __restrict__ vhf, int n, long npack, int nspin) {
  const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const long p = i >= j ? (long)i * (i + 1) / 2 + j : (long)j * (j + 1) / 2 + i;
  const long e = (long)i * n + j, nn = (long)n * n;
  const double jj = JKp[p];
  for (int s = 0; s < nspin; ++s) {
    const double v = jj - kscale * JKp[(long)(1 + s) * npack + p];
    vhf[(long)s * nn + e] = v;
    F[(long)s * nn + e] = heff[(long)s * nn + e] + v;
  }
}

// Per-spin traces of one SCF cycle in a single pass over the matrices (all symmetric, so sum_ij A_ij D_ji is
// taken element-wise):  part[blk][4*s + 0..2] = sum a_s.D_s, sum b_s.D_s, sum c_s.D_s ; [4*s+3] = sum (D_s - Dold_s)^2
// (nbed/scf/huzinaga_scf.py:182-194; nbed/scf/embedded_hcore_funcs.py:38-41).  Null operands are skipped.
__global__ void scf_traces_partial_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                          const double* __restrict__ c, const double* __restrict__ D,
                                          const double* __restrict__ Dold, long nn, int nspin,
                                          double* __restrict__ part) {
  __shared__ double red[32];
  for (int s = 0; s < nspin; ++s) {
    double va = 0.0, vb = 0.0, vc = 0.0, vd = 0.0;
    const long o = (long)s * nn;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += (long)gridDim.x * blockDim.x) {
      const double d = D[o + i];
      if (a) va = fma(a[o + i], d, va);
      if (b) vb = fma(b[o + i], d, vb);
      if (c) vc = fma(c[o + i], d, vc);
      if (Dold) { const double t = d - Dold[o + i]; vd = fma(t, t, vd); }
    }
    va = block_sum(va, red);
    if (threadIdx.x == 0) part[(long)blockIdx.x * 8 + 4 * s + 0] = va;
    vb = block_sum(vb, red);
    if (threadIdx.x == 0) part[(long)blockIdx.x * 8 + 4 * s + 1] = vb;
    vc = block_sum(vc, red);
    if (threadIdx.x == 0) part[(long)blockIdx.x * 8 + 4 * s + 2] = vc;
    vd = block_sum(vd, red);
    if (threadIdx.x == 0) part[(long)blockIdx.x * 8 + 4 * s + 3] = vd;
  }
}
__global__ void scf_traces_final_kernel(const double* __restrict__ part, int nparts, int nvals, double* __restrict__ out) {
  __shared__ double red[32];
  for (int k = 0; k < nvals; ++k) {
    double v = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) v += part[(long)i * 8 + k];
    v = block_sum(v, red);
    if (threadIdx.x == 0) out[k] = v;
  }
}

// several dot products against one vector in one launch: out[k] = sum_i x[i] * ys[k][i]   (DIIS Gram row)
struct MultiDotArgs {
  const double* ys[9];
  int nd;
};
__global__ void multidot_partial_kernel(const double* __restrict__ x, MultiDotArgs a, long cnt, double* __restrict__ part) {
  __shared__ double red[32];
  double v[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) v[k] = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (long)gridDim.x * blockDim.x) {
    const double xv = x[i];
#pragma unroll
    for (int k = 0; k < 9; ++k)
      if (k < a.nd) v[k] = fma(xv, a.ys[k][i], v[k]);
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    if (k < a.nd) {
      const double r = block_sum(v[k], red);
      if (threadIdx.x == 0) part[(long)blockIdx.x * 9 + k] = r;
    }
  }
}
__global__ void multidot_final_kernel(const double* __restrict__ part, int nparts, int nd, double* __restrict__ out) {
  __shared__ double red[32];
  for (int k = 0; k < nd; ++k) {
    double v = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) v += part[(long)i * 9 + k];
    v = block_sum(v, red);
    if (threadIdx.x == 0) out[k] = v;
  }
}

// ---- ao2mo epilogue + spin-orbital scatter (nbed/ham_builder.py:131-133,158-216,254) -----------------
// out[p][r][s][q] = eri[(p*m+q)][(r*m+s)]     (chemist (pq|rs) -> openfermion physicist order)
__global__ void chem_to_phys_kernel(const double* __restrict__ eri, double* __restrict__ out, int m, int transpose_in) {
  const long m2 = (long)m * m, m4 = m2 * m2;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m4) return;
  const int q = i % m, s = (i / m) % m, r = (i / m2) % m, p = i / (m2 * m);
  const long pq = (long)p * m + q, rs = (long)r * m + s;
  out[i] = transpose_in ? eri[rs * m2 + pq] : eri[pq * m2 + rs];
}
// h2[2p+a][2q+b][2r+c][2s+d] from two[blk][p][q][r][s]; blocks aaaa, bbbb, abba(<-aabb), baab(<-bbaa)
__global__ void spinorb_two_kernel(const double* __restrict__ two, double* __restrict__ h2, int m, double tol,
                                   double scale) {
  const int nq = 2 * m;
  const long n4 = (long)nq * nq * nq * nq;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int S = i % nq, R = (i / nq) % nq, Q = (i / ((long)nq * nq)) % nq, P = i / ((long)nq * nq * nq);
  const int a = P & 1, b = Q & 1, c = R & 1, d = S & 1;
  int blk = -1;
  if (a == 0 && b == 0 && c == 0 && d == 0) blk = 0;
  else if (a == 1 && b == 1 && c == 1 && d == 1) blk = 1;
  else if (a == 0 && b == 1 && c == 1 && d == 0) blk = 2;
  else if (a == 1 && b == 0 && c == 0 && d == 1) blk = 3;
  double v = 0.0;
  if (blk >= 0) {
    const long m2 = (long)m * m;
    v = two[(long)blk * m2 * m2 + ((long)(P >> 1) * m + (Q >> 1)) * m2 + (long)(R >> 1) * m + (S >> 1)];
    if (fabs(v) < tol) v = 0.0;
    v *= scale;
  }
  h2[i] = v;
}
__global__ void spinorb_one_kernel(const double* __restrict__ one, double* __restrict__ h1, int m, double tol) {
  const int nq = 2 * m;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq * nq) return;
  const int Q = i % nq, P = i / nq;
  double v = 0.0;
  if ((P & 1) == (Q & 1)) {
    v = one[(long)(P & 1) * m * m + (long)(P >> 1) * m + (Q >> 1)];
    if (fabs(v) < tol) v = 0.0;
  }
  h1[i] = v;
}

}  // namespace nbd
