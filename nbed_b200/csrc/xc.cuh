// Exchange-correlation on an atom-centred grid (rows a3 / f3): what the reference's
// Kohn-Sham objects reach through pyscf.dft.numint + libxc in every get_veff of the embedded loop
// (nbed/scf/huzinaga_scf.py:55,156 with RKS/UKS objects; nbed/driver.py:289-313, 1154-1166).
//
//   eval_ao        phi_mu(r_g), grad phi_mu(r_g)            (numint.eval_ao, deriv = 1; once per geometry, resident)
//   T = Phi D      FP64 tensor-core GEMM (gemm.cuh)          (numint: rho = einsum('gi,ij,gj->g'))
//   rho, grad rho  one warp per grid point                   (warp-shuffle row reductions)
//   functional     one thread per grid point, forward-mode automatic differentiation of the energy density
//   M = w (1/2 vrho Phi + g . grad Phi) ; V = Phi^T M + (Phi^T M)^T   (GEMM + fused transpose-add)
//
// The grid (coords, weights) is an INPUT: the caller passes PySCF's mf.grids.coords / mf.grids.weights.
#pragma once
#include "common.cuh"
#include "integrals.cuh"

namespace nbd {

enum { NBD_XC_NONE = 0, NBD_XC_B3LYP = 1, NBD_XC_LDA = 2 };

// ---- forward-mode AD: value + derivatives with respect to (rho_a, rho_b, sigma_aa, sigma_ab, sigma_bb) ----------
struct Jet5 {
  double v, d[5];
};
__device__ __forceinline__ Jet5 jconst(double c) { return Jet5{c, {0, 0, 0, 0, 0}}; }
__device__ __forceinline__ Jet5 jvar(double x, int i) {
  Jet5 r = jconst(x);
  r.d[i] = 1.0;
  return r;
}
__device__ __forceinline__ Jet5 operator+(Jet5 a, Jet5 b) {
  Jet5 r;
  r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < 5; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
__device__ __forceinline__ Jet5 operator-(Jet5 a, Jet5 b) {
  Jet5 r;
  r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < 5; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
__device__ __forceinline__ Jet5 operator-(Jet5 a) {
  Jet5 r;
  r.v = -a.v;
#pragma unroll
  for (int i = 0; i < 5; ++i) r.d[i] = -a.d[i];
  return r;
}
__device__ __forceinline__ Jet5 operator*(Jet5 a, Jet5 b) {
  Jet5 r;
  r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < 5; ++i) r.d[i] = a.d[i] * b.v + b.d[i] * a.v;
  return r;
}
__device__ __forceinline__ Jet5 operator/(Jet5 a, Jet5 b) {
  Jet5 r;
  const double inv = 1.0 / b.v;
  r.v = a.v * inv;
#pragma unroll
  for (int i = 0; i < 5; ++i) r.d[i] = (a.d[i] - b.d[i] * r.v) * inv;
  return r;
}
__device__ __forceinline__ Jet5 operator+(Jet5 a, double c) { a.v += c; return a; }
__device__ __forceinline__ Jet5 operator+(double c, Jet5 a) { a.v += c; return a; }
__device__ __forceinline__ Jet5 operator-(Jet5 a, double c) { a.v -= c; return a; }
__device__ __forceinline__ Jet5 operator-(double c, Jet5 a) { return jconst(c) - a; }
__device__ __forceinline__ Jet5 operator*(Jet5 a, double c) {
  a.v *= c;
#pragma unroll
  for (int i = 0; i < 5; ++i) a.d[i] *= c;
  return a;
}
__device__ __forceinline__ Jet5 operator*(double c, Jet5 a) { return a * c; }
__device__ __forceinline__ Jet5 operator/(Jet5 a, double c) { return a * (1.0 / c); }
__device__ __forceinline__ Jet5 operator/(double c, Jet5 a) { return jconst(c) / a; }
__device__ __forceinline__ Jet5 jchain(Jet5 a, double f, double df) {
  Jet5 r;
  r.v = f;
#pragma unroll
  for (int i = 0; i < 5; ++i) r.d[i] = a.d[i] * df;
  return r;
}
__device__ __forceinline__ Jet5 jpow(Jet5 a, double p) {
  const double vp1 = pow(a.v, p - 1.0);
  return jchain(a, vp1 * a.v, p * vp1);
}
__device__ __forceinline__ Jet5 jlog(Jet5 a) { return jchain(a, log(a.v), 1.0 / a.v); }
__device__ __forceinline__ Jet5 jexp(Jet5 a) {
  const double e = exp(a.v);
  return jchain(a, e, e);
}
__device__ __forceinline__ Jet5 jsqrt(Jet5 a) {
  const double s = sqrt(a.v);
  return jchain(a, s, 0.5 / s);
}
__device__ __forceinline__ Jet5 jatan(Jet5 a) { return jchain(a, atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
__device__ __forceinline__ Jet5 jasinh(Jet5 a) { return jchain(a, asinh(a.v), 1.0 / sqrt(1.0 + a.v * a.v)); }

// ---- functionals: energy per unit volume (libxc definitions; see oracle/xc_restatement.py for the pin) -----------
__device__ inline Jet5 xc_slater_x(Jet5 ra, Jet5 rb) {
  const double cx = 0.75 * pow(3.0 / 3.14159265358979323846, 1.0 / 3.0);
  return (jpow(ra, 4.0 / 3.0) + jpow(rb, 4.0 / 3.0)) * (-cx * pow(2.0, 1.0 / 3.0));
}
__device__ inline Jet5 xc_b88_x(Jet5 ra, Jet5 rb, Jet5 saa, Jet5 sbb) {
  const double beta = 0.0042;
  Jet5 out = xc_slater_x(ra, rb);
  for (int s = 0; s < 2; ++s) {
    const Jet5 r = s == 0 ? ra : rb, sg = s == 0 ? saa : sbb;
    const Jet5 r43 = jpow(r, 4.0 / 3.0);
    const Jet5 x = jsqrt(sg) / r43;
    out = out - beta * r43 * x * x / (1.0 + 6.0 * beta * x * jasinh(x));
  }
  return out;
}
__device__ inline Jet5 xc_vwn_aux(double a, double b, double c, double x0, Jet5 rs) {
  const Jet5 x = jsqrt(rs);
  const Jet5 X = x * x + b * x + c;
  const double q = sqrt(4.0 * c - b * b);
  const double x0x = x0 * x0 + b * x0 + c;
  const Jet5 at = jatan(q / (2.0 * x + b));
  return a * (jlog(x * x / X) + (2.0 * b / q) * at -
              (b * x0 / x0x) * (jlog((x - x0) * (x - x0) / X) + (2.0 * (b + 2.0 * x0) / q) * at));
}
__device__ inline Jet5 xc_vwn_rpa_c(Jet5 ra, Jet5 rb) {
  const Jet5 rho = ra + rb;
  const Jet5 rs = pow(3.0 / (4.0 * 3.14159265358979323846), 1.0 / 3.0) * jpow(rho, -1.0 / 3.0);
  const Jet5 z = (ra - rb) / rho;
  const Jet5 fz = (jpow(1.0 + z, 4.0 / 3.0) + jpow(1.0 - z, 4.0 / 3.0) - 2.0) / (pow(2.0, 4.0 / 3.0) - 2.0);
  const Jet5 ep = xc_vwn_aux(0.0310907, 13.0720, 42.7198, -0.409286, rs);
  const Jet5 ef = xc_vwn_aux(0.01554535, 20.1231, 101.578, -0.743294, rs);
  return rho * (ep * (1.0 - fz) + ef * fz);
}
__device__ inline Jet5 xc_lyp_c(Jet5 ra, Jet5 rb, Jet5 saa, Jet5 sab, Jet5 sbb) {
  const double a = 0.04918, b = 0.132, c = 0.2533, d = 0.349;
  const double pi = 3.14159265358979323846;
  const Jet5 rho = ra + rb;
  const Jet5 rm13 = jpow(rho, -1.0 / 3.0);
  const Jet5 den = 1.0 + d * rm13;
  const Jet5 omega = jexp(-c * rm13) / den * jpow(rho, -11.0 / 3.0);
  const Jet5 delta = c * rm13 + d * rm13 / den;
  const double cf = 0.3 * pow(3.0 * pi * pi, 2.0 / 3.0);
  const Jet5 sig = saa + 2.0 * sab + sbb;
  const Jet5 t1 = -a * 4.0 / den * ra * rb / rho;
  Jet5 inner = ra * rb * (pow(2.0, 11.0 / 3.0) * cf * (jpow(ra, 8.0 / 3.0) + jpow(rb, 8.0 / 3.0)) +
                          (47.0 / 18.0 - 7.0 * delta / 18.0) * sig - (2.5 - delta / 18.0) * (saa + sbb) -
                          (delta - 11.0) / 9.0 * (ra / rho * saa + rb / rho * sbb));
  inner = inner - (2.0 / 3.0) * rho * rho * sig + ((2.0 / 3.0) * rho * rho - ra * ra) * sbb + ((2.0 / 3.0) * rho * rho - rb * rb) * saa;
  return t1 - a * b * omega * inner;
}

constexpr double XC_RHO_CUT = 1e-10;

// per grid point: f (energy density), vrho[2], vsigma[3]; out layout [6][ng]
__global__ void xc_functional_kernel(int code, const double* __restrict__ rho, const double* __restrict__ sigma, int ng,
                                     double* __restrict__ out) {
  const int gidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= ng) return;
  const double ra = rho[gidx], rb = rho[ng + gidx];
  double res[6] = {0, 0, 0, 0, 0, 0};
  if (ra > XC_RHO_CUT && rb > XC_RHO_CUT) {
    const Jet5 a = jvar(ra, 0), b = jvar(rb, 1);
    const Jet5 saa = jvar(fmax(sigma[gidx], 1e-40), 2), sab = jvar(sigma[ng + gidx], 3), sbb = jvar(fmax(sigma[2 * ng + gidx], 1e-40), 4);
    Jet5 e;
    if (code == NBD_XC_B3LYP)
      e = 0.08 * xc_slater_x(a, b) + 0.72 * xc_b88_x(a, b, saa, sbb) + 0.19 * xc_vwn_rpa_c(a, b) + 0.81 * xc_lyp_c(a, b, saa, sab, sbb);
    else
      e = xc_slater_x(a, b) + xc_vwn_rpa_c(a, b);
    res[0] = e.v;
    for (int i = 0; i < 5; ++i) res[1 + i] = e.d[i];
  }
  for (int i = 0; i < 6; ++i) out[(long)i * ng + gidx] = res[i];
}

// ---- AO values and gradients: one thread per grid point, all shells; ao layout [4][ng][nao] -------------------------
__global__ void xc_eval_ao_kernel(const IntShell* __restrict__ shells, int nsh, const double* __restrict__ env,
                                  const double* __restrict__ c2s, const double* __restrict__ coords, int ng, int nao,
                                  double* __restrict__ ao) {
  const int gidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= ng) return;
  const double gx = coords[3 * gidx], gy = coords[3 * gidx + 1], gz = coords[3 * gidx + 2];
  const long plane = (long)ng * nao;
  for (int is = 0; is < nsh; ++is) {
    const IntShell S = shells[is];
    const double dx = gx - S.x, dy = gy - S.y, dz = gz - S.z;
    const double r2 = dx * dx + dy * dy + dz * dz;
    double rad = 0.0, drad = 0.0;
    for (int ip = 0; ip < S.nprim; ++ip) {
      const double a = env[S.ptr_exp + ip];
      const double e = env[S.ptr_coef + ip] * exp(-a * r2);
      rad += e;
      drad += -a * e;
    }
    const int l = S.l, nc = ncart(l);
    double cart[4][ncart(INT_LMAX_AO)];
    double px[INT_LMAX_AO + 1], py[INT_LMAX_AO + 1], pz[INT_LMAX_AO + 1];
    px[0] = py[0] = pz[0] = 1.0;
    for (int k = 1; k <= l; ++k) {
      px[k] = px[k - 1] * dx;
      py[k] = py[k - 1] * dy;
      pz[k] = pz[k - 1] * dz;
    }
    int k = 0;
    for (int lx = l; lx >= 0; --lx)
      for (int ly = l - lx; ly >= 0; --ly, ++k) {
        const int lz = l - lx - ly;
        const double poly = px[lx] * py[ly] * pz[lz];
        cart[0][k] = poly * rad;
        const double dpx = lx > 0 ? lx * px[lx - 1] * (py[ly] * pz[lz]) : 0.0;
        const double dpy = ly > 0 ? ly * py[ly - 1] * (px[lx] * pz[lz]) : 0.0;
        const double dpz = lz > 0 ? lz * pz[lz - 1] * (px[lx] * py[ly]) : 0.0;
        cart[1][k] = dpx * rad + poly * 2.0 * dx * drad;
        cart[2][k] = dpy * rad + poly * 2.0 * dy * drad;
        cart[3][k] = dpz * rad + poly * 2.0 * dz * drad;
      }
    const double* T = c2s + c2s_off(l);
    for (int m = 0; m < 2 * l + 1; ++m) {
      double v[4] = {0, 0, 0, 0};
      for (int c = 0; c < nc; ++c) {
        const double t = T[m * nc + c];
        if (t == 0.0) continue;
        for (int q = 0; q < 4; ++q) v[q] += cart[q][c] * t;
      }
      for (int q = 0; q < 4; ++q) ao[q * plane + (long)gidx * nao + S.ao_off + m] = v[q];
    }
  }
}

// rho_s(g) = sum_mu T_s[g][mu] phi[g][mu] ; grad_k rho_s(g) = 2 sum_mu T_s[g][mu] d_k phi[g][mu]   (one warp per point)
// rho [2][ng], grad [2][3][ng]; T [2][np][nao] for the chunk starting at grid point g0 (np points)
__global__ void xc_density_kernel(const double* __restrict__ T, const double* __restrict__ ao, long plane, int g0, int np,
                                  int ng, int nao, double* __restrict__ rho, double* __restrict__ grad) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= np) return;
  const long row = (long)(g0 + warp) * nao;
  for (int s = 0; s < 2; ++s) {
    const double* t = T + ((long)s * np + warp) * nao;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int m = lane; m < nao; m += 32) {
      const double tv = t[m];
      a0 = fma(tv, ao[row + m], a0);
      a1 = fma(tv, ao[plane + row + m], a1);
      a2 = fma(tv, ao[2 * plane + row + m], a2);
      a3 = fma(tv, ao[3 * plane + row + m], a3);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
    if (lane == 0) {
      rho[(long)s * ng + g0 + warp] = a0;
      grad[((long)s * 3 + 0) * ng + g0 + warp] = 2.0 * a1;
      grad[((long)s * 3 + 1) * ng + g0 + warp] = 2.0 * a2;
      grad[((long)s * 3 + 2) * ng + g0 + warp] = 2.0 * a3;
    }
  }
}
// sigma [3][ng] = (aa, ab, bb) from grad [2][3][ng]
__global__ void xc_sigma_kernel(const double* __restrict__ grad, int ng, double* __restrict__ sigma) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  double aa = 0.0, ab = 0.0, bb = 0.0;
  for (int k = 0; k < 3; ++k) {
    const double ga = grad[(long)k * ng + g], gb = grad[((long)3 + k) * ng + g];
    aa += ga * ga; ab += ga * gb; bb += gb * gb;
  }
  sigma[g] = aa; sigma[ng + g] = ab; sigma[2L * ng + g] = bb;
}
// M_s[g][mu] = phi (1/2 w vrho_s) + sum_k d_k phi  w (2 vsigma_ss grad_k rho_s + vsigma_ab grad_k rho_s')
// fx [6][ng] = (f, vra, vrb, vsaa, vsab, vsbb);  M [2][ng][nao]; grid points g0 + blockIdx.y
__global__ void xc_potential_rows_kernel(const double* __restrict__ ao, long plane, const double* __restrict__ w,
                                         const double* __restrict__ fx, const double* __restrict__ grad, int g0, int ng,
                                         int nao, double* __restrict__ M) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = g0 + blockIdx.y, s = blockIdx.z;
  if (m >= nao) return;
  const double wg = w[g];
  const double vr = fx[(long)(1 + s) * ng + g], vss = fx[(long)(s == 0 ? 3 : 5) * ng + g], vab = fx[4L * ng + g];
  const long row = (long)g * nao + m;
  double v = ao[row] * (0.5 * wg * vr);
  for (int k = 0; k < 3; ++k) {
    const double gv = 2.0 * vss * grad[((long)s * 3 + k) * ng + g] + vab * grad[((long)(1 - s) * 3 + k) * ng + g];
    v += ao[(k + 1) * plane + row] * (wg * gv);
  }
  M[((long)s * ng + g) * nao + m] = v;
}
// V[b] = P[b] + P[b]^T  (in place on P's lower/upper pairs; batch = gridDim.z)
__global__ void add_transpose_kernel(double* __restrict__ P, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  double* a = P + (long)blockIdx.z * n * n;
  if (j < n && j >= i) {
    const double s = a[(long)i * n + j] + a[(long)j * n + i];
    a[(long)i * n + j] = s;
    a[(long)j * n + i] = s;
  }
}
// out[0] = sum_g w f ; out[1 + s] = sum_g w rho_s   (two-stage deterministic reduction: partials [blocks][3])
__global__ void xc_integrate_partial_kernel(const double* __restrict__ w, const double* __restrict__ fx, const double* __restrict__ rho,
                                            int ng, double* __restrict__ part) {
  __shared__ double red[32];
  double a = 0.0, b = 0.0, c = 0.0;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += gridDim.x * blockDim.x) {
    const double wg = w[g];
    a = fma(wg, fx[g], a);
    b = fma(wg, rho[g], b);
    c = fma(wg, rho[ng + g], c);
  }
  a = block_sum(a, red);
  b = block_sum(b, red);
  c = block_sum(c, red);
  if (threadIdx.x == 0) {
    part[3 * blockIdx.x] = a;
    part[3 * blockIdx.x + 1] = b;
    part[3 * blockIdx.x + 2] = c;
  }
}
__global__ void xc_integrate_final_kernel(const double* __restrict__ part, int nparts, double* __restrict__ out) {
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int k = 0; k < nparts; ++k) s += part[3 * k + threadIdx.x];
    out[threadIdx.x] = s;
  }
}
// out[e] = sum_q part[q][e]  (fixed order)
__global__ void sum_partials_kernel(const double* __restrict__ part, int nparts, long cnt, double* __restrict__ out) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cnt) return;
  double s = 0.0;
  for (int q = 0; q < nparts; ++q) s += part[(long)q * cnt + e];
  out[e] = s;
}
// F_s += V_s, vhf_s += V_s
// pair[0] = pair[1] = D / 2: the spin densities of a restricted Kohn-Sham object
__global__ void halve_to_pair_kernel(const double* __restrict__ D, double* __restrict__ pair, long cnt) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cnt) {
    const double v = 0.5 * D[i];
    pair[i] = v;
    pair[cnt + i] = v;
  }
}
__global__ void xc_add_potential_kernel(const double* __restrict__ V, double* __restrict__ F, double* __restrict__ vhf, long cnt) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  F[i] += V[i];
  vhf[i] += V[i];
}

}  // namespace nbd
