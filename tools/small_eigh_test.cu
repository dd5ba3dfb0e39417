// small_eigh_kernel (one-CTA Jacobi, n <= 32) against its definition: residual |A v - lambda v|, orthonormality of the
// eigenvectors, ascending order, and the eigenvalues of the host Householder / QL solver, on random, degenerate, diagonal and
// zero matrices of every size 1 .. 32; microseconds per launch for a batch of two.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/small_eigh_test tools/small_eigh_test.cu
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../nbed_b200/csrc/small_eigh.cuh"
#include "../nbed_b200/csrc/host_linalg.h"
using namespace nbd;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s -> %s line %d\n", #x, cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

int main() {
  double *dA, *dw;
  CK(cudaMalloc(&dA, 2 * 32 * 32 * 8));
  CK(cudaMalloc(&dw, 2 * 32 * 8));
  double worst_res = 0, worst_orth = 0, worst_val = 0;
  int bad = 0;
  srand(3);
  for (int n = 1; n <= 32; ++n)
    for (int kind = 0; kind < 5; ++kind) {
      std::vector<double> a((size_t)n * n, 0.0);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
          double v = 0.0;
          if (kind == 0) v = rand() / (double)RAND_MAX - 0.5 + (i == j ? -3.0 + 6.0 * i / n : 0.0);
          if (kind == 1) v = (i == j) ? 1.0 : 0.0;                                  // identity: all degenerate
          if (kind == 2) v = (i == j) ? (double)(i / 2) : ((i / 2 == j / 2) ? 1e-3 : 0.0);  // pairs of close levels
          if (kind == 3) v = 0.0;                                                  // zero matrix
          if (kind == 4) v = 1e6 * (rand() / (double)RAND_MAX - 0.5) * ((i + j) % 3 == 0 ? 1.0 : 1e-6);  // wide range
          a[(size_t)i * n + j] = v;
          a[(size_t)j * n + i] = kind == 0 ? 777.0 : v;  // kind 0: the upper triangle is garbage and must be ignored
        }
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j)
          if (kind == 0) a[(size_t)j * n + i] = 777.0;
      std::vector<double> sym((size_t)n * n);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) sym[(size_t)i * n + j] = a[(size_t)std::max(i, j) * n + std::min(i, j)];
      CK(cudaMemcpy(dA, a.data(), (size_t)n * n * 8, cudaMemcpyHostToDevice));
      small_eigh_kernel<<<1, SE_THREADS>>>(dA, dw, n);
      CK(cudaDeviceSynchronize());
      std::vector<double> v((size_t)n * n), w(n);
      CK(cudaMemcpy(v.data(), dA, (size_t)n * n * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(w.data(), dw, (size_t)n * 8, cudaMemcpyDeviceToHost));
      double nrm = 0;
      for (double x : sym) nrm = std::max(nrm, std::fabs(x));
      nrm = std::max(nrm, 1e-300);
      for (int k = 0; k < n; ++k) {
        if (k > 0 && w[k] < w[k - 1]) { ++bad; printf("order n=%d kind=%d\n", n, kind); }
        for (int i = 0; i < n; ++i) {
          double r = -w[k] * v[(size_t)k * n + i];
          for (int j = 0; j < n; ++j) r += sym[(size_t)i * n + j] * v[(size_t)k * n + j];
          worst_res = std::max(worst_res, std::fabs(r) / nrm);
        }
        for (int l = 0; l <= k; ++l) {
          double d = 0;
          for (int i = 0; i < n; ++i) d += v[(size_t)k * n + i] * v[(size_t)l * n + i];
          worst_orth = std::max(worst_orth, std::fabs(d - (k == l ? 1.0 : 0.0)));
        }
      }
      std::vector<double> hw, hv, t = sym;
      if (!householder_ql_eigh(n, t, hw, hv)) jacobi_eigh(n, t, hw, hv);
      std::sort(hw.begin(), hw.end());
      for (int k = 0; k < n; ++k) worst_val = std::max(worst_val, std::fabs(hw[k] - w[k]) / nrm);
    }
  // warm start: solve A, then A + small symmetric perturbation starting from the eigenvectors of A
  {
    double* dW;
    CK(cudaMalloc(&dW, 32 * 32 * 8));
    for (int n : {7, 24, 31}) {
      std::vector<double> a((size_t)n * n), b((size_t)n * n);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
          const double v = rand() / (double)RAND_MAX - 0.5 + (i == j ? -3.0 + 6.0 * i / n : 0.0);
          const double d = 1e-3 * (rand() / (double)RAND_MAX - 0.5);
          a[(size_t)i * n + j] = a[(size_t)j * n + i] = v;
          b[(size_t)i * n + j] = b[(size_t)j * n + i] = v + d;
        }
      CK(cudaMemcpy(dA, a.data(), (size_t)n * n * 8, cudaMemcpyHostToDevice));
      small_eigh_kernel<<<1, SE_THREADS>>>(dA, dw, n, dW, 0);
      CK(cudaMemcpy(dA, b.data(), (size_t)n * n * 8, cudaMemcpyHostToDevice));
      small_eigh_kernel<<<1, SE_THREADS>>>(dA, dw, n, dW, 1);
      CK(cudaDeviceSynchronize());
      std::vector<double> v((size_t)n * n), w(n);
      CK(cudaMemcpy(v.data(), dA, (size_t)n * n * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(w.data(), dw, (size_t)n * 8, cudaMemcpyDeviceToHost));
      for (int k = 0; k < n; ++k) {
        if (k > 0 && w[k] < w[k - 1]) { ++bad; printf("warm order n=%d\n", n); }
        for (int i = 0; i < n; ++i) {
          double r = -w[k] * v[(size_t)k * n + i];
          for (int j = 0; j < n; ++j) r += b[(size_t)i * n + j] * v[(size_t)k * n + j];
          worst_res = std::max(worst_res, std::fabs(r) / 3.0);
        }
        for (int l = 0; l <= k; ++l) {
          double d = 0;
          for (int i = 0; i < n; ++i) d += v[(size_t)k * n + i] * v[(size_t)l * n + i];
          worst_orth = std::max(worst_orth, std::fabs(d - (k == l ? 1.0 : 0.0)));
        }
      }
    }
    cudaFree(dW);
  }
  printf("sizes 1..32 x 5 kinds: max residual / |A| = %.2e, max |V^T V - 1| = %.2e, max eigenvalue deviation / |A| = %.2e, order errors %d\n",
         worst_res, worst_orth, worst_val, bad);
  // timing: batch of two, n = 7 and 24
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int n : {7, 24, 32}) {
    std::vector<double> a((size_t)2 * n * n);
    for (int b = 0; b < 2; ++b)
      for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
          const double v = rand() / (double)RAND_MAX - 0.5 + (i == j ? -3.0 + 6.0 * i / n : 0.0);
          a[(size_t)b * n * n + (size_t)i * n + j] = a[(size_t)b * n * n + (size_t)j * n + i] = v;
        }
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaMemcpy(dA, a.data(), a.size() * 8, cudaMemcpyHostToDevice));
      cudaEventRecord(e0);
      small_eigh_kernel<<<2, SE_THREADS>>>(dA, dw, n);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = std::min(best, ms);
    }
    double* dW2;
    CK(cudaMalloc(&dW2, 2 * 32 * 32 * 8));
    CK(cudaMemcpy(dA, a.data(), a.size() * 8, cudaMemcpyHostToDevice));
    small_eigh_kernel<<<2, SE_THREADS>>>(dA, dw, n, dW2, 0);
    float bestw = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      std::vector<double> b = a;
      for (int bz = 0; bz < 2; ++bz)
        for (int i = 0; i < n; ++i)
          for (int j = 0; j <= i; ++j) {
            const double d = 1e-4 * (rand() / (double)RAND_MAX - 0.5);
            b[(size_t)bz * n * n + (size_t)i * n + j] += d;
            if (i != j) b[(size_t)bz * n * n + (size_t)j * n + i] += d;
          }
      CK(cudaMemcpy(dA, b.data(), b.size() * 8, cudaMemcpyHostToDevice));
      cudaEventRecord(e0);
      small_eigh_kernel<<<2, SE_THREADS>>>(dA, dw, n, dW2, 1);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      bestw = std::min(bestw, ms);
    }
    cudaFree(dW2);
    printf("n=%d batch 2: %.1f us per launch cold, %.1f us warm-started (matrix changed by 1e-4)\n", n, 1e3 * best, 1e3 * bestw);
  }
  return (bad || worst_res > 1e-13 || worst_orth > 1e-13 || worst_val > 1e-13) ? 1 : 0;
}
