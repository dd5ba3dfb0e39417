// Which cuSOLVER symmetric eigensolver path is fastest for the per-iteration Fock diagonalisation?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/eig_bench tools/eig_bench.cu -lcusolver -lcublas
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#define CK(x) do { auto e_ = (x); if (e_ != 0) { printf("FAIL %s -> %d line %d\n", #x, (int)e_, __LINE__); exit(1);} } while (0)

static std::vector<double> make_sym(int n, unsigned seed) {
  std::vector<double> a((size_t)n * n);
  srand(seed);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double v = (rand() / (double)RAND_MAX - 0.5) * 0.1;
      if (i == j) v += -10.0 + 9.5 * i / n;
      a[(size_t)i * n + j] = a[(size_t)j * n + i] = v;
    }
  return a;
}

struct Timer {
  cudaEvent_t a, b;
  Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
  void start(cudaStream_t s = 0) { cudaEventRecord(a, s); }
  float stop(cudaStream_t s = 0) { cudaEventRecord(b, s); cudaEventSynchronize(b); float t; cudaEventElapsedTime(&t, a, b); return t; }
};

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 1376;
  int nocc = argc > 2 ? atoi(argv[2]) : 5;
  auto h = make_sym(n, 1);
  const size_t nn = (size_t)n * n;
  double *A, *A0, *W, *A2, *W2;
  CK(cudaMalloc(&A, 2 * nn * 8)); CK(cudaMalloc(&A0, nn * 8)); CK(cudaMalloc(&W, 2 * n * 8));
  A2 = A + nn; W2 = W + n;
  CK(cudaMemcpy(A0, h.data(), nn * 8, cudaMemcpyHostToDevice));
  int* info; CK(cudaMalloc(&info, 16));
  cusolverDnHandle_t H, H2; CK(cusolverDnCreate(&H)); CK(cusolverDnCreate(&H2));
  cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  CK(cusolverDnSetStream(H, s1)); CK(cusolverDnSetStream(H2, s2));
  Timer T;
  auto reset = [&]() { cudaMemcpy(A, A0, nn * 8, cudaMemcpyDeviceToDevice); cudaMemcpy(A2, A0, nn * 8, cudaMemcpyDeviceToDevice); cudaDeviceSynchronize(); };

  // 1. Dsyevd
  {
    int lw; CK(cusolverDnDsyevd_bufferSize(H, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, A, n, W, &lw));
    double *wk, *wk2; CK(cudaMalloc(&wk, (size_t)lw * 8)); CK(cudaMalloc(&wk2, (size_t)lw * 8));
    for (int rep = 0; rep < 3; ++rep) {
      reset(); T.start(s1);
      CK(cusolverDnDsyevd(H, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, A, n, W, wk, lw, info));
      printf("Dsyevd n=%d: %.3f ms\n", n, T.stop(s1));
    }
    // two solves issued back to back on two streams from one thread
    for (int rep = 0; rep < 2; ++rep) {
      reset(); cudaDeviceSynchronize(); T.start(0); cudaStreamSynchronize(0);
      auto t0 = std::chrono::steady_clock::now();
      CK(cusolverDnDsyevd(H, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, A, n, W, wk, lw, info));
      CK(cusolverDnDsyevd(H2, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, A2, n, W2, wk2, lw, info + 1));
      cudaDeviceSynchronize();
      printf("2x Dsyevd, two streams, one thread: %.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    // two solves from two host threads
    for (int rep = 0; rep < 3; ++rep) {
      reset(); cudaDeviceSynchronize();
      auto t0 = std::chrono::steady_clock::now();
      std::thread th([&] { cudaSetDevice(0); cusolverDnDsyevd(H2, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, A2, n, W2, wk2, lw, info + 1); cudaStreamSynchronize(s2); });
      CK(cusolverDnDsyevd(H, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, A, n, W, wk, lw, info));
      cudaStreamSynchronize(s1); th.join();
      printf("2x Dsyevd, two streams, two threads: %.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    cudaFree(wk); cudaFree(wk2);
  }
  // 2. Dsyevdx, lowest nocc by index
  {
    int lw, meig; CK(cusolverDnDsyevdx_bufferSize(H, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_I, CUBLAS_FILL_MODE_UPPER, n, A, n, 0, 0, 1, nocc, &meig, W, &lw));
    double* wk; CK(cudaMalloc(&wk, (size_t)lw * 8));
    for (int rep = 0; rep < 3; ++rep) {
      reset(); T.start(s1);
      CK(cusolverDnDsyevdx(H, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_I, CUBLAS_FILL_MODE_UPPER, n, A, n, 0, 0, 1, nocc, &meig, W, wk, lw, info));
      printf("Dsyevdx lowest %d: %.3f ms (meig %d)\n", nocc, T.stop(s1), meig);
    }
    cudaFree(wk);
  }
  // 3. Xsyevd (64-bit generic API)
  {
    cusolverDnParams_t par; CK(cusolverDnCreateParams(&par));
    size_t lwd, lwh; CK(cusolverDnXsyevd_bufferSize(H, par, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, A, n, CUDA_R_64F, W, CUDA_R_64F, &lwd, &lwh));
    void* wd; CK(cudaMalloc(&wd, lwd)); std::vector<char> wh(lwh + 1);
    for (int rep = 0; rep < 3; ++rep) {
      reset(); T.start(s1);
      CK(cusolverDnXsyevd(H, par, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, A, n, CUDA_R_64F, W, CUDA_R_64F, wd, lwd, wh.data(), lwh, info));
      printf("Xsyevd: %.3f ms\n", T.stop(s1));
    }
    cudaFree(wd);
  }
  // 4. XsyevBatched, batch = 2
  {
    cusolverDnParams_t par; CK(cusolverDnCreateParams(&par));
    size_t lwd, lwh;
    auto st = cusolverDnXsyevBatched_bufferSize(H, par, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, A, n, CUDA_R_64F, W, CUDA_R_64F, &lwd, &lwh, 2);
    if (st != CUSOLVER_STATUS_SUCCESS) printf("XsyevBatched_bufferSize status %d\n", (int)st);
    else {
      void* wd; CK(cudaMalloc(&wd, lwd)); std::vector<char> wh(lwh + 1);
      for (int rep = 0; rep < 3; ++rep) {
        reset(); T.start(s1);
        st = cusolverDnXsyevBatched(H, par, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, A, n, CUDA_R_64F, W, CUDA_R_64F, wd, lwd, wh.data(), lwh, info, 2);
        printf("XsyevBatched x2: %.3f ms (status %d)\n", T.stop(s1), (int)st);
      }
      cudaFree(wd);
    }
  }
  // 5. Dsyevj
  {
    syevjInfo_t ji; CK(cusolverDnCreateSyevjInfo(&ji)); CK(cusolverDnXsyevjSetTolerance(ji, 1e-14)); CK(cusolverDnXsyevjSetMaxSweeps(ji, 30));
    int lw; CK(cusolverDnDsyevj_bufferSize(H, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, A, n, W, &lw, ji));
    double* wk; CK(cudaMalloc(&wk, (size_t)lw * 8));
    reset(); T.start(s1);
    CK(cusolverDnDsyevj(H, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, A, n, W, wk, lw, info, ji));
    printf("Dsyevj: %.3f ms\n", T.stop(s1));
    cudaFree(wk);
  }
  // 6. generalised Dsygvd for reference
  {
    auto hs = make_sym(n, 2);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) hs[(size_t)i * n + j] = (i == j) ? 1.0 : hs[(size_t)i * n + j] * 0.02;
    double* S; CK(cudaMalloc(&S, nn * 8));
    int lw; CK(cusolverDnDsygvd_bufferSize(H, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, A, n, S, n, W, &lw));
    double* wk; CK(cudaMalloc(&wk, (size_t)lw * 8));
    for (int rep = 0; rep < 2; ++rep) {
      reset(); cudaMemcpy(S, hs.data(), nn * 8, cudaMemcpyHostToDevice); T.start(s1);
      CK(cusolverDnDsygvd(H, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, A, n, S, n, W, wk, lw, info));
      printf("Dsygvd: %.3f ms\n", T.stop(s1));
    }
  }
  printf("done\n");
  return 0;
}
