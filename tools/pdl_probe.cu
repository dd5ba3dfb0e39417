// Probe: does programmatic dependent launch overlap two independent kernels of one stream on this system?
// nvcc -gencode arch=compute_100a,code=sm_100a -o pdl_probe pdl_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void spin_kernel(long long cycles, int trigger, int* sink) {
  if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {
  }
  if (sink && threadIdx.x == 0 && blockIdx.x == 0) *sink = 1;
}

static float run(int trigger, int pdl, int primary_ex, size_t smem_a, size_t smem_b, int grid_a, int grid_b) {
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  int* sink;
  cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const long long cyc = 4000000;  // ~2 ms
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a, st);
    if (primary_ex) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid_a); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem_a; cfg.stream = st;
      cudaLaunchKernelEx(&cfg, spin_kernel, cyc, trigger, sink);
    } else {
      spin_kernel<<<grid_a, 128, smem_a, st>>>(cyc, trigger, sink);
    }
    if (pdl) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid_b); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem_b; cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, spin_kernel, cyc, 0, sink);
    } else {
      spin_kernel<<<grid_b, 128, smem_b, st>>>(cyc, 0, sink);
    }
    cudaEventRecord(b, st);
    cudaStreamSynchronize(st);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    best = ms < best ? ms : best;
  }
  printf("trigger=%d pdl=%d primary_ex=%d smem=%zu/%zu grid=%d/%d : %.3f ms  (%s)\n", trigger, pdl, primary_ex, smem_a, smem_b, grid_a, grid_b,
         best, cudaGetErrorString(cudaGetLastError()));
  return best;
}

int main() {
  run(0, 0, 0, 0, 0, 148, 148);
  run(1, 1, 0, 0, 0, 148, 148);
  run(1, 1, 1, 0, 0, 148, 148);
  run(0, 1, 0, 0, 0, 148, 148);
  run(1, 1, 0, 57 * 1024, 57 * 1024, 296, 506);
  run(1, 1, 0, 57 * 1024, 57 * 1024, 296, 148);
  run(1, 0, 0, 57 * 1024, 57 * 1024, 296, 148);
  return 0;
}
