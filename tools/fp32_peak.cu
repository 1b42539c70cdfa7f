// fp32_peak.cu — micro-benchmark of the FP32 pipe on the box (SURVEY.md §8d asks for the measured figure beside hbm_gbs):
// (a) FFMA throughput (2 flop each), (b) separately rounded FMUL+FADD throughput — the parity build's arithmetic
// (--fmad=false) — as thread-operations per second.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp32_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <bool FMA>
__global__ void k(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.0f, x2 = x0 + 2.0f, x3 = x0 + 3.0f, x4 = x0 + 4.0f, x5 = x0 + 5.0f, x6 = x0 + 6.0f, x7 = x0 + 7.0f;
  for (int i = 0; i < iters; i++) {
    if (FMA) {
      x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
      x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
    } else {
      x0 = __fadd_rn(__fmul_rn(x0, a), b); x1 = __fadd_rn(__fmul_rn(x1, a), b); x2 = __fadd_rn(__fmul_rn(x2, a), b); x3 = __fadd_rn(__fmul_rn(x3, a), b);
      x4 = __fadd_rn(__fmul_rn(x4, a), b); x5 = __fadd_rn(__fmul_rn(x5, a), b); x6 = __fadd_rn(__fmul_rn(x6, a), b); x7 = __fadd_rn(__fmul_rn(x7, a), b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
  float* out;
  cudaMalloc(&out, (size_t)blocks * threads * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; mode++) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) k<true><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
      else k<false><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const double inst = (double)blocks * threads * iters * 8.0 * (mode == 0 ? 1.0 : 2.0);
    printf("{\"mode\": \"%s\", \"thread_inst_per_s\": %.4g, \"flop_per_s\": %.4g, \"ms\": %.3f, \"sms\": %d}\n", mode == 0 ? "ffma" : "fmul+fadd (no contraction)",
           inst / (best * 1e-3), inst * (mode == 0 ? 2.0 : 1.0) / (best * 1e-3), best, p.multiProcessorCount);
  }
  return cudaGetLastError() != cudaSuccess;
}
