// Throughput of MUFU.TANH / MUFU.EX2 / MUFU.RCP / FFMA per SM per clock on this GPU (one-off microbenchmark).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/mufu_bench.cu -o scripts/bin/mufu_bench
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters, long long* cyc) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i * 0.1f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
      if (OP == 4) { unsigned u = __float_as_uint(x[i]); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u)); x[i] = __uint_as_float(u); }
      if (OP == 5) { unsigned u = __float_as_uint(x[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u)); x[i] = __uint_as_float(u); }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8);
  const char* names[6] = {"tanh.approx.f32", "ex2.approx.f32", "rcp.approx.f32", "fma.f32", "tanh.bf16x2 (2 results)", "ex2.bf16x2 (2 results)"};
  const int iters = 2000;
  for (int op = 0; op < 6; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      if (op == 0) k<0><<<148, 1024>>>(out, iters, cyc);
      if (op == 1) k<1><<<148, 1024>>>(out, iters, cyc);
      if (op == 2) k<2><<<148, 1024>>>(out, iters, cyc);
      if (op == 3) k<3><<<148, 1024>>>(out, iters, cyc);
      if (op == 4) k<4><<<148, 1024>>>(out, iters, cyc);
      if (op == 5) k<5><<<148, 1024>>>(out, iters, cyc);
      cudaDeviceSynchronize();
    }
    printf("%-26s %.2f instr-lanes/clk/SM\n", names[op], 1024.0 * 8 * iters / (double)*cyc);
  }
  return 0;
}
