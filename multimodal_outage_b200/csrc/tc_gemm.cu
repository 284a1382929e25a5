// Weight image builder for the tcgen05 position GEMM (see tc_gemm.cuh).
#include "tc_gemm.cuh"

namespace gwn {

__global__ void wprep_kernel(WPrepParams w) {
  const int total = w.K * w.N;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / w.N, n = i % w.N;
    const int q = k >> 5, kk = k & 31;
    float val = w.transposed ? w.W[w.w_off[q] + (long long)n * w.ld + kk] : w.W[w.w_off[q] + (long long)kk * w.ld + n];
    if (w.scale) val *= w.scale[kk];
    w.img[((long long)(k >> 3) * w.N + n) * 8 + (k & 7)] = __float2bfloat16_rn(val);
  }
  if (w.bias_out) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < w.N; n += gridDim.x * blockDim.x) {
      float b = w.bias ? w.bias[n] : 0.f;
      if (w.shift) {
        for (int k = 0; k < w.K; ++k) {
          const int q = k >> 5, kk = k & 31;
          const float val = w.transposed ? w.W[w.w_off[q] + (long long)n * w.ld + kk]
                                         : w.W[w.w_off[q] + (long long)kk * w.ld + n];
          b = fmaf(w.shift[kk], val, b);
        }
      }
      w.bias_out[n] = b;
    }
  }
}

int launch_wprep(const WPrepParams& w, cudaStream_t st) {
  GWN_REQUIRE(w.W && w.img && w.K % 32 == 0 && w.K / 32 <= PG_TC_MAX_CHUNKS && w.N >= 1, "wprep: bad argument");
  int blocks = (int)cdiv((long long)w.K * w.N, 256);
  wprep_kernel<<<blocks, 256, 0, st>>>(w);
  GWN_LAUNCHED();
  return 0;
}

}  // namespace gwn
