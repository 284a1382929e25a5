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
    if (w.half_odd && (n & 1)) val *= 0.5f;
    w.img[((long long)(k >> 3) * w.N + n) * 8 + (k & 7)] = __float2bfloat16_rn(val);
  }
  if (w.bias_out || w.bias_chunk) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < w.N; n += gridDim.x * blockDim.x) {
      float b = w.bias ? w.bias[n] : 0.f;
      if (w.shift) {
#pragma unroll 16
        for (int k = 0; k < w.K; ++k) {      // (unrolled: 16 independent loads in flight instead of one L2 round trip per step)
          const int q = k >> 5, kk = k & 31;
          const float val = w.transposed ? w.W[w.w_off[q] + (long long)n * w.ld + kk]
                                         : w.W[w.w_off[q] + (long long)kk * w.ld + n];
          b = fmaf(w.shift[kk], val, b);
        }
      }
      if (w.half_odd && (n & 1)) b *= 0.5f;
      if (w.bias_out) w.bias_out[n] = b;
      if (w.bias_chunk) {       // split-precision bias rows of the extra K=16 chunk
        const bf16 hi = __float2bfloat16_rn(b);
        const bf16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
        bf16* d0 = w.img + ((long long)(w.K >> 3) * w.N + n) * 8;
        bf16* d1 = w.img + ((long long)((w.K >> 3) + 1) * w.N + n) * 8;
        const bf16 z = __float2bfloat16_rn(0.f);
        d0[0] = hi; d0[1] = lo;
#pragma unroll
        for (int i = 2; i < 8; ++i) d0[i] = z;
#pragma unroll
        for (int i = 0; i < 8; ++i) d1[i] = z;
      }
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
