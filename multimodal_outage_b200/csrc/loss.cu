// Mean-squared-error loss of the training step (lit.py:24 `nn.MSELoss()`, applied at lit.py:41/50) as two kernels.
//
// Between the head's forward and backward the stock path runs six tiny dependent launches (squared difference, a
// single-pass mean reduction of 8.7 us, the root gradient fill, zeros_like, the elementwise backward, ...) while the
// GPU is otherwise idle.  Forward here is ONE launch without a workspace: a cluster of eight CTAs, each reducing an
// eighth of the elements; the leader adds the eight partial sums through distributed shared memory in a fixed order
// (deterministic, no atomics, nothing to zero - safe under CUDA-graph replay).  Backward is one elementwise launch
// that writes every element (no zero fill).
#include "common.cuh"

namespace gwn {

constexpr int MSE_CTAS = 8, MSE_THREADS = 1024;

__global__ void __cluster_dims__(MSE_CTAS, 1, 1) __launch_bounds__(MSE_THREADS)
mse_loss_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float* __restrict__ loss) {
  pdl_wait();
  pdl_trigger();
  __shared__ float wsum[32];
  __shared__ float part;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
  const long long n4 = vec ? (n >> 2) : 0;
  float s0 = 0.f, s1 = 0.f;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  long long i = (long long)rank * MSE_THREADS + tid;
  const long long stride = (long long)MSE_CTAS * MSE_THREADS;
  for (; i + stride < n4; i += 2 * stride) {        // two independent 16-byte load pairs in flight
    const float4 x0 = __ldg(a4 + i), y0 = __ldg(b4 + i), x1 = __ldg(a4 + i + stride), y1 = __ldg(b4 + i + stride);
    float d;
    d = x0.x - y0.x; s0 = fmaf(d, d, s0); d = x0.y - y0.y; s0 = fmaf(d, d, s0);
    d = x0.z - y0.z; s0 = fmaf(d, d, s0); d = x0.w - y0.w; s0 = fmaf(d, d, s0);
    d = x1.x - y1.x; s1 = fmaf(d, d, s1); d = x1.y - y1.y; s1 = fmaf(d, d, s1);
    d = x1.z - y1.z; s1 = fmaf(d, d, s1); d = x1.w - y1.w; s1 = fmaf(d, d, s1);
  }
  if (i < n4) {
    const float4 x0 = __ldg(a4 + i), y0 = __ldg(b4 + i);
    float d;
    d = x0.x - y0.x; s0 = fmaf(d, d, s0); d = x0.y - y0.y; s0 = fmaf(d, d, s0);
    d = x0.z - y0.z; s0 = fmaf(d, d, s0); d = x0.w - y0.w; s0 = fmaf(d, d, s0);
  }
  for (long long j = 4 * n4 + (long long)rank * MSE_THREADS + tid; j < n; j += stride) {      // tail / unaligned input
    const float d = a[j] - b[j];
    s1 = fmaf(d, d, s1);
  }
  float s = warp_sum(s0 + s1);
  if (lane == 0) wsum[warp] = s;
  __syncthreads();
  if (warp == 0) {
    s = warp_sum(wsum[lane]);
    if (lane == 0) part = s;
  }
  // partial sums visible across the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (rank == 0 && tid == 0) {
    double tot = 0.0;
    const uint32_t local = static_cast<uint32_t>(__cvta_generic_to_shared(&part));
    for (uint32_t r = 0; r < (uint32_t)MSE_CTAS; ++r) {
      uint32_t remote;
      float v;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
      tot += (double)v;
    }
    *loss = (float)(tot / (double)n);
  }
  // nobody leaves while the leader may still read its shared memory
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// d_a[i] = (a[i] - b[i]) * 2 g / n   (g = the incoming gradient of the scalar loss, read on the device)
__global__ void __launch_bounds__(256) mse_loss_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           const float* __restrict__ g, long long n, float* __restrict__ da) {
  pdl_wait();
  pdl_trigger();
  const float k = 2.f * __ldg(g) / (float)n;
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(da)) & 15) == 0;
  const long long n4 = vec ? (n >> 2) : 0;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x, stride = (long long)gridDim.x * 256;
  for (long long i = t; i < n4; i += stride) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(a) + i), y = __ldg(reinterpret_cast<const float4*>(b) + i);
    reinterpret_cast<float4*>(da)[i] = make_float4((x.x - y.x) * k, (x.y - y.y) * k, (x.z - y.z) * k, (x.w - y.w) * k);
  }
  for (long long j = 4 * n4 + t; j < n; j += stride) da[j] = (a[j] - b[j]) * k;
}

}  // namespace gwn

using namespace gwn;

extern "C" int gwn_mse_loss_fwd(const float* a, const float* b, long long n, float* loss, void* stream) {
  GWN_REQUIRE(a && b && loss && n >= 1, "mse_loss_fwd: bad argument");
  GWN_CUDA(launch_pdl(mse_loss_fwd_kernel, dim3(MSE_CTAS), dim3(MSE_THREADS), 0, reinterpret_cast<cudaStream_t>(stream), a, b, n, loss));
  GWN_LAUNCHED();
  return 0;
}

extern "C" int gwn_mse_loss_bwd(const float* a, const float* b, const float* grad_loss, long long n, float* d_a, void* stream) {
  GWN_REQUIRE(a && b && grad_loss && d_a && n >= 1, "mse_loss_bwd: bad argument");
  long long blocks = cdiv(cdiv(n, 4), 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  GWN_CUDA(launch_pdl(mse_loss_bwd_kernel, dim3((unsigned)blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), a, b, grad_loss, n, d_a));
  GWN_LAUNCHED();
  return 0;
}
