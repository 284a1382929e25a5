// Microbenchmark: how fast can ONE thread issue tcgen05.mma (M=128, K=16, bf16) for N = 32..256?
// Operands are whatever is in shared memory (values irrelevant).  Prints cycles per MMA.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../multimodal_outage_b200/csrc/tc.cuh"
using namespace gwn::tc;

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int n_mma, int N, long long* out, int side) {
  __shared__ volatile int stop;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 0 && lane == 0) {
    const bool a_mn = side >= 10;
    const uint32_t idesc = make_idesc_bf16(128, N, a_mn, true);
    const uint32_t sb = smem_u32(smem);
    long long t0 = clock64();
    if (MODE == 0) {          // same descriptors every time (best case for the issue loop)
      const uint64_t ad = a_mn ? make_smem_desc(sb, 128, 1024) : make_smem_desc(sb, 2048, 128), bd = make_smem_desc(sb + 32768, 128, 1280);
      for (int i = 0; i < n_mma; ++i) umma_bf16(tb, ad, bd, idesc, 1u);
    } else {                  // descriptors recomputed per MMA like the fused kernel does
      const uint64_t adm = make_smem_desc(0, 1280, 128), bdu = make_smem_desc(0, 128, 1280);
      for (int i = 0; i < n_mma; i += 30)
        for (int m = 0; m < 6; ++m)
          for (int ks = 0; ks < 5; ++ks) {
            const uint64_t ad = adm + (uint64_t)((sb + m * 12800 + 2 * ks * 1280) >> 4);
            const uint64_t bd = bdu + (uint64_t)((sb + 32768 + m * 5120 + ks * 256) >> 4);
            umma_bf16(tb, ad, bd, idesc, 1u);
          }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
    stop = 1;
  } else if (warp > 0 && (side % 10)) {
    // side traffic: side=1 TMEM loads (x32) of this warp's quadrant; side=2 shared-memory 16-byte stores
    float acc = 0.f; int cnt = 0;
    while (!stop) {
      if (side % 10 == 1) {
        float v[32];
        tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + 256u, v);
        acc += v[lane];
      } else {
        *reinterpret_cast<uint4*>(smem + 65536 + warp * 4096 + lane * 16 + (cnt & 7) * 512) = make_uint4(cnt, 0, 0, 0);
      }
      ++cnt;
    }
    if (acc == 123.f) out[1] = cnt;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  for (int side : {0, 1, 2, 10, 11, 12})
  for (int mode = 0; mode < 1; ++mode)
    for (int N : {32, 80, 128, 256}) {
      const int n = 3000;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, 128, 128 * 1024>>>(n, N, d, side); else k<1><<<1, 128, 128 * 1024>>>(n, N, d, side);
        cudaDeviceSynchronize();
      }
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("side %d mode %d N=%3d: issue %.1f cyc/MMA, issue+drain %.1f cyc/MMA (floor %d)  err=%s\n", side, mode, N, (double)h[0] / n,
             (double)h[1] / n, 128 * N / 256, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
