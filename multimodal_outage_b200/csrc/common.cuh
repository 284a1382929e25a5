// Shared helpers for libgwn (sm_100a only).  See include/gwn.h for the ABI.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gwn.h"

namespace gwn {

void set_error(const char* fmt, ...);

#define GWN_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::gwn::set_error(__VA_ARGS__);  \
      return -1;                      \
    }                                 \
  } while (0)

#define GWN_CUDA(expr)                                                         \
  do {                                                                         \
    cudaError_t e__ = (expr);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      ::gwn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,            \
                       cudaGetErrorString(e__));                               \
      return -2;                                                               \
    }                                                                          \
  } while (0)

void count_launch();
#define GWN_LAUNCHED()                  \
  do {                                  \
    ::gwn::count_launch();              \
    GWN_CUDA(cudaPeekAtLastError());    \
  } while (0)

typedef __nv_bfloat16 bf16;

// ---- programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its predecessor in the
// stream is still draining - its CTAs become resident as the predecessor's CTAs exit and run their prologue (barrier /
// TMEM set-up, resident weight and support images: data that is constant within a step).  pdl_wait() blocks until the
// predecessor has completed and its memory is visible: it must precede every access to data an earlier kernel of the
// step produces or still reads.  pdl_trigger() (placed AFTER the kernel's own pdl_wait: at most two kernels overlap)
// allows the successor's launch.  Both are no-ops for normal launches.  GWN_PDL=0 turns the launch attribute off (A/B).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Debug timelines (scripts/gpu_*_trace.py): a device pointer handed over in an environment variable.  Only builds made
// with -DGWN_TRACE (GWN_TRACE=1 python -m multimodal_outage_b200.build) look at the environment at all: the release
// library never parses an environment-supplied pointer and pays no getenv per launch.
long long* trace_ptr(const char* env_name);

static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

// position -> (sample, row in sample) with 32-bit arithmetic: a 64-bit div/mod costs several hundred cycles
// per call on the device and used to dominate the producer loops; launchers guarantee P < 2^31.
__device__ __forceinline__ void split_pos(long long pp, long long rows_per_n, long long& n, long long& rem) {
  const uint32_t q = (uint32_t)pp / (uint32_t)rows_per_n;
  n = q;
  rem = (uint32_t)pp - q * (uint32_t)rows_per_n;
}

// ---- 4-wide loads/stores of 32-channel rows (16 B fp32 / 8 B bf16) ----
__device__ __forceinline__ void load4(const float* p, float v[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float v[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.x));
  float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4(float* p, const float v[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float v[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}
// 8 channels: 32 B fp32 (two 16-byte accesses) / one 16-byte access bf16
__device__ __forceinline__ void load8(const float* p, float v[8]) { load4(p, v); load4(p + 4, v + 4); }
__device__ __forceinline__ void load8(const bf16* p, float v[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
}
__device__ __forceinline__ void store8(float* p, const float v[8]) { store4(p, v); store4(p + 4, v + 4); }
__device__ __forceinline__ void store8(bf16* p, const float v[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void load2(const float* p, float v[2]) {
  float2 t = *reinterpret_cast<const float2*>(p);
  v[0] = t.x; v[1] = t.y;
}
__device__ __forceinline__ void load2(const bf16* p, float v[2]) {
  float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
  v[0] = a.x; v[1] = a.y;
}
__device__ __forceinline__ void store2(float* p, float a, float b) {
  *reinterpret_cast<float2*>(p) = make_float2(a, b);
}
__device__ __forceinline__ void store2(bf16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- Philox4x32-7 (counter-based RNG for the fused dropout mask) ----
// Seven rounds (the minimum Salmon et al. report as passing BigCrush; the mask generator is the largest single item in
// the SIMT budget of the fused diffusion kernels, and every round is 4 multiplies on a serial chain).
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t offset, uint64_t counter) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32);
  uint32_t c2 = (uint32_t)offset, c3 = (uint32_t)(offset >> 32);
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// Dropout keep-mask for 16 consecutive channels: element e = row*32 + col -> group idx16 = e/16 (= row*2 + col/16).
// One Philox call yields sixteen 8-bit uniforms; an element is dropped when its uniform < thr = round(p * 256), and
// kept elements are scaled by 256 / (256 - thr): the mask is exactly unbiased for the drop rate thr/256 it realises
// (p = 0.3 -> 77/256 = 0.3008).
__device__ __forceinline__ void dropout16(uint64_t seed, uint64_t offset, uint64_t idx16, float p, float m[16]) {
  const uint4 r = philox4x32(seed, offset, idx16);
  uint32_t thr = (uint32_t)(p * 256.0f + 0.5f);
  thr = thr > 255u ? 255u : thr;
  const float inv = 256.0f / (256.0f - (float)thr);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[4 * i] = ((w[i] & 0xFFu) >= thr) ? inv : 0.f;
    m[4 * i + 1] = (((w[i] >> 8) & 0xFFu) >= thr) ? inv : 0.f;
    m[4 * i + 2] = (((w[i] >> 16) & 0xFFu) >= thr) ? inv : 0.f;
    m[4 * i + 3] = ((w[i] >> 24) >= thr) ? inv : 0.f;
  }
}
// the 8 channels of group idx8 = e/8 out of the same stream
__device__ __forceinline__ void dropout8(uint64_t seed, uint64_t offset, uint64_t idx8, float p, float m[8]) {
  float t[16];
  dropout16(seed, offset, idx8 >> 1, p, t);
  const bool h = (idx8 & 1) != 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = h ? t[8 + i] : t[i];
}
// the 4 channels [col, col+4) of row `row` (col % 4 == 0) out of the same stream
__device__ __forceinline__ void dropout4(uint64_t seed, uint64_t offset, uint64_t row, int col, float p, float m[4]) {
  float t[8];
  dropout8(seed, offset, row * 4 + (uint64_t)(col >> 3), p, t);
  const int h = (col >> 2) & 1;
  m[0] = h ? t[4] : t[0]; m[1] = h ? t[5] : t[1]; m[2] = h ? t[6] : t[2]; m[3] = h ? t[7] : t[3];
}

// ---- end-of-kernel accumulation of a CTA's partial result into a global fp32 array ----
// Same-address reductions serialise in L2 (~3.5 cycles each): when all CTAs of a persistent grid finish together and
// walk the same addresses in the same order, every address sees gridDim.x back-to-back reductions and a scalar,
// row-strided flush of a few thousand values takes 15-50 us.  The partial is therefore staged in shared memory and
// flushed with coalesced 16-byte vector reductions, each CTA starting at its own rotation of the array.
__device__ __forceinline__ void red_add_v4(float* dst, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// dst[r * ld + c] += src[r * src_ld + c] for r < rows, c < cols (cols % 4 == 0, src 16-byte aligned, src_ld % 4 == 0).
// Called by nthr threads with ids t = 0..nthr-1 after the staging writes were made visible (barrier).
__device__ __forceinline__ void red_flush_2d(float* dst, int ld, const float* src, int src_ld, int rows, int cols, int t,
                                             int nthr) {
  const int c4 = cols >> 2, n4 = rows * c4;
  const int rot = (int)(((long long)blockIdx.x * n4) / gridDim.x);
  const bool vec = ((reinterpret_cast<unsigned long long>(dst) & 15ull) == 0) && (ld & 3) == 0;
  for (int i = t; i < n4; i += nthr) {
    int k = i + rot;
    if (k >= n4) k -= n4;
    const int r = k / c4, c = (k - r * c4) << 2;
    const float4 v = *reinterpret_cast<const float4*>(src + (size_t)r * src_ld + c);
    float* d = dst + (size_t)r * ld + c;
    if (vec) red_add_v4(d, v);
    else { atomicAdd(d, v.x); atomicAdd(d + 1, v.y); atomicAdd(d + 2, v.z); atomicAdd(d + 3, v.w); }
  }
}
// flat variant for arrays whose length is not a multiple of 4 (dst[i] += src[i], i < n)
__device__ __forceinline__ void red_flush_1d(float* dst, const float* src, int n, int t, int nthr) {
  const bool vec = (reinterpret_cast<unsigned long long>(dst) & 15ull) == 0;
  const int n4 = vec ? (n >> 2) : 0;
  if (n4 > 0) {
    const int rot = (int)(((long long)blockIdx.x * n4) / gridDim.x);
    for (int i = t; i < n4; i += nthr) {
      int k = i + rot;
      if (k >= n4) k -= n4;
      red_add_v4(dst + 4 * k, *reinterpret_cast<const float4*>(src + 4 * k));
    }
  }
  for (int i = (n4 << 2) + t; i < n; i += nthr) atomicAdd(dst + i, src[i]);
}

}  // namespace gwn
