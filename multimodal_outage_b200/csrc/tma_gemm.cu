// Host side of the TMA-fed tcgen05 GEMM (tma_gemm.cuh): tensor-map construction through the driver entry
// point (no -lcuda link dependency), operand descriptor constants, and the diffusion hop for graphs whose
// supports do not fit on chip (V > 128; the 3,100-node configurations):
//
//   Y[s, w, c] (+)= sum_v Aop[w, v] * X[s, v, c]           nconv, graph_wavenet.py:60-66
//
//   GEMM view: M = w (node), K = v (node), N = (slab, channel).  A operand = bf16 image of the support,
//   [w][v] row-major (K-major, 128B swizzle; Aop = A^T for the forward hop, A for the backward hop);
//   B operand = the channels-last activation slot itself (MN-major, 64B swizzle, 3-D TMA box of 8 slabs).
#include <mutex>

#include "tma_gemm.cuh"
#include "tma_gemm2.cuh"
#include "tma_hops.cuh"

namespace gwn {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int tg_map_2d(CUtensorMap* map, const void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes, uint32_t box0,
              uint32_t box1) {
  EncodeTiledFn fn = encode_fn();
  GWN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  GWN_REQUIRE(stride1_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(base) % 16) == 0 && box0 * 2 <= 128 && box1 <= 256,
              "tma map: unaligned tensor (stride %llu)", (unsigned long long)stride1_bytes);
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {stride1_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GWN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed: %d", (int)r);
  return 0;
}

int tg_map_slabs(CUtensorMap* map, const void* base, uint64_t V, uint64_t slabs, uint32_t box_v, uint32_t box_slabs) {
  EncodeTiledFn fn = encode_fn();
  GWN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  GWN_REQUIRE((reinterpret_cast<uintptr_t>(base) % 16) == 0 && box_v <= 256 && box_slabs <= 256, "tma slab map: bad box");
  cuuint64_t dims[3] = {32, V, slabs};
  cuuint64_t strides[2] = {64, V * 64};
  cuuint32_t box[3] = {32, box_v, box_slabs};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GWN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
  return 0;
}

int tg_map_rows3d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t n, uint64_t pitch, uint32_t box_rows) {
  return tg_map_rows3d_box(map, base, rows, n, pitch, box_rows, 1);
}

int tg_map_rows3d_box(CUtensorMap* map, const void* base, uint64_t rows, uint64_t n, uint64_t pitch, uint32_t box_rows,
                      uint32_t box_n) {
  EncodeTiledFn fn = encode_fn();
  GWN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  GWN_REQUIRE((reinterpret_cast<uintptr_t>(base) % 16) == 0 && (pitch * 2) % 16 == 0 && box_rows <= 256,
              "tma rows map: unaligned tensor");
  cuuint64_t dims[3] = {32, rows, n};
  cuuint64_t strides[2] = {pitch * 2, rows * pitch * 2};
  cuuint32_t box[3] = {32, box_rows, box_n};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GWN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(rows3d) failed: %d", (int)r);
  return 0;
}

void tg_operand(TgOperand& o, int mode, int rows) {
  o.mode = mode;
  if (mode == TG_K_SW128) {            // [rows][64 k] : row = 128 B, 8-row atoms of 1024 B
    o.tile_bytes = (uint32_t)rows * 128u;
    o.lbo = 16; o.sbo = 1024; o.layout = 2;
    for (int i = 0; i < 4; ++i) o.koff[i] = 32u * i;
    o.n_boxes = 1; o.box_bytes = o.tile_bytes; o.box_mn = rows;
  } else if (mode == TG_MN_SW128) {    // boxes of [64 k][64 mn]: 8 KB each
    o.n_boxes = rows / 64; o.box_bytes = 64u * 128u; o.box_mn = 64;
    o.tile_bytes = (uint32_t)o.n_boxes * o.box_bytes;
    o.lbo = o.box_bytes; o.sbo = 1024; o.layout = 2;
    for (int i = 0; i < 4; ++i) o.koff[i] = 2048u * i;
  } else if (mode == TG_MN_SW64) {     // [rows/32 slabs][64 nodes][64 B]
    o.n_boxes = 1; o.box_mn = rows; o.tile_bytes = (uint32_t)(rows / 32) * 64u * 64u; o.box_bytes = o.tile_bytes;
    o.lbo = 64u * 64u; o.sbo = 512; o.layout = 4;
    for (int i = 0; i < 4; ++i) o.koff[i] = 1024u * i;
  } else {                             // TG_K_SW64: [2 slabs][rows nodes][64 B]
    o.n_boxes = 1; o.box_mn = rows; o.tile_bytes = 2u * (uint32_t)rows * 64u; o.box_bytes = o.tile_bytes;
    o.lbo = 16; o.sbo = 512; o.layout = 4;
    o.koff[0] = 0; o.koff[1] = 32; o.koff[2] = (uint32_t)rows * 64u; o.koff[3] = (uint32_t)rows * 64u + 32u;
  }
  o.tile_bytes = (o.tile_bytes + 1023u) & ~1023u;
}

int tg_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

bool tg2_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("GWN_TG2");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

// ------------------------------------------------------------------------------------------ epilogues
struct EpiHopBig {   // accumulator row = node w, column = (slab, channel)
  bf16* out; const bf16* add; int V; long long slabs;
  __device__ __forceinline__ void chunk(int m, bool m_ok, int n0, float v[32]) const {
    const long long slab = n0 >> 5;
    if (!m_ok || slab >= slabs) return;
    const long long off = (slab * V + m) * 32;
    if (add) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t[4];
        load4(add + off + 4 * j, t);
        v[4 * j] += t[0]; v[4 * j + 1] += t[1]; v[4 * j + 2] += t[2]; v[4 * j + 3] += t[3];
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 pk;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]);
      __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
      __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
      pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(out + off + 8 * j) = pk;
    }
  }
};

struct EpiStoreF32 {  // C[m][n] fp32 row-major (tests); atomic when split-K
  float* C; int ldc, N, atomic;
  __device__ __forceinline__ void chunk(int m, bool m_ok, int n0, float v[32]) const {
    if (!m_ok) return;
    float* dst = C + (long long)m * ldc + n0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n0 + j < N) {
        if (atomic) atomicAdd(dst + j, v[j]); else dst[j] = v[j];
      }
  }
};

// ------------------------------------------------------------------------------------------ big-V hop
int launch_hop_big(const bf16* img, int Vp, const bf16* X, bf16* Y, const bf16* add, long long slabs, int V,
                   cudaStream_t st) {
  if (slabs <= 0) return 0;
  GWN_REQUIRE(Vp % 8 == 0 && Vp >= V, "hop_big: image pitch %d must be a multiple of 8 and >= V", Vp);
  CUtensorMap ma, mb;
  if (int rc = tg_map_2d(&ma, img, (uint64_t)V, (uint64_t)V, (uint64_t)Vp * 2, 64, 128)) return rc;
  const int bn = slabs >= 8 ? 256 : (int)(32 * slabs);
  if (int rc = tg_map_slabs(&mb, X, (uint64_t)V, (uint64_t)slabs, 64, (uint32_t)(bn / 32))) return rc;
  TgParams p{};
  p.M = V; p.K = V; p.N = (int)(slabs * 32); p.bn = bn; p.splits = 1;
  GWN_REQUIRE(slabs * 32 < (1ll << 31), "hop_big: too many slabs");
  tg_operand(p.a, TG_K_SW128, 128);
  EpiHopBig e{Y, add, V, slabs};
  if (bn == 256 && V > 128 && tg2_enabled()) {        // CTA pairs: 256 nodes x 8 slabs per tile, 4 slabs staged per CTA
    if (int rc = tg_map_slabs(&mb, X, (uint64_t)V, (uint64_t)slabs, 64, 4)) return rc;
    tg_operand(p.b, TG_MN_SW64, 128);
    return launch_tma_gemm2(ma, mb, p, e, st);
  }
  tg_operand(p.b, TG_MN_SW64, bn);
  return launch_tma_gemm(ma, mb, p, e, st);
}

// ------------------------------------------------------------------------------------------ sparse hop (ELL rows)
// acc[0..7] += a * (the eight bf16 channels of q), fp32 throughout.  Packed fp32 FMA (fma.rn.f32x2: FFMA2 on sm_100, two
// IEEE fp32 fused multiply-adds per instruction, the same results as two FFMAs) - the sparse hop is instruction-issue bound and
// per channel pair this is shift + mask + one FFMA2 instead of shift + mask + two FFMAs.
__device__ __forceinline__ void ell_fma8(uint64_t acc[4], float a, const uint4& q) {
  const uint32_t u[4] = {q.x, q.y, q.z, q.w};
  uint64_t a2;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint64_t b2;
    asm("mov.b64 %0, {%1, %2};" : "=l"(b2) : "r"(u[i] << 16), "r"(u[i] & 0xFFFF0000u));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(a2), "l"(b2));
  }
}
__device__ __forceinline__ void ell_acc_init(uint64_t acc[4], const uint4& q) {   // acc = the eight bf16 channels of q
  const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "r"(u[i] << 16), "r"(u[i] & 0xFFFF0000u));
}
__device__ __forceinline__ uint4 ell_acc_pack(const uint64_t acc[4]) {             // eight fp32 sums -> eight bf16 (rn)
  uint4 o;
  uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    ow[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return o;
}
// the four (index, value) pairs k0..k0+3 of an ELL row: one 16-byte load each when the row pitch allows it
__device__ __forceinline__ void ell_row4(const int* ir, const float* vr, int k0, int W, bool vec, int v[4], float a[4]) {
  if (vec) {
    const int4 vi = __ldg(reinterpret_cast<const int4*>(ir + k0));
    const float4 va = __ldg(reinterpret_cast<const float4*>(vr + k0));
    v[0] = vi.x; v[1] = vi.y; v[2] = vi.z; v[3] = vi.w;
    a[0] = va.x; a[1] = va.y; a[2] = va.z; a[3] = va.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool in = k0 + j < W;
      v[j] = in ? __ldg(ir + k0 + j) : -1;
      a[j] = in ? __ldg(vr + k0 + j) : 0.f;
    }
  }
}
// y[s, w, :] = sum_k val[w][k] * x[s, idx[w][k], :] (+ add[s, w, :]); thread = (row w, 8-channel piece), 64 rows per CTA,
// blockIdx.y = slab.  The index / value rows are shared by every slab (L1 / L2 hits); x rows are 64-byte gathers that stay
// in L2 (a slab of 3,100 nodes is 198 KB).  HBM traffic: x once, y once - a dense V x V GEMM at 0.3 % density does 300x the
// arithmetic for the same numbers.
__global__ void __launch_bounds__(256) hop_ell_kernel(const int* __restrict__ idx, const float* __restrict__ val, int W,
                                                      const bf16* __restrict__ X, bf16* Y,
                                                      const bf16* add, int V) {      // (add may alias Y: y += A x)
  pdl_wait();
  pdl_trigger();
  const int w = blockIdx.x * 64 + (threadIdx.x >> 2), pc = threadIdx.x & 3;
  if (w >= V) return;
  const long long base = (long long)blockIdx.y * V;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (add) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(add + (base + w) * 32) + pc);
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[2 * i] = __uint_as_float(u[i] << 16); acc[2 * i + 1] = __uint_as_float(u[i] & 0xFFFF0000u); }
  }
  const int* ir = idx + (long long)w * W;
  const float* vr = val + (long long)w * W;
  // four neighbours at a time: their index / value loads, then the four independent 16-byte gathers, are in flight together
  // (one neighbour per iteration is a chain of two dependent loads); rows hold their non-zeros first, so a chunk that starts
  // with a pad ends the row; pads inside a chunk contribute zeros (nothing is loaded for them)
  for (int k0 = 0; k0 < W; k0 += 4) {
    int v[4]; float a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool in = k0 + j < W;
      v[j] = in ? __ldg(ir + k0 + j) : -1;
      a[j] = in ? __ldg(vr + k0 + j) : 0.f;
    }
    if (v[0] < 0) break;
    uint4 q[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      q[j] = v[j] >= 0 ? __ldg(reinterpret_cast<const uint4*>(X + (base + v[j]) * 32) + pc) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t u[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] = fmaf(a[j], __uint_as_float(u[i] << 16), acc[2 * i]);
        acc[2 * i + 1] = fmaf(a[j], __uint_as_float(u[i] & 0xFFFF0000u), acc[2 * i + 1]);
      }
    }
  }
  uint4 o;
  uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
    ow[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *(reinterpret_cast<uint4*>(Y + (base + w) * 32) + pc) = o;
}

// The same hop with the slab staged in shared memory (V x 64 B <= 220 KB: every 3,100-node slab is 198 KB): one CTA per
// slab reads x[s] once, coalesced, and the ~8 gathers per output row hit shared memory instead of L2 (the global version
// moves 8x the slab through the L2 -> SM path and is bound by it).  Persistent over slabs.
__global__ void __launch_bounds__(1024) hop_ell_smem_kernel(const int* __restrict__ idx, const float* __restrict__ val, int W,
                                                            const bf16* __restrict__ X, bf16* Y, const bf16* add, int V,
                                                            int slabs) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ uint4 xs[];                    // [V][4] pieces of 8 channels
  const int tid = threadIdx.x, pc = tid & 3;
  const bool vec = (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(val)) & 15) == 0;
  for (int s = blockIdx.x; s < slabs; s += gridDim.x) {
    const long long base = (long long)s * V;
    const uint4* src = reinterpret_cast<const uint4*>(X + base * 32);
    for (int i = tid; i < V * 4; i += 1024) xs[i] = __ldg(src + i);
    __syncthreads();
    for (int w = tid >> 2; w < V; w += 256) {
      uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
      if (add) ell_acc_init(acc, *(reinterpret_cast<const uint4*>(add + (base + w) * 32) + pc));
      const int* ir = idx + (long long)w * W;
      const float* vr = val + (long long)w * W;
      for (int k0 = 0; k0 < W; k0 += 4) {
        int v[4]; float a[4];
        ell_row4(ir, vr, k0, W, vec, v, a);
        if (v[0] < 0) break;
#pragma unroll
        for (int j = 0; j < 4; ++j)      // pads inside a chunk: nothing is loaded, zeros are added
          ell_fma8(acc, a[j], v[j] >= 0 ? xs[v[j] * 4 + pc] : make_uint4(0u, 0u, 0u, 0u));
      }
      *(reinterpret_cast<uint4*>(Y + (base + w) * 32) + pc) = ell_acc_pack(acc);
    }
    __syncthreads();                               // the slab buffer is reused by the next slab of this CTA
  }
}

int launch_hop_ell(const int* idx, const float* val, int W, const bf16* X, bf16* Y, const bf16* add, long long slabs, int V,
                   cudaStream_t st) {
  if (slabs <= 0) return 0;
  GWN_REQUIRE(idx && val && W >= 1 && X && Y && X != Y, "hop_ell: bad argument");
  const size_t slab_bytes = (size_t)V * 64;
  if (slab_bytes <= 220 * 1024 && slabs < (1ll << 31)) {       // the slab fits in shared memory
    static bool attr = false;
    static int sms = 148;
    if (!attr) {
      GWN_CUDA(cudaFuncSetAttribute(hop_ell_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      int dev = 0;
      GWN_CUDA(cudaGetDevice(&dev));
      GWN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      attr = true;
    }
    const int per_sm = slab_bytes <= 100 * 1024 ? 2 : 1;
    const long long grid = slabs < (long long)sms * per_sm ? slabs : (long long)sms * per_sm;
    GWN_CUDA(launch_pdl(hop_ell_smem_kernel, dim3((unsigned)grid), dim3(1024), slab_bytes, st, idx, val, W, X, Y, add, V, (int)slabs));
    GWN_LAUNCHED();
    return 0;
  }
  GWN_REQUIRE(slabs <= 65535, "hop_ell: too many slabs (%lld)", slabs);
  GWN_CUDA(launch_pdl(hop_ell_kernel, dim3((unsigned)cdiv(V, 64), (unsigned)slabs), dim3(256), 0, st, idx, val, W, X, Y, add, V));
  GWN_LAUNCHED();
  return 0;
}

// dA[v, w] += sum_{s,c} X[s,v,c] * G[s,w,c]   (nconv gradient wrt the support; adaptive adjacency only)
//
// Split-K: at V = 3100 the output is 25 x 13 = 325 tiles of 128 x 256 - 2.2 rounds of work for 148 persistent CTAs that take
// 3 (each tile contracts over every slab: ~100 us).  With the slab range cut into `splits` parts the CTAs walk
// 325 * splits smaller tiles (5 parts: 10.98 rounds of a fifth each) and add their partial sums with 16-byte vector
// reductions; tiles of one part run together, so the CTAs still share operand boxes through L2.
struct EpiAccF32 {
  float* C; int ldc, N, atomic;
  __device__ __forceinline__ void chunk(int m, bool m_ok, int n0, float v[32]) const {
    if (!m_ok) return;
    float* dst = C + (long long)m * ldc + n0;
    if (atomic) {
      if (n0 + 32 <= N && (ldc & 3) == 0 && (reinterpret_cast<unsigned long long>(C) & 15ull) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) red_add_v4(dst + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
        return;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) atomicAdd(dst + j, v[j]);
      return;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n0 + j < N) dst[j] += v[j];
  }
};

// Number of K parts that minimises the persistent grid's makespan, from a cost model fitted to B200 timings at V = 3100
// (scripts/gpu_big_graph_bench.py): a CTA's share is ceil(tiles * s / sms) tiles of k_blocks / s K blocks at ~0.52 us each
// plus ~5 us per tile (accumulator hand-off and the reductions' L2 traffic); the single-part form pays ~30 us per tile for its
// row-per-thread read-modify-write epilogue instead.  768 slabs: 663 -> 508 us, 192 slabs: 287 -> 159 us.
// GWN_DADJ_SPLITS overrides (A/B runs; 1 = the deterministic single-part form).
static int dadj_splits(long long tiles, int k_blocks, int sms, double us_per_kb = 0.52) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("GWN_DADJ_SPLITS");
    forced = e ? atoi(e) : 0;
  }
  if (forced > 0) return forced;
  if (tiles % sms == 0 || k_blocks < 32) return 1;    // nothing to balance / too little work per tile to cut
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= 8; ++s) {
    if (s > 1 && k_blocks / s < 16) break;            // keep >= 16 K blocks (64 MMAs) per tile
    const long long rounds = (tiles * s + sms - 1) / sms;
    const double cost = (double)rounds * (us_per_kb * k_blocks / s + (s == 1 ? 30.0 : 5.0));
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}

int launch_dadj_big(const bf16* X, const bf16* G, float* dA, long long slabs, int V, cudaStream_t st) {
  if (slabs <= 0) return 0;
  CUtensorMap ma, mb;
  const int bn = V >= 256 ? 256 : 32 * (int)cdiv(V, 32);
  if (int rc = tg_map_slabs(&ma, X, (uint64_t)V, (uint64_t)slabs, 128, 2)) return rc;
  if (int rc = tg_map_slabs(&mb, G, (uint64_t)V, (uint64_t)slabs, (uint32_t)bn, 2)) return rc;
  TgParams p{};
  GWN_REQUIRE(slabs * 32 < (1ll << 31), "dadj_big: too many slabs");
  p.M = V; p.N = V; p.K = (int)(slabs * 32); p.bn = bn;
  tg_operand(p.a, TG_K_SW64, 128);
  if (bn == 256 && tg2_enabled()) {                     // CTA pairs: 256 x 256 tiles, each CTA stages 128 rows of X and of G
    if (int rc = tg_map_slabs(&mb, G, (uint64_t)V, (uint64_t)slabs, 128, 2)) return rc;
    p.splits = dadj_splits(cdiv(V, 2 * TG_BM) * cdiv(V, bn), (int)cdiv(p.K, TG_BK), tg_sm_count() / 2, 0.75);
    tg_operand(p.b, TG_K_SW64, 128);
    EpiAccF32 e{dA, V, V, p.splits > 1 ? 1 : 0};
    return launch_tma_gemm2(ma, mb, p, e, st);
  }
  p.splits = dadj_splits(cdiv(V, TG_BM) * cdiv(V, bn), (int)cdiv(p.K, TG_BK), tg_sm_count());
  tg_operand(p.b, TG_K_SW64, bn);
  EpiAccF32 e{dA, V, V, p.splits > 1 ? 1 : 0};
  return launch_tma_gemm(ma, mb, p, e, st);
}

// supports fp32 [V][V] -> bf16 images [2][V][Vp]: image 0 = A^T (forward hop operand), image 1 = A (backward)
__global__ void support_images_kernel(const float* __restrict__ A, bf16* __restrict__ out, int V, int Vp) {
  __shared__ float t[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;    // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int r = by + j, c = bx + tx;
    const float v = (r < V && c < V) ? A[(long long)r * V + c] : 0.f;
    t[j][tx] = v;
    if (r < V && c < Vp) out[(long long)V * Vp + (long long)r * Vp + c] = __float2bfloat16_rn(v);   // image 1 = A
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int r = bx + j, c = by + tx;     // transposed: out0[r][c] = A[c][r]
    if (r < V && c < Vp) out[(long long)r * Vp + c] = __float2bfloat16_rn(c < V ? t[tx][j] : 0.f);
  }
}

}  // namespace gwn

using namespace gwn;

extern "C" long long gwn_support_images_bytes(int V, int n_supports) {
  const long long Vp = ((V + 7) / 8) * 8;
  return (long long)n_supports * 2 * V * Vp * 2;
}

extern "C" int gwn_support_images_prep(const float* const* supports, int n_supports, int V, void* out, void* stream) {
  GWN_REQUIRE(supports && out && n_supports >= 1 && n_supports <= GWN_MAX_SUPPORTS && V >= 1, "support_images: bad argument");
  const int Vp = ((V + 7) / 8) * 8;
  dim3 grid((unsigned)cdiv(Vp, 32), (unsigned)cdiv(V, 32)), block(32, 8);
  for (int s = 0; s < n_supports; ++s) {
    support_images_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        supports[s], reinterpret_cast<bf16*>(out) + (long long)s * 2 * V * Vp, V, Vp);
    GWN_LAUNCHED();
  }
  return 0;
}

// One hop over a slot-major activation buffer: y = Aop * x (+ add), Aop = image `which` (0: A^T, 1: A) of support s.
extern "C" int gwn_hop_big(const void* images, int n_supports, int support, int which, const void* x, void* y,
                           const void* add, long long slabs, int V, void* stream) {
  GWN_REQUIRE(images && x && y && support >= 0 && support < n_supports && (which == 0 || which == 1), "hop_big: bad argument");
  const int Vp = ((V + 7) / 8) * 8;
  const bf16* img = reinterpret_cast<const bf16*>(images) + ((long long)support * 2 + which) * V * Vp;
  return launch_hop_big(img, Vp, reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y),
                        reinterpret_cast<const bf16*>(add), slabs, V, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int gwn_hop_ell(const int* idx, const float* val, int width, const void* x, void* y, const void* add,
                           long long slabs, int V, void* stream) {
  return launch_hop_ell(idx, val, width, reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y),
                        reinterpret_cast<const bf16*>(add), slabs, V, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int gwn_dadj_big(const void* x, const void* g, float* dA, long long slabs, int V, void* stream) {
  GWN_REQUIRE(x && g && dA && V >= 1, "dadj_big: bad argument");
  return launch_dadj_big(reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(g), dA, slabs, V,
                         reinterpret_cast<cudaStream_t>(stream));
}

// Test entry: C[M][N] fp32 = A (.) B in every staging mode (see tma_gemm.cuh).
//   a_mode 0: A [M][lda] (K contiguous)      a_mode 1: A [K][lda] (M contiguous)
//   b_mode 0: B [N][ldb] (K contiguous)      b_mode 1: B [K][ldb] (N contiguous)     b_mode 2: B [N/32][K][32]
extern "C" int gwn_gemm_test(const void* A, const void* B, float* C, int M, int N, int K, int a_mode, int b_mode,
                             int lda, int ldb, int bn, int splits, void* stream) {
  GWN_REQUIRE(A && B && C, "gemm_test: NULL argument");
  CUtensorMap ma, mb;
  TgParams p{};
  p.M = M; p.N = N; p.K = K; p.bn = bn; p.splits = splits;
  if (a_mode == 0) {
    if (int rc = tg_map_2d(&ma, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, 64, 128)) return rc;
    tg_operand(p.a, TG_K_SW128, 128);
  } else {
    if (int rc = tg_map_2d(&ma, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, 64)) return rc;
    tg_operand(p.a, TG_MN_SW128, 128);
  }
  if (b_mode == 0) {
    if (int rc = tg_map_2d(&mb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, 64, (uint32_t)bn)) return rc;
    tg_operand(p.b, TG_K_SW128, bn);
  } else if (b_mode == 1) {
    GWN_REQUIRE(bn % 64 == 0, "gemm_test: MN-major B needs bn %% 64 == 0");
    if (int rc = tg_map_2d(&mb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, 64)) return rc;
    tg_operand(p.b, TG_MN_SW128, bn);
  } else {
    if (int rc = tg_map_slabs(&mb, B, (uint64_t)K, (uint64_t)(N / 32), 64, (uint32_t)(bn / 32))) return rc;
    tg_operand(p.b, TG_MN_SW64, bn);
  }
  EpiStoreF32 e{C, N, N, splits > 1 ? 1 : 0};
  return launch_tma_gemm(ma, mb, p, e, reinterpret_cast<cudaStream_t>(stream));
}
