// Weight-gradient style contractions (reduction over positions) on tcgen05.
//
// wgrad_tc: dW[(chunk,c), n] += sum_p A[p,(chunk,c)] * G[p,n]                (conv / mlp weight grads)
//   MMA orientation: M = weight row (128 per M-tile, up to 3 tiles), N = output channel, K = position.
//   Both operands are channels-last rows, i.e. MN-major: 16-byte pieces (8 channels of one position) are
//   cp.async'ed straight into the MN-major no-swizzle canonical layout (8 K-rows x 16 B core matrices).
//   The accumulator stays in TMEM over the CTA's whole share of positions; one epilogue at the end adds it to
//   the global fp32 gradient with atomics.  A resident "ones" row (M index 32*n_chunks) yields the bias
//   gradient sum_p G[p,n] for free.  Optional affine (folded BatchNorm of the layer input):
//   dW = scale_c * D + shift_c * db.
//
// dadj_tc: dA[v,w] += sum_{s,c} X[s,v,c] * G[s,w,c]                           (nconv grad wrt the support)
//   M = v, N = w, K = channel; both operands K-major (channels contiguous): per slab 2 K-steps.
#include <cstdlib>

#include "tc.cuh"
#include "tc_wgrad.cuh"

namespace gwn {

constexpr int WG_PT = 64;            // positions per tile (4 MMA K-steps)
constexpr int WG_STAGES = 3;
constexpr int WG_PRODUCERS = 128;    // warps 0-3
constexpr int WG_MMA_WARP = 4;
constexpr int WG_THREADS = 32 * 9;   // warps 5-8 epilogue

#define WG_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && g < 48 && (tid & 31) == 0) p.trace[g * 8 + (slot)] = clock64(); } while (0)

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mrows = 32 * p.n_chunks;                 // real weight rows; row `mrows` is the ones row
  const int mt = (mrows + 1 + 127) / 128;            // M tiles
  const int N = p.N;
  const uint32_t a_bytes = (uint32_t)mt * 16u * WG_PT * 16u;      // 16 m8-blocks per tile, [PT x 16 B] each
  const uint32_t g_bytes = (uint32_t)(N / 8) * WG_PT * 16u;
  const uint32_t stage_bytes = a_bytes + g_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WG_STAGES * stage_bytes);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + WG_STAGES;    // [STAGES]
  uint64_t* tfull = bars + 2 * WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);
  float* db_s = reinterpret_cast<float*>(bars + 2 * WG_STAGES + 2);   // [N]

  if (tid == 0) {
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], WG_PRODUCERS); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  const uint32_t tmem_cols = (mt * N <= 32) ? 32u : (mt * N <= 64) ? 64u : (mt * N <= 128) ? 128u : 256u;
  if (warp == WG_MMA_WARP) tmem_alloc(tmem_slot, tmem_cols);
  {  // zero the A regions once (padding rows / unused chunk blocks stay zero), then plant the ones row
    for (int s = 0; s < WG_STAGES; ++s) {
      uint4* a = reinterpret_cast<uint4*>(smem + (size_t)s * stage_bytes);
      for (int i = tid; i < (int)(a_bytes / 16); i += WG_THREADS) a[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    const int m8 = mrows / 8;   // mrows is a multiple of 32 -> the ones row is element 0 of block m8
    for (int s = 0; s < WG_STAGES; ++s) {
      uint4* blk = reinterpret_cast<uint4*>(smem + (size_t)s * stage_bytes + (size_t)m8 * WG_PT * 16);
      for (int i = tid; i < WG_PT; i += WG_THREADS) blk[i] = make_uint4(0x00003F80u, 0u, 0u, 0u);  // bf16 1.0
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < WG_MMA_WARP) {
    // ===================== producers =====================
    const int cg = tid & 3, r0 = tid >> 2;   // rows r0 and r0+32 of the tile, channel group cg
    const uint32_t ro32 = (uint32_t)p.rows_per_n_out;
    int g = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int stage = g % WG_STAGES, phase = (g / WG_STAGES) & 1;
      mbar_wait(&empty[stage], (uint32_t)(phase ^ 1));
      if (warp == 0) WG_TRACE(0);
      const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
      const uint32_t sg = sa + a_bytes;
      // 32-bit row arithmetic (launcher guarantees every source has < 2^31 rows); one IMAD.WIDE per copy
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = r0 + 32 * h;
        const uint32_t pp = (uint32_t)tile * WG_PT + (uint32_t)i;
        const bool pv = pp < (uint32_t)p.P;
        const uint32_t n = pp / ro32, rem = pp - n * ro32;
        const uint32_t sdst = sa + (uint32_t)(cg * WG_PT + i) * 16u;
#pragma unroll
        for (int q = 0; q < WG_MAX_CHUNKS; ++q) {
          if (q < p.n_chunks) {
            const int sr = (int)rem + (int)p.ch[q].row_off;
            const bool ok = pv && sr >= 0 && sr < (int)p.ch[q].rows_per_n;
            const uint32_t row = n * (uint32_t)p.ch[q].rows_per_n + (uint32_t)sr;
            const bf16* src = p.ch[q].base + (ok ? (size_t)row * (uint32_t)p.ch[q].pitch + cg * 8 : 0);
            cp_async16(sdst + (uint32_t)(q * 4 * WG_PT) * 16u, src, ok ? 16u : 0u);
          }
        }
        const bf16* gsrc = p.G + (pv ? (size_t)pp * (uint32_t)p.g_pitch + cg * 8 : 0);
        const uint32_t gdst = sg + (uint32_t)(cg * WG_PT + i) * 16u;
        cp_async16(gdst, gsrc, pv ? 16u : 0u);
        if (N > 32) cp_async16(gdst + (uint32_t)(4 * WG_PT) * 16u, gsrc + 32, pv ? 16u : 0u);
      }
      cp_async_commit();
      if (warp == 0) WG_TRACE(1);
      if (g >= WG_STAGES - 1) {
        cp_async_wait<WG_STAGES - 1>();
        fence_proxy_async();
        mbar_arrive(&full[(g - (WG_STAGES - 1)) % WG_STAGES]);
      }
      if (warp == 0) WG_TRACE(2);
      ++g;
    }
    // drain: groups g-2, g-1 (for 3 stages) are still pending
    cp_async_wait<0>();
    fence_proxy_async();
    for (int k = (g >= WG_STAGES - 1 ? g - (WG_STAGES - 1) : 0); k < g; ++k) mbar_arrive(&full[k % WG_STAGES]);
  } else if (warp == WG_MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, N, /*a_mn=*/true, /*b_mn=*/true);
      int g = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int stage = g % WG_STAGES;
        mbar_wait(&full[stage], (uint32_t)((g / WG_STAGES) & 1));
        WG_TRACE(3);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes), sg = sa + a_bytes;
        for (int t = 0; t < mt; ++t)
          for (int ks = 0; ks < WG_PT / 16; ++ks) {
            const uint64_t adesc = make_smem_desc(sa + (uint32_t)t * 16u * WG_PT * 16u + (uint32_t)ks * 256u, 128u,
                                                  WG_PT * 16u);
            const uint64_t bdesc = make_smem_desc(sg + (uint32_t)ks * 256u, 128u, WG_PT * 16u);
            umma_bf16(tmem_base + (uint32_t)(t * N), adesc, bdesc, idesc, (g == 0 && ks == 0) ? 0u : 1u);
          }
        umma_commit(&empty[stage]);
        WG_TRACE(4);
        ++g;
      }
      umma_commit(tfull);
    }
    __syncwarp();
  } else {
    // ===================== epilogue: one pass at the end =====================
    const int quad = warp & 3;
    const bool any = blockIdx.x < p.n_tiles;
    if (any) {
      mbar_wait(tfull, 0u);
      tc_fence_after();
      // pass 1: the ones row -> bias gradient (and into smem for the affine correction)
      const int ones_t = mrows / 128, ones_r = mrows % 128;
      if (quad == ones_r / 32) {
        for (int c0 = 0; c0 < N; c0 += 32) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ones_t * N + c0), v);
          if (lane == ones_r % 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              db_s[c0 + j] = v[j];
              if (p.db) atomicAdd(p.db + c0 + j, v[j]);
            }
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // epilogue warps only
      for (int t = 0; t < mt; ++t) {
        const int m = t * 128 + quad * 32 + lane;
        for (int c0 = 0; c0 < N; c0 += 32) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * N + c0), v);
          if (m < mrows) {
            float sc = 1.f, sh = 0.f;
            if (p.scale) { sc = __ldg(p.scale + (m & 31)); sh = __ldg(p.shift + (m & 31)); }
            float* dst = p.dW + (long long)m * p.ldw + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + j, fmaf(sc, v[j], sh * db_s[c0 + j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WG_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------ dadj
constexpr int DJ_SLABS = 4;          // slabs per stage
constexpr int DJ_STAGES = 3;

__global__ void __launch_bounds__(WG_THREADS, 1) dadj_tc_kernel(const __grid_constant__ DadjParams p) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = p.V;
  const int Np = ((V + 15) / 16) * 16;                   // MMA N (w), multiple of 16
  // per slab: X image [4 kc][128 rows][16 B] = 8 KB, G image [4 kc][Np rows][16 B]
  const uint32_t x_bytes = 4u * 128u * 16u, g_bytes = 4u * (uint32_t)Np * 16u;
  const uint32_t slab_bytes = x_bytes + g_bytes;
  const uint32_t stage_bytes = DJ_SLABS * slab_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)DJ_STAGES * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + DJ_STAGES;
  uint64_t* tfull = bars + 2 * DJ_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * DJ_STAGES + 1);

  if (tid == 0) {
    for (int i = 0; i < DJ_STAGES; ++i) { mbar_init(&full[i], WG_PRODUCERS); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  const uint32_t tmem_cols = Np <= 32 ? 32u : Np <= 64 ? 64u : 128u;
  if (warp == WG_MMA_WARP) tmem_alloc(tmem_slot, tmem_cols);
  {  // zero everything once: rows >= V of every image must stay zero
    uint4* a = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < (int)(DJ_STAGES * stage_bytes / 16); i += WG_THREADS) a[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int steps_total_per_tile = p.n_terms;   // each (tile, term) is one pipeline stage fill

  if (warp < WG_MMA_WARP) {
    const int kc = tid & 3, r0 = tid >> 2;   // 16-byte K piece kc of rows r0, r0+32, ...
    int g = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      for (int term = 0; term < steps_total_per_tile; ++term) {
        const int stage = g % DJ_STAGES, phase = (g / DJ_STAGES) & 1;
        mbar_wait(&empty[stage], (uint32_t)(phase ^ 1));
        const uint32_t sbase = smem_u32(smem + (size_t)stage * stage_bytes);
        const bf16* X = p.t[term].X;
        const bf16* G = p.t[term].G;
        for (int sl = 0; sl < DJ_SLABS; ++sl) {
          const long long slab = (long long)tile * DJ_SLABS + sl;
          const bool sok = slab < p.slabs;
          const uint32_t sx = sbase + (uint32_t)sl * slab_bytes, sg = sx + x_bytes;
          for (int v = r0; v < V; v += WG_PRODUCERS / 4) {
            const long long off = (slab * V + v) * 32 + kc * 8;
            cp_async16(sx + (uint32_t)(kc * 128 + v) * 16u, sok ? X + off : X, sok ? 16u : 0u);
            cp_async16(sg + (uint32_t)(kc * Np + v) * 16u, sok ? G + off : G, sok ? 16u : 0u);
          }
        }
        cp_async_commit();
        if (g >= DJ_STAGES - 1) {
          cp_async_wait<DJ_STAGES - 1>();
          fence_proxy_async();
          mbar_arrive(&full[(g - (DJ_STAGES - 1)) % DJ_STAGES]);
        }
        ++g;
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int k = (g >= DJ_STAGES - 1 ? g - (DJ_STAGES - 1) : 0); k < g; ++k) mbar_arrive(&full[k % DJ_STAGES]);
  } else if (warp == WG_MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, Np, /*a_mn=*/false, /*b_mn=*/false);
      int g = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int term = 0; term < steps_total_per_tile; ++term) {
          const int stage = g % DJ_STAGES;
          mbar_wait(&full[stage], (uint32_t)((g / DJ_STAGES) & 1));
          tc_fence_after();
          const uint32_t sbase = smem_u32(smem + (size_t)stage * stage_bytes);
          for (int sl = 0; sl < DJ_SLABS; ++sl) {
            const uint32_t sx = sbase + (uint32_t)sl * slab_bytes, sg = sx + x_bytes;
            for (int ks = 0; ks < 2; ++ks) {
              // K-major: the two 16-byte K pieces of a K-step are LBO apart, 8-row groups SBO = 128 B apart
              const uint64_t adesc = make_smem_desc(sx + (uint32_t)ks * 2u * 128u * 16u, 128u * 16u, 128u);
              const uint64_t bdesc = make_smem_desc(sg + (uint32_t)ks * 2u * (uint32_t)Np * 16u, (uint32_t)Np * 16u, 128u);
              umma_bf16(tmem_base, adesc, bdesc, idesc, (g == 0 && sl == 0 && ks == 0) ? 0u : 1u);
            }
          }
          umma_commit(&empty[stage]);
          ++g;
        }
      }
      umma_commit(tfull);
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;
    if (blockIdx.x < p.n_tiles) {
      mbar_wait(tfull, 0u);
      tc_fence_after();
      const int v = quad * 32 + lane;
      for (int c0 = 0; c0 < Np; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0)
            : "memory");
        tmem_ld_wait();
        if (v < V) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < V) atomicAdd(p.dA + (long long)v * V + c0 + j, __uint_as_float(r[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WG_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

static int g_sms = 0;
static int sm_count() {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(dadj_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  return g_sms;
}

int wgrad_tc_supported(int n_chunks, int N) {
  if (n_chunks < 1 || n_chunks > WG_MAX_CHUNKS || (N != 32 && N != 64)) return 0;
  int mt = (32 * n_chunks + 1 + 127) / 128;
  size_t stage = (size_t)mt * 16 * WG_PT * 16 + (size_t)(N / 8) * WG_PT * 16;
  return (mt * N <= 256 && WG_STAGES * stage + 512 <= 227 * 1024) ? 1 : 0;
}

int launch_wgrad_tc(WgParams& p, cudaStream_t st) {
  if (p.P <= 0) return 0;
  GWN_REQUIRE(wgrad_tc_supported(p.n_chunks, p.N), "wgrad_tc: unsupported shape (chunks=%d, N=%d)", p.n_chunks, p.N);
  GWN_REQUIRE(p.P < (1ll << 31), "wgrad_tc: too many positions");
  p.n_tiles = (int)cdiv(p.P, WG_PT);
  {
    const char* e = getenv("GWN_WG_TRACE");
    p.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr;
  }
  int mt = (32 * p.n_chunks + 1 + 127) / 128;
  size_t stage = (size_t)mt * 16 * WG_PT * 16 + (size_t)(p.N / 8) * WG_PT * 16;
  size_t smem = WG_STAGES * stage + 512;
  int sms = sm_count();
  int grid = p.n_tiles < sms ? p.n_tiles : sms;
  wgrad_tc_kernel<<<grid, WG_THREADS, smem, st>>>(p);
  GWN_LAUNCHED();
  return 0;
}

int launch_dadj_tc(DadjParams& p, cudaStream_t st) {
  if (p.slabs <= 0) return 0;
  GWN_REQUIRE(p.V <= 128 && p.n_terms >= 1 && p.n_terms <= 4, "dadj_tc: V=%d unsupported", p.V);
  p.n_tiles = (int)cdiv(p.slabs, DJ_SLABS);
  int Np = ((p.V + 15) / 16) * 16;
  size_t stage = (size_t)DJ_SLABS * (4 * 128 * 16 + 4 * Np * 16);
  size_t smem = DJ_STAGES * stage + 256;
  GWN_REQUIRE(smem <= 227 * 1024, "dadj_tc: smem");
  int sms = sm_count();
  int grid = p.n_tiles < sms ? p.n_tiles : sms;
  dadj_tc_kernel<<<grid, WG_THREADS, smem, st>>>(p);
  GWN_LAUNCHED();
  return 0;
}

}  // namespace gwn
