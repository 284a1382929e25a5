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
#include "tma_gemm.cuh"

namespace gwn {

constexpr int WG_PT = 128;           // positions per K tile (8 MMA K-steps): 8 KB per TMA box
constexpr int WG_MAX_STAGES = 4;
constexpr int WG_MMA_WARP = 4;
constexpr int WG_THREADS = 32 * 9;   // warp 0 TMA producer, warp 4 MMA, warps 5-8 epilogue
constexpr uint32_t WG_ATOM = WG_PT * 64u;     // one 32-channel operand atom: [128 positions][64 B], 64B-swizzled

struct WgMaps { CUtensorMap a[WG_MAX_CHUNKS]; CUtensorMap g[2]; };

#define WG_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && g < 48 && (tid & 31) == 0) p.trace[g * 8 + (slot)] = clock64(); } while (0)

// Operands arrive by TMA: every 32-channel group of A (a temporal tap / concat slot) and of G is one box
// {32 ch, 128 rows, 1 sample} -> [128][64 B] 64B-swizzled = one MN-major SWIZZLE_64B atom (M or N = channel,
// K = position); rows outside the sample are zero-filled by TMA, so they add nothing to the sums.
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgMaps maps,
                                                                 const __grid_constant__ WgParams p, int WG_STAGES) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#define WG_MARK(i) do { if (p.trace && blockIdx.x == 0 && tid == (i == 2 || i == 3 ? 32 * (WG_MMA_WARP + 1) : 0)) p.trace[47 * 8 + (i)] = clock64(); } while (0)
  WG_MARK(0);
  const int mrows = 32 * p.n_chunks;                 // real weight rows; row `mrows` is the ones row
  const int mt = (mrows + 1 + 127) / 128;            // M tiles
  const int N = p.N;
  const uint32_t a_bytes = (uint32_t)mt * 4u * WG_ATOM;           // 4 atoms per M tile
  const uint32_t g_bytes = (uint32_t)(N / 32) * WG_ATOM;
  const uint32_t stage_bytes = a_bytes + g_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WG_STAGES * stage_bytes);
  uint64_t* full = bars;                     // [STAGES]
  uint64_t* empty = bars + WG_MAX_STAGES;    // [STAGES]
  uint64_t* tfull = bars + 2 * WG_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_MAX_STAGES + 1);
  float* db_s = reinterpret_cast<float*>(bars + 2 * WG_MAX_STAGES + 2);   // [N]

  if (tid == 0) {
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  const uint32_t tmem_cols = (mt * N <= 32) ? 32u : (mt * N <= 64) ? 64u : (mt * N <= 128) ? 128u : 256u;
  if (warp == WG_MMA_WARP) tmem_alloc(tmem_slot, tmem_cols);
  {  // atoms TMA never writes (the ones atom and the padding atoms of the last M tile): zero, then plant the ones row
    for (int s = 0; s < WG_STAGES; ++s) {
      uint4* a = reinterpret_cast<uint4*>(smem + (size_t)s * stage_bytes + (size_t)p.n_chunks * WG_ATOM);
      const int n16 = (int)((a_bytes - (uint32_t)p.n_chunks * WG_ATOM) / 16);
      for (int i = tid; i < n16; i += WG_THREADS) a[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    // ones atom = atom index n_chunks: element (position k, channel 0) = 1.0; 64B swizzle puts logical 16-byte chunk 0
    // of row k at physical chunk ((k >> 1) & 3)
    for (int s = 0; s < WG_STAGES; ++s) {
      uint8_t* atom = smem + (size_t)s * stage_bytes + (size_t)p.n_chunks * WG_ATOM;
      for (int k = tid; k < WG_PT; k += WG_THREADS)
        *reinterpret_cast<uint16_t*>(atom + k * 64 + ((k >> 1) & 3) * 16) = 0x3F80u;   // bf16 1.0
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  WG_MARK(1);

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
    {
      int g = 0, stage = 0, phase = 0;
      const uint32_t tx = (uint32_t)p.n_chunks * WG_ATOM + g_bytes;
      int n = (int)blockIdx.x / p.tiles_per_n, rt = (int)blockIdx.x - n * p.tiles_per_n;
      const int dn = (int)gridDim.x / p.tiles_per_n, dr = (int)gridDim.x - dn * p.tiles_per_n;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++g) {
        const int r0 = rt * WG_PT;
        mbar_wait(&empty[stage], (uint32_t)(phase ^ 1));
        WG_TRACE(0);
        if (elect_one()) {
          tg::mbar_expect_tx(&full[stage], tx);
          const uint32_t sa = base + (uint32_t)stage * stage_bytes, sg = sa + a_bytes;
          for (int q = 0; q < p.n_chunks; ++q)
            tg::tma_3d(sa + (uint32_t)q * WG_ATOM, &maps.a[p.map_of[q]], 0, r0 + p.row_off[q], n, &full[stage]);
          for (int j = 0; j < N / 32; ++j) tg::tma_3d(sg + (uint32_t)j * WG_ATOM, &maps.g[j], 0, r0, n, &full[stage]);
        }
        __syncwarp();
        WG_TRACE(1);
        n += dn; rt += dr;
        if (rt >= p.tiles_per_n) { rt -= p.tiles_per_n; ++n; }
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == WG_MMA_WARP) {
    {
      const uint32_t idesc = make_idesc_bf16(128, N, /*a_mn=*/true, /*b_mn=*/true);
      // MN-major SW64: LBO = next 32-channel atom, SBO = next 8 positions, K=16 step = 1024 B
      const uint64_t dt = tg::make_desc_sw(0, WG_ATOM, 512u, 4u);
      int g = 0, stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        mbar_wait(&full[stage], (uint32_t)phase);
        WG_TRACE(3);
        tc_fence_after();
        const uint32_t sa = base + (uint32_t)stage * stage_bytes, sg = sa + a_bytes;
        if (elect_one()) {
          for (int t = 0; t < mt; ++t)
#pragma unroll
            for (int ks = 0; ks < WG_PT / 16; ++ks) {
              const uint64_t adesc = dt + (uint64_t)((sa + (uint32_t)t * 4u * WG_ATOM + (uint32_t)ks * 1024u) >> 4);
              const uint64_t bdesc = dt + (uint64_t)((sg + (uint32_t)ks * 1024u) >> 4);
              umma_bf16(tmem_base + (uint32_t)(t * N), adesc, bdesc, idesc, (g == 0 && ks == 0) ? 0u : 1u);
            }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        WG_TRACE(4);
        ++g;
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(tfull);
    }
    __syncwarp();
  } else if (warp > WG_MMA_WARP) {
    // ===================== epilogue: one pass at the end =====================
    const int quad = warp & 3;
    const bool any = blockIdx.x < p.n_tiles;
    if (any) {
      mbar_wait(tfull, 0u);
      tc_fence_after();
      WG_MARK(2);
      // The CTA's partial dW (and db) is staged in shared memory (the TMA stages are free once tfull has fired) and
      // flushed with rotated, coalesced vector reductions (red_flush_2d): the direct per-row scalar atomics of 148 CTAs
      // finishing together serialised in L2 and cost ~17 us per launch.
      float* stg = reinterpret_cast<float*>(smem);           // [mrows][N + 4] fp32, then [N] for db
      const int sld = N + 4;
      const int et = tid - 32 * (WG_MMA_WARP + 1);           // 0..127 among the epilogue threads
      // pass 1: the ones row -> bias gradient (into smem for the affine correction)
      const int ones_t = mrows / 128, ones_r = mrows % 128;
      if (quad == ones_r / 32) {
        for (int c0 = 0; c0 < N; c0 += 32) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ones_t * N + c0), v);
          if (lane == ones_r % 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) db_s[c0 + j] = v[j];
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
            float* dst = stg + (size_t)m * sld + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) =
                  make_float4(fmaf(sc, v[j], sh * db_s[c0 + j]), fmaf(sc, v[j + 1], sh * db_s[c0 + j + 1]),
                              fmaf(sc, v[j + 2], sh * db_s[c0 + j + 2]), fmaf(sc, v[j + 3], sh * db_s[c0 + j + 3]));
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      red_flush_2d(p.dW, p.ldw, stg, sld, mrows, N, et, 128);
      if (p.db) red_flush_1d(p.db, db_s, N, et, 128);
      WG_MARK(3);
    }
  }
  tc_fence_before();
  __syncthreads();
  WG_MARK(4);
  if (warp == WG_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------ dadj
constexpr int DJ_SLABS = 4;          // slabs per stage
constexpr int DJ_STAGES = 3;

struct DjMaps { CUtensorMap x[4]; CUtensorMap g[4]; };

// Both operands K-major (channels contiguous): per stage two TMA boxes {32 ch, rows, 4 slabs} -> [4][rows][64 B],
// 64B-swizzled; nodes >= V and slabs past the end are zero-filled by TMA.
__global__ void __launch_bounds__(WG_THREADS, 1) dadj_tc_kernel(const __grid_constant__ DjMaps maps,
                                                                const __grid_constant__ DadjParams p) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = p.V;
  const int Np = ((V + 15) / 16) * 16;                   // MMA N (w), multiple of 16
  const uint32_t x_slab = 128u * 64u, g_slab = (uint32_t)Np * 64u;
  const uint32_t x_bytes = DJ_SLABS * x_slab, g_bytes = (DJ_SLABS * g_slab + 1023u) & ~1023u;
  const uint32_t stage_bytes = x_bytes + g_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)DJ_STAGES * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + DJ_STAGES;
  uint64_t* tfull = bars + 2 * DJ_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * DJ_STAGES + 1);

  if (tid == 0) {
    for (int i = 0; i < DJ_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  const uint32_t tmem_cols = Np <= 32 ? 32u : Np <= 64 ? 64u : 128u;
  if (warp == WG_MMA_WARP) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int g = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int term = 0; term < p.n_terms; ++term, ++g) {
          const int stage = g % DJ_STAGES, phase = (g / DJ_STAGES) & 1;
          mbar_wait(&empty[stage], (uint32_t)(phase ^ 1));
          tg::mbar_expect_tx(&full[stage], x_bytes + DJ_SLABS * g_slab);
          const uint32_t sx = base + (uint32_t)stage * stage_bytes;
          tg::tma_3d(sx, &maps.x[term], 0, 0, tile * DJ_SLABS, &full[stage]);
          tg::tma_3d(sx + x_bytes, &maps.g[term], 0, 0, tile * DJ_SLABS, &full[stage]);
        }
      }
    }
    __syncwarp();
  } else if (warp == WG_MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, Np, /*a_mn=*/false, /*b_mn=*/false);
      int g = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int term = 0; term < p.n_terms; ++term, ++g) {
          const int stage = g % DJ_STAGES;
          mbar_wait(&full[stage], (uint32_t)((g / DJ_STAGES) & 1));
          tc_fence_after();
          const uint32_t sx = base + (uint32_t)stage * stage_bytes, sg = sx + x_bytes;
#pragma unroll
          for (int sl = 0; sl < DJ_SLABS; ++sl)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t adesc = tg::make_desc_sw(sx + (uint32_t)sl * x_slab + (uint32_t)ks * 32u, 16u, 512u, 4u);
              const uint64_t bdesc = tg::make_desc_sw(sg + (uint32_t)sl * g_slab + (uint32_t)ks * 32u, 16u, 512u, 4u);
              umma_bf16(tmem_base, adesc, bdesc, idesc, (g == 0 && sl == 0 && ks == 0) ? 0u : 1u);
            }
          umma_commit(&empty[stage]);
        }
      }
      umma_commit(tfull);
    }
    __syncwarp();
  } else if (warp > WG_MMA_WARP) {
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
  size_t stage = (size_t)mt * 4 * WG_ATOM + (size_t)(N / 32) * WG_ATOM;
  return (mt * N <= 256 && 2 * stage + 2048 <= 224 * 1024) ? 1 : 0;
}

int launch_wgrad_tc(WgParams& p, cudaStream_t st) {
  if (p.P <= 0) return 0;
  GWN_REQUIRE(wgrad_tc_supported(p.n_chunks, p.N), "wgrad_tc: unsupported shape (chunks=%d, N=%d)", p.n_chunks, p.N);
  GWN_REQUIRE(p.P < (1ll << 31), "wgrad_tc: too many positions");
  {
    p.trace = trace_ptr("GWN_WG_TRACE");
  }
  bool flat = true;    // every chunk has the output's own row structure: tile the flat position axis
  for (int q = 0; q < p.n_chunks; ++q)
    if (p.ch[q].rows_per_n != p.rows_per_n_out || p.ch[q].row_off != 0) flat = false;
  const long long n_real = p.P / p.rows_per_n_out;
  GWN_REQUIRE(n_real * p.rows_per_n_out == p.P, "wgrad_tc: P is not a whole number of samples");
  const long long rows_out = flat ? p.P : p.rows_per_n_out;
  p.n_samples = flat ? 1 : (int)n_real;
  p.tiles_per_n = (int)cdiv(rows_out, WG_PT);
  p.n_tiles = p.n_samples * p.tiles_per_n;
  WgMaps maps;
  int n_maps = 0;
  struct Key { const bf16* base; long long rows; int pitch; } keys[WG_MAX_CHUNKS];
  for (int q = 0; q < p.n_chunks; ++q) {
    const WgChunk& c = p.ch[q];
    const long long rows = flat ? p.P : c.rows_per_n;
    int m = -1;
    for (int i = 0; i < n_maps; ++i)
      if (keys[i].base == c.base && keys[i].rows == rows && keys[i].pitch == c.pitch) m = i;
    if (m < 0) {
      m = n_maps++;
      keys[m] = Key{c.base, rows, c.pitch};
      if (int rc = tg_map_rows3d(&maps.a[m], c.base, (uint64_t)rows, (uint64_t)(flat ? 1 : n_real), (uint64_t)c.pitch, WG_PT))
        return rc;
    }
    p.map_of[q] = m;
    GWN_REQUIRE(c.row_off > -(1ll << 30) && c.row_off < (1ll << 30), "wgrad_tc: row offset out of range");
    p.row_off[q] = (int)c.row_off;
  }
  for (int i = n_maps; i < WG_MAX_CHUNKS; ++i) maps.a[i] = maps.a[0];
  for (int j = 0; j < 2; ++j) {
    const int jj = j < p.N / 32 ? j : 0;
    if (int rc = tg_map_rows3d(&maps.g[j], p.G + 32 * jj, (uint64_t)rows_out, (uint64_t)(flat ? 1 : n_real),
                               (uint64_t)p.g_pitch, WG_PT))
      return rc;
  }
  int mt = (32 * p.n_chunks + 1 + 127) / 128;
  size_t stage = (size_t)mt * 4 * WG_ATOM + (size_t)(p.N / 32) * WG_ATOM;
  int stages = (int)((224 * 1024 - 2048) / stage);
  if (stages > WG_MAX_STAGES) stages = WG_MAX_STAGES;
  size_t smem = stages * stage + 1024 + 1024;
  int sms = sm_count();
  int grid = p.n_tiles < sms ? p.n_tiles : sms;
  wgrad_tc_kernel<<<grid, WG_THREADS, smem, st>>>(maps, p, stages);
  GWN_LAUNCHED();
  return 0;
}

int launch_dadj_tc(DadjParams& p, cudaStream_t st) {
  if (p.slabs <= 0) return 0;
  GWN_REQUIRE(p.V <= 128 && p.n_terms >= 1 && p.n_terms <= 4, "dadj_tc: V=%d unsupported", p.V);
  p.n_tiles = (int)cdiv(p.slabs, DJ_SLABS);
  const int Np = ((p.V + 15) / 16) * 16;
  const size_t x_bytes = (size_t)DJ_SLABS * 128 * 64, g_bytes = ((size_t)DJ_SLABS * Np * 64 + 1023) & ~(size_t)1023;
  const size_t smem = DJ_STAGES * (x_bytes + g_bytes) + 1024 + 256;
  GWN_REQUIRE(smem <= 227 * 1024, "dadj_tc: smem");
  DjMaps maps;
  for (int t = 0; t < 4; ++t) {
    const int tt = t < p.n_terms ? t : 0;
    if (int rc = tg_map_rows3d_box(&maps.x[t], p.t[tt].X, (uint64_t)p.V, (uint64_t)p.slabs, 32, 128, DJ_SLABS)) return rc;
    if (int rc = tg_map_rows3d_box(&maps.g[t], p.t[tt].G, (uint64_t)p.V, (uint64_t)p.slabs, 32, (uint32_t)Np, DJ_SLABS)) return rc;
  }
  int sms = sm_count();
  int grid = p.n_tiles < sms ? p.n_tiles : sms;
  dadj_tc_kernel<<<grid, WG_THREADS, smem, st>>>(maps, p);
  GWN_LAUNCHED();
  return 0;
}

}  // namespace gwn
