// Kernel template of the tcgen05 position GEMM (see tc_gemm.cuh).  Included where the epilogue functors live.
//
// Epilogue functor contract (one thread = one output row = one TMEM lane):
//   __device__ void chunk(long long p, long long n, long long rem, bool valid, int c0, float v[32]);
//        32 consecutive output columns [c0, c0+32) of row p (valid == p < P); called by every lane so the
//        functor may use warp collectives (the BN-statistics reduction does).
//   __device__ void finish(float* red_s);   called once per epilogue warp after its last tile: add the warp's statistics
//        into the CTA's 64-float shared scratch (zeroed by the kernel)
//   __device__ void flush(const float* red_s, int lane);   called by ONE warp after the CTA-wide barrier: one global
//        atomic per statistic per CTA (per-warp global atomics put 16 x 148 serialised updates on each address)
#pragma once
#include <cstdlib>
#include <type_traits>

#include "tc.cuh"
#include "tc_gemm.cuh"
#include "tma_gemm.cuh"

namespace gwn {

// Optional fused weight gradient (Epi::kWgrad, used by the gate data-gradient GEMM): the tile's chunk boxes (dfg of tap j,
// half h: [128 rows][32 cols]) and its second "extra" box (u_prev rows of the same tile) are exactly the operands of
//     dW[(j, c), n] += sum_rows u_prev[row][c] * dfg_j[row][n]              (graph_wavenet.py:222-224 weight gradient)
// so the MMA warp issues, per chunk, 8 more MMAs (K = 128 rows) with both boxes viewed MN-major (SWIZZLE_64B atoms) into a
// 128-column accumulator that stays in TMEM for the CTA's whole share; a resident "ones" atom behind the u_prev box adds
// the bias gradient as accumulator row 32.  The CTA's partial is flushed once at the end (staged, rotated, vectorised).
// warps: 0 and 2 = TMA producers (one thread each, on different schedulers; they split a tile's boxes),
// 1 = MMA issuer, 3 idle, 4-19 = epilogue (four per TMEM lane quadrant = per scheduler)
constexpr int PGT_PROD_A = 0, PGT_MMA_WARP = 1, PGT_PROD_B = 2;
constexpr int PGT_EPI_WARP0 = 4;
constexpr int PGT_EPI_WARPS = 16;
constexpr int PGT_EPI_RANKS = PGT_EPI_WARPS / 4;
constexpr int PGT_THREADS = 32 * (PGT_EPI_WARP0 + PGT_EPI_WARPS);
constexpr uint32_t PGT_ONES_BYTES = 4096;   // resident "ones" A tile of the bias MMA: [2 K pieces][128 rows][16 B]

// lane c of the warp ends up with sum over the 32 lanes of v[c]  (31 shuffles; v is destroyed)
__device__ __forceinline__ float warp_column_sums(float v[32], int lane) {
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool up = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < step; ++i) {
      const float keep = up ? v[i + step] : v[i];
      const float send = up ? v[i] : v[i + step];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return v[0];
}

// 16 columns: lanes 2c and 2c+1 both end up with the sum over the 32 lanes of v[c]  (16 shuffles; v is destroyed)
__device__ __forceinline__ float warp_column_sums16(float v[16], int lane) {
#pragma unroll
  for (int step = 16; step >= 2; step >>= 1) {
    const bool up = (lane & step) != 0;
    const int half = step >> 1;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = up ? v[i + half] : v[i];
      const float send = up ? v[i] : v[i + half];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

struct PgMaps { CUtensorMap m[PG_TC_MAX_MAPS]; };

// Optional staged outputs (Epi::kTmaOut = K > 0, used by the gated conv): the epilogue threads write their bf16 output rows
// into 64B-swizzled [128 rows][64 B] shared-memory tiles (conflict-free 16-byte stores) and the otherwise idle warp 3
// sends each tile to global memory with ONE bulk tensor store per output (rows past the sample's end are clipped by the
// tensor map) - instead of 32-row x 16-byte scattered st.global per warp instruction.  Two staging slots.
//   host:   int n_out() const;  bf16* out_ptr(int i) const;          (outputs actually written by this launch, <= K)
//   device: void chunk_st(p, n, rem, valid, c0, v, slot_smem, tile_row, sempty_bar, parity);   tile i of output o lives at
//           slot_smem + o * 8192; the functor waits for (sempty_bar, parity) before its first shared-memory store
template <typename E, typename = void> struct epi_tma_out { static constexpr int value = 0; };
template <typename E> struct epi_tma_out<E, std::void_t<decltype(E::kTmaOut)>> { static constexpr int value = E::kTmaOut; };
constexpr int PGT_STORE_WARP = 3;
constexpr int PGT_OUT_SLOTS = 2;
constexpr int PGT_OUT_MAP0 = PG_TC_MAX_MAPS - 3;       // output maps live in the last three PgMaps entries

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
#ifdef GWN_TRACE     // clock64 timeline of CTA 0, 16 slots per tile (scripts/gpu_gate_trace.py); compiled out of release builds
#define PG_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && g < 47) p.trace[g * 16 + (slot)] = clock64(); } while (0)
#define PG_MARK(k) do { if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[47 * 16 + (k)] = clock64(); } while (0)   // kernel phases
#else
#define PG_TRACE(slot) do { } while (0)
#define PG_MARK(k) do { } while (0)
#endif

// (sample, 128-row tile inside the sample) of a CTA's current macro tile, advanced without divisions: a 32-bit
// division costs a few hundred cycles of latency on the rarely-scheduled producer / per-item epilogue paths
struct PgWalk {
  int n, rt, dn, dr, tpn;
  __device__ __forceinline__ void init(int first_sub, int step_sub, int tiles_per_n) {
    tpn = tiles_per_n;
    n = first_sub / tpn; rt = first_sub - n * tpn;
    dn = step_sub / tpn; dr = step_sub - dn * tpn;
  }
  __device__ __forceinline__ void advance() {
    n += dn; rt += dr;
    if (rt >= tpn) { rt -= tpn; ++n; }
  }
  __device__ __forceinline__ void sub(int t, int& ns, int& r0) const {   // sub-tile t of the macro tile
    ns = n; int r = rt + t;
    while (r >= tpn) { r -= tpn; ++ns; }
    r0 = r * 128;
  }
};

// A operand: per chunk one TMA box -> [128 rows][64 B] 64B-swizzled (K-major SW64: 8-row atoms of 512 B, the two
// K=16 halves of a chunk 32 B apart); B operand: resident no-swizzle weight image; D: 2 or 4 TMEM accumulators.
// NCH > 0: the chunk count is a compile-time constant (the MMA issue loop unrolls to immediates).
template <typename Epi, int NCH>
__global__ void __launch_bounds__(PGT_THREADS, 1) pos_gemm_tc_kernel(const __grid_constant__ PgMaps maps,
                                                                      const __grid_constant__ PgParams p, Epi epi,
                                                                      int stages, int n_acc) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  PG_MARK(0);
  const int N = p.N;
  const int n_chunks = NCH > 0 ? NCH : p.n_chunks;
  const int K8 = n_chunks * 4 + (p.has_bias ? 2 : 0);        // 16-byte K pieces per weight row
  const int SUB = p.sub;
  const int NB = n_chunks + p.n_extra;                         // TMA boxes per sub-tile (MMA chunks + epilogue extras)
  // one stage: SUB x NB boxes of 128 x 64 B (+ the resident ones atom of the fused weight gradient, SUB == 1 there)
  const uint32_t a_bytes = (uint32_t)NB * (uint32_t)SUB * 8192u;
  const uint32_t w_bytes = ((uint32_t)K8 * (uint32_t)N * 16u + 1023u) & ~1023u;
  uint8_t* a_s = smem;                                        // stages first (1024-aligned boxes)
  // (kWgrad: ONE 8 KB "ones" atom behind the ring, shared by every stage - the B descriptor's leading byte offset reaches it)
  constexpr int KOUT = epi_tma_out<Epi>::value;
  const int n_out = KOUT > 0 ? p.n_out : 0;                   // staged outputs of this launch (0: direct stores)
  constexpr uint32_t out_slot_bytes = (uint32_t)KOUT * 8192u;
  uint8_t* wones_s = smem + (size_t)stages * a_bytes;
  uint8_t* out_s = wones_s + (Epi::kWgrad ? 8192 : 0);           // [slots][KOUT][128 rows][64 B]
  uint8_t* w_s = out_s + (size_t)PGT_OUT_SLOTS * out_slot_bytes;
  uint8_t* ones_s = w_s + w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones_s + PGT_ONES_BYTES);
  uint64_t* full = bars;               // [stages] (<= 8)
  uint64_t* empty = bars + 8;
  uint64_t* tfull = bars + 16;         // [n_acc] (<= 4)
  uint64_t* tempty = bars + 20;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  uint64_t* wfull = bars + 25;         // fused weight gradient: all MMAs of the CTA completed
  uint64_t* sfull = bars + 26;         // [2] staged outputs: every epilogue warp of the tile has written its rows
  uint64_t* sempty = bars + 28;        // [2] the bulk stores of the slot have read it
  float* red_s = reinterpret_cast<float*>(bars + 32);        // [64] CTA-level statistics scratch
  (void)a_s;
  if (tid < 64) red_s[tid] = 0.f;
  // epilogue warps that take part in one macro tile: its SUB x ceil(N/32) items go to consecutive ranks of every lane
  // quadrant, so min(items, ranks) warps per quadrant have work; the others skip the tile (no wait, no arrival)
  const int epi_items = SUB * ((N + 31) / 32);
  const uint32_t epi_arrivals = 4u * (uint32_t)(epi_items < PGT_EPI_RANKS ? epi_items : PGT_EPI_RANKS);

  if (tid == 0) {
    // a stage is free when the MMAs have read it and (with extras) every epilogue warp with work in the tile is done with it
    // (ONE arrival per warp: 512 per-thread arrivals per tile serialise on the barrier word)
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], p.n_extra ? 1 + epi_arrivals : 1); }
    for (int i = 0; i < n_acc; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], epi_arrivals); }
    mbar_init(wfull, 1);
    for (int i = 0; i < PGT_OUT_SLOTS; ++i) { mbar_init(&sfull[i], epi_arrivals); mbar_init(&sempty[i], 1); }
    fence_barrier_init();
  }
  const uint32_t acc_cols = N <= 32 ? 32u : N <= 64 ? 64u : N <= 128 ? 128u : 256u;
  const uint32_t buf_cols = acc_cols * (uint32_t)SUB;          // one accumulator buffer: SUB sub-tiles side by side
  const uint32_t tmem_cols = Epi::kWgrad ? 256u : (uint32_t)n_acc * buf_cols;      // (kWgrad: 4 x 32 accumulators + 128 dW columns)
  if (warp == PGT_MMA_WARP) tmem_alloc(tmem_slot, tmem_cols);
  {
    // ones tile: element (row, k) = 1 for k < 2 (the two bias rows), else 0
    uint4* od = reinterpret_cast<uint4*>(ones_s);
    for (int i = tid; i < 256; i += PGT_THREADS) od[i] = i < 128 ? make_uint4(0x3F803F80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  }
  if constexpr (Epi::kWgrad) {
    // ones atom: element (position k, channel 0) = 1.0; 64B swizzle puts logical 16-byte chunk 0 of row k at physical
    // chunk ((k >> 1) & 3).
    {
      uint4* atom = reinterpret_cast<uint4*>(wones_s);
      for (int i = tid; i < 512; i += PGT_THREADS) atom[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    for (int k = tid; k < 128; k += PGT_THREADS)
      *reinterpret_cast<uint16_t*>(wones_s + k * 64 + ((k >> 1) & 3) * 16) = 0x3F80u;   // bf16 1.0
  }
  // Programmatic launch: barrier / TMEM / ones-atom set-up above overlapped the predecessor's tail.  The weight image below
  // is step-constant too UNLESS it folds the batch statistics the predecessor just produced (wsrc.bn == 1, training) or is
  // a prepared image (w_img): only then the wait comes first; otherwise it follows the image build.
  const bool pdl_early = p.wsrc.W == nullptr || p.wsrc.bn == 1;
  // 16-byte loads of the fp32 weights (a thread owns 4 consecutive output columns - or, transposed, 4 consecutive K rows -
  // of one row), all issued BEFORE the wait: the parameters are step-constant even when the fold is not
  constexpr int MAXIT = 4;                                            // K*N <= 4*4*PGT_THREADS elements (launcher checks)
  float4 wv[MAXIT];
  if (p.wsrc.W != nullptr) {
    const PgWsrc& ws = p.wsrc;
    const int N4 = N >> 2, total4 = (32 * n_chunks * N) >> 2;
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int i = tid + it * PGT_THREADS;
      wv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < total4) {
        if (ws.transposed) { const int k4 = i & 7, r = i >> 3, n = r % N, q = r / N; wv[it] = __ldg(reinterpret_cast<const float4*>(ws.W + ws.w_off[q] + n * ws.ld + 4 * k4)); }
        else { const int n4 = i % N4, k = i / N4; wv[it] = __ldg(reinterpret_cast<const float4*>(ws.W + ws.w_off[k >> 5] + (k & 31) * ws.ld + 4 * n4)); }
      }
    }
  }
  PG_MARK(1);
  if (pdl_early) { pdl_wait(); pdl_trigger(); }
  PG_MARK(2);
  // ===================== TMA producers (warps PROD_A / PROD_B; boxes of a macro tile dealt alternately to the two warps; the
  // whole warp walks the loop so addresses / coordinates stay in uniform registers, one elected lane issues).  A resumable
  // routine: the gated conv's prologue calls it for the first ring of stages BEFORE the weight image is built (the loads'
  // HBM round trip then runs under the BatchNorm fold), the role dispatch below for the rest. =====================
  const bool is_prod = warp == PGT_PROD_A || warp == PGT_PROD_B;
  PgWalk pw;
  int p_g = 0, p_stage = 0, p_phase = 0, p_tile = (int)blockIdx.x;
  if (is_prod) pw.init((int)blockIdx.x * SUB, (int)gridDim.x * SUB, p.tiles_per_n);
  auto produce = [&](int max_tiles) {
    const int me = warp == PGT_PROD_A ? 0 : 1;
    const int boxes = SUB * NB;
    const uint32_t my_bytes = (uint32_t)((boxes + 1 - me) >> 1) * 8192u;
    for (; p_tile < p.n_tiles && max_tiles > 0; p_tile += gridDim.x, ++p_g, --max_tiles) {
      const int g = p_g, stage = p_stage;
      (void)g;
      mbar_wait(&empty[stage], (uint32_t)(p_phase ^ 1));
      if (me == 0 && lane == 0) PG_TRACE(0);
      if (elect_one()) {
        if (my_bytes) tg::mbar_expect_tx(&full[stage], my_bytes); else mbar_arrive(&full[stage]);
        const uint32_t sa = base + (uint32_t)stage * a_bytes;
        int i = 0;
        for (int t = 0; t < SUB; ++t) {
          // sub-tile = 128 consecutive rows of ONE sample; past the last sub-tile the sample index is out of range
          // and TMA zero-fills the box
          int n, r0;
          pw.sub(t, n, r0);
          for (int q = 0; q < NB; ++q, ++i)
            if ((i & 1) == me)
              tg::tma_3d(sa + (uint32_t)i * 8192u, &maps.m[p.map_of[q]], 0, r0 + p.row_off[q], n, &full[stage]);
        }
      }
      __syncwarp();
      if (me == 0 && lane == 0) PG_TRACE(1);
      pw.advance();
      if (++p_stage == stages) { p_stage = 0; p_phase ^= 1; }
    }
  };
  // head start (staged-output epilogues only: their staging slots hold the fold's scratch, so the TMA stages are free)
  const bool head_start = KOUT > 0 && pdl_early && p.wsrc.W != nullptr;
  if (head_start) {
    __syncthreads();                                             // barriers initialised
    if (is_prod) produce(stages);
  }
  if (p.wsrc.W == nullptr) {
    const uint4* src = reinterpret_cast<const uint4*>(p.w_img);
    uint4* dst = reinterpret_cast<uint4*>(w_s);
    for (int i = tid; i < K8 * N; i += PGT_THREADS) dst[i] = __ldg(src + i);
  } else {
    // ---- weight image built here from the fp32 weights (tc_gemm.cuh: PgWsrc) ----
    const PgWsrc& ws = p.wsrc;
    float* sc_s = red_s + 64;            // [32] scale | [32] shift
    const int K = 32 * n_chunks;
    if (tid < 32) {
      float sc = 1.f, sh = 0.f;
      if (ws.bn == 2) {
        sc = ws.scale_in[tid]; sh = ws.shift_in[tid];
      } else if (ws.bn) {
        float mean, var;
        if (ws.bn_training) {
          const double m = ws.bn_stats[tid] / ws.bn_count;
          double v = ws.bn_stats[32 + tid] / ws.bn_count - m * m;
          if (v < 0) v = 0;
          mean = (float)m; var = (float)v;
          if (blockIdx.x == 0) {
            const double unbiased = ws.bn_count > 1 ? v * (ws.bn_count / (ws.bn_count - 1.0)) : v;
            ws.running_mean[tid] = (1.f - ws.momentum) * ws.running_mean[tid] + ws.momentum * mean;
            ws.running_var[tid] = (1.f - ws.momentum) * ws.running_var[tid] + ws.momentum * (float)unbiased;
          }
        } else {
          mean = ws.running_mean[tid]; var = ws.running_var[tid];
        }
        const float rstd = rsqrtf(var + ws.eps);
        sc = ws.gamma[tid] * rstd;
        sh = ws.beta[tid] - mean * sc;
        if (blockIdx.x == 0) {
          ws.scale_out[tid] = sc; ws.shift_out[tid] = sh;
          if (ws.mean_out) { ws.mean_out[tid] = mean; ws.rstd_out[tid] = rstd; }
        }
      }
      sc_s[tid] = sc; sc_s[32 + tid] = sh;
    }
    if (p.has_bias) {                    // zero the bias chunk (two K pieces)
      uint4* bz = reinterpret_cast<uint4*>(w_s) + (size_t)(K8 - 2) * N;
      for (int i = tid; i < 2 * N; i += PGT_THREADS) bz[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    bf16* img = reinterpret_cast<bf16*>(w_s);
    // The shift-into-bias partial sums go through a scratch array in the (still idle, contiguous) TMA stages and are added
    // in a FIXED order: every CTA must build bit-identical images (atomics would make a sample's result depend on which
    // CTA computed it).
    float* part_s = reinterpret_cast<float*>(head_start ? out_s : a_s);      // [PGT_THREADS][4]
    const int N4 = N >> 2, total4 = (K * N) >> 2;
    float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int i = tid + it * PGT_THREADS;
      if (i < total4) {
        const float raw[4] = {wv[it].x, wv[it].y, wv[it].z, wv[it].w};
        if (ws.transposed) {
          const int k4 = i & 7, r = i >> 3, n = r % N, q = r / N, k = q * 32 + 4 * k4;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(raw[0] * sc_s[4 * k4], raw[1] * sc_s[4 * k4 + 1]);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(raw[2] * sc_s[4 * k4 + 2], raw[3] * sc_s[4 * k4 + 3]);
          uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(img + ((size_t)(k >> 3) * N + n) * 8 + (k & 7)) = pk;
        } else {
          const int n4 = i % N4, k = i / N4, kk = k & 31;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = 4 * n4 + j;
            float val = raw[j] * sc_s[kk];
            if (ws.half_odd && (n & 1)) val *= 0.5f;
            img[((size_t)(k >> 3) * N + n) * 8 + (k & 7)] = __float2bfloat16_rn(val);
            part[j] = fmaf(sc_s[32 + kk], raw[j], part[j]);
          }
        }
      }
    }
    if (ws.bn && !ws.transposed) *reinterpret_cast<float4*>(part_s + 4 * tid) = make_float4(part[0], part[1], part[2], part[3]);
    __syncthreads();
    if (p.has_bias) {
      for (int n = tid; n < N; n += PGT_THREADS) {
        float b = ws.bias ? __ldg(ws.bias + n) : 0.f;
        if (ws.bn && !ws.transposed) {
          // threads t with t % N4 == n/4 hold the partials of column n (PGT_THREADS % N4 == 0: launcher checks)
          float acc = 0.f;
          for (int t = n >> 2; t < PGT_THREADS; t += N4) acc += part_s[4 * t + (n & 3)];
          b += acc;
        }
        if (ws.half_odd && (n & 1)) b *= 0.5f;
        const bf16 hi = __float2bfloat16_rn(b);
        const bf16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
        bf16* d0 = img + ((size_t)(K8 - 2) * N + n) * 8;
        d0[0] = hi; d0[1] = lo;
      }
    }
    __syncthreads();      // part_s lives in the first TMA stage: the producer may only start after the bias is read
  }
  if (!pdl_early) { pdl_wait(); pdl_trigger(); }
  // (global writes only after the wait: the zeroed statistics buffer is memory an earlier kernel may still be using)
  if (p.wsrc.W != nullptr && blockIdx.x == 0 && p.wsrc.zero64 && tid >= 64 && tid < 128) p.wsrc.zero64[tid - 64] = 0.0;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  PG_MARK(3);
  const uint32_t tmem_base = *tmem_slot;
  const int acc_mask = n_acc - 1, acc_shift = n_acc == 4 ? 2 : 1;

  if (is_prod) {
    produce(0x7fffffff);                                       // (the tiles not requested by the prologue's head start)
  } else if (warp == PGT_MMA_WARP) {
    // ===================== MMA issuer: the whole warp walks the loop (uniform registers), one elected lane issues =====================
    {
      const uint32_t idesc = make_idesc_bf16(128, N, false, false);
      const uint32_t w_addr = smem_u32(w_s);
      // descriptor templates: only the 14-bit start-address field changes per MMA (addresses < 256 KB: no carry)
      const uint64_t at = tg::make_desc_sw(0, 16u, 512u, 4u);
      const uint64_t bt = make_smem_desc(w_addr, (uint32_t)N * 16u, 128u);
      const uint64_t ones_d = make_smem_desc(smem_u32(ones_s), 2048u, 128u);
      const uint32_t bstep = ((uint32_t)N * 32u) >> 4;            // two K pieces of the weight image per K=16 step
      const uint64_t bias_d = bt + (uint64_t)(2u * bstep) * (uint32_t)n_chunks;
      const bool has_bias = p.has_bias != 0;
      int g = 0, stage = 0, sphase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int acc = g & acc_mask;
        mbar_wait(&tempty[acc], (uint32_t)(((g >> acc_shift) & 1) ^ 1));
        if (lane == 0) PG_TRACE(2);
        mbar_wait(&full[stage], (uint32_t)sphase);
        if (lane == 0) PG_TRACE(3);
        tc_fence_after();
        if (elect_one()) {
          for (int t = 0; t < SUB; ++t) {
            uint64_t ad = at + (uint64_t)((base + (uint32_t)stage * a_bytes + (uint32_t)(t * NB) * 8192u) >> 4);
            uint64_t bd = bt;
            const uint32_t d = tmem_base + (uint32_t)acc * buf_cols + (uint32_t)t * acc_cols;
            if (has_bias) umma_bf16(d, ones_d, bias_d, idesc, 0u);
            if constexpr (NCH > 0) {
#pragma unroll
              for (int q = 0; q < NCH; ++q) {
                umma_bf16(d, ad + (uint64_t)(q * 512), bd + (uint64_t)(2u * q) * bstep, idesc, (q == 0 && !has_bias) ? 0u : 1u);
                umma_bf16(d, ad + (uint64_t)(q * 512 + 2), bd + (uint64_t)(2u * q + 1u) * bstep, idesc, 1u);
              }
            } else {
#pragma unroll 2
              for (int q = 0; q < n_chunks; ++q) {
                umma_bf16(d, ad, bd, idesc, (q == 0 && !has_bias) ? 0u : 1u);
                umma_bf16(d, ad + 2u, bd + bstep, idesc, 1u);            // second K=16 half of the chunk: +32 B
                ad += 8192u >> 4;
                bd += 2u * bstep;
              }
            }
          }
          umma_commit(&tfull[acc]);
          if constexpr (Epi::kWgrad) {
            // dWt[(chunk q, n), c] += dfg_q^T u_prev over the tile's 128 rows - ONE chain of 8 MMAs (K = 16 rows each) for all four
            // chunks: A = the tile's four chunk boxes viewed MN-major (M = 4 x 32 dfg columns: SW64 atoms, LBO = next box, SBO =
            // next 8 rows, K step = 1024 B), B = the u_prev box and the ones atom behind it (N = 48: 32 channels, the ones
            // column = bias gradient, 15 zero columns).  The earlier orientation (M = channels padded to 128, N = 32 dfg
            // columns, one chain per chunk) issued 32 MMAs whose 4 KB A read (45 cycles each) made the MMA warp the
            // bottleneck of the kernel: 1,780 of 2,250 cycles per tile.
            const uint32_t sb = base + (uint32_t)stage * a_bytes;
            const uint64_t wt = tg::make_desc_sw(0, 8192u, 512u, 4u);
            const uint32_t idw = make_idesc_bf16(128, 48, true, true);
            const uint64_t aw = wt + (uint64_t)(sb >> 4);                                       // chunk boxes 0..3
            const uint32_t ub = sb + (uint32_t)(NB - 1) * 8192u;                               // u_prev box; next MN atom = the ones atom
            const uint64_t bw = tg::make_desc_sw(ub, smem_u32(wones_s) - ub, 512u, 4u);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_bf16(tmem_base + 128u, aw + (uint64_t)(ks * 64), bw + (uint64_t)(ks * 64), idw, (g == 0 && ks == 0) ? 0u : 1u);
          }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (lane == 0) PG_TRACE(4);
        ++g;
        if (++stage == stages) { stage = 0; sphase ^= 1; }
      }
      if constexpr (Epi::kWgrad) {
        if (elect_one()) umma_commit(wfull);
        __syncwarp();
      }
    }
  } else if (warp == PGT_STORE_WARP) {
    // ===================== store warp: staged output tiles -> global memory, one bulk tensor store per output =====================
    if constexpr (KOUT > 0) {
      if (n_out > 0) {
        PgWalk w; w.init((int)blockIdx.x, (int)gridDim.x, p.tiles_per_n);      // (staged outputs: SUB == 1)
        const uint32_t so = smem_u32(out_s);
        int g = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++g) {
          const int slot = g & 1;
          mbar_wait_lazy(&sfull[slot], (uint32_t)((g >> 1) & 1));
          if (lane == 0) PG_TRACE(8);
          if (lane == 0) {                                      // (bulk groups belong to the issuing thread: always lane 0)
            int ns, r0;
            w.sub(0, ns, r0);
            for (int o = 0; o < n_out; ++o)
              tma_store_3d(&maps.m[PGT_OUT_MAP0 + o], so + (uint32_t)slot * out_slot_bytes + (uint32_t)o * 8192u, 0, r0, ns);
            bulk_commit();
            bulk_wait_read0();
            mbar_arrive(&sempty[slot]);
            PG_TRACE(9);
          }
          __syncwarp();
          w.advance();
        }
        if (lane == 0) bulk_wait0();
        __syncwarp();
      }
    }
  } else if (warp >= PGT_EPI_WARP0) {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int rank = (warp - PGT_EPI_WARP0) >> 2;               // 0..3: which of the quadrant's four warps
    const int n_c32 = (N + 31) / 32;                            // 32-column chunks per sub-tile
    const int IT = SUB * n_c32;                                 // work items (sub-tile, chunk) per macro tile
    const bool nc_pow2 = (n_c32 & (n_c32 - 1)) == 0;
    const int nc_shift = n_c32 >= 8 ? 3 : n_c32 >= 4 ? 2 : n_c32 >= 2 ? 1 : 0;
    PgWalk w; w.init((int)blockIdx.x * SUB, (int)gridDim.x * SUB, p.tiles_per_n);
    int g = 0, stage = 0;
    int first = rank;                                           // items are dealt round-robin ACROSS tiles, so that
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {   // narrow tiles (IT < 4) still use every warp
      if (first < IT) {                                         // this warp has work in the tile
        const int acc = g & acc_mask;
        const bool tr = quad == 0 && first == 0 && lane == 0;      // (trace builds: the warp that takes the tile's first item)
        (void)tr;
        mbar_wait_lazy(&tfull[acc], (uint32_t)((g >> acc_shift) & 1));
        if (tr) PG_TRACE(5);
        tc_fence_after();
        int item = first;
        for (; item < IT; item += PGT_EPI_RANKS) {
          int t, ci;
          if (nc_pow2) { t = item >> nc_shift; ci = item & (n_c32 - 1); } else { t = item / n_c32; ci = item - t * n_c32; }
          const int c0 = ci * 32;
          int ns, r0;
          w.sub(t, ns, r0);
          const int r = r0 + quad * 32 + lane;
          const bool pv = r < p.rows_out && ns < p.n_samples;
          // (sample, row) of the caller's view: virtual samples (position-wise GEMMs tile the flat position axis)
          long long pp = (long long)ns * p.rows_out + r, n = ns, rem = r;
          if (p.rows_out != (int)p.rows_per_n_out && pv) split_pos(pp, p.rows_per_n_out, n, rem);
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * buf_cols + (uint32_t)t * acc_cols + (uint32_t)c0, v);
          if (tr && item == 0) PG_TRACE(7);
          if constexpr (Epi::kExtra) {
            epi.chunk_ex(pp, n, rem, pv, c0, v, smem + (size_t)stage * a_bytes + (size_t)(t * NB + n_chunks) * 8192, quad * 32 + lane);
          } else if constexpr (KOUT > 0) {
            if (n_out > 0)
              epi.chunk_st(pp, n, rem, pv, c0, v, out_s + (size_t)(g & 1) * out_slot_bytes, quad * 32 + lane, &sempty[g & 1],
                           (uint32_t)(((g >> 1) & 1) ^ 1));
            else
              epi.chunk(pp, n, rem, pv, c0, v);
          } else {
            epi.chunk(pp, n, rem, pv, c0, v);
          }
        }
        first = item - IT;                                      // where this warp starts in the next tile
        if (tr) PG_TRACE(10);
        tc_fence_before();
        if constexpr (KOUT > 0) { if (n_out > 0) fence_proxy_async(); }     // staged rows -> visible to the bulk store
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&tempty[acc]);
          if (p.n_extra) mbar_arrive(&empty[stage]);
          if constexpr (KOUT > 0) { if (n_out > 0) mbar_arrive(&sfull[g & 1]); }
        }
        if (tr) PG_TRACE(6);
      } else {
        first -= IT;
      }
      ++g;
      w.advance();
      if (++stage == stages) stage = 0;
    }
    if constexpr (Epi::kWgrad) {
      // ---- flush of the fused weight gradient: lane n of quadrant q holds row (chunk q, column n): 32 channels and the ones
      // column (bias gradient).  One warp per quadrant stages it (BatchNorm fold applied) in the idle first TMA stage, then
      // every epilogue thread takes part in one rotated vector flush.
      asm volatile("bar.sync 4, 512;" ::: "memory");              // every epilogue warp is done with the stages' extra boxes
      float* stg = reinterpret_cast<float*>(a_s);                   // [(tap j) * 32 + c][64] fp32
      float* db_s = stg + 4096 * 2;                                 // [64] bias gradient (taps <= 4 -> <= 8192 floats above)
      if (rank == 0) {
        mbar_wait(wfull, 0u);
        tc_fence_after();
        float v[32], o[32];
        const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + 128u;
        tmem_ld32(ta + 32u, o);                                   // (columns 32..47 are live; 48..63 allocated, unused)
        const float db = o[0];
        tmem_ld32(ta, v);
        epi.wgrad_row_t(quad, lane, v, db, db_s, stg);
      }
      asm volatile("bar.sync 4, 512;" ::: "memory");
      epi.wgrad_flush(stg, db_s, n_chunks, (warp - PGT_EPI_WARP0) * 32 + lane, 32 * PGT_EPI_WARPS);
    }
    epi.finish(red_s);
  }
  tc_fence_before();
  __syncthreads();
  PG_MARK(4);
  if (warp == 0) epi.flush(red_s, lane);
  if (warp == PGT_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

template <typename Epi, int NCH>
static int pos_gemm_tc_run(const PgMaps& maps, const PgParams& p, const Epi& epi, int stages, int n_acc, int grid, size_t smem,
                           cudaStream_t st) {
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    GWN_CUDA(cudaFuncSetAttribute(pos_gemm_tc_kernel<Epi, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  GWN_CUDA(launch_pdl(pos_gemm_tc_kernel<Epi, NCH>, dim3(grid), dim3(PGT_THREADS), smem, st, maps, p, epi, stages, n_acc));
  GWN_LAUNCHED();
  return 0;
}

template <typename Epi>
int launch_pos_gemm_tc(PgParams& p, const Epi& epi, cudaStream_t st) {
  if (p.P <= 0) return 0;
  GWN_REQUIRE(p.P < (1ll << 31), "pos_gemm_tc: too many positions");
  GWN_REQUIRE(p.n_chunks >= 1 && p.n_chunks <= PG_TC_MAX_CHUNKS && p.N % 16 == 0 && p.N >= 16 && p.N <= 256,
              "pos_gemm_tc: unsupported shape (chunks=%d, N=%d)", p.n_chunks, p.N);
  // position-wise GEMM (every chunk has the output's own row structure): tile the flat position axis
  bool flat = true;
  const int NB = p.n_chunks + p.n_extra;
  GWN_REQUIRE(NB <= PG_TC_MAX_CHUNKS, "pos_gemm_tc: too many chunks");
  for (int q = 0; q < NB; ++q)
    if (p.ch[q].rows_per_n != p.rows_per_n_out || p.ch[q].row_off != 0) flat = false;
  const long long n_real = p.P / p.rows_per_n_out;
  GWN_REQUIRE(n_real * p.rows_per_n_out == p.P, "pos_gemm_tc: P is not a whole number of samples");
  p.rows_out = flat ? (int)p.P : (int)p.rows_per_n_out;
  p.n_samples = flat ? 1 : (int)n_real;
  const int sms = tg_sm_count();
  const int acc_c = p.N <= 32 ? 32 : p.N <= 64 ? 64 : p.N <= 128 ? 128 : 256;
  const size_t w_bytes = ((size_t)(p.n_chunks * 4 + (p.has_bias ? 2 : 0)) * p.N * 16 + 1023) & ~(size_t)1023;
  constexpr int KOUT = epi_tma_out<Epi>::value;
  int n_out = 0;
  if constexpr (KOUT > 0) n_out = epi.n_out();
  p.n_out = n_out;
  const size_t fixed = w_bytes + PGT_ONES_BYTES + 1024 + 1024 + (Epi::kWgrad ? 8192 : 0) +    // alignment slack + barriers / scratch
                       (size_t)PGT_OUT_SLOTS * KOUT * 8192;
  GWN_REQUIRE(fixed + 2 * (size_t)NB * 8192 <= 227 * 1024,
              "pos_gemm_tc: K=%d does not fit 2 stages in shared memory", 32 * p.n_chunks);
  if (Epi::kWgrad)
    GWN_REQUIRE(p.n_chunks == Epi::kFast && p.n_extra == 2 && p.N == 32 && p.n_chunks <= 4,
                "pos_gemm_tc: the fused weight gradient needs %d chunks of 32 columns and the (du, u_prev) extras", Epi::kFast);
  p.tiles_per_n = (int)cdiv(p.rows_out, 128);                     // 128-row sub-tiles per (virtual) sample
  const long long sub_tiles = (long long)p.n_samples * p.tiles_per_n;
  // macro tiles: several 128-row sub-tiles per pipeline step amortise the hand-offs; the choice also looks at the
  // wave quantisation of the persistent grid (tiles / (ceil(tiles / SMs) * SMs))
  int subs[3] = {1, 2, 4}, stg_of[3] = {0, 0, 0};
  double eff_of[3] = {-1, -1, -1}, best_eff = -1;
  for (int i = 0; i < 3; ++i) {
    const int sub = subs[i];
    if (2 * acc_c * sub > 512 || ((Epi::kWgrad || n_out > 0) && sub > 1)) continue;
    const size_t ab = (size_t)NB * 8192 * sub;
    int stg = (int)((227 * 1024 - fixed) / ab);
    if (stg > 8) stg = 8;
    if (stg < (sub == 1 ? 2 : 3)) continue;
    const long long nt = cdiv(sub_tiles, sub);
    const long long waves = cdiv(nt, sms);
    stg_of[i] = stg; eff_of[i] = (double)sub_tiles / (double)(waves * sms * sub);
    if (eff_of[i] > best_eff) best_eff = eff_of[i];
  }
  int pick = 0;
  for (int i = 0; i < 3; ++i)
    if (eff_of[i] >= best_eff - 0.02) pick = i;
  {   // GWN_PG_SUB=<1|2|4>: force the macro-tile size (A/B measurements)
    const char* e = getenv("GWN_PG_SUB");
    if (e) for (int i = 0; i < 3; ++i) if (subs[i] == atoi(e) && eff_of[i] > 0) pick = i;
  }
  p.sub = subs[pick];
  const int stages = stg_of[pick];
  const int n_acc = acc_c * p.sub <= 128 ? 4 : 2;
  p.n_tiles = (int)cdiv(sub_tiles, p.sub);
  PgMaps maps;
  int n_maps = 0;
  struct Key { const bf16* base; long long rows; int pitch, col; } keys[PG_TC_MAX_MAPS];
  for (int q = 0; q < NB; ++q) {
    const PgChunk& c = p.ch[q];
    const long long rows = flat ? p.P : c.rows_per_n;
    int m = -1;
    for (int i = 0; i < n_maps; ++i)
      if (keys[i].base == c.base && keys[i].rows == rows && keys[i].pitch == c.pitch && keys[i].col == c.col_off) m = i;
    if (m < 0) {
      GWN_REQUIRE(n_maps < PG_TC_MAX_MAPS, "pos_gemm_tc: too many distinct sources");
      m = n_maps++;
      keys[m] = Key{c.base, rows, c.pitch, c.col_off};
      if (int rc = tg_map_rows3d(&maps.m[m], c.base + c.col_off, (uint64_t)rows, (uint64_t)(flat ? 1 : n_real),
                                 (uint64_t)c.pitch, 128))
        return rc;
    }
    p.map_of[q] = m;
    GWN_REQUIRE(c.row_off > -(1ll << 30) && c.row_off < (1ll << 30), "pos_gemm_tc: row offset out of range");
    p.row_off[q] = (int)c.row_off;
  }
  for (int i = n_maps; i < PG_TC_MAX_MAPS; ++i) maps.m[i] = maps.m[0];
  if constexpr (KOUT > 0) {
    GWN_REQUIRE(n_out <= KOUT && KOUT <= 3 && n_maps <= PGT_OUT_MAP0 && p.N == 64, "pos_gemm_tc: staged outputs need <= %d input maps", PGT_OUT_MAP0);
    for (int o = 0; o < n_out; ++o)
      if (int rc = tg_map_rows3d(&maps.m[PGT_OUT_MAP0 + o], epi.out_ptr(o), (uint64_t)p.rows_out, (uint64_t)p.n_samples, 32, 128))
        return rc;
  }
  {
    p.trace = trace_ptr("GWN_PG_TRACE");
  }
  const size_t a_bytes = (size_t)NB * 8192 * p.sub;
  const size_t smem = fixed + stages * a_bytes;
  if (p.wsrc.W) {
    GWN_REQUIRE(32 * p.n_chunks * p.N <= 16 * PGT_THREADS && p.N % 4 == 0 && p.wsrc.ld % 4 == 0,
                "pos_gemm_tc: weight image %d x %d too large for the in-kernel build", 32 * p.n_chunks, p.N);
    GWN_REQUIRE(!p.wsrc.bn || (!p.wsrc.transposed && PGT_THREADS % (p.N / 4) == 0 && stages * a_bytes >= 16 * PGT_THREADS),
                "pos_gemm_tc: BatchNorm fold needs N/4 to divide the CTA size");
    GWN_REQUIRE((reinterpret_cast<unsigned long long>(p.wsrc.W) & 15ull) == 0, "pos_gemm_tc: weights must be 16-byte aligned");
  }
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;
  if constexpr (Epi::kFast > 0)
    if (p.n_chunks == Epi::kFast) return pos_gemm_tc_run<Epi, Epi::kFast>(maps, p, epi, stages, n_acc, grid, smem, st);
  return pos_gemm_tc_run<Epi, 0>(maps, p, epi, stages, n_acc, grid, smem, st);
}

}  // namespace gwn
