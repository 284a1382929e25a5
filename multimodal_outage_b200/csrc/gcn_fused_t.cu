// Fused diffusion graph convolution, forward, TRANSPOSED ("T-form") contraction for small graphs (see gcn_fused.cuh
// for the math).  gcn_fused.cu puts a slab's V nodes on the 128 accumulator rows: at V = 67 half of every MMA, two of
// the four warp schedulers and half of every epilogue warp are idle, and the hop GEMM is 30 MMAs of 128x32x16 per slab
// (bound by the shared-memory read of the 128-row A operand).  Here a group of FOUR slabs is contracted at once with
//     M = (slab, channel) = 128      N = output node w (NP = 16 ceil(V/16))      K = (hop j, input node v), j = 0..H
//     h^T[(s,c), w] = sum_{j,v} U_j[(s,v), c] * Mt_j[v, w]          Mt_0 = I, Mt_{2s+1} = A_s, Mt_{2s+2} = A_s^2
// i.e. 30 MMAs of 128x80x16 per FOUR slabs (4.5x less tensor time), every TMEM lane / scheduler carries epilogue work,
// the per-channel statistics are per-thread sums and the bias is one register.
//   GEMM 1 (as before, but over flat 128-position tiles): U = z W, [128 pos, 32] x [32, 32(1+H)], D in TMEM (224 cols)
//   stage warps: U (lane = position) -> bf16 -> the MN-major A operand of GEMM 2 in shared memory, row k = j*V + v,
//                M group = (slab, channel/8)
//   GEMM 2: A = U^T (MN-major, no swizzle), B = the stacked support image [KT/8][NP][8] (K-major, resident),
//           D = two 80-column accumulators
//   epilogue (lane = channel of slab q): + bias, dropout, residual (folded BN), bf16 store, statistics.
// Shared memory (V = 67, H = 6): z tiles 16 KB | A 120 KB | B 75 KB | W image 14 KB | keep-bits 1.5 KB = 226.6 KB.
#include "gcn_fused.cuh"
#include "tc.cuh"
#include "tma_gemm.cuh"

namespace gwn {

constexpr int GT_THREADS = 640;        // 20 warps: producer, MMA, 2 mask, 8 stage, 8 epilogue
constexpr int GT_TILES = 3;          // 128-position tiles per group of 4 slabs (4 V <= 384)

struct GtLayout { uint32_t z_off, a_off, b_off, w_off, m_off, s_off, bar_off, total; };
__host__ __device__ inline GtLayout gt_layout(int KT, int NP, int NU) {
  GtLayout L;
  L.z_off = 0;
  L.a_off = 2u * 8192u;
  L.b_off = L.a_off + (uint32_t)KT * 256u;                 // [16 M groups][KT rows][16 B]
  L.w_off = L.b_off + (uint32_t)KT * (uint32_t)NP * 2u;    // [KT/8][NP][16 B]
  L.m_off = L.w_off + 4u * (uint32_t)NU * 16u;             // [4][NU][16 B]
  L.s_off = L.m_off + 32u * 12u * 4u;                      // keep-bits [32 ch][12 words]
  L.bar_off = L.s_off + 256u;                              // statistics scratch [64]
  L.total = L.bar_off + 160u;
  return L;
}

// 16 keep-bits (bit i = channel i of the group is kept) out of the same stream as dropout16: byte-parallel compare
// (u8 >= thr on four bytes of a word at once) and a multiply "movemask" instead of 16 extract/compare/select chains
__device__ __forceinline__ uint32_t dropout16_bits(uint64_t seed, uint64_t offset, uint64_t idx16, uint32_t thr) {
  const uint4 r = philox4x32(seed, offset, idx16);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  const uint32_t H = 0x80808080u;
  uint32_t bits = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t y;
    if (thr <= 128u) y = (w[i] | ((w[i] | H) - thr * 0x01010101u)) & H;                 // bit 7 of each byte: byte >= thr
    else y = w[i] & (((w[i] & ~H) | H) - (thr - 128u) * 0x01010101u) & H;
    bits |= ((((y >> 7) & 0x01010101u) * 0x01020408u) >> 24) << (4 * i);
  }
  return bits;
}

// 32 x 32 bit-matrix transpose across the warp: in: lane n holds word x with bit c; out: lane c holds bit n
__device__ __forceinline__ uint32_t warp_bit_transpose(uint32_t x, int lane) {
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
    x = (lane & j) ? ((x & ~m) | ((y & ~m) >> j)) : ((x & m) | ((y & m) << j));
  }
  return x;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ uint32_t gt_pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

#define GT_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && gi < 63 && lane == 0) p.trace[gi * 8 + (slot)] = clock64(); } while (0)

template <int NM>
__global__ void __launch_bounds__(GT_THREADS, 1) gcn_fwd_t_kernel(const __grid_constant__ CUtensorMap zmap,
                                                                   const __grid_constant__ GcnFwdParams p) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NU = 32 * (1 + NM);
  const int V = p.V, KT = p.KT, NP = p.NP;
  const GtLayout L = gt_layout(KT, NP, NU);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* z_full = bars;            // [2]
  uint64_t* z_empty = bars + 2;       // [2]
  uint64_t* ua_full = bars + 4;       // U tile, column half a = chunks 0..3 (TMEM columns [0,128))
  uint64_t* ua_empty = bars + 5;
  uint64_t* a_full = bars + 6;
  uint64_t* a_empty = bars + 7;
  uint64_t* d_full = bars + 8;        // [2]
  uint64_t* d_empty = bars + 10;      // [2]
  uint64_t* m_free = bars + 12;
  uint64_t* m_full = bars + 13;
  uint64_t* ub_full = bars + 14;      // U tile, column half b = chunks 4..NM (TMEM columns [128,224)): GEMM 1 of one half
  uint64_t* ub_empty = bars + 15;     // runs under the staging of the other
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  uint64_t* a_half = bars + 17;       // chunks 0..3 of the whole group are staged: GEMM 2 may start on K rows [0, 256)
  uint32_t* kbits = reinterpret_cast<uint32_t*>(smem + L.m_off);       // [32 ch][12 words over the group's 384 rows]
  float* sscr = reinterpret_cast<float*>(smem + L.s_off);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&z_full[i], 1); mbar_init(&z_empty[i], 1);
      mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 8);
    }
    // consumer -> producer hand-offs count ONE arrival per warp (lane 0 after __syncwarp): per-thread arrivals put
    // ~2,400 serialised updates of the barrier words on the shared-memory pipe per group
    mbar_init(m_free, 8); mbar_init(m_full, 2);
    mbar_init(ua_full, 1); mbar_init(ua_empty, 8); mbar_init(ub_full, 1); mbar_init(ub_empty, 8);
    mbar_init(a_full, 8); mbar_init(a_empty, 1); mbar_init(a_half, 8);
    fence_barrier_init();
    tg::tma_prefetch_desc(&zmap);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  {  // resident operands: stacked support image (B of GEMM 2), mlp weight image (B of GEMM 1); A zeroed (K padding
     // rows and the rows of absent slabs must be finite)
    const uint4* src = reinterpret_cast<const uint4*>(p.mats_t);
    uint4* dst = reinterpret_cast<uint4*>(smem + L.b_off);
    const int nb = KT * NP / 8;
    for (int i0 = tid; i0 < nb; i0 += 4 * GT_THREADS) {
      uint4 v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) if (i0 + q * GT_THREADS < nb) v[q] = __ldg(src + i0 + q * GT_THREADS);
#pragma unroll
      for (int q = 0; q < 4; ++q) if (i0 + q * GT_THREADS < nb) dst[i0 + q * GT_THREADS] = v[q];
    }
    uint4* az = reinterpret_cast<uint4*>(smem + L.a_off);
    for (int i = tid; i < KT * 16; i += GT_THREADS) az[i] = make_uint4(0u, 0u, 0u, 0u);
    bf16* wimg = reinterpret_cast<bf16*>(smem + L.w_off);
    constexpr int T4 = 8 * NU, ITS = (T4 + GT_THREADS - 1) / GT_THREADS;
    float4 wv[ITS];
#pragma unroll
    for (int it = 0; it < ITS; ++it) {
      const int i = tid + it * GT_THREADS;
      wv[it] = i < T4 ? __ldg(reinterpret_cast<const float4*>(p.w_src) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < ITS; ++it) {
      const int i = tid + it * GT_THREADS;
      if (i < T4) {
        const int co = (i & 7) * 4, c = (i >> 3) & 31, n = (i >> 8) * 32 + co;
        bf16* d = wimg + ((c >> 3) * NU + n) * 8 + (c & 7);
        d[0] = __float2bfloat16_rn(wv[it].x); d[8] = __float2bfloat16_rn(wv[it].y);
        d[16] = __float2bfloat16_rn(wv[it].z); d[24] = __float2bfloat16_rn(wv[it].w);
      }
    }
    if (tid < 64) sscr[tid] = 0.f;
    fence_proxy_async();
  }
  // the prologue read step-constant data only (support image, mlp weights): under a programmatic launch it runs while the
  // gate kernel that produces z / scale / shift drains; those, the residual and all outputs come after this point
  pdl_wait();
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  if (tid == 0 && (sbase & 1023u)) __trap();     // the 64B-swizzled z tiles assume a 1024-byte aligned window
  const int n_groups = (p.slabs + 3) >> 2;
  constexpr uint32_t TU = 0u, TD = 256u;       // TMEM columns: U tile [0,224), D2 buffers at 256 and 384

  if (warp == 0) {
    // ===================== TMA producer: z tiles of 128 positions =====================
    int tt = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
      for (int t = 0; t < GT_TILES; ++t, ++tt) {
        const int zb = tt & 1;
        mbar_wait(&z_empty[zb], (uint32_t)(((tt >> 1) & 1) ^ 1));
        if (elect_one()) {
          tg::mbar_expect_tx(&z_full[zb], 8192u);
          tg::tma_3d(sbase + L.z_off + (uint32_t)zb * 8192u, &zmap, 0, g * 4 * V + 128 * t, 0, &z_full[zb]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the loop; one elected lane issues) =====================
    constexpr int NA_COLS = NU < 128 ? NU : 128, NB_COLS = NU - NA_COLS;          // GEMM 1 in two column halves
    const uint32_t idesc1a = make_idesc_bf16(128, NA_COLS, false, false);
    const uint32_t idesc1b = make_idesc_bf16(128, NB_COLS > 0 ? NB_COLS : 16, false, false);
    const uint32_t idesc2 = make_idesc_bf16(128, NP, true, false);
    const uint64_t az = tg::make_desc_sw(0, 16u, 512u, 4u);                          // z tile: K-major SW64
    const uint64_t bw = make_smem_desc(sbase + L.w_off, (uint32_t)NU * 16u, 128u);   // W image: K-major
    const uint64_t au = make_smem_desc(sbase + L.a_off, 128u, (uint32_t)KT * 16u);   // U^T: MN-major, 16 groups
    const uint64_t bm = make_smem_desc(sbase + L.b_off, (uint32_t)NP * 16u, 128u);   // stacked supports: K-major
    const uint32_t bstep = ((uint32_t)NP * 32u) >> 4;                                // two K pieces per K = 16 step
    const int ksteps = KT >> 4;
    int tt = 0, gi = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++gi) {
      GT_TRACE(0);
      for (int t = 0; t < GT_TILES; ++t, ++tt) {
        const int zb = tt & 1;
        mbar_wait(&z_full[zb], (uint32_t)((tt >> 1) & 1));
        const uint64_t ad = az + (uint64_t)((sbase + L.z_off + (uint32_t)zb * 8192u) >> 4);
        const uint64_t w2 = (uint64_t)(((uint32_t)NU * 32u) >> 4);         // second K = 16 half of the weight image
        mbar_wait(ua_empty, (uint32_t)((tt & 1) ^ 1));
        tc_fence_after();
        if (elect_one()) {
          umma_bf16(tmem_base + TU, ad, bw, idesc1a, 0u);
          umma_bf16(tmem_base + TU, ad + 2u, bw + w2, idesc1a, 1u);
          if (NB_COLS == 0) umma_commit(&z_empty[zb]);
          umma_commit(ua_full);
        }
        __syncwarp();
        if (NB_COLS > 0) {
          mbar_wait(ub_empty, (uint32_t)((tt & 1) ^ 1));
          tc_fence_after();
          if (elect_one()) {
            umma_bf16(tmem_base + TU + 128u, ad, bw + 128u, idesc1b, 0u);           // weight rows n >= 128: + 128 x 16 B
            umma_bf16(tmem_base + TU + 128u, ad + 2u, bw + 128u + w2, idesc1b, 1u);
            umma_commit(&z_empty[zb]);
            umma_commit(ub_full);
          }
          __syncwarp();
        }
      }
      const int b = gi & 1;
      GT_TRACE(1);
      // GEMM 2 in two K parts: rows k = j V + v < 256 belong to chunks j <= 3, whose staging (column half a) ends one hand-off
      // before the group's last one - the first 16 K steps run under the staging of the last tile's column half b
      const int ks_half = (NB_COLS > 0 && 4 * V >= 256 && ksteps > 16) ? 16 : 0;
      const uint32_t d = tmem_base + TD + (uint32_t)b * 128u;
      if (ks_half > 0) {
        mbar_wait(a_half, (uint32_t)(gi & 1));
        mbar_wait(&d_empty[b], (uint32_t)(((gi >> 1) & 1) ^ 1));
        tc_fence_after();
        if (elect_one()) {
          for (int ks = 0; ks < ks_half; ++ks)
            umma_bf16(d, au + (uint64_t)(16 * ks), bm + (uint64_t)ks * bstep, idesc2, ks == 0 ? 0u : 1u);
        }
        __syncwarp();
      }
      mbar_wait(a_full, (uint32_t)(gi & 1));
      GT_TRACE(2);
      if (ks_half == 0) mbar_wait(&d_empty[b], (uint32_t)(((gi >> 1) & 1) ^ 1));
      tc_fence_after();
      if (elect_one()) {
        for (int ks = ks_half; ks < ksteps; ++ks)
          umma_bf16(d, au + (uint64_t)(16 * ks), bm + (uint64_t)ks * bstep, idesc2, ks == 0 ? 0u : 1u);
        umma_commit(a_empty);
        umma_commit(&d_full[b]);
      }
      __syncwarp();
      GT_TRACE(3);
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== stage warps: TMEM(U) -> bf16 -> A operand of GEMM 2; dropout keep-bits =====================
    const int q = warp & 3, set = (warp - 4) >> 2;           // set 0: chunks 0,2,4,6; set 1: chunks 1,3,5
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int tt = 0, gi = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++gi) {
      const int ns = min(4, p.slabs - 4 * g);                // slabs of this group
      if (warp == 8) GT_TRACE(4);
      for (int t = 0; t < GT_TILES; ++t, ++tt) {
        const int rr = 128 * t + q * 32 + lane;              // row inside the group
        const bool valid = rr < ns * V;
        const int s = rr / V, v = rr - s * V;
        uint8_t* arow = smem + L.a_off + (size_t)(s * 4) * KT * 16 + (size_t)v * 16;
        const bool warp_has_rows = 128 * t + q * 32 < ns * V;       // (the third tile holds 4V - 256 rows: most warps skip it)
        auto put = [&](int j, const uint32_t (&r)[32]) {
#pragma unroll
          for (int cg = 0; cg < 4; ++cg) {
            uint4 pk;
            pk.x = gt_pack(__uint_as_float(r[8 * cg]), __uint_as_float(r[8 * cg + 1]));
            pk.y = gt_pack(__uint_as_float(r[8 * cg + 2]), __uint_as_float(r[8 * cg + 3]));
            pk.z = gt_pack(__uint_as_float(r[8 * cg + 4]), __uint_as_float(r[8 * cg + 5]));
            pk.w = gt_pack(__uint_as_float(r[8 * cg + 6]), __uint_as_float(r[8 * cg + 7]));
            *reinterpret_cast<uint4*>(arow + (size_t)cg * KT * 16 + (size_t)(j * V) * 16) = pk;
          }
        };
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half == 1 && NM < 4) break;                            // (no second column half)
          const int j = 4 * half + set, j2 = j + 2;                  // this warp's chunks of the half
          const int jmax = half == 0 ? (NM < 3 ? NM : 3) : NM;
          mbar_wait_lazy2(half == 0 ? ua_full : ub_full, (uint32_t)(tt & 1));
          if (t == 0 && half == 0) mbar_wait_lazy2(a_empty, (uint32_t)((gi & 1) ^ 1));      // GEMM 2 of the previous group has read A
          tc_fence_after();
          uint32_t ra[32], rb[32];
          const bool l1 = warp_has_rows && j <= jmax, l2 = warp_has_rows && j2 <= jmax;
          if (l1) tmem_ld32_issue(tmem_base + lane_off + TU + (uint32_t)j * 32u, ra);
          if (l2) tmem_ld32_issue(tmem_base + lane_off + TU + (uint32_t)j2 * 32u, rb);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(half == 0 ? ua_empty : ub_empty);   // the half is in registers: GEMM 1 may overwrite it
          if (valid && l1) put(j, ra);
          if (valid && l2) put(j2, rb);
          if (half == 0 && t == GT_TILES - 1) {                      // column half a of the whole group is in shared memory
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_half);
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
      if (warp == 8) GT_TRACE(5);
    }
  } else if (warp == 2 || warp == 3) {
    // ===================== mask warps: dropout keep-bits of the group's 384 rows =====================
    // drawn position-major (lane = position: the generator's native order, identical to every other kernel path) and
    // transposed to channel-major words with a 5-step shuffle butterfly; no TMEM access, so any warp can do it
    const bool philox = (p.mask == nullptr) && p.drop_p > 0.f;
    uint64_t sd = 0, of = 0;
    if (philox) { sd = p.rng ? __ldg(p.rng) : p.seed; of = p.rng ? p.offset + __ldg(p.rng + 1) : p.offset; }
    uint32_t thr = (uint32_t)(p.drop_p * 256.0f + 0.5f);
    thr = thr > 255u ? 255u : thr;
    int gi = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++gi) {
      const int ns = min(4, p.slabs - 4 * g);
      mbar_wait_lazy(m_free, (uint32_t)((gi & 1) ^ 1));           // the epilogue of the previous group has read its bits
      if (philox) {
        for (int wi = warp - 2; wi < 12; wi += 2) {
          const int rr = 32 * wi + lane;
          uint32_t keep = 0;
          if (rr < ns * V) {
            const long long pp = ((long long)g * 4) * V + rr;
            keep = dropout16_bits(sd, of, (uint64_t)(pp * 2), thr) | (dropout16_bits(sd, of, (uint64_t)(pp * 2 + 1), thr) << 16);
          }
          kbits[lane * 12 + wi] = warp_bit_transpose(keep, lane);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(m_full);
    }
  } else if (warp >= 12) {
    // ===================== epilogue: lane = channel c of slab q of the group =====================
    // two warps per quadrant: set 0 takes nodes [0, 32), set 1 nodes [32, V).  All of a warp's residual loads are
    // requested BEFORE waiting for the accumulator (they are the only HBM latency on this path).
    const int q = warp & 3, c = lane, e = (warp - 12) >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const bool philox = (p.mask == nullptr) && p.drop_p > 0.f;
    uint32_t thr = (uint32_t)(p.drop_p * 256.0f + 0.5f);
    thr = thr > 255u ? 255u : thr;
    const float inv = philox ? 256.0f / (256.0f - (float)thr) : 1.f;
    const float bias = __ldg(p.bias + c);
    const float sc = p.scale ? __ldg(p.scale + c) : 1.f, sh = p.scale ? __ldg(p.shift + c) : 0.f;
    const int cb = 32 * e;                                   // first node / accumulator column of this warp
    const int c_end = e == 0 ? min(32, NP) : NP;             // accumulator columns [cb, c_end) belong to this warp
    float s1 = 0.f, s2 = 0.f;
    int gi = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++gi) {
      const int b = gi & 1;
      const int slab = 4 * g + q;
      const bool has = slab < p.slabs;
      const long long p0 = (long long)slab * V;
      const bf16* rp = p.u_prev; bf16* up = p.u; const bf16* mp = p.mask;
      if (has) {
        long long n, rem;
        split_pos(p0, p.RO, n, rem);
        rp = p.u_prev + (n * p.RI + rem + p.crop) * 32 + c;
        up = p.u + p0 * 32 + c;
        if (mp) mp = p.mask + p0 * 32 + c;
      }
      uint32_t res[48];
#pragma unroll
      for (int i = 0; i < 48; ++i)
        if (has && cb + i < V && (e == 1 || i < 32)) res[i] = __ldg(reinterpret_cast<const uint16_t*>(rp) + (size_t)(cb + i) * 32);
      mbar_wait_lazy(&d_full[b], (uint32_t)((gi >> 1) & 1));
      if (warp == 12) GT_TRACE(6);
      tc_fence_after();
      // keep-bits of this channel for the slab's rows [q V, q V + V) of the group, aligned to bit 0 (three scalars: a
      // dynamically indexed array would live in local memory)
      uint32_t kb0 = 0xffffffffu, kb1 = 0xffffffffu, kb2 = 0xffffffffu;
      mbar_wait_lazy(m_full, (uint32_t)(gi & 1));
      if (philox) {
        const int r0 = q * V, w0 = r0 >> 5, o = r0 & 31;
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = (w0 + k < 12) ? kbits[c * 12 + w0 + k] : 0u;
        kb0 = __funnelshift_r(w[0], w[1], o); kb1 = __funnelshift_r(w[1], w[2], o); kb2 = __funnelshift_r(w[2], w[3], o);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(m_free);
      const uint32_t td = tmem_base + lane_off + TD + (uint32_t)b * 128u;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int c0 = cb + 16 * k;
        if (c0 < c_end) {
          uint32_t r[16];
          tmem_ld16(td + (uint32_t)c0, r);
          tmem_ld_wait();
          if (c0 + 16 >= c_end) {                 // this warp's last read: the accumulator may be overwritten
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[b]);
          }
          if (has) {
            const int nv = V - c0;                // valid nodes of this sub-chunk (may be <= 0 or > 16)
            const uint32_t kw = (c0 < 32 ? kb0 : c0 < 64 ? kb1 : kb2) >> (c0 & 31);
            bf16* upk = up + (size_t)c0 * 32;
            if (mp == nullptr && nv >= 16) {      // common case: a full sub-chunk, no explicit mask - branch free
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float m = ((kw >> i) & 1u) ? inv : 0.f;
                const float rv = __uint_as_float(res[16 * k + i] << 16);
                const float val = fmaf(__uint_as_float(r[i]) + bias, m, fmaf(rv, sc, sh));
                upk[i * 32] = __float2bfloat16_rn(val);
                s1 += val; s2 = fmaf(val, val, s2);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (i < nv) {
                  float m;
                  if (mp) m = __bfloat162float(mp[(size_t)(c0 + i) * 32]);
                  else m = ((kw >> i) & 1u) ? inv : 0.f;
                  const float rv = __uint_as_float(res[16 * k + i] << 16);
                  const float val = fmaf(__uint_as_float(r[i]) + bias, m, fmaf(rv, sc, sh));
                  upk[i * 32] = __float2bfloat16_rn(val);
                  s1 += val; s2 = fmaf(val, val, s2);
                }
              }
            }
          }
        }
      }
      if (cb >= c_end) {                          // (V <= 32: set 1 has no columns but still takes part in the hand-offs)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d_empty[b]);
      }
      if (warp == 12) GT_TRACE(7);
    }
    atomicAdd(sscr + c, s1);
    atomicAdd(sscr + 32 + c, s2);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    atomicAdd(p.stats + lane, (double)sscr[lane]);
    atomicAdd(p.stats + 32 + lane, (double)sscr[32 + lane]);
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int gcn_fused_t_supported(int V, int n_mats) {
  if (V < 1 || 4 * V > 128 * GT_TILES || (n_mats != 2 && n_mats != 4 && n_mats != 6)) return 0;
  const int KT = (((1 + n_mats) * V + 15) / 16) * 16, NP = ((V + 15) / 16) * 16;
  if (NP > 80) return 0;      // epilogue column split: set 0 = [0,32), set 1 = [32,80)
  return gt_layout(KT, NP, 32 * (1 + n_mats)).total <= 227u * 1024u ? 1 : 0;
}

int launch_gcn_fwd_t(GcnFwdParams& p, cudaStream_t st) {
  if (p.slabs <= 0) return 0;
  GWN_REQUIRE(gcn_fused_t_supported(p.V, p.n_mats) && p.w_src && p.mats_t, "gcn_fwd_t: unsupported shape (V=%d, %d matrices)",
              p.V, p.n_mats);
  p.KT = (((1 + p.n_mats) * p.V + 15) / 16) * 16;
  p.NP = ((p.V + 15) / 16) * 16;
  const GtLayout L = gt_layout(p.KT, p.NP, 32 * (1 + p.n_mats));
  CUtensorMap zmap;
  if (int rc = tg_map_rows3d(&zmap, p.z, (uint64_t)p.slabs * p.V, 1, 32, 128)) return rc;
  const int sms = tg_sm_count();
  const int groups = (p.slabs + 3) / 4;
  const int grid = groups < sms ? groups : sms;
#define GT_CASE(NM_)                                                                                                  \
  if (p.n_mats == NM_) {                                                                                              \
    static bool attr = false;                                                                                         \
    if (!attr) {                                                                                                      \
      GWN_CUDA(cudaFuncSetAttribute(gcn_fwd_t_kernel<NM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      attr = true;                                                                                                    \
    }                                                                                                                 \
    GWN_CUDA(launch_pdl(gcn_fwd_t_kernel<NM_>, dim3(grid), dim3(GT_THREADS), L.total, st, zmap, p));                  \
    GWN_LAUNCHED();                                                                                                   \
    return 0;                                                                                                         \
  }
  GT_CASE(6) GT_CASE(4) GT_CASE(2)
#undef GT_CASE
  GWN_REQUIRE(false, "gcn_fwd_t: no kernel instance for %d matrices", p.n_mats);
  return -1;
}

}  // namespace gwn
