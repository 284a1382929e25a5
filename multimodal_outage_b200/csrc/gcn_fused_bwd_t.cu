// Fused diffusion graph convolution, backward, TRANSPOSED ("T-form") hops over groups of FOUR slabs (math in
// gcn_fused.cuh; graph_wavenet.py:76-98 backward, :222-226 gate backward).
//
// gcn_fused_bwd.cu puts a slab's V nodes on the 128 accumulator rows: 65 MMAs of 128 x 32 x 16 per slab, each bound by
// the shared-memory read of its 128-row A operand, with the SIMT work of 67 nodes on two of the four schedulers.
// Here the hops of four slabs run at once with the CHANNELS on the accumulator rows:
//   GEMM H   dU^T[(s,c'), (j,v)] = sum_w dh^T[(s,c'), w] * Mt_j[v, w]        M = (slab, channel) = 128, K = node w,
//            N = (hop slot j, node v); slot 0 = identity (dU_0 = dh), Mt_{2s+1} = A_s, Mt_{2s+2} = A_s A_s
//            - 6 hops of 4 slabs are 30 MMAs of 128 x <=128 x 16 instead of 120 of 128 x 32 x 16.
// The accumulator (lane = (slab, channel), column = (hop, node)) is drained by eight stage warps - every TMEM lane
// quadrant and every scheduler carries the same load - into the POSITION-major operand of the two weight GEMMs:
//   GEMM Z   dz[pos, c]      = sum_{j,c'} dU_j[pos, c'] W[(j,c), c']         M = position, K = (hop, channel)
//   GEMM W   dW[(j,c'), c]  += sum_pos dU_j[pos, c'] z[pos, c]               M = (hop, channel), K = position
// (the SAME staged bytes, viewed MN-major for Z and K-major for W; z = a.b rebuilt on chip).  Work moves through the
// kernel in ITEMS = (node range r of 32 nodes) x (hop half h of 4 slots): one item is 128 TMEM columns and a 32 KB
// staging slot; two of each form a ring, so GEMM H of item i+1 and the staging of item i+1 run under GEMMs Z / W of item i.
//
// Adaptive-support gradient without a second hop: with U5 = z W_{2sa+1}, U6 = z W_{2sa+2} (position GEMMs, staged
// through the same ring as two more items per group), Q5 = sum U5[v] . dh[w] and Q6 = sum U6[v] . dh[w],
//   dA = Q5 + A^T Q6 + Q6 A^T            (the A^2 hop is linear in Q6: the two products are applied ONCE per step, in
// fp32, by gwn_adp_pair_bwd, instead of one hop + one more product per slab here).  Q5 accumulates into dA, Q6 into dQ6.
//
// TMEM (512 columns): ring 2 x 128 | dz 32 | dW 2 x 32 | Q5 80 | Q6 80.
// Shared memory (V = 67, 3 supports): stacked image 80 KB | dh^T x2 40 KB | z 18 KB | W^T image 14 KB | W56 4 KB |
// staging ring 64 KB = 220 KB.
// Warps (24): w0, w22 MMA issue (stage 1 / stage 2 of every item) | w4-w11 stage (two per TMEM quadrant) | w12-w15 epilogue (gate backward, one per quadrant) |
// w1-w3, w16-w21 prep (one position per thread: dh = du . mask, z = a . b).
#include <type_traits>

#include "gcn_fused.cuh"
#include "tc.cuh"

namespace gwn {

constexpr int BT_THREADS = 768;      // 24 warps (80 registers per thread either way)
constexpr int BT_PREP_WARPS = 9;
constexpr int BT_ISSUE_Z = 22;       // MMA-issuing warps besides warp 0: GEMM Z,
constexpr int BT_ISSUE_W = 23;       // GEMM W + the Q products

struct BtLayout { uint32_t bh_off, dh_off, dh_bytes, z_off, wt_off, w56_off, ring_off, bar_off, total; };
__host__ __device__ constexpr BtLayout bt_layout(const BtGeom& G, bool has_da) {
  BtLayout L{};
  L.bh_off = 0;
  L.dh_off = (uint32_t)(G.KW / 8) * (uint32_t)G.NTOT * 16u;
  L.dh_bytes = 16u * (uint32_t)G.KW * 16u;                    // [(slab, c'/8) 16 groups][KW rows w][16 B]
  L.z_off = L.dh_off + 2u * L.dh_bytes;                       // [c/8 4 groups][4 NPD rows][16 B]
  L.wt_off = L.z_off + 4u * 4u * (uint32_t)G.NPD * 16u;       // (M = 128 reads past the last slab land in the images)
  L.w56_off = L.wt_off + 4u * (uint32_t)G.NH * 32u * 16u;
  L.ring_off = L.w56_off + (has_da ? 4096u : 0u);
  L.bar_off = L.ring_off + 2u * 32768u;
  L.total = L.bar_off + 384u;
  return L;
}

__device__ __forceinline__ uint32_t bt_pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void bt_unpack8(const uint4& q, float v[8]) {
  v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xFFFF0000u);
  v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xFFFF0000u);
  v[4] = __uint_as_float(q.z << 16); v[5] = __uint_as_float(q.z & 0xFFFF0000u);
  v[6] = __uint_as_float(q.w << 16); v[7] = __uint_as_float(q.w & 0xFFFF0000u);
}
__device__ __forceinline__ uint4 bt_pack8(const uint32_t* r) {
  return make_uint4(bt_pack(__uint_as_float(r[0]), __uint_as_float(r[1])), bt_pack(__uint_as_float(r[2]), __uint_as_float(r[3])),
                    bt_pack(__uint_as_float(r[4]), __uint_as_float(r[5])), bt_pack(__uint_as_float(r[6]), __uint_as_float(r[7])));
}
__device__ __forceinline__ void bt_tmem_ld8(uint32_t taddr, uint32_t r[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void bt_tmem_ld16(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// tcgen05.mma with the descriptors given as (lo, hi) words: only lo (start address, LBO) changes between the MMAs of
// this kernel, so an operand costs one 32-bit uniform add instead of a 64-bit add pair
__device__ __forceinline__ void bt_mma(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int I, int N, typename F>
__device__ __forceinline__ void bt_static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    bt_static_for<I + 1, N>(f);
  }
}

// item table of one group: D items (r, h) in order, the two U items (slab pairs; DA only) after ranges 0 and 1
template <int NM, int NPD, bool DA>
struct BtItems {
  static constexpr int NR = BtGeom(NM, NPD).NR, NHALF = BtGeom(NM, NPD).NHALF;
  static constexpr int ND = NR * NHALF;
  static constexpr int rs(int r) { return BtGeom(NM, NPD).rs(r); }
  static constexpr int nh(int h) { return BtGeom(NM, NPD).nh(h); }
  static constexpr int n0(int r, int h) { return BtGeom(NM, NPD).n0(r, h); }
  static constexpr int nmma(int r, int h) { return BtGeom(NM, NPD).nmma(r, h); }
  static constexpr int NI = ND + (DA ? 2 : 0);
  // kind: 0 = D, 1 = U;  a = r (D) or slab pair p (U);  b = h (D)
  static constexpr int kind(int i) { return decode(i, 0); }
  static constexpr int arg_a(int i) { return decode(i, 1); }
  static constexpr int arg_b(int i) { return decode(i, 2); }
  static constexpr int decode(int i, int what) {
    int k = 0, up = 0;
    for (int r = 0; r < NR; ++r) {
      for (int h = 0; h < NHALF; ++h, ++k)
        if (k == i) return what == 0 ? 0 : what == 1 ? r : h;
      if (DA && up < 2) {
        if (k == i) return what == 0 ? 1 : what == 1 ? up : 0;
        ++k; ++up;
      }
    }
    while (DA && up < 2) {
      if (k == i) return what == 0 ? 1 : what == 1 ? up : 0;
      ++k; ++up;
    }
    return -1;
  }
};

// debug timeline of CTA 0 (GWN_TRACE builds, scripts/gpu_gcn_bwd_t_trace.py): p.trace[item * 16 + slot] = clock64()
#ifdef GWN_TRACE
#define BT_TRACE(item, slot) do { if (p.trace && blockIdx.x == 0 && (item) < 250 && lane == 0) p.trace[(item) * 16 + (slot)] = clock64(); } while (0)
#else
#define BT_TRACE(item, slot) do { } while (0)
#endif

template <int NM, int NPD, bool DA>
__global__ void __launch_bounds__(BT_THREADS, 1) gcn_bwd_t_kernel(const __grid_constant__ GcnBwdParams p) {
  using namespace tc;
  using IT = BtItems<NM, NPD, DA>;
  constexpr BtGeom G(NM, NPD);
  constexpr BtLayout L = bt_layout(G, DA);
  constexpr int NI = IT::NI, NH = G.NH, KW = G.KW, NR = G.NR, NHALF = G.NHALF, NTOT = G.NTOT;
  constexpr int NU = 32 * NH;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = p.V;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* t_full = bars;            // [2] stage-1 accumulator of the ring slot complete
  uint64_t* t_empty = bars + 2;       // [2] ... drained into registers
  uint64_t* s_full = bars + 4;        // [2] staging slot written
  uint64_t* s_empty = bars + 6;       // [2] ... consumed by the stage-2 MMAs
  uint64_t* dh_full = bars + 8;       // [2]
  uint64_t* dh_empty = bars + 10;     // [2]
  uint64_t* z_full = bars + 12;
  uint64_t* z_empty = bars + 13;
  uint64_t* dz_full = bars + 14;
  uint64_t* dz_empty = bars + 15;
  uint64_t* w_full = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  float* dbs = reinterpret_cast<float*>(bars + 18);     // [32] bias-gradient partial of the CTA

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 8); mbar_init(&s_full[i], 8); mbar_init(&s_empty[i], 2);
      mbar_init(&dh_full[i], BT_PREP_WARPS); mbar_init(&dh_empty[i], 2);
    }
    mbar_init(z_full, BT_PREP_WARPS); mbar_init(z_empty, 2);
    mbar_init(dz_full, 1); mbar_init(dz_empty, 4);
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  {  // resident operands
    const uint4* src = reinterpret_cast<const uint4*>(p.mats_bt);
    uint4* dst = reinterpret_cast<uint4*>(smem + L.bh_off);
    constexpr int nb = (KW / 8) * NTOT;
    for (int i0 = tid; i0 < nb; i0 += 4 * BT_THREADS) {
      uint4 v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) if (i0 + q * BT_THREADS < nb) v[q] = __ldg(src + i0 + q * BT_THREADS);
#pragma unroll
      for (int q = 0; q < 4; ++q) if (i0 + q * BT_THREADS < nb) dst[i0 + q * BT_THREADS] = v[q];
    }
    // dh^T / z buffers: padding rows (w >= V, v >= V) are never written again and must be zero
    uint4* zz = reinterpret_cast<uint4*>(smem + L.dh_off);
    for (int i = tid; i < (int)((L.wt_off - L.dh_off) / 16); i += BT_THREADS) zz[i] = make_uint4(0u, 0u, 0u, 0u);
    // bf16 UMMA images of the mlp weight: wt[kc = (j, c'>>3)][n = c][c' & 7] = W[j*32 + c][c'];
    //                                     w56[kc = c>>3][n = (which, c')][c & 7] = W[(2sa+1+which)*32 + c][c']
    bf16* wt = reinterpret_cast<bf16*>(smem + L.wt_off);
    bf16* w56 = reinterpret_cast<bf16*>(smem + L.w56_off);
    constexpr int T4 = 8 * NU, ITS = (T4 + BT_THREADS - 1) / BT_THREADS;
    float4 wv[ITS];
#pragma unroll
    for (int it = 0; it < ITS; ++it) {
      const int i = tid + it * BT_THREADS;
      wv[it] = i < T4 ? __ldg(reinterpret_cast<const float4*>(p.w_src) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < ITS; ++it) {
      const int i = tid + it * BT_THREADS;
      if (i < T4) {
        const int j = i >> 8, c = (i >> 3) & 31, co = (i & 7) * 4;
        const float v4[4] = {wv[it].x, wv[it].y, wv[it].z, wv[it].w};
        uint2 pk; pk.x = bt_pack(v4[0], v4[1]); pk.y = bt_pack(v4[2], v4[3]);
        *reinterpret_cast<uint2*>(wt + ((j * 4 + (co >> 3)) * 32 + c) * 8 + (co & 7)) = pk;
        if (DA && (j == 2 * p.sa + 1 || j == 2 * p.sa + 2)) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            w56[((c >> 3) * 64 + (j - (2 * p.sa + 1)) * 32 + co + q) * 8 + (c & 7)] = __float2bfloat16_rn(v4[q]);
        }
      }
    }
    if (tid < 32) dbs[tid] = 0.f;
    fence_proxy_async();
  }
  // everything above read step-constant data only (support image, mlp weights) and wrote shared memory: it may run while
  // the previous kernel of the stream drains.  du / a / b / dz_last and every output are touched after this point.
  pdl_wait();
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  const int n_groups = (p.slabs + 3) >> 2;
  constexpr uint32_t TDZ = 256u, TDW = 288u, TA5 = 352u, TA6 = 432u;
  // fused dropout (common.cuh: dropout16): keep iff the element's Philox byte >= thr; kept values are scaled by inv,
  // which this kernel applies in fp32 to what leaves it (dz, dW, db, Q) instead of to every element of dh
  const bool philox = (p.mask == nullptr) && p.drop_p > 0.f;
  uint32_t thr = (uint32_t)(p.drop_p * 256.0f + 0.5f);
  thr = thr > 255u ? 255u : thr;
  const float inv = philox ? 256.0f / (256.0f - (float)thr) : 1.f;

  if (warp == 0 || warp == BT_ISSUE_Z || warp == BT_ISSUE_W) {
    // ===================== MMA issuers =====================
    // THREE issuing warps: A (warp 0) issues stage 1 of every item (GEMM H / U56 into the ring); Z and W issue stage 2
    // (GEMM Z; GEMM W and the Q products) out of the staged slot.  A single issuer runs in lock-step with the tensor
    // pipe - its barrier polls and descriptor arithmetic between bursts (~300 cycles, 16 bursts per group) were dead
    // time for the pipe; split up, one warp's bookkeeping runs under the others' MMAs.  Dependencies between the
    // streams are the ring's mbarriers; each warp's tcgen05.commit covers its own MMAs, so the "free" barriers of a
    // buffer count its reading issuers (dh: A + W, z: A + W, staging slot: Z + W).  The whole warp walks the program
    // (descriptors stay in uniform registers), one elected lane issues.  Only the low 32 bits of a descriptor
    // (address, LBO) ever change: one uniform add per operand.
    auto hi_of = [](uint32_t sbo) -> uint32_t { return (uint32_t)(make_smem_desc(0, 0u, sbo) >> 32); };
    auto lo_of = [](uint32_t addr, uint32_t lbo) -> uint32_t { return (uint32_t)make_smem_desc(addr, lbo, 0u); };
    if (warp == 0) {
      // ---- issuer A: GEMM H (D item) or U5|U6 = z W56 (U item) into ring slot `slot`
      const uint32_t bh_lo = lo_of(sbase + L.bh_off, (uint32_t)NTOT * 16u), bh_hi = hi_of(128u);
      const uint32_t dh_lo0 = lo_of(sbase + L.dh_off, 128u), dh_hi = hi_of((uint32_t)KW * 16u);
      const uint32_t zk_lo = lo_of(sbase + L.z_off, 4u * NPD * 16u), zk_hi = hi_of(128u);
      const uint32_t w56_lo = lo_of(sbase + L.w56_off, 64u * 16u), w56_hi = hi_of(128u);
      int it = 0, gi = 0;
      for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++gi) {
        bt_static_for<0, NI>([&](auto I_) {
          constexpr int I = decltype(I_)::value;
          constexpr int kind = IT::kind(I), a = IT::arg_a(I), hb = IT::arg_b(I);
          const int slot = it & 1, b = gi & 1;
          BT_TRACE(it, 0);
          mbar_wait(&t_empty[slot], (uint32_t)(((it >> 1) & 1) ^ 1));
          if (I == 0) mbar_wait(&dh_full[b], (uint32_t)((gi >> 1) & 1));
          if (kind == 1) mbar_wait(z_full, (uint32_t)(gi & 1));
          tc_fence_after();
          BT_TRACE(it, 1);
          // (everything that depends on run-time values is computed OUTSIDE the elected block: inside it - a divergent
          //  region - the compiler keeps such values in vector registers and pays an R2UR per MMA operand)
          const uint32_t td = tmem_base + (uint32_t)slot * 128u;
          const uint32_t alo = dh_lo0 + (((uint32_t)b * L.dh_bytes) >> 4);
          if (elect_one()) {
            if constexpr (kind == 0) {
              constexpr uint32_t idesc = make_idesc_bf16(128, IT::nmma(a, hb), true, false);
              const uint32_t blo = bh_lo + (uint32_t)IT::n0(a, hb);
#pragma unroll
              for (int ks = 0; ks < KW / 16; ++ks)
                bt_mma(td, alo + (uint32_t)ks * 16u, dh_hi, blo + (uint32_t)(2 * ks) * (uint32_t)NTOT, bh_hi, idesc,
                       ks == 0 ? 0u : 1u);
            } else {
              constexpr uint32_t idesc = make_idesc_bf16(128, 64, false, false);
#pragma unroll
              for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                  bt_mma(td + 64u * sl, zk_lo + (uint32_t)((2 * a + sl) * NPD) + (uint32_t)(2 * ks) * (4u * NPD), zk_hi,
                         w56_lo + (uint32_t)(2 * ks) * 64u, w56_hi, idesc, ks == 0 ? 0u : 1u);
            }
            umma_commit(&t_full[slot]);
            if (I == NI - 1) { umma_commit(z_empty); umma_commit(&dh_empty[b]); }
          }
          __syncwarp();
          BT_TRACE(it, 2);
          ++it;
        });
      }
    } else if (warp == BT_ISSUE_Z) {
      // ---- issuer Z: GEMM Z (D items) out of staging slot `slot`; hands dz to the epilogue after a range's last half
      const uint32_t wt_lo = lo_of(sbase + L.wt_off, 32u * 16u), wt_hi = hi_of(128u);
      const uint32_t smn_lo0 = lo_of(sbase + L.ring_off, 128u), smn_hi = hi_of(2048u);      // staged item, MN-major
      int it = 0, gi = 0;
      for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++gi) {
        bt_static_for<0, NI>([&](auto I_) {
          constexpr int I = decltype(I_)::value;
          constexpr int kind = IT::kind(I), a = IT::arg_a(I), hb = IT::arg_b(I);
          const int slot = it & 1;
          if constexpr (kind == 0) {
            constexpr int r = a, h = hb, nh = IT::nh(h);
            BT_TRACE(it, 3);
            mbar_wait(&s_full[slot], (uint32_t)((it >> 1) & 1));
            if (h == 0) mbar_wait(dz_empty, (uint32_t)(((gi * NR + r) & 1) ^ 1));
            tc_fence_after();
            BT_TRACE(it, 4);
            const uint32_t a_z = smn_lo0 + (uint32_t)slot * (32768u >> 4);
            if (elect_one()) {
              // dz[(s, v_l), c] (+)= item[(s, v_l), (j_l, c')] . W^T[(4h + j_l, c'), c]
              constexpr uint32_t idz = make_idesc_bf16(128, 32, true, false);
#pragma unroll
              for (int ks = 0; ks < 2 * nh; ++ks)
                bt_mma(tmem_base + TDZ, a_z + (uint32_t)ks * 16u, smn_hi, wt_lo + (uint32_t)(4 * h * 4 + 2 * ks) * 32u, wt_hi,
                       idz, (h == 0 && ks == 0) ? 0u : 1u);
              if (h == NHALF - 1) umma_commit(dz_full);
              umma_commit(&s_empty[slot]);
            }
            __syncwarp();
            BT_TRACE(it, 5);
          } else {
            // U items are consumed by issuer W alone; this warp still waits for the slot's data phase before it gives its
            // share of the "slot free" count - an early arrival could complete the PREVIOUS item's phase of this slot
            mbar_wait(&s_full[slot], (uint32_t)((it >> 1) & 1));
            if (elect_one()) mbar_arrive(&s_empty[slot]);
            __syncwarp();
          }
          ++it;
        });
      }
    } else {
      // ---- issuer W: GEMM W (D items) and the Q products (U items) out of staging slot `slot`
      const uint32_t sk_lo0 = lo_of(sbase + L.ring_off, 2048u), sk_hi = hi_of(128u);        // staged item, K-major
      const uint32_t uk_lo0 = lo_of(sbase + L.ring_off, (uint32_t)NPD * 16u), uk_hi = hi_of(128u);   // staged U5 / U6, K-major
      const uint32_t zmn_base = (sbase + L.z_off) >> 4, zmn_hi = hi_of(4u * NPD * 16u);     // z, MN-major (LBO per MMA)
      const uint32_t dhk_lo0 = lo_of(sbase + L.dh_off, (uint32_t)KW * 16u), dhk_hi = hi_of(128u);   // dh, K-major (Q products)
      int it = 0, gi = 0;
      for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++gi) {
        bt_static_for<0, NI>([&](auto I_) {
          constexpr int I = decltype(I_)::value;
          constexpr int kind = IT::kind(I), a = IT::arg_a(I), hb = IT::arg_b(I);
          const int slot = it & 1, b = gi & 1;
          mbar_wait(&s_full[slot], (uint32_t)((it >> 1) & 1));
          if (kind == 0) mbar_wait(z_full, (uint32_t)(gi & 1));
          tc_fence_after();
          const uint32_t so = (uint32_t)slot * (32768u >> 4);
          const uint32_t a_w = sk_lo0 + so, a_u = uk_lo0 + so;                          // run-time parts, outside the
          const uint32_t dlo = dhk_lo0 + (((uint32_t)b * L.dh_bytes) >> 4);             // elected blocks (see issuer A)
          const uint32_t not_first = gi == 0 ? 0u : 1u;
          if (elect_one()) {
            if constexpr (kind == 0) {
              constexpr int r = a, h = hb, rs = IT::rs(r);
              // dW_h[(j_l, c'), c] += item[pos, (j_l, c')]^T . z[pos, c], K = the item's positions, 16 per step = staging
              // position groups (2kk, 2kk+1) = z rows row(pg) = s NPD + 32 r + 8 vg, pg = s (rs/8) + vg
              constexpr uint32_t idw = make_idesc_bf16(128, 32, false, true);
              constexpr int gps = rs / 8;             // position groups per slab in this item
#pragma unroll
              for (int kk = 0; kk < (4 * rs) / 16; ++kk) {
                const int pg0 = 2 * kk, pg1 = 2 * kk + 1;
                const int row0 = (pg0 / gps) * NPD + 32 * r + 8 * (pg0 % gps);
                const int row1 = (pg1 / gps) * NPD + 32 * r + 8 * (pg1 % gps);
                bt_mma(tmem_base + TDW + 32u * h, a_w + (uint32_t)pg0 * 128u, sk_hi,
                       zmn_base + (uint32_t)row0 + ((uint32_t)(row1 - row0) << 16), zmn_hi, idw,
                       (r == 0 && kk == 0) ? not_first : 1u);
              }
            } else {
              // Q5[v, w] += U5_s[v, c'] dh_s[w, c'],  Q6[v, w] += U6_s[v, c'] dh_s[w, c']   (K = 32 channels of slab s)
              constexpr uint32_t ida = make_idesc_bf16(128, KW, false, false);
#pragma unroll
              for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                for (int which = 0; which < 2; ++which)
#pragma unroll
                  for (int ch = 0; ch < 2; ++ch)
                    bt_mma(tmem_base + (which == 0 ? TA5 : TA6), a_u + (uint32_t)((sl * 8 + which * 4 + 2 * ch) * NPD), uk_hi,
                           dlo + (uint32_t)(((2 * a + sl) * 4 + 2 * ch) * KW), dhk_hi, ida,
                           (a == 0 && sl == 0 && ch == 0) ? not_first : 1u);
            }
            umma_commit(&s_empty[slot]);
            if (I == NI - 1) { umma_commit(z_empty); umma_commit(&dh_empty[b]); }
          }
          __syncwarp();
          BT_TRACE(it, 6);
          ++it;
        });
      }
      if (elect_one()) umma_commit(w_full);
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== stage warps: ring accumulator -> bf16 -> staging slot =====================
    const int q = warp & 3, set = (warp - 4) >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    float db_acc = 0.f;                 // bias gradient of channel c' = lane, slab q of every group (set 0 only)
    // per-thread byte offsets inside a staging slot (everything else in the store addresses is an immediate):
    //   D item: ((q gps + vg) 128 + (set + 2k) 32 + lane) 16 = q gps 2048 + set 512 + lane 16  +  vg 2048 + k 1024
    //   U item: ((set 8 + which 4 + cg) NPD + v) 16         = set 128 NPD + v 16               +  (which 4 + cg) NPD 16
    const uint32_t tb_lane = (uint32_t)(set * 512 + lane * 16);
    const uint32_t tb_u = (uint32_t)(set * 128 * NPD + (q * 32 + lane) * 16);
    int it = 0, gi = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++gi) {
      bt_static_for<0, NI>([&](auto I_) {
        constexpr int I = decltype(I_)::value;
        constexpr int kind = IT::kind(I), a = IT::arg_a(I), hb = IT::arg_b(I);
        const int slot = it & 1;
        mbar_wait(&t_full[slot], (uint32_t)((it >> 1) & 1));
        if (warp == 4) BT_TRACE(it, 7);
        mbar_wait(&s_empty[slot], (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        if (warp == 4) BT_TRACE(it, 8);
        const uint32_t tsrc = tmem_base + lane_off + (uint32_t)slot * 128u;
        uint8_t* sdst = smem + L.ring_off + (size_t)slot * 32768u;
        if constexpr (kind == 0) {
          constexpr int r = a, h = hb, rs = IT::rs(r), nh = IT::nh(h), gps = rs / 8;
          // lane = (slab q, channel c' = lane); columns (j_l, v_l).  set 0 takes slots j_l = 0, 2; set 1 slots 1, 3
          // one 32-column load in flight at a time: two would need 64 registers (the CTA's 24 warps leave 80 per thread)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int jl = set + 2 * k;
            const bool two = set + 2 < nh;
            if (jl >= nh) {                        // (warp-uniform) nothing of this warp's in the item / second slot
              if (k == 0) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_empty[slot]);
              }
              continue;
            }
            uint32_t ra[8 * gps];
            if constexpr (rs == 32) {
              tmem_ld32_issue(tsrc + (uint32_t)(jl * 32), ra);
            } else {
#pragma unroll
              for (int vg = 0; vg < gps; ++vg) bt_tmem_ld8(tsrc + (uint32_t)(jl * rs + 8 * vg), ra + 8 * vg);
            }
            tmem_ld_wait();
            if (k == (two ? 1 : 0)) {        // this warp's last read of the slot: the accumulator may be overwritten
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&t_empty[slot]);
            }
            if (h == 0 && jl == 0) {
#pragma unroll
              for (int e = 0; e < 8 * gps; ++e) db_acc += __uint_as_float(ra[e]);
            }
#pragma unroll
            for (int vg = 0; vg < gps; ++vg)
              *reinterpret_cast<uint4*>(sdst + tb_lane + (uint32_t)(q * (gps * 2048)) + (uint32_t)(vg * 2048 + k * 1024)) =
                  bt_pack8(ra + 8 * vg);
          }
        } else {
          // lane = node v = 32 q + lane of slab 2a + set; columns (which, c')
          const int v = q * 32 + lane;
          if (q * 32 >= NPD) {                   // (warp-uniform) this quadrant holds no node rows: nothing to read
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[slot]);
          } else {
#pragma unroll
            for (int which = 0; which < 2; ++which) {
              uint32_t ra[32];
              tmem_ld32_issue(tsrc + (uint32_t)(64 * set + 32 * which), ra);
              tmem_ld_wait();
              if (which == 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_empty[slot]);
              }
              if (v < NPD) {
#pragma unroll
                for (int cg = 0; cg < 4; ++cg)
                  *reinterpret_cast<uint4*>(sdst + tb_u + (uint32_t)((which * 4 + cg) * NPD * 16)) = bt_pack8(ra + 8 * cg);
              }
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (warp == 4) BT_TRACE(it, 9);
        if (lane == 0) mbar_arrive(&s_full[slot]);
        ++it;
      });
    }
    if (set == 0) atomicAdd(dbs + lane, db_acc * inv);
  } else if (warp >= 12 && warp < 16) {
    // ===================== epilogue warps: dz (+ dz_last) -> gate backward -> dfg =====================
    const int q = warp & 3;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int gi = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++gi) {
      bt_static_for<0, NR>([&](auto R_) {
        constexpr int r = decltype(R_)::value;
        constexpr int rs = IT::rs(r);
        const int m = q * 32 + lane;
        const int s = m / rs, v = 32 * r + (m - s * rs);
        const int slab = 4 * g + s;
        const bool valid = m < 4 * rs && v < V && slab < p.slabs;
        const long long pp = (long long)slab * V + v;
        uint4 qa[4], qb[4];
        bool tailrow = false;
        long long tail_off = 0;
        {   // unconditional loads (rows without a position re-read position 0 and drop the result)
          const long long pl = (valid && !(p.debug & 2)) ? pp : 0;
          const uint4* s1 = reinterpret_cast<const uint4*>(p.a + pl * 32);
          const uint4* s2 = reinterpret_cast<const uint4*>(p.b + pl * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) { qa[j] = __ldg(s1 + j); qb[j] = __ldg(s2 + j); }
          if (valid && p.dz_last) {
            long long n, rem;
            split_pos(pp, p.RO, n, rem);
            if (rem >= p.last_begin) { tailrow = true; tail_off = (n * p.last_rows + rem - p.last_begin) * 32; }
          }
        }
        const int rc = gi * NR + r;
        if (warp == 12) BT_TRACE(gi * NI + r, 10);
        mbar_wait_park(dz_full, (uint32_t)(rc & 1));
        tc_fence_after();
        if (warp == 12) BT_TRACE(gi * NI + r, 11);
        uint4* out = reinterpret_cast<uint4*>(p.dfg + pp * 64);
        // the whole accumulator row goes to registers first and dz is handed back at once: the next node range's GEMM Z
        // waits for exactly this (a single dz buffer - TMEM is full)
        uint32_t dzr[32];
        bt_tmem_ld16(tmem_base + lane_off + TDZ, dzr);
        bt_tmem_ld16(tmem_base + lane_off + TDZ + 16u, dzr + 16);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dz_empty);
        if (warp == 12) BT_TRACE(gi * NI + r, 12);
        if (valid) {
          uint4 tq[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) tq[j] = make_uint4(0u, 0u, 0u, 0u);
          if (tailrow) {                           // the four loads of a tail row go out together (one latency, not four)
#pragma unroll
            for (int j = 0; j < 4; ++j) tq[j] = __ldg(reinterpret_cast<const uint4*>(p.dz_last + tail_off) + j);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t aw[4] = {qa[j].x, qa[j].y, qa[j].z, qa[j].w}, bw[4] = {qb[j].x, qb[j].y, qb[j].z, qb[j].w};
            const uint32_t tw[4] = {tq[j].x, tq[j].y, tq[j].z, tq[j].w};
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int w = i >> 1;
              const float av = (i & 1) ? __uint_as_float(aw[w] & 0xFFFF0000u) : __uint_as_float(aw[w] << 16);
              const float bv = (i & 1) ? __uint_as_float(bw[w] & 0xFFFF0000u) : __uint_as_float(bw[w] << 16);
              const float tv = (i & 1) ? __uint_as_float(tw[w] & 0xFFFF0000u) : __uint_as_float(tw[w] << 16);
              const float gb = fmaf(__uint_as_float(dzr[8 * j + i]), inv, tv) * bv;     // (dz inv + dz_last) b
              const float ga = gb * av;
              o[i] = bt_pack(gb * fmaf(-av, av, 1.f), fmaf(-ga, bv, ga));               // df = g b (1 - a^2), dg = g a b (1 - b)
            }
            if (!(p.debug & 8) || o[0] == 0x12345u) {
            out[2 * j] = make_uint4(o[0], o[1], o[2], o[3]);
            out[2 * j + 1] = make_uint4(o[4], o[5], o[6], o[7]);
            }
          }
        }
        if (warp == 12) BT_TRACE(gi * NI + r, 13);
      });
    }
    // ---- gradient flush: dW tiles, Q5 / Q6 (all MMAs of the CTA complete)
    // (every operand buffer behind the stacked image is idle now: dh^T, z, the weight images and the ring, 140 KB)
    float* stg_w = reinterpret_cast<float*>(smem + L.dh_off);            // [NU][32]
    float* stg_5 = stg_w + 32 * NU;                                      // [V][V], 16-byte aligned starts
    float* stg_6 = stg_5 + ((V * V + 3) & ~3);
    static_assert((size_t)32 * NU * 4 + 2 * (size_t)(NPD * NPD + 4) * 4 <= L.bar_off - L.dh_off, "flush staging does not fit");
    const int et = q * 32 + lane;
    mbar_wait_park(w_full, 0u);
    tc_fence_after();
#pragma unroll 1
    for (int t = 0; t < NHALF; ++t) {
      float vv[32];
      tmem_ld32(tmem_base + lane_off + TDW + 32u * t, vv);
      const int m = t * 128 + et;                                        // (j, c') = (4 t + et / 32, et % 32)
      if (m < NU) {
        float* dst = stg_w + (size_t)(m >> 5) * 32 * 32 + (m & 31);
#pragma unroll
        for (int c = 0; c < 32; ++c) dst[c * 32] = vv[c] * inv;
      }
    }
    if (DA) {
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {
        float* stg = which == 0 ? stg_5 : stg_6;
#pragma unroll 1
        for (int c0 = 0; c0 < KW; c0 += 16) {
          uint32_t r16[16];
          bt_tmem_ld16(tmem_base + lane_off + (which == 0 ? TA5 : TA6) + (uint32_t)c0, r16);
          tmem_ld_wait();
          if (et < V) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < V) stg[et * V + c0 + j] = __uint_as_float(r16[j]) * inv;
          }
        }
      }
    }
    asm volatile("bar.sync 2, 128;" ::: "memory");
    red_flush_1d(p.dw_mlp, stg_w, 32 * NU, et, 128);
    if (DA) {
      red_flush_1d(p.dA, stg_5, V * V, et, 128);
      red_flush_1d(p.dQ6, stg_6, V * V, et, 128);
    }
  } else if (warp < 22) {
    // ===================== prep warps: one position (slab s, node v) per thread =====================
    const int pi = warp < 4 ? warp - 1 : warp - 13;                      // 0 .. 8
    const int t = pi * 32 + lane;
    const int s = t / V, v = t - s * V;
    const bool mine = t < 4 * V;
    uint64_t sd = 0, of = 0;
    if (philox) { sd = p.rng ? __ldg(p.rng) : p.seed; of = p.rng ? p.offset + __ldg(p.rng + 1) : p.offset; }
    int gi = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++gi) {
      const int b = gi & 1;
      const int slab = 4 * g + s;
      const bool real = mine && slab < p.slabs;
      // L2 prefetch of the rows the CTA reads two groups from now (one bulk request per array, issued by one thread):
      // the row loads below then hit L2 - with HBM misses in flight for ~1 us the loads of 288 threads keep the SM's
      // load/store queue full and every shared-memory store of the stage warps waits behind them
      if (pi == 0 && lane == 0 && !(p.debug & 16)) {
        const int g2 = g + 2 * (int)gridDim.x;
        if (g2 < n_groups) {
          const long long p0 = (long long)g2 * 4 * V;
          const int ns2 = min(4, p.slabs - 4 * g2);
          const uint32_t bytes = (uint32_t)(ns2 * V) * 64u;
          const bf16* srcs[3] = {p.du + p0 * 32, p.a + p0 * 32, p.b + p0 * 32};
#pragma unroll
          for (int k = 0; k < 3; ++k)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(srcs[k]), "r"(bytes) : "memory");
        }
      }
      const long long pp = (long long)slab * V + v;
      uint4 od[4], oz[4];
      {   // unconditional (threads without a position re-read position 0 and write zeros)
        // Packed bf16x2 arithmetic: z = a . b is one HMUL2 per channel pair (the exact product rounded once - the same
        // bits as the fp32 product rounded to bf16); the fused dropout multiplies by the KEEP flag (0 / 1, exact) and the
        // scale 1 / (1 - p) is applied downstream in fp32 (`inv`: dz in the epilogue, dW / db / Q at the flush) - the
        // SIMT issue slots, not the tensor pipe, bound this kernel.
        const long long pl = real ? pp : 0;
        if (p.debug & 4) __nanosleep((unsigned)pi * 250u);
        const uint4* s0 = reinterpret_cast<const uint4*>(p.du + ((p.debug & 1) ? 0 : pl) * 32);
        const uint4* s1 = reinterpret_cast<const uint4*>(p.a + ((p.debug & 1) ? 0 : pl) * 32);
        const uint4* s2 = reinterpret_cast<const uint4*>(p.b + ((p.debug & 1) ? 0 : pl) * 32);
        uint4 qd[4], qa[4], qb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { qd[j] = __ldg(s0 + j); qa[j] = __ldg(s1 + j); qb[j] = __ldg(s2 + j); }
        if (p.debug & 1) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { qd[j] = make_uint4(0x3F803F80u, 0x3F003F00u, 0x3F803F80u, 0x3F003F00u); qa[j] = qd[j]; qb[j] = qd[j]; }
        }
        auto mul2 = [](uint32_t x, uint32_t y) -> uint32_t {
          __nv_bfloat162 r = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&x), *reinterpret_cast<__nv_bfloat162*>(&y));
          return *reinterpret_cast<uint32_t*>(&r);
        };
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t dw[4] = {qd[j].x, qd[j].y, qd[j].z, qd[j].w};
          const uint32_t aw[4] = {qa[j].x, qa[j].y, qa[j].z, qa[j].w};
          const uint32_t bw[4] = {qb[j].x, qb[j].y, qb[j].z, qb[j].w};
          uint32_t mw[4] = {0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u};      // bf16x2 (1, 1)
          if (p.mask) {
            const uint4 mk = __ldg(reinterpret_cast<const uint4*>(p.mask + pl * 32) + j);
            mw[0] = mk.x; mw[1] = mk.y; mw[2] = mk.z; mw[3] = mk.w;
          }
          uint32_t o[4], z[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) { o[t] = mul2(dw[t], mw[t]); z[t] = mul2(aw[t], bw[t]); }
          od[j] = make_uint4(o[0], o[1], o[2], o[3]);
          oz[j] = make_uint4(z[0], z[1], z[2], z[3]);
        }
        if (philox) {
          // keep flags of the row's 32 channels: Philox word i of call h = channels 16h + 4i .. + 3, one byte each
          // (the stream of dropout16); byte >= thr on four bytes at once, expanded to bf16x2 multipliers 0 / 1
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint4 rr = philox4x32(sd, of, (uint64_t)(pl * 2 + h));
            const uint32_t w[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t Hb = 0x80808080u;
              uint32_t y;
              if (thr <= 128u) y = (w[i] | ((w[i] | Hb) - thr * 0x01010101u)) & Hb;
              else y = w[i] & (((w[i] & ~Hb) | Hb) - (thr - 128u) * 0x01010101u) & Hb;
              const uint32_t f = y >> 7;                                       // 0 / 1 in every byte
              const uint32_t m0 = __byte_perm(f, 0u, 0x4140) * 0x3F80u;        // channels 4i, 4i+1
              const uint32_t m1 = __byte_perm(f, 0u, 0x4342) * 0x3F80u;        // channels 4i+2, 4i+3
              // channel 16h + 4i + k lives in uint4 j = 2h + (i >> 1), word 2 (i & 1) + (k >> 1)
              uint32_t* ow = reinterpret_cast<uint32_t*>(&od[2 * h + (i >> 1)]) + 2 * (i & 1);
              ow[0] = mul2(ow[0], m0);
              ow[1] = mul2(ow[1], m1);
            }
          }
        }
        if (!real) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { od[j] = make_uint4(0u, 0u, 0u, 0u); oz[j] = od[j]; }
        }
      }
      if (warp == 1) BT_TRACE(gi * NI + 4, 10);
      mbar_wait_park(&dh_empty[b], (uint32_t)(((gi >> 1) & 1) ^ 1));
      if (mine) {
        uint8_t* dd = smem + L.dh_off + (size_t)b * L.dh_bytes + ((size_t)(s * 4) * KW + v) * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dd + (size_t)j * KW * 16) = od[j];
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dh_full[b]);
      if (warp == 1) BT_TRACE(gi * NI + 4, 11);
      mbar_wait_park(z_empty, (uint32_t)((gi & 1) ^ 1));
      if (warp == 1) BT_TRACE(gi * NI + 4, 12);
      if (mine) {
        uint8_t* zd = smem + L.z_off + ((size_t)s * NPD + v) * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(zd + (size_t)j * 4 * NPD * 16) = oz[j];
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(z_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) atomicAdd(p.db_mlp + lane, dbs[lane]);
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
static int bt_npd(int V) { return ((V + 7) / 8) * 8; }

int gcn_bwd_t_image_elems(int V, int n_mats) {
  const BtGeom G(n_mats, bt_npd(V));
  return (G.KW / 8) * G.NTOT * 8;
}

// kernel instances (NM, NPD): every graph of 65..72 nodes (the 67-county configurations) and the small shapes the
// tests use; other shapes keep the node-major kernel (gcn_fused_bwd.cu)
#define BT_INSTANCES(X) X(6, 72) X(4, 72) X(2, 72) X(6, 64) X(6, 40) X(6, 32) X(6, 24) X(6, 8) X(4, 40) X(2, 40)

int gcn_bwd_t_supported(int V, int n_mats, bool has_da) {
  if (V < 1 || 4 * V > 32 * BT_PREP_WARPS) return 0;
  const int npd = bt_npd(V);
  bool inst = false;
#define BT_HAS(NM_, NPD_) if (n_mats == NM_ && npd == NPD_) inst = true;
  BT_INSTANCES(BT_HAS)
#undef BT_HAS
  if (!inst) return 0;
  const BtGeom G(n_mats, npd);
  if (G.KW > 80) return 0;                                  // Q accumulators: 80 TMEM columns each
  return bt_layout(G, has_da).total <= 227u * 1024u ? 1 : 0;
}

int launch_gcn_bwd_t(GcnBwdParams& p, cudaStream_t st) {
  if (p.slabs <= 0) return 0;
  const bool has_da = p.sa >= 0;
  GWN_REQUIRE(gcn_bwd_t_supported(p.V, p.n_mats, has_da) && p.w_src && p.mats_bt,
              "gcn_bwd_t: unsupported shape (V=%d, %d matrices)", p.V, p.n_mats);
  GWN_REQUIRE((long long)p.slabs * p.V < (1ll << 31), "gcn_bwd_t: too many positions");
  GWN_REQUIRE(!has_da || (p.dA && p.dQ6 && p.sa < p.n_mats / 2), "gcn_bwd_t: bad support-gradient arguments");
  const int npd = bt_npd(p.V);
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    GWN_CUDA(cudaGetDevice(&dev));
    GWN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int groups = (p.slabs + 3) / 4;
  const int grid = groups < sms ? groups : sms;
#define BT_CASE(NM_, NPD_)                                                                                              \
  if (p.n_mats == NM_ && npd == NPD_) {                                                                                 \
    static bool attr = false;                                                                                           \
    if (!attr) {                                                                                                        \
      GWN_CUDA(cudaFuncSetAttribute(gcn_bwd_t_kernel<NM_, NPD_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      GWN_CUDA(cudaFuncSetAttribute(gcn_bwd_t_kernel<NM_, NPD_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));  \
      attr = true;                                                                                                      \
    }                                                                                                                   \
    constexpr BtGeom G_(NM_, NPD_);                                                                                     \
    if (has_da) GWN_CUDA(launch_pdl(gcn_bwd_t_kernel<NM_, NPD_, true>, dim3(grid), dim3(BT_THREADS), bt_layout(G_, true).total, st, p)); \
    else GWN_CUDA(launch_pdl(gcn_bwd_t_kernel<NM_, NPD_, false>, dim3(grid), dim3(BT_THREADS), bt_layout(G_, false).total, st, p));  \
    GWN_LAUNCHED();                                                                                                     \
    return 0;                                                                                                           \
  }
  BT_INSTANCES(BT_CASE)
#undef BT_CASE
  GWN_REQUIRE(false, "gcn_bwd_t: no kernel instance for %d matrices, V=%d", p.n_mats, p.V);
  return -1;
}

}  // namespace gwn
