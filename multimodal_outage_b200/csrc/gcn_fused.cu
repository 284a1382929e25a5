// Fused diffusion graph convolution, forward (see gcn_fused.cuh for the math and the data flow).
//
// One persistent CTA per SM walks slabs (slab = one (n,t) pair = V nodes x 32 channels).  16 warps; a warp may
// only touch the TMEM lane quadrant (warp % 4), and node w of a slab lives in lane w, so the roles are laid out
// per quadrant q (= the warp scheduler that has to issue that quadrant's epilogue work):
//   q0: warps 0,4 = stage A/B, 8,12 = epilogue group 0/1        q1: warps 1,5 / 9,13 likewise
//   q2: warp 2 = producer, 6 = stage, 10,14 = epilogue 0/1      q3: warp 3 = MMA issuer, 7 = stage, 11,15 = epilogue
//   producer : z slab -> smem ring (cp.async, 16-byte pieces, K-major A-operand layout of GEMM 1)
//   MMA      : lane 0 issues GEMM 1 of slab k, then GEMM 2 of slab k-1 (software pipelined)
//   stage    : TMEM(U) -> bf16 -> smem ring (B operand of GEMM 2); U_0 + bias -> TMEM(h) (tcgen05.st).  The two
//              stage warps of q0/q1 split the 32-column chunks (even / odd).
//   epilogue : TMEM(h) -> dropout, residual (folded BN), bf16 store, BN statistics; the two groups alternate
//              slabs (each owns one of the two h accumulators).
// Quadrants without node rows (V <= 64: q2, q3; V <= 96: q3) run no stage / epilogue work at all.
// TMEM (512 columns): U accumulators at columns [0,224) and [256,480), h accumulators at [224,256), [480,512).
#include "gcn_fused.cuh"
#include "tc.cuh"
#include "tc_gemm_impl.cuh"   // warp_column_sums

namespace gwn {

constexpr int GF_ZST = 4;          // z ring stages
constexpr int GF_UST = 3;          // U ring stages
constexpr int GF_THREADS = 32 * 16;
constexpr int GF_PRODUCER = 2, GF_MMA = 3;

struct GfLayout {
  uint32_t mat_bytes, w_off, w_bytes, z_off, z_piece, z_stage, u_off, u_slot, u_stage, bar_off, total;
};
__host__ __device__ inline GfLayout gf_layout(int Kp, int n_mats) {
  GfLayout L;
  L.mat_bytes = (uint32_t)(Kp / 8) * (uint32_t)Kp * 16u;       // [Kp/8][Kp rows][16 B]
  L.w_off = (uint32_t)n_mats * L.mat_bytes;
  L.w_bytes = 4u * 32u * (uint32_t)(1 + n_mats) * 16u;         // [4][NU][16 B]
  L.z_off = L.w_off + L.w_bytes;
  L.z_piece = (uint32_t)(Kp + 2) * 16u;                        // +2 rows: the 4 channel groups hit different banks
  L.z_stage = 4u * L.z_piece;
  L.u_off = L.z_off + GF_ZST * L.z_stage;
  L.u_slot = 4u * (uint32_t)Kp * 16u;                          // [4 cg][Kp nodes][16 B]
  L.u_stage = (uint32_t)n_mats * L.u_slot;
  L.bar_off = (L.u_off + GF_UST * L.u_stage + 2048u + 127u) & ~127u;   // 2 KB slack: M=128 operand rows past Kp
  L.total = L.bar_off + 256u + 640u;                           // barriers + {bias, scale, shift} + statistics scratch
  return L;
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t r[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void unpack_bf16x8(const uint4& q, float v[8]) {
  v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xFFFF0000u);
  v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xFFFF0000u);
  v[4] = __uint_as_float(q.z << 16); v[5] = __uint_as_float(q.z & 0xFFFF0000u);
  v[6] = __uint_as_float(q.w << 16); v[7] = __uint_as_float(q.w & 0xFFFF0000u);
}

#define GF_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && k < 64 && lane == 0) p.trace[k * 8 + (slot)] = clock64(); } while (0)

// NM = resident hop matrices, KSTEPS = Kp/16: compile-time so the MMA issue loop unrolls to immediates (the
// single issuing thread is the critical resource: ~110 cycles per MMA with runtime loop bounds, ~59 unrolled).
template <int NM, int KSTEPS>
__global__ void __launch_bounds__(GF_THREADS, 1) gcn_fwd_kernel(const __grid_constant__ GcnFwdParams p) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#define GF_MARK(i) do { if (p.trace && blockIdx.x == 0 && tid == 0) p.trace[63 * 8 + (i)] = clock64(); } while (0)
  GF_MARK(0);
  constexpr int Kp = 16 * KSTEPS, nm = NM, NU = 32 * (1 + NM);
  const int V = p.V;
  const GfLayout L = gf_layout(Kp, nm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* z_full = bars;              // [4]
  uint64_t* z_empty = bars + 4;         // [4]
  uint64_t* ut_full = bars + 8;         // [2]
  uint64_t* ut_empty = bars + 10;       // [2]
  uint64_t* us_full = bars + 12;        // [3]
  uint64_t* us_empty = bars + 15;       // [3]
  uint64_t* ht_full = bars + 18;        // [2]
  uint64_t* ht_empty = bars + 20;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
  float* cst = reinterpret_cast<float*>(smem + L.bar_off + 256);     // [32] bias | [32] scale | [32] shift

  const int nq_stage = (Kp + 31) / 32;          // quadrants holding (zero-padded) node rows of U
  const int nq_epi = (V + 31) / 32;             // quadrants holding real node rows
  const int n_stage_warps = 2 * min(nq_stage, 2) + max(0, nq_stage - 2);

  if (tid == 0) {
    for (int i = 0; i < GF_ZST; ++i) { mbar_init(&z_full[i], 32); mbar_init(&z_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ut_full[i], 1); mbar_init(&ut_empty[i], 32 * n_stage_warps);
      mbar_init(&ht_full[i], 1); mbar_init(&ht_empty[i], 32 * nq_epi);
    }
    for (int i = 0; i < GF_UST; ++i) { mbar_init(&us_full[i], 32 * n_stage_warps); mbar_init(&us_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == GF_MMA) tmem_alloc(tmem_slot, 512);
  {  // resident operands: support images (rows < Kp of each K piece) and the mlp weight image
    constexpr int per_piece = Kp;                   // 16-byte rows kept per K piece
    constexpr int pieces = Kp / 8;
    {   // all matrices in one flat loop, 4 independent 16-byte loads in flight per thread
      constexpr int per_mat = pieces * per_piece, total = NM * per_mat;
      const uint4* src0 = reinterpret_cast<const uint4*>(p.mats);
      uint4* dst0 = reinterpret_cast<uint4*>(smem);
      for (int i0 = tid; i0 < total; i0 += 4 * GF_THREADS) {
        uint4 v[4]; int di[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = i0 + q * GF_THREADS;
          di[q] = -1;
          if (i < total) {
            const int m = i / per_mat, e = i - m * per_mat, kc = e / per_piece, r = e - kc * per_piece;
            v[q] = __ldg(src0 + (size_t)p.mat_src[m] * pieces * 128 + kc * 128 + r);
            di[q] = m * (int)(L.mat_bytes / 16) + kc * per_piece + r;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (di[q] >= 0) dst0[di[q]] = v[q];
      }
    }
    if (p.w_src) {       // bf16 UMMA image of the mlp weight built here: (k = c, n = (j, c')) = W[j*32 + c][c']
      bf16* wimg = reinterpret_cast<bf16*>(smem + L.w_off);
      constexpr int T4 = 8 * NU, ITS = (T4 + GF_THREADS - 1) / GF_THREADS;      // 16-byte loads, all issued first
      float4 wv[ITS];
#pragma unroll
      for (int it = 0; it < ITS; ++it) {
        const int i = tid + it * GF_THREADS;
        wv[it] = i < T4 ? __ldg(reinterpret_cast<const float4*>(p.w_src) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int it = 0; it < ITS; ++it) {
        const int i = tid + it * GF_THREADS;
        if (i < T4) {
          const int co = (i & 7) * 4, c = (i >> 3) & 31, n = (i >> 8) * 32 + co;
          bf16* d = wimg + ((c >> 3) * NU + n) * 8 + (c & 7);
          d[0] = __float2bfloat16_rn(wv[it].x); d[8] = __float2bfloat16_rn(wv[it].y);
          d[16] = __float2bfloat16_rn(wv[it].z); d[24] = __float2bfloat16_rn(wv[it].w);
        }
      }
    } else {
      const uint4* wsrc = reinterpret_cast<const uint4*>(p.w_img);
      uint4* wdst = reinterpret_cast<uint4*>(smem + L.w_off);
      for (int i = tid; i < (int)(L.w_bytes / 16); i += GF_THREADS) wdst[i] = __ldg(wsrc + i);
    }
    if (tid >= 64 && tid < 128) cst[96 + tid - 64] = 0.f;     // CTA-level (sum, sum^2) scratch
    if (tid < 32) {
      cst[tid] = __ldg(p.bias + tid);
      cst[32 + tid] = p.scale ? __ldg(p.scale + tid) : 1.f;
      cst[64 + tid] = p.scale ? __ldg(p.shift + tid) : 0.f;
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  const int quad = warp & 3, wq = warp >> 2;      // quadrant and index of this warp inside its quadrant
  GF_MARK(1);

  if (warp == GF_PRODUCER) {
    // ===================== producer =====================
    int k = 0;
    for (long long slab = blockIdx.x; slab < p.slabs; slab += gridDim.x, ++k) {
      const int zs = k % GF_ZST;
      mbar_wait(&z_empty[zs], (uint32_t)(((k / GF_ZST) & 1) ^ 1));
      const bf16* src = p.z + slab * V * 32;
      const uint32_t dst = sbase + L.z_off + (uint32_t)zs * L.z_stage;
      for (int i = lane; i < Kp * 4; i += 32) {
        const int v = i >> 2, cg = i & 3;
        const bool ok = v < V;
        cp_async16(dst + (uint32_t)cg * L.z_piece + (uint32_t)v * 16u, ok ? src + v * 32 + cg * 8 : p.z, ok ? 16u : 0u);
      }
      cp_async_commit();
      GF_TRACE(0);
      if (k >= GF_ZST - 2) {       // ZST-2 groups may stay in flight
        cp_async_wait<GF_ZST - 2>();
        fence_proxy_async();
        mbar_arrive(&z_full[(k - (GF_ZST - 2)) % GF_ZST]);
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int kk = (k >= GF_ZST - 2 ? k - (GF_ZST - 2) : 0); kk < k; ++kk) mbar_arrive(&z_full[kk % GF_ZST]);
  } else if (warp == GF_MMA) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop (so descriptors live in uniform registers and the MMAs issue back to back); only
    // the tcgen05 instructions themselves run on one elected lane.
    {
      const uint32_t idescU = make_idesc_bf16(128, NU, false, false);
      const uint32_t idescH = make_idesc_bf16(128, 32, false, true);
      // descriptor templates (start address added per use; addresses are < 256 KB so no field carry)
      const uint64_t adz = make_smem_desc(0, L.z_piece, 128u);                 // z slab: K-major, K piece stride z_piece
      const uint64_t bdw = make_smem_desc(0, (uint32_t)NU * 16u, 128u);        // W image: K-major
      const uint64_t adm = make_smem_desc(0, (uint32_t)Kp * 16u, 128u);        // support image: K-major
      const uint64_t bdu = make_smem_desc(0, 128u, (uint32_t)Kp * 16u);        // U slot: MN-major
      auto issue_hops = [&](int kk) {
        const int us = kk % GF_UST, hb = kk & 1;
        mbar_wait(&us_full[us], (uint32_t)((kk / GF_UST) & 1));
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)hb * 256u + 224u;
        const uint32_t ub = sbase + L.u_off + (uint32_t)us * L.u_stage;
        constexpr uint32_t mat_bytes = (uint32_t)(Kp / 8) * (uint32_t)Kp * 16u, u_slot = 4u * (uint32_t)Kp * 16u;
        const uint64_t ad0 = adm + (uint64_t)(sbase >> 4), bd0 = bdu + (uint64_t)(ub >> 4);
        if (elect_one()) {
#pragma unroll
          for (int m = 0; m < NM; ++m) {
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
              const uint64_t ad = ad0 + (uint64_t)(((uint32_t)m * mat_bytes + (uint32_t)(2 * ks) * (uint32_t)Kp * 16u) >> 4);
              const uint64_t bd = bd0 + (uint64_t)(((uint32_t)m * u_slot + (uint32_t)ks * 256u) >> 4);
              umma_bf16(d, ad, bd, idescH, 1u);
            }
          }
          umma_commit(&us_empty[us]);
          umma_commit(&ht_full[hb]);
        }
        __syncwarp();
      };
      int k = 0;
      for (long long slab = blockIdx.x; slab < p.slabs; slab += gridDim.x, ++k) {
        const int zs = k % GF_ZST, ub = k & 1;
        mbar_wait(&z_full[zs], (uint32_t)((k / GF_ZST) & 1));
        mbar_wait(&ut_empty[ub], (uint32_t)(((k >> 1) & 1) ^ 1));
        tc_fence_after();
        GF_TRACE(1);
        const uint32_t za = sbase + L.z_off + (uint32_t)zs * L.z_stage;
        const uint32_t wa = sbase + L.w_off;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t ad = adz + (uint64_t)((za + (uint32_t)(2 * ks) * L.z_piece) >> 4);
            const uint64_t bd = bdw + (uint64_t)((wa + (uint32_t)(2 * ks) * (uint32_t)NU * 16u) >> 4);
            umma_bf16(tmem_base + (uint32_t)ub * 256u, ad, bd, idescU, ks == 0 ? 0u : 1u);
          }
          umma_commit(&z_empty[zs]);
          umma_commit(&ut_full[ub]);
        }
        __syncwarp();
        GF_TRACE(2);
        if (k > 0) issue_hops(k - 1);
        GF_TRACE(3);
      }
      if (k > 0) issue_hops(k - 1);
    }
    __syncwarp();
  } else if ((quad < 2 && wq < 2) || (quad >= 2 && wq == 1)) {
    // ===================== stage warps: U -> smem (bf16), U_0 + bias -> h accumulator =====================
    if (quad < nq_stage) {
      const int row = quad * 32 + lane;
      const int first = (quad < 2) ? wq : 0, step = (quad < 2) ? 2 : 1;   // chunk split between the two warps of q0/q1
      const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
      int k = 0;
      for (long long slab = blockIdx.x; slab < p.slabs; slab += gridDim.x, ++k) {
        const int ub = k & 1, us = k % GF_UST, hb = k & 1;
        mbar_wait(&ut_full[ub], (uint32_t)((k >> 1) & 1));
        mbar_wait(&us_empty[us], (uint32_t)(((k / GF_UST) & 1) ^ 1));
        if (warp == 0) GF_TRACE(4);
        if (first == 0) mbar_wait(&ht_empty[hb], (uint32_t)(((k >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t tU = tmem_base + lane_off + (uint32_t)ub * 256u;
        uint8_t* ust = smem + L.u_off + (size_t)us * L.u_stage;
        uint32_t r[32];
        for (int j = first; j <= nm; j += step) {
          tmem_ld32_issue(tU + (uint32_t)j * 32u, r);
          tmem_ld_wait();
          if (j == 0) {
#pragma unroll
            for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) + cst[c]);
            tmem_st32(tmem_base + lane_off + (uint32_t)hb * 256u + 224u, r);
          } else if (row < Kp) {
            uint8_t* dst = ust + (size_t)(j - 1) * L.u_slot + (size_t)row * 16;
#pragma unroll
            for (int cg = 0; cg < 4; ++cg) {
              uint4 pk;
              pk.x = pack_bf16(__uint_as_float(r[8 * cg]), __uint_as_float(r[8 * cg + 1]));
              pk.y = pack_bf16(__uint_as_float(r[8 * cg + 2]), __uint_as_float(r[8 * cg + 3]));
              pk.z = pack_bf16(__uint_as_float(r[8 * cg + 4]), __uint_as_float(r[8 * cg + 5]));
              pk.w = pack_bf16(__uint_as_float(r[8 * cg + 6]), __uint_as_float(r[8 * cg + 7]));
              *reinterpret_cast<uint4*>(dst + (size_t)cg * Kp * 16) = pk;
            }
          }
        }
        if (first == 0) tmem_st_wait();
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&ut_empty[ub]);
        mbar_arrive(&us_full[us]);
        if (warp == 0) GF_TRACE(5);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int grp = (quad < 2) ? wq - 2 : wq - 2;       // wq 2 -> group 0, wq 3 -> group 1
    const int w = quad * 32 + lane;
    if (quad < nq_epi) {
      const bool valid = w < V;
      float sa[32], sb[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) { sa[c] = 0.f; sb[c] = 0.f; }
      uint64_t sd = 0, of = 0;
      const bool philox = (p.mask == nullptr) && p.drop_p > 0.f;
      if (philox) { sd = p.rng ? __ldg(p.rng) : p.seed; of = p.rng ? p.offset + __ldg(p.rng + 1) : p.offset; }
      const float4* sc4 = reinterpret_cast<const float4*>(cst + 32);
      const float4* sh4 = reinterpret_cast<const float4*>(cst + 64);
      int k = 0;
      for (long long slab = blockIdx.x; slab < p.slabs; slab += gridDim.x, ++k) {
        if ((k & 1) != grp) continue;
        const int hb = grp;
        // the residual row is requested BEFORE waiting for the accumulator, so the
        // HBM round trip overlaps the MMA pipeline latency
        const long long pp = slab * V + w;
        uint4 rr[4];
        if (valid) {
          long long n, rem;
          split_pos(pp, p.RO, n, rem);
          const uint4* rp = reinterpret_cast<const uint4*>(p.u_prev + (n * p.RI + rem + p.crop) * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) rr[j] = __ldg(rp + j);
        }
        mbar_wait(&ht_full[hb], (uint32_t)((k >> 1) & 1));
        if (quad == 0) GF_TRACE(6);
        tc_fence_after();
        float h[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)hb * 256u + 224u, h);
        tc_fence_before();
        mbar_arrive(&ht_empty[hb]);
        if (valid) {
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            float m16[16];
            if (philox) dropout16(sd, of, (uint64_t)(pp * 2 + jj), p.drop_p, m16);
#pragma unroll
            for (int jh = 0; jh < 2; ++jh) {
              const int j = 2 * jj + jh;
              float m[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f}, r[8];
              if (p.mask) unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(p.mask + pp * 32) + j), m);
              else if (philox) {
#pragma unroll
                for (int i = 0; i < 8; ++i) m[i] = m16[8 * jh + i];
              }
              unpack_bf16x8(rr[j], r);
              const float4 s0 = sc4[2 * j], s1 = sc4[2 * j + 1], t0 = sh4[2 * j], t1 = sh4[2 * j + 1];
              const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
              const float sh[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) h[8 * j + i] = fmaf(h[8 * j + i], m[i], fmaf(r[i], sc[i], sh[i]));
            }
          }
          bf16* up = p.u + pp * 32;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 pk;
            pk.x = pack_bf16(h[8 * j], h[8 * j + 1]); pk.y = pack_bf16(h[8 * j + 2], h[8 * j + 3]);
            pk.z = pack_bf16(h[8 * j + 4], h[8 * j + 5]); pk.w = pack_bf16(h[8 * j + 6], h[8 * j + 7]);
            *reinterpret_cast<uint4*>(up + 8 * j) = pk;
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) { sa[c] += h[c]; sb[c] = fmaf(h[c], h[c], sb[c]); }
        }
        if (quad == 0) GF_TRACE(7);
      }
      const float s1 = warp_column_sums(sa, lane);
      const float s2 = warp_column_sums(sb, lane);
      atomicAdd(cst + 96 + lane, s1);            // CTA-level pre-reduction: one global atomic per statistic per CTA
      atomicAdd(cst + 128 + lane, s2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    atomicAdd(p.stats + lane, (double)cst[96 + lane]);
    atomicAdd(p.stats + 32 + lane, (double)cst[128 + lane]);
  }
  GF_MARK(2);
  if (warp == GF_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int gcn_fused_supported(int V, int n_mats) {
  if (V < 1 || V > 80 || (n_mats != 2 && n_mats != 4 && n_mats != 6)) return 0;   // kernel instances: see launch_gcn_fwd
  const int Kp = ((V + 15) / 16) * 16;
  return gf_layout(Kp, n_mats).total <= 227u * 1024u ? 1 : 0;
}

int launch_gcn_fwd(GcnFwdParams& p, cudaStream_t st) {
  if (p.slabs <= 0) return 0;
  p.Kp = ((p.V + 15) / 16) * 16;
  GWN_REQUIRE(gcn_fused_supported(p.V, p.n_mats), "gcn_fwd: V=%d with %d resident matrices does not fit on chip", p.V,
              p.n_mats);
  const GfLayout L = gf_layout(p.Kp, p.n_mats);
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    GWN_CUDA(cudaGetDevice(&dev));
    GWN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = p.slabs < sms ? p.slabs : sms;
  const int ks = p.Kp / 16;
#define GF_CASE(NM_, KS_)                                                                                         \
  if (p.n_mats == NM_ && ks == KS_) {                                                                             \
    static bool attr = false;                                                                                     \
    if (!attr) {                                                                                                  \
      GWN_CUDA(cudaFuncSetAttribute(gcn_fwd_kernel<NM_, KS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      attr = true;                                                                                                \
    }                                                                                                             \
    gcn_fwd_kernel<NM_, KS_><<<grid, GF_THREADS, L.total, st>>>(p);                                               \
    GWN_LAUNCHED();                                                                                               \
    return 0;                                                                                                     \
  }
  GF_CASE(6, 5) GF_CASE(4, 5) GF_CASE(2, 5) GF_CASE(6, 4) GF_CASE(4, 4) GF_CASE(2, 4) GF_CASE(6, 3) GF_CASE(4, 3)
  GF_CASE(2, 3) GF_CASE(6, 2) GF_CASE(4, 2) GF_CASE(2, 2) GF_CASE(6, 1) GF_CASE(4, 1) GF_CASE(2, 1)
#undef GF_CASE
  GWN_REQUIRE(false, "gcn_fwd: no kernel instance for %d matrices, Kp=%d", p.n_mats, p.Kp);
  return -1;
}

}  // namespace gwn
