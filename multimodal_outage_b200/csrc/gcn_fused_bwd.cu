// Fused diffusion graph convolution, backward (math in gcn_fused.cuh).  Same machinery as the forward kernel:
// one persistent CTA per SM walks slabs (one (n,t) pair = V nodes x 32 channels), every support image and the
// mlp weight image resident in shared memory, intermediates in TMEM / shared memory.
//
// Per slab k (buffer b = k & 1):
//   prep     : du, a, b rows -> dh = du.mask (bf16) into slot 0 of concat buffer b, z = a.b into z tile b  (+ db sums)
//   MMA      : GEMM H  dU_j = M_j^T dh        (A = transposed support image, B = slot 0, D = 32 TMEM columns per j)
//   stage    : TMEM(dU_j) -> bf16 -> slots 1..H of concat buffer b
//   MMA      : GEMM Z  dz = [dh | dU] W^T      (A = concat buffer b as K-major [node][(j,c')], B = W^T image)
//              GEMM W  dW += [dh | dU]^T z     (A = the SAME bytes as MN-major [(j,c')][node], B = z tile; the
//                                               accumulator stays in TMEM for the CTA's whole share of slabs)
//   epilogue : TMEM(dz) + dz_last -> gate backward with a, b -> dfg (bf16, 128 B per node)
// TMEM columns: dU [0, 32H), dz [192, 224), dW tiles [224, 256) and [256, 288).
// Warps (16): q0/q1: stage A (w0,w1), stage B (w4,w5), epilogue (w8,w9); q2: stage w6,w14, epilogue w10; q3: MMA w3 (stage w7,w11 when Kp > 96).
// The three prep warps (w12, w13, w15) are not tied to a TMEM lane quadrant; they finish ALL their arithmetic (mask,
// products, bf16 packing) in registers BEFORE waiting for the buffer, so a freed buffer is refilled in ~100 cycles.
#include "gcn_fused.cuh"
#include "tc.cuh"
#include "tc_gemm_impl.cuh"   // warp_column_sums

namespace gwn {

constexpr int GB_THREADS = 32 * 16;
constexpr int GB_MMA = 3;

struct GbLayout {
  uint32_t mat_bytes, w_off, w_bytes, cat_off, slot_bytes, cat_bytes, z_off, z_bytes, bar_off, total;
  uint32_t fwd_off, w56_off, u6_off, t1_off;     // dA extras (has_da): forward image, W56 image, U6 / T1 staging x2
};
__host__ __device__ inline GbLayout gb_layout(int Kp, int n_mats, bool has_da = false) {
  GbLayout L;
  L.mat_bytes = (uint32_t)(Kp / 8) * (uint32_t)Kp * 16u;       // [Kp/8][Kp rows][16 B]
  L.w_off = (uint32_t)n_mats * L.mat_bytes;
  L.w_bytes = 4u * (uint32_t)(1 + n_mats) * 32u * 16u;         // [4(1+H)][32][16 B]
  L.cat_off = L.w_off + L.w_bytes;
  L.slot_bytes = 4u * (uint32_t)Kp * 16u;                      // [4 cg][Kp nodes][16 B]
  L.cat_bytes = (uint32_t)(1 + n_mats) * L.slot_bytes;
  L.z_off = L.cat_off + 2u * L.cat_bytes;
  L.z_bytes = L.slot_bytes;
  uint32_t end = L.z_off + 2u * L.z_bytes;
  L.fwd_off = L.w56_off = L.u6_off = L.t1_off = end;
  if (has_da) {
    L.u6_off = end; end += 2u * L.slot_bytes;
    L.t1_off = end; end += 2u * L.slot_bytes;
    L.fwd_off = end; end += L.mat_bytes;
    L.w56_off = end; end += 4u * 64u * 16u;
  }
  L.bar_off = (end + 4096u + 127u) & ~127u;     // slack: M=128 operand rows past Kp / past slot H
  L.total = L.bar_off + 256u;
  return L;
}

__device__ __forceinline__ uint32_t gb_pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void gb_unpack8(const uint4& q, float v[8]) {
  v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xFFFF0000u);
  v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xFFFF0000u);
  v[4] = __uint_as_float(q.z << 16); v[5] = __uint_as_float(q.z & 0xFFFF0000u);
  v[6] = __uint_as_float(q.w << 16); v[7] = __uint_as_float(q.w & 0xFFFF0000u);
}

#define GB_TRACE(kk, slot) do { if (p.trace && blockIdx.x == 0 && (kk) < 64 && lane == 0) p.trace[(kk) * 8 + (slot)] = clock64(); } while (0)

template <int NM, int KSTEPS, bool DA>
__global__ void __launch_bounds__(GB_THREADS, 1) gcn_bwd_kernel(const __grid_constant__ GcnBwdParams p) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#define GB_MARK(i) do { if (p.trace && blockIdx.x == 0 && tid == 0) p.trace[63 * 8 + (i)] = clock64(); } while (0)
  GB_MARK(0);
  constexpr int Kp = 16 * KSTEPS, NU = 32 * (1 + NM);
  const int V = p.V;
  const GbLayout L = gb_layout(Kp, NM, DA);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* buf_empty = bars;           // [2] concat buffer + z tile free (tail MMAs of slab k-2 completed)
  uint64_t* in_full = bars + 2;         // [2] dh and z tiles written
  uint64_t* ut_full = bars + 4;         // dU accumulators complete
  uint64_t* ut_empty = bars + 5;        // dU accumulators drained
  uint64_t* us_full = bars + 6;         // [2] dU staged in shared memory
  uint64_t* dz_full = bars + 8;
  uint64_t* dz_empty = bars + 9;
  uint64_t* w_full = bars + 10;         // all MMAs of the CTA completed (dW accumulators final)
  uint64_t* u56_full = bars + 11;       // dA: U5|U6 accumulators complete
  uint64_t* u6s_full = bars + 12;       // [2] U6 staged in shared memory
  uint64_t* t1_full = bars + 14;        // T1 = U5 + hop(U6) accumulator complete
  uint64_t* t1_empty = bars + 15;       // T1 accumulator drained
  uint64_t* t1s_full = bars + 16;       // [2] T1 staged in shared memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int nq_stage = (Kp + 31) / 32;
  const int nq_epi = (V + 31) / 32;
  const int n_stage_warps = 2 * nq_stage;        // two per quadrant that holds node rows

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&buf_empty[i], 1); mbar_init(&in_full[i], 96); mbar_init(&us_full[i], 32 * n_stage_warps);
    }
    mbar_init(ut_full, 1); mbar_init(ut_empty, 32 * n_stage_warps);
    mbar_init(dz_full, 1); mbar_init(dz_empty, 32 * nq_epi);
    mbar_init(w_full, 1);
    mbar_init(u56_full, 1); mbar_init(t1_full, 1); mbar_init(t1_empty, 32 * nq_stage);
    for (int i = 0; i < 2; ++i) { mbar_init(&u6s_full[i], 32 * nq_stage); mbar_init(&t1s_full[i], 32 * nq_stage); }
    fence_barrier_init();
  }
  if (warp == GB_MMA) tmem_alloc(tmem_slot, 512);
  {  // resident operands
    constexpr int pieces = Kp / 8;
    {   // all matrices in one flat loop, 4 independent 16-byte loads in flight per thread
      constexpr int per_mat = pieces * Kp, total = NM * per_mat;
      const uint4* src0 = reinterpret_cast<const uint4*>(p.mats);
      uint4* dst0 = reinterpret_cast<uint4*>(smem);
      for (int i0 = tid; i0 < total; i0 += 4 * GB_THREADS) {
        uint4 v[4]; int di[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = i0 + q * GB_THREADS;
          di[q] = -1;
          if (i < total) {
            const int m = i / per_mat, e = i - m * per_mat, kc = e / Kp, r = e - kc * Kp;
            v[q] = __ldg(src0 + (size_t)p.mat_src[m] * pieces * 128 + kc * 128 + r);
            di[q] = m * (int)(L.mat_bytes / 16) + kc * Kp + r;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (di[q] >= 0) dst0[di[q]] = v[q];
      }
    }
    if (p.w_src) {       // bf16 UMMA images of the mlp weight built here (layouts: see the comment above gcn_bwd_fused_supported)
      bf16* wt = reinterpret_cast<bf16*>(smem + L.w_off);
      bf16* w56 = reinterpret_cast<bf16*>(smem + L.w56_off);
      constexpr int T4 = 8 * 32 * (1 + NM), ITS = (T4 + GB_THREADS - 1) / GB_THREADS;   // 16-byte loads, all issued first
      float4 wv[ITS];
#pragma unroll
      for (int it = 0; it < ITS; ++it) {
        const int i = tid + it * GB_THREADS;
        wv[it] = i < T4 ? __ldg(reinterpret_cast<const float4*>(p.w_src) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int it = 0; it < ITS; ++it) {
        const int i = tid + it * GB_THREADS;
        if (i < T4) {
          const int j = i >> 8, c = (i >> 3) & 31, co = (i & 7) * 4;
          const float v4[4] = {wv[it].x, wv[it].y, wv[it].z, wv[it].w};
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v4[0], v4[1]), h1 = __floats2bfloat162_rn(v4[2], v4[3]);
          uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(wt + ((j * 4 + (co >> 3)) * 32 + c) * 8 + (co & 7)) = pk;     // 4 consecutive c'
          if (DA && (j == 2 * p.sa + 1 || j == 2 * p.sa + 2)) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              w56[((c >> 3) * 64 + (j - (2 * p.sa + 1)) * 32 + co + q) * 8 + (c & 7)] = __float2bfloat16_rn(v4[q]);
          }
        }
      }
    } else {
      const uint4* wsrc = reinterpret_cast<const uint4*>(p.wt_img);
      uint4* wdst = reinterpret_cast<uint4*>(smem + L.w_off);
      for (int i = tid; i < (int)(L.w_bytes / 16); i += GB_THREADS) wdst[i] = __ldg(wsrc + i);
    }
    if (DA) {
      const uint4* src = reinterpret_cast<const uint4*>(p.mats) + (size_t)p.mat_fwd * pieces * 128;
      uint4* dst = reinterpret_cast<uint4*>(smem + L.fwd_off);
      for (int i = tid; i < pieces * Kp; i += GB_THREADS) dst[(i / Kp) * Kp + (i % Kp)] = __ldg(src + (i / Kp) * 128 + (i % Kp));
      if (!p.w_src) {
        const uint4* w5 = reinterpret_cast<const uint4*>(p.w56_img);
        uint4* d5 = reinterpret_cast<uint4*>(smem + L.w56_off);
        for (int i = tid; i < 4 * 64; i += GB_THREADS) d5[i] = __ldg(w5 + i);
      }
    }
    // slack past the buffers is read (as ignored accumulator rows) by the M=128 operands: keep it finite
    uint4* sl = reinterpret_cast<uint4*>(smem + L.bar_off - 4096u);
    for (int i = tid; i < 4096 / 16; i += GB_THREADS) sl[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  const int quad = warp & 3, wq = warp >> 2;
  GB_MARK(1);
  constexpr uint32_t TZ = 192u, TW = 224u;      // TMEM columns of dz and dW
  constexpr uint32_t TU5 = 288u, TU6 = 320u, TDA = 352u;   // dA: U5 -> T1, U6, dA accumulator (Kp columns)

  if (warp == GB_MMA) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop (so descriptors live in uniform registers and the MMAs issue back to back); only
    // the tcgen05 instructions themselves run on one elected lane.
    {
      const uint32_t idescH = make_idesc_bf16(128, 32, false, true);    // matT (K-major) x dh (MN-major)
      const uint32_t idescZ = make_idesc_bf16(128, 32, false, false);   // cat (K-major) x W^T image (K-major)
      const uint32_t idescW = make_idesc_bf16(128, 32, true, true);     // cat (MN-major) x z (MN-major)
      const uint64_t adm = make_smem_desc(0, (uint32_t)Kp * 16u, 128u);           // K-major, K piece stride Kp*16
      const uint64_t bmn = make_smem_desc(0, 128u, (uint32_t)Kp * 16u);           // MN-major [group][node][16 B]
      const uint64_t bdw = make_smem_desc(0, 32u * 16u, 128u);                    // W^T image: K piece stride 512 B
      const uint64_t b56 = make_smem_desc(0, 64u * 16u, 128u);                    // W56 image: K piece stride 1024 B
      const uint32_t idescU = make_idesc_bf16(128, 64, false, false);   // z (K-major) x W56 image (K-major)
      const uint32_t idescA = make_idesc_bf16(128, Kp, false, false);   // T1|U6 (K-major) x dh|dU5 (K-major, N = node)
      auto tail = [&](int kk) {
        const int bb = kk & 1;
        mbar_wait(&us_full[bb], (uint32_t)((kk >> 1) & 1));
        mbar_wait(dz_empty, (uint32_t)((kk & 1) ^ 1));
        GB_TRACE(kk + 1, 4);
        tc_fence_after();
        const uint32_t cat = sbase + L.cat_off + (uint32_t)bb * L.cat_bytes;
        const uint32_t zt = sbase + L.z_off + (uint32_t)bb * L.z_bytes;
        const uint64_t a0 = adm + (uint64_t)(cat >> 4), w0 = bdw + (uint64_t)((sbase + L.w_off) >> 4);
        const uint64_t am = bmn + (uint64_t)(cat >> 4), bz = bmn + (uint64_t)(zt >> 4);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 2 * (1 + NM); ++ks)
            umma_bf16(tmem_base + TZ, a0 + (uint64_t)(((uint32_t)(2 * ks) * (uint32_t)Kp * 16u) >> 4),
                      w0 + (uint64_t)(((uint32_t)(2 * ks) * 512u) >> 4), idescZ, ks == 0 ? 0u : 1u);
          umma_commit(dz_full);
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks)
              umma_bf16(tmem_base + TW + 32u * t, am + (uint64_t)(((uint32_t)(16 * t) * (uint32_t)Kp * 16u + (uint32_t)ks * 256u) >> 4),
                        bz + (uint64_t)(((uint32_t)ks * 256u) >> 4), idescW, (kk == 0 && ks == 0) ? 0u : 1u);
        }
        __syncwarp();
        if (DA) {
          // dA += T1^T-style products: D[v, w] += sum_c T1[v,c] dh[w,c] + U6[v,c] dU5[w,c]   (all operands K-major)
          GB_TRACE(kk + 1, 5);
          mbar_wait(&t1s_full[bb], (uint32_t)((kk >> 1) & 1));
          tc_fence_after();
          const uint64_t at1 = adm + (uint64_t)((sbase + L.t1_off + (uint32_t)bb * L.slot_bytes) >> 4);
          const uint64_t au6 = adm + (uint64_t)((sbase + L.u6_off + (uint32_t)bb * L.slot_bytes) >> 4);
          const uint64_t bdh = adm + (uint64_t)(cat >> 4);
          const uint64_t bd5 = adm + (uint64_t)((cat + (uint32_t)(2 * p.sa + 1) * L.slot_bytes) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma_bf16(tmem_base + TDA, at1 + (uint64_t)(((uint32_t)(2 * ks) * (uint32_t)Kp * 16u) >> 4),
                        bdh + (uint64_t)(((uint32_t)(2 * ks) * (uint32_t)Kp * 16u) >> 4), idescA, (kk == 0 && ks == 0) ? 0u : 1u);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma_bf16(tmem_base + TDA, au6 + (uint64_t)(((uint32_t)(2 * ks) * (uint32_t)Kp * 16u) >> 4),
                        bd5 + (uint64_t)(((uint32_t)(2 * ks) * (uint32_t)Kp * 16u) >> 4), idescA, 1u);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&buf_empty[bb]);
        __syncwarp();
      };
      int k = 0;
      for (long long slab = blockIdx.x; slab < p.slabs; slab += gridDim.x, ++k) {
        const int bb = k & 1;
        mbar_wait(&in_full[bb], (uint32_t)((k >> 1) & 1));
        GB_TRACE(k, 0);
        if (DA) {
          // U5 | U6 = z [W_{2sa+1} | W_{2sa+2}]  (issued first: its staging overlaps the hop MMAs below)
          mbar_wait(t1_empty, (uint32_t)((k & 1) ^ 1));
          tc_fence_after();
          const uint64_t az = adm + (uint64_t)((sbase + L.z_off + (uint32_t)bb * L.z_bytes) >> 4);
          const uint64_t bw = b56 + (uint64_t)((sbase + L.w56_off) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma_bf16(tmem_base + TU5, az + (uint64_t)(((uint32_t)(2 * ks) * (uint32_t)Kp * 16u) >> 4),
                        bw + (uint64_t)(((uint32_t)(2 * ks) * 1024u) >> 4), idescU, ks == 0 ? 0u : 1u);
            umma_commit(u56_full);
          }
          __syncwarp();
        }
        mbar_wait(ut_empty, (uint32_t)((k & 1) ^ 1));
        GB_TRACE(k, 1);
        tc_fence_after();
        const uint32_t cat = sbase + L.cat_off + (uint32_t)bb * L.cat_bytes;
        const uint64_t a0 = adm + (uint64_t)(sbase >> 4), b0 = bmn + (uint64_t)(cat >> 4);
        if (elect_one()) {
#pragma unroll
          for (int m = 0; m < NM; ++m)
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks)
              umma_bf16(tmem_base + 32u * m,
                        a0 + (uint64_t)(((uint32_t)m * ((uint32_t)(Kp / 8) * (uint32_t)Kp * 16u) + (uint32_t)(2 * ks) * (uint32_t)Kp * 16u) >> 4),
                        b0 + (uint64_t)(((uint32_t)ks * 256u) >> 4), idescH, ks == 0 ? 0u : 1u);
          umma_commit(ut_full);
        }
        __syncwarp();
        GB_TRACE(k, 2);
        // tail of the previous slab BEFORE this slab's T1 hop: its commit frees the concat / z buffers the prep warps
        // are waiting for, so the next slab's operands are being written while the T1 hop is issued
        if (k > 0) tail(k - 1);
        GB_TRACE(k, 3);
        if (DA) {
          // T1 = U5 + A^T-hop(U6): accumulate the forward hop of the staged U6 onto the U5 columns
          mbar_wait(&u6s_full[bb], (uint32_t)((k >> 1) & 1));
          tc_fence_after();
          const uint64_t af = adm + (uint64_t)((sbase + L.fwd_off) >> 4);
          const uint64_t bu = bmn + (uint64_t)((sbase + L.u6_off + (uint32_t)bb * L.slot_bytes) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks)
              umma_bf16(tmem_base + TU5, af + (uint64_t)(((uint32_t)(2 * ks) * (uint32_t)Kp * 16u) >> 4),
                        bu + (uint64_t)(((uint32_t)ks * 256u) >> 4), idescH, 1u);
            umma_commit(t1_full);
          }
          __syncwarp();
        }
        GB_TRACE(k, 6);
      }
      if (k > 0) tail(k - 1);
      if (elect_one()) umma_commit(w_full);
    }
    __syncwarp();
  } else if (warp == 12 || warp == 13 || warp == 15) {
    // ===================== prep warps: dh = du . mask -> slot 0, z = a . b -> z tile =====================
    const int v = (warp == 15 ? 2 : warp - 12) * 32 + lane;          // node row
    const bool has_row = v < Kp, real = v < V;
    float dbs[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) dbs[c] = 0.f;
    uint64_t sd = 0, of = 0;
    const bool philox = (p.mask == nullptr) && p.drop_p > 0.f;
    if (philox) { sd = p.rng ? __ldg(p.rng) : p.seed; of = p.rng ? p.offset + __ldg(p.rng + 1) : p.offset; }
    int k = 0;
    for (long long slab = blockIdx.x; slab < p.slabs; slab += gridDim.x, ++k) {
      const int bb = k & 1;
      const long long pp = slab * V + v;
      uint4 od[4], oz[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { od[j] = make_uint4(0u, 0u, 0u, 0u); oz[j] = od[j]; }
      if (real) {
        const uint4* s0 = reinterpret_cast<const uint4*>(p.du + pp * 32);
        const uint4* s1 = reinterpret_cast<const uint4*>(p.a + pp * 32);
        const uint4* s2 = reinterpret_cast<const uint4*>(p.b + pp * 32);
        uint4 qd[4], qa[4], qb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { qd[j] = __ldg(s0 + j); qa[j] = __ldg(s1 + j); qb[j] = __ldg(s2 + j); }
        float m16[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float d[8], m[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f}, av[8], bv[8];
          gb_unpack8(qd[j], d); gb_unpack8(qa[j], av); gb_unpack8(qb[j], bv);
          if (p.mask) gb_unpack8(__ldg(reinterpret_cast<const uint4*>(p.mask + pp * 32) + j), m);
          else if (philox) {
            if ((j & 1) == 0) dropout16(sd, of, (uint64_t)(pp * 2 + (j >> 1)), p.drop_p, m16);   // one call per 16 channels
#pragma unroll
            for (int i = 0; i < 8; ++i) m[i] = m16[8 * (j & 1) + i];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) { d[i] *= m[i]; dbs[8 * j + i] += d[i]; av[i] *= bv[i]; }
          od[j] = make_uint4(gb_pack(d[0], d[1]), gb_pack(d[2], d[3]), gb_pack(d[4], d[5]), gb_pack(d[6], d[7]));
          oz[j] = make_uint4(gb_pack(av[0], av[1]), gb_pack(av[2], av[3]), gb_pack(av[4], av[5]), gb_pack(av[6], av[7]));
        }
      }
      mbar_wait(&buf_empty[bb], (uint32_t)(((k >> 1) & 1) ^ 1));
      if (has_row) {
        uint8_t* dslot = smem + L.cat_off + (size_t)bb * L.cat_bytes + (size_t)v * 16;
        uint8_t* zslot = smem + L.z_off + (size_t)bb * L.z_bytes + (size_t)v * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          *reinterpret_cast<uint4*>(dslot + (size_t)j * Kp * 16) = od[j];
          *reinterpret_cast<uint4*>(zslot + (size_t)j * Kp * 16) = oz[j];
        }
      }
      fence_proxy_async();
      mbar_arrive(&in_full[bb]);
    }
    const float s = warp_column_sums(dbs, lane);
    atomicAdd(p.db_mlp + lane, s);
  } else if ((quad < 2 && wq < 2) || warp == 6 || warp == 14 || warp == 7 || warp == 11) {
    // ===================== stage warps: TMEM(dU_j) -> bf16 -> slots 1..H =====================
    // two per quadrant (q0: w0,w4  q1: w1,w5  q2: w6,w14  q3: w7,w11), splitting the hop chunks even / odd: the
    // TMEM round trips of one warp are serial, so a lone warp doing all H chunks of its quadrant was the last to arrive
    if (quad < nq_stage) {
      const int row = quad * 32 + lane;
      const int sidx = (quad < 2) ? wq : ((warp == 6 || warp == 7) ? 0 : 1);
      const int first = 1 + sidx, step = 2;
      const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
      int k = 0;
      for (long long slab = blockIdx.x; slab < p.slabs; slab += gridDim.x, ++k) {
        const int bb = k & 1;
        uint32_t r[32];
        auto stage_tile = [&](uint32_t tcol, uint8_t* dst0) {     // 32 TMEM columns of this row -> [4 cg][Kp][16 B]
          tmem_ld32_issue(tmem_base + lane_off + tcol, r);
          tmem_ld_wait();
          if (row < Kp) {
            uint8_t* dst = dst0 + (size_t)row * 16;
#pragma unroll
            for (int cg = 0; cg < 4; ++cg) {
              uint4 pk;
              pk.x = gb_pack(__uint_as_float(r[8 * cg]), __uint_as_float(r[8 * cg + 1]));
              pk.y = gb_pack(__uint_as_float(r[8 * cg + 2]), __uint_as_float(r[8 * cg + 3]));
              pk.z = gb_pack(__uint_as_float(r[8 * cg + 4]), __uint_as_float(r[8 * cg + 5]));
              pk.w = gb_pack(__uint_as_float(r[8 * cg + 6]), __uint_as_float(r[8 * cg + 7]));
              *reinterpret_cast<uint4*>(dst + (size_t)cg * Kp * 16) = pk;
            }
          }
        };
        const bool does_u6 = DA && sidx == 0, does_t1 = DA && sidx == 1;
        if (does_u6) {       // U6 first: it is ready before the hops of this slab finish
          mbar_wait(u56_full, (uint32_t)(k & 1));
          tc_fence_after();
          stage_tile(TU6, smem + L.u6_off + (size_t)bb * L.slot_bytes);
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(&u6s_full[bb]);
        }
        mbar_wait(ut_full, (uint32_t)(k & 1));
        mbar_wait(&buf_empty[bb], (uint32_t)(((k >> 1) & 1) ^ 1));
        tc_fence_after();
        uint8_t* cat = smem + L.cat_off + (size_t)bb * L.cat_bytes;
        for (int j = first; j <= NM; j += step) {
          tmem_ld32_issue(tmem_base + lane_off + 32u * (uint32_t)(j - 1), r);
          tmem_ld_wait();
          if (j + step > NM) { tc_fence_before(); mbar_arrive(ut_empty); }     // my last read of the accumulators
          if (row < Kp) {
            uint8_t* dst = cat + (size_t)j * L.slot_bytes + (size_t)row * 16;
#pragma unroll
            for (int cg = 0; cg < 4; ++cg) {
              uint4 pk;
              pk.x = gb_pack(__uint_as_float(r[8 * cg]), __uint_as_float(r[8 * cg + 1]));
              pk.y = gb_pack(__uint_as_float(r[8 * cg + 2]), __uint_as_float(r[8 * cg + 3]));
              pk.z = gb_pack(__uint_as_float(r[8 * cg + 4]), __uint_as_float(r[8 * cg + 5]));
              pk.w = gb_pack(__uint_as_float(r[8 * cg + 6]), __uint_as_float(r[8 * cg + 7]));
              *reinterpret_cast<uint4*>(dst + (size_t)cg * Kp * 16) = pk;
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(&us_full[bb]);
        if (does_t1) {
          mbar_wait(t1_full, (uint32_t)(k & 1));
          tc_fence_after();
          stage_tile(TU5, smem + L.t1_off + (size_t)bb * L.slot_bytes);
          tc_fence_before();
          mbar_arrive(t1_empty);
          fence_proxy_async();
          mbar_arrive(&t1s_full[bb]);
        }
      }
    }
  } else if (wq == 2 && quad < 3) {
    // ===================== epilogue warps (w8, w9, w10): dz -> gate backward -> dfg; then the dW flush =====================
    const int w = quad * 32 + lane;
    if (quad < nq_epi) {
      const bool valid = w < V;
      int k = 0;
      for (long long slab = blockIdx.x; slab < p.slabs; slab += gridDim.x, ++k) {
        const long long pp = slab * V + w;
        uint4 qa[4], qb[4], ql[4];
        bool tailrow = false;
        if (valid) {
          const uint4* s1 = reinterpret_cast<const uint4*>(p.a + pp * 32);
          const uint4* s2 = reinterpret_cast<const uint4*>(p.b + pp * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) { qa[j] = __ldg(s1 + j); qb[j] = __ldg(s2 + j); }
          if (p.dz_last) {
            long long n, rem;
            split_pos(pp, p.RO, n, rem);
            if (rem >= p.last_begin) {
              tailrow = true;
              const uint4* s3 = reinterpret_cast<const uint4*>(p.dz_last + (n * p.last_rows + rem - p.last_begin) * 32);
#pragma unroll
              for (int j = 0; j < 4; ++j) ql[j] = __ldg(s3 + j);
            }
          }
        }
        mbar_wait(dz_full, (uint32_t)(k & 1));
        tc_fence_after();
        float dz[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + TZ, dz);
        tc_fence_before();
        mbar_arrive(dz_empty);
        if (valid) {
          uint4* out = reinterpret_cast<uint4*>(p.dfg + pp * 64);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float av[8], bv[8], o[16];
            gb_unpack8(qa[j], av); gb_unpack8(qb[j], bv);
            if (tailrow) {
              float t[8];
              gb_unpack8(ql[j], t);
#pragma unroll
              for (int i = 0; i < 8; ++i) dz[8 * j + i] += t[i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float g = dz[8 * j + i];
              o[2 * i] = g * bv[i] * (1.f - av[i] * av[i]);
              o[2 * i + 1] = g * av[i] * bv[i] * (1.f - bv[i]);
            }
            out[2 * j] = make_uint4(gb_pack(o[0], o[1]), gb_pack(o[2], o[3]), gb_pack(o[4], o[5]), gb_pack(o[6], o[7]));
            out[2 * j + 1] = make_uint4(gb_pack(o[8], o[9]), gb_pack(o[10], o[11]), gb_pack(o[12], o[13]), gb_pack(o[14], o[15]));
          }
        }
      }
    }
  }
  // ===================== gradient flush (w8, w9, w10 and w15: one warp per TMEM lane quadrant) =====================
  // dW: D_W[(j, c'), c] -> dw_mlp[(j*32 + c), c'];  dA: accumulator row = node v, column = node w.  The CTA's partials
  // are staged in shared memory (every operand buffer is idle once w_full has fired) and flushed with rotated,
  // coalesced vector reductions (common.cuh: red_flush_1d) - direct scalar atomics from 148 CTAs finishing together
  // serialise in L2 and cost tens of microseconds per launch.
  if ((wq == 2 && quad < 3) || warp == 15) {
    float* stg_w = reinterpret_cast<float*>(smem + L.cat_off);             // [NU][32]
    float* stg_a = stg_w + 32 * NU;                                        // [V][V]
    const int et = quad * 32 + lane;
    mbar_wait(w_full, 0u);
    tc_fence_after();
#pragma unroll 1
    for (int t = 0; t < 2; ++t) {
      const int m = t * 128 + quad * 32 + lane;
      float vv[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + TW + 32u * t, vv);
      if (m < NU) {
        float* dst = stg_w + (size_t)(m >> 5) * 32 * 32 + (m & 31);
#pragma unroll
        for (int c = 0; c < 32; ++c) dst[c * 32] = vv[c];
      }
    }
    if (DA) {
      const int w = quad * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < Kp; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem_base + ((uint32_t)(quad * 32) << 16) + TDA + (uint32_t)c0)
            : "memory");
        tmem_ld_wait();
        if (w < V) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < V) stg_a[w * V + c0 + j] = __uint_as_float(r[j]);
        }
      }
    }
    asm volatile("bar.sync 2, 128;" ::: "memory");
    red_flush_1d(p.dw_mlp, stg_w, 32 * NU, et, 128);
    if (DA) red_flush_1d(p.dA, stg_a, V * V, et, 128);
  }
  tc_fence_before();
  __syncthreads();
  GB_MARK(2);
  if (warp == GB_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// image layouts built in the kernel prologue: wt_img[kc = (j, c'>>3)][n = c][c' & 7] = W_mlp[j*32 + c][c'];
//   w56[kc = c>>3][n = (h, c')][c & 7] = W_mlp[(2sa+1+h)*32 + c][c']
int gcn_bwd_fused_supported(int V, int n_mats) {
  if (V < 1 || V > 80 || (n_mats != 2 && n_mats != 4 && n_mats != 6)) return 0;
  const int Kp = ((V + 15) / 16) * 16;
  return gb_layout(Kp, n_mats, true).total <= 227u * 1024u ? 1 : 0;
}


int launch_gcn_bwd(GcnBwdParams& p, cudaStream_t st) {
  if (p.slabs <= 0) return 0;
  p.Kp = ((p.V + 15) / 16) * 16;
  GWN_REQUIRE(gcn_bwd_fused_supported(p.V, p.n_mats), "gcn_bwd: V=%d with %d resident matrices is not supported", p.V,
              p.n_mats);
  GWN_REQUIRE((long long)p.slabs * p.V < (1ll << 31), "gcn_bwd: too many positions");
  const bool has_da = p.sa >= 0;
  GWN_REQUIRE(!has_da || (p.dA && (p.w56_img || p.w_src) && p.sa < p.n_mats / 2), "gcn_bwd: bad support-gradient arguments");
  const GbLayout L = gb_layout(p.Kp, p.n_mats, has_da);
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    GWN_CUDA(cudaGetDevice(&dev));
    GWN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = p.slabs < sms ? p.slabs : sms;
  const int ks = p.Kp / 16;
#define GB_CASE(NM_, KS_)                                                                                         \
  if (p.n_mats == NM_ && ks == KS_) {                                                                             \
    static bool attr = false;                                                                                     \
    if (!attr) {                                                                                                  \
      GWN_CUDA(cudaFuncSetAttribute(gcn_bwd_kernel<NM_, KS_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      GWN_CUDA(cudaFuncSetAttribute(gcn_bwd_kernel<NM_, KS_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      attr = true;                                                                                                \
    }                                                                                                             \
    if (has_da) gcn_bwd_kernel<NM_, KS_, true><<<grid, GB_THREADS, L.total, st>>>(p);                             \
    else gcn_bwd_kernel<NM_, KS_, false><<<grid, GB_THREADS, L.total, st>>>(p);                                   \
    GWN_LAUNCHED();                                                                                               \
    return 0;                                                                                                     \
  }
  GB_CASE(6, 5) GB_CASE(4, 5) GB_CASE(2, 5) GB_CASE(6, 4) GB_CASE(4, 4) GB_CASE(2, 4) GB_CASE(6, 3) GB_CASE(4, 3)
  GB_CASE(2, 3) GB_CASE(6, 2) GB_CASE(4, 2) GB_CASE(2, 2) GB_CASE(6, 1) GB_CASE(4, 1) GB_CASE(2, 1)
#undef GB_CASE
  GWN_REQUIRE(false, "gcn_bwd: no kernel instance for %d matrices, Kp=%d", p.n_mats, p.Kp);
  return -1;
}

}  // namespace gwn
