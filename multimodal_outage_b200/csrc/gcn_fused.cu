// Fused diffusion graph convolution, forward (see gcn_fused.cuh for the math and the data flow).
//
// One persistent CTA per SM walks slabs (slab = one (n,t) pair = V nodes x 32 channels).  Warp roles:
//   warp 0        producer : z slab -> smem ring (cp.async, 16-byte pieces, K-major A-operand layout of GEMM 1)
//   warp 1        MMA      : lane 0 issues GEMM 1 of slab k, then GEMM 2 of slab k-1 (software pipelined)
//   warps 2-5     stage    : TMEM(U) -> bf16 -> smem ring (B operand of GEMM 2); U_0 + bias -> TMEM(h) (tcgen05.st)
//   warps 6-13    epilogue : TMEM(h) -> dropout, residual (folded BN), bf16 store, BN statistics; two groups of
//                            4 warps alternate slabs (each owns one of the two h accumulators)
// TMEM (512 columns): U accumulators at columns [0,224) and [256,480), h accumulators at [224,256), [480,512).
#include "gcn_fused.cuh"
#include "tc.cuh"
#include "tc_gemm_impl.cuh"   // warp_column_sums

namespace gwn {

constexpr int GF_ZST = 4;          // z ring stages
constexpr int GF_UST = 3;          // U ring stages
constexpr int GF_STAGE_WARP0 = 2, GF_EPI_WARP0 = 6;
constexpr int GF_THREADS = 32 * 14;

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
  L.total = L.bar_off + 256u;
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

__global__ void __launch_bounds__(GF_THREADS, 1) gcn_fwd_kernel(const __grid_constant__ GcnFwdParams p) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = p.V, Kp = p.Kp, nm = p.n_mats, NU = 32 * (1 + nm);
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

  if (tid == 0) {
    for (int i = 0; i < GF_ZST; ++i) { mbar_init(&z_full[i], 32); mbar_init(&z_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ut_full[i], 1); mbar_init(&ut_empty[i], 128);
      mbar_init(&ht_full[i], 1); mbar_init(&ht_empty[i], 128);
    }
    for (int i = 0; i < GF_UST; ++i) { mbar_init(&us_full[i], 128); mbar_init(&us_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  {  // resident operands: support images (rows < Kp of each K piece) and the mlp weight image
    const int per_piece = Kp;                       // 16-byte rows kept per K piece
    const int pieces = Kp / 8;
    for (int m = 0; m < nm; ++m) {
      const uint4* src = reinterpret_cast<const uint4*>(p.mats) + (size_t)p.mat_src[m] * pieces * 128;
      uint4* dst = reinterpret_cast<uint4*>(smem + (size_t)m * L.mat_bytes);
      for (int i = tid; i < pieces * per_piece; i += GF_THREADS) {
        const int kc = i / per_piece, r = i % per_piece;
        dst[kc * per_piece + r] = __ldg(src + kc * 128 + r);
      }
    }
    const uint4* wsrc = reinterpret_cast<const uint4*>(p.w_img);
    uint4* wdst = reinterpret_cast<uint4*>(smem + L.w_off);
    for (int i = tid; i < (int)(L.w_bytes / 16); i += GF_THREADS) wdst[i] = __ldg(wsrc + i);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sbase = smem_u32(smem);

  if (warp == 0) {
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
      if (k >= GF_ZST - 2) {       // ZST-2 groups may stay in flight
        cp_async_wait<GF_ZST - 2>();
        fence_proxy_async();
        mbar_arrive(&z_full[(k - (GF_ZST - 2)) % GF_ZST]);
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int kk = (k >= GF_ZST - 2 ? k - (GF_ZST - 2) : 0); kk < k; ++kk) mbar_arrive(&z_full[kk % GF_ZST]);
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idescU = make_idesc_bf16(128, NU, false, false);
      const uint32_t idescH = make_idesc_bf16(128, 32, false, true);
      const int ksteps = Kp / 16;
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
        for (int m = 0; m < nm; ++m) {
          const uint32_t a0 = sbase + (uint32_t)m * L.mat_bytes;
          const uint32_t b0 = ub + (uint32_t)m * L.u_slot;
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t ad = adm + (uint64_t)((a0 + (uint32_t)(2 * ks) * (uint32_t)Kp * 16u) >> 4);
            const uint64_t bd = bdu + (uint64_t)((b0 + (uint32_t)ks * 256u) >> 4);
            umma_bf16(d, ad, bd, idescH, 1u);
          }
        }
        umma_commit(&us_empty[us]);
        umma_commit(&ht_full[hb]);
      };
      int k = 0;
      for (long long slab = blockIdx.x; slab < p.slabs; slab += gridDim.x, ++k) {
        const int zs = k % GF_ZST, ub = k & 1;
        mbar_wait(&z_full[zs], (uint32_t)((k / GF_ZST) & 1));
        mbar_wait(&ut_empty[ub], (uint32_t)(((k >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t za = sbase + L.z_off + (uint32_t)zs * L.z_stage;
        const uint32_t wa = sbase + L.w_off;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t ad = adz + (uint64_t)((za + (uint32_t)(2 * ks) * L.z_piece) >> 4);
          const uint64_t bd = bdw + (uint64_t)((wa + (uint32_t)(2 * ks) * (uint32_t)NU * 16u) >> 4);
          umma_bf16(tmem_base + (uint32_t)ub * 256u, ad, bd, idescU, ks == 0 ? 0u : 1u);
        }
        umma_commit(&z_empty[zs]);
        umma_commit(&ut_full[ub]);
        if (k > 0) issue_hops(k - 1);
      }
      if (k > 0) issue_hops(k - 1);
    }
    __syncwarp();
  } else if (warp < GF_EPI_WARP0) {
    // ===================== stage warps: U -> smem (bf16), U_0 + bias -> h accumulator =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const bool active = quad * 32 < Kp;        // this quadrant holds real (or zero-padding) node rows
    float bias[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) bias[c] = __ldg(p.bias + c);
    int k = 0;
    for (long long slab = blockIdx.x; slab < p.slabs; slab += gridDim.x, ++k) {
      const int ub = k & 1, us = k % GF_UST, hb = k & 1;
      mbar_wait(&ut_full[ub], (uint32_t)((k >> 1) & 1));
      mbar_wait(&us_empty[us], (uint32_t)(((k / GF_UST) & 1) ^ 1));
      mbar_wait(&ht_empty[hb], (uint32_t)(((k >> 1) & 1) ^ 1));
      tc_fence_after();
      if (active) {
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const uint32_t tU = tmem_base + lane_off + (uint32_t)ub * 256u;
        uint32_t r[32];
        tmem_ld32_issue(tU, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) + bias[c]);
        tmem_st32(tmem_base + lane_off + (uint32_t)hb * 256u + 224u, r);
        uint8_t* ust = smem + L.u_off + (size_t)us * L.u_stage;
        for (int j = 1; j <= nm; ++j) {
          tmem_ld32_issue(tU + (uint32_t)j * 32u, r);
          tmem_ld_wait();
          if (row < Kp) {
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
        tmem_st_wait();
        fence_proxy_async();
      }
      tc_fence_before();
      mbar_arrive(&ut_empty[ub]);
      mbar_arrive(&us_full[us]);
    }
  } else {
    // ===================== epilogue warps =====================
    const int quad = warp & 3;
    const int grp = (warp - GF_EPI_WARP0) >> 2;
    const int w = quad * 32 + lane;
    const bool valid = w < V;
    float sa[32], sb[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) { sa[c] = 0.f; sb[c] = 0.f; }
    uint64_t sd = 0, of = 0;
    const bool philox = (p.mask == nullptr) && p.drop_p > 0.f;
    if (philox) { sd = p.rng ? __ldg(p.rng) : p.seed; of = p.rng ? p.offset + __ldg(p.rng + 1) : p.offset; }
    int k = 0;
    for (long long slab = blockIdx.x; slab < p.slabs; slab += gridDim.x, ++k) {
      if ((k & 1) != grp) continue;
      const int hb = grp;
      mbar_wait(&ht_full[hb], (uint32_t)((k >> 1) & 1));
      tc_fence_after();
      float h[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)hb * 256u + 224u, h);
      tc_fence_before();
      mbar_arrive(&ht_empty[hb]);
      if (valid) {
        const long long pp = slab * V + w;
        const long long n = pp / p.RO, rem = pp % p.RO;
        const bf16* rp = p.u_prev + (n * p.RI + rem + p.crop) * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float m[4] = {1.f, 1.f, 1.f, 1.f}, r[4];
          if (p.mask) load4(p.mask + pp * 32 + 4 * j, m);
          else if (philox) dropout4(sd, of, (uint64_t)(pp * 8 + j), p.drop_p, m);
          load4(rp + 4 * j, r);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = 4 * j + i;
            const float rr = p.scale ? fmaf(r[i], __ldg(p.scale + c), __ldg(p.shift + c)) : r[i];
            h[c] = fmaf(h[c], m[i], rr);
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
    }
    const float s1 = warp_column_sums(sa, lane);
    const float s2 = warp_column_sums(sb, lane);
    atomicAdd(p.stats + lane, (double)s1);
    atomicAdd(p.stats + 32 + lane, (double)s2);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// w_img[(c>>3)][n = (j, c')][c & 7] = W_mlp[j*32 + c][c']
__global__ void gcn_wprep_kernel(const float* __restrict__ w, int n_mats, bf16* __restrict__ img) {
  const int NU = 32 * (1 + n_mats);
  const int total = 32 * NU;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i / NU, n = i % NU;
    const int j = n >> 5, co = n & 31;
    img[((c >> 3) * NU + n) * 8 + (c & 7)] = __float2bfloat16_rn(w[(j * 32 + c) * 32 + co]);
  }
}

int gcn_fused_supported(int V, int n_mats) {
  if (V < 1 || V > 128 || n_mats < 1 || n_mats > GF_MAX_MATS) return 0;
  const int Kp = ((V + 15) / 16) * 16;
  return gf_layout(Kp, n_mats).total <= 227u * 1024u ? 1 : 0;
}

int launch_gcn_wprep(const float* w_mlp, int n_mats, bf16* w_img, cudaStream_t st) {
  gcn_wprep_kernel<<<8, 256, 0, st>>>(w_mlp, n_mats, w_img);
  GWN_LAUNCHED();
  return 0;
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
    GWN_CUDA(cudaFuncSetAttribute(gcn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const int grid = p.slabs < sms ? p.slabs : sms;
  gcn_fwd_kernel<<<grid, GF_THREADS, L.total, st>>>(p);
  GWN_LAUNCHED();
  return 0;
}

}  // namespace gwn
