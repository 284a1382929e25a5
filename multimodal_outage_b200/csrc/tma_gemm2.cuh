// Two-SM form of the TMA-fed tcgen05 GEMM (tma_gemm.cuh) for the 3,100-node diffusion GEMMs: CTA PAIRS (clusters of two)
// work on 256 x 256 output tiles with `tcgen05.mma.cta_group::2`.
//
//   D[m, n] = sum_k A(m, k) * B(n, k)        pair tile 256 x bn (bn <= 256), K-block 64, bf16 x bf16 -> fp32
//
// Why: with one CTA per 128 x 256 tile a K block costs 48 KB of TMA writes plus 48 KB of operand reads on ONE SM's shared
// memory (192 cycles at 128 B/clk for 4 MMAs worth 128 cycles each) - the single-SM kernel measures 0.50 us per K block,
// 0.68 of the cuBLAS rate, in both the hop GEMM and the support gradient.  In a pair each CTA stages its own 128 rows of A
// and only HALF of the B tile (bn/2 columns); the MMA reads both halves across the pair: 32 KB written + 32 KB read per SM
// and K block.
//
// Protocol (rank = %cluster_ctarank; the even CTA is the leader):
//   * both CTAs run a TMA producer warp; every load signals the LEADER's full barrier (peer bit of the barrier address
//     cleared), which expects the bytes of both CTAs' stages;
//   * only the leader's MMA warp issues MMAs; tcgen05.commit multicasts its arrival to BOTH CTAs' empty / accumulator-full
//     barriers (same shared-memory offsets in both CTAs);
//   * each CTA's 16 epilogue warps drain the CTA's own 128 accumulator rows (all bn columns) and arrive on the LEADER's
//     accumulator-empty barrier (count = 2 x epilogue warps);
//   * TMEM is allocated / freed with the cta_group::2 forms by warp 1 of both CTAs; cluster barriers around set-up and
//     before teardown (the leader's MMAs read the peer's shared memory until the last commit).
#pragma once
#include "tma_gemm.cuh"

namespace gwn {

namespace tg2 {
using namespace tc;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;    // shared::cluster address of the same offset in the pair's even CTA

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {  // one full warp in BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair when the MMAs issued so far have completed
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}
__device__ __forceinline__ void tma2_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar) & kPeerMask)
      : "memory");
}
__device__ __forceinline__ void tma2_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
          "r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar) & kPeerMask)
      : "memory");
}
__device__ __forceinline__ void load_operand2(const TgOperand& o, const CUtensorMap* map, uint32_t dst, int mn0, int k0,
                                              uint64_t* bar) {
  if (o.mode == TG_K_SW128) {
    tma2_2d(dst, map, k0, mn0, bar);
  } else if (o.mode == TG_MN_SW128) {
    for (int i = 0; i < o.n_boxes; ++i) tma2_2d(dst + (uint32_t)i * o.box_bytes, map, mn0 + i * o.box_mn, k0, bar);
  } else if (o.mode == TG_MN_SW64) {
    tma2_3d(dst, map, 0, k0, mn0 >> 5, bar);
  } else {
    tma2_3d(dst, map, 0, mn0, k0 >> 5, bar);
  }
}
}  // namespace tg2

// p.a describes ONE CTA's 128 rows of A, p.b one CTA's bn/2 columns of B (tile_bytes, descriptor constants, boxes);
// p.m_tiles counts 256-row pair tiles.  Epilogue functor contract as in tma_gemm.cuh (`chunk`).
template <typename Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TG_THREADS, 1)
tma_gemm2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ TgParams p, Epi epi) {
  using namespace tc;
  using namespace tg;
  using namespace tg2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_rank();
  const int ST = p.stages;
  const uint32_t stage_bytes = p.a.tile_bytes + p.b.tile_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)ST * stage_bytes);
  uint64_t* full = bars;            // [ST]  leader's are used
  uint64_t* empty = bars + 8;       // [ST]  per CTA
  uint64_t* tfull = bars + 16;      // [2]   per CTA
  uint64_t* tempty = bars + 18;     // [2]   leader's are used
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  if (tid == 0) {
    for (int i = 0; i < ST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * TG_EPI_WARPS); }
    fence_barrier_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  const uint32_t acc_cols = p.bn <= 32 ? 32u : p.bn <= 64 ? 64u : p.bn <= 128 ? 128u : 256u;
  cluster_sync();                   // both CTAs resident, barriers initialised, before anything crosses the pair
  if (warp == 1) tmem_alloc2(tmem_slot, 2 * acc_cols);
  pdl_wait();
  pdl_trigger();
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int half_bn = p.bn >> 1;

  if (warp == 0) {
    int s = 0, ph = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs) {
      const int mt = tile % p.m_tiles, rest = tile / p.m_tiles;
      const int nt = rest % p.n_tiles, sp = rest / p.n_tiles;
      const int kb0 = sp * p.kb_per_split;
      const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[s], (uint32_t)(ph ^ 1));
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(&full[s], 2u * stage_bytes);
          const uint32_t sa = base + (uint32_t)s * stage_bytes, sb = sa + p.a.tile_bytes;
          load_operand2(p.a, &mapA, sa, mt * 2 * TG_BM + rank * TG_BM, kb * TG_BK, &full[s]);
          load_operand2(p.b, &mapB, sb, nt * p.bn + rank * half_bn, kb * TG_BK, &full[s]);
        }
        __syncwarp();
        if (++s == ST) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(256, p.bn, p.a.mode == TG_MN_SW128 || p.a.mode == TG_MN_SW64,
                                             p.b.mode == TG_MN_SW128 || p.b.mode == TG_MN_SW64);
      const uint64_t at = make_desc_sw(0, p.a.lbo, p.a.sbo, p.a.layout);
      const uint64_t bt = make_desc_sw(0, p.b.lbo, p.b.sbo, p.b.layout);
      int s = 0, ph = 0, tcount = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs, ++tcount) {
        const int rest = tile / p.m_tiles;
        const int sp = rest / p.n_tiles;
        const int kb0 = sp * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        const int acc = tcount & 1;
        mbar_wait(&tempty[acc], (uint32_t)(((tcount >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)acc * acc_cols;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[s], (uint32_t)ph);
          tc_fence_after();
          const uint32_t sa = base + (uint32_t)s * stage_bytes, sb = sa + p.a.tile_bytes;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < TG_BK / 16; ++ks) {
              const uint64_t adesc = at + (uint64_t)((sa + p.a.koff[ks]) >> 4);
              const uint64_t bdesc = bt + (uint64_t)((sb + p.b.koff[ks]) >> 4);
              umma2_bf16(d, adesc, bdesc, idesc, (kb == kb0 && ks == 0) ? 0u : 1u);
            }
            umma2_commit(&empty[s]);
            if (kb == kb1 - 1) umma2_commit(&tfull[acc]);
          }
          __syncwarp();
          if (++s == ST) { s = 0; ph ^= 1; }
        }
        if (kb1 <= kb0) {
          if (elect_one()) umma2_commit(&tfull[acc]);
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else {
    const int quad = warp & 3, erank = (warp - 2) >> 2;
    int tcount = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++tcount) {
      const int mt = tile % p.m_tiles, rest = tile / p.m_tiles;
      const int nt = rest % p.n_tiles;
      const int acc = tcount & 1;
      mbar_wait(&tfull[acc], (uint32_t)((tcount >> 1) & 1));
      tc_fence_after();
      const int m = mt * 2 * TG_BM + rank * TG_BM + quad * 32 + lane;
      const bool m_ok = m < p.M;
      const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * acc_cols;
      for (int c0 = 32 * erank; c0 < p.bn; c0 += 32 * TG_EPI_RANKS) {
        float v[32];
        tmem_ld32(t0 + (uint32_t)c0, v);
        epi.chunk(m, m_ok, nt * p.bn + c0, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty[acc]);
    }
  }
  tc_fence_before();
  cluster_sync();                   // the leader's MMAs and commits no longer touch the peer
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 2 * acc_cols);
  }
}

// 2-SM path on / off (GWN_TG2=0: every big GEMM through the single-SM kernel; A/B runs)
bool tg2_enabled();

template <typename Epi>
int launch_tma_gemm2(const CUtensorMap& mapA, const CUtensorMap& mapB, TgParams& p, const Epi& epi, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return 0;
  GWN_REQUIRE(p.bn % 64 == 0 && p.bn >= 64 && p.bn <= 256, "tma_gemm2: bad tile width %d", p.bn);
  p.m_tiles = (int)cdiv(p.M, 2 * TG_BM);
  p.n_tiles = (int)cdiv(p.N, p.bn);
  p.k_blocks = (int)cdiv(p.K, TG_BK);
  if (p.splits < 1) p.splits = 1;
  p.kb_per_split = (int)cdiv(p.k_blocks, p.splits);
  p.splits = (int)cdiv(p.k_blocks, p.kb_per_split);
  const size_t stage = (size_t)p.a.tile_bytes + p.b.tile_bytes;
  int stages = (int)((size_t)(224 * 1024 - 1024 - 256) / stage);
  if (stages > 6) stages = 6;
  GWN_REQUIRE(stages >= 2, "tma_gemm2: tile does not fit 2 stages");
  p.stages = stages;
  const size_t smem = stages * stage + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    GWN_CUDA(cudaFuncSetAttribute(tma_gemm2_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int pairs_hw = tg_sm_count() / 2;
  const long long tiles = (long long)p.m_tiles * p.n_tiles * p.splits;
  const int pairs = (int)(tiles < pairs_hw ? tiles : pairs_hw);
  GWN_CUDA(launch_pdl(tma_gemm2_kernel<Epi>, dim3(2 * pairs), dim3(TG_THREADS), smem, st, mapA, mapB, p, epi));
  GWN_LAUNCHED();
  return 0;
}

}  // namespace gwn
