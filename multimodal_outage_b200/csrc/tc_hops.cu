// Diffusion hops on the 5th-gen tensor cores (tcgen05), for graphs whose supports fit on chip (V <= 128).
//
//   out[s, w, c] = sum_terms sum_v Mop_t[w, v] * in_t[s, v, c]   (+ add-in)      (nconv, graph_wavenet.py:60-66)
//
// Orientation: M = output node w (128 TMEM lanes, V of them used), N = (slab, channel) = 8 slabs x 32 = 256,
// K = input node v (padded to a multiple of 16).
//   A operand  : the support / its square / their transposes, bf16, RESIDENT in shared memory for the whole
//                kernel, K-major no-swizzle canonical layout (built once per forward by hop_mats_prep_kernel).
//   B operand  : one channels-last activation tile [8 slabs][Kp nodes][32 ch] brought in by ONE TMA box per load
//                (3-D tensor map {32 ch, V nodes, slabs}: nodes >= V and slabs past the end are zero-filled), 64B
//                swizzled = MN-major SWIZZLE_64B operand, one swizzle atom column per slab.
//   D          : fp32 in TMEM, two 256-column accumulators so the epilogue of one output overlaps the MMAs of
//                the next.  Epilogue: tcgen05.ld -> (+ add-in) -> bf16 -> 64-byte stores into the concat slot.
// Roles (416 threads): warp 0 lane 0 TMA producer, warp 4 lane 0 MMA issuer (+TMEM alloc), warps 5-12 epilogue
// (two warps per TMEM lane quadrant, each draining 4 of the 8 slabs of an accumulator, loads issued in pairs).
// Persistent: grid = min(tiles, SMs); each CTA walks tiles of 8 slabs.
#include "tc.cuh"
#include "hop_prep.cuh"
#include "tc_hops.cuh"
#include "gcn_fused.cuh"
#include "tma_gemm.cuh"

namespace gwn {

struct HopMaps { CUtensorMap m[TH_MAX_STEPS]; };

__global__ void __launch_bounds__(TH_THREADS, 1) hops_tc_kernel(const __grid_constant__ HopMaps maps,
                                                                const __grid_constant__ HopParams p) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Kp = p.Kp, V = p.V;
  const uint32_t mat_bytes = (uint32_t)(Kp / 8) * 2048u;
  const uint32_t stage_bytes = 8u * (uint32_t)Kp * 64u;      // [8 slabs][Kp nodes][64 B]
  uint8_t* mats_s = smem;
  uint8_t* stage_s = smem + (size_t)p.n_mats * mat_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_s + 2 * (size_t)stage_bytes);
  uint64_t* full = bars;          // [2]
  uint64_t* empty = bars + 2;     // [2]
  uint64_t* tfull = bars + 4;     // [2]
  uint64_t* tempty = bars + 6;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], TH_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == TH_MMA_WARP) tmem_alloc(tmem_slot, 512);
  {  // resident A operands: global image -> smem (generic proxy writes, then made visible to the async proxy)
    const int per16 = (int)(mat_bytes / 16);
    for (int m = 0; m < p.n_mats; ++m) {
      const uint4* src = reinterpret_cast<const uint4*>(p.mats) + (size_t)p.mat_src[m] * per16;
      uint4* dst = reinterpret_cast<uint4*>(mats_s + (size_t)m * mat_bytes);
      for (int i = tid; i < per16; i += TH_THREADS) dst[i] = __ldg(src + i);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      int g = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int si = 0; si < p.n_steps; ++si) {
          if (!(p.steps[si].flags & TH_LOAD)) continue;
          const int stage = g & 1, phase = (g >> 1) & 1;
          mbar_wait(&empty[stage], (uint32_t)(phase ^ 1));
          tg::mbar_expect_tx(&full[stage], stage_bytes);
          tg::tma_3d(smem_u32(stage_s + (size_t)stage * stage_bytes), &maps.m[si], 0, 0, tile * 8, &full[stage]);
          ++g;
        }
      }
    }
    __syncwarp();
  } else if (warp == TH_MMA_WARP) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 256, /*a_mn=*/false, /*b_mn=*/true);
      const uint32_t mats_addr = smem_u32(mats_s), stage_addr = smem_u32(stage_s);
      int g = 0, stage = 0;
      uint32_t acc_uses[2] = {0u, 0u};
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int si = 0; si < p.n_steps; ++si) {
          const HopStep st = p.steps[si];
          if (st.flags & TH_LOAD) {
            stage = g & 1;
            mbar_wait(&full[stage], (uint32_t)((g >> 1) & 1));
            ++g;
          }
          if (st.flags & TH_FIRST) mbar_wait(&tempty[st.acc], (acc_uses[st.acc] & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t a0 = mats_addr + (uint32_t)st.mat * mat_bytes;
          const uint32_t b0 = stage_addr + (uint32_t)stage * stage_bytes;
          const uint32_t d = tmem_base + (uint32_t)st.acc * 256u;
          for (int ks = 0; ks < Kp / 16; ++ks) {
            const uint64_t adesc = make_smem_desc(a0 + (uint32_t)ks * 4096u, 2048u, 128u);
            const uint64_t bdesc = tg::make_desc_sw(b0 + (uint32_t)ks * 1024u, (uint32_t)Kp * 64u, 512u, 4u);
            umma_bf16(d, adesc, bdesc, idesc, ((st.flags & TH_FIRST) && ks == 0) ? 0u : 1u);
          }
          if (st.flags & TH_RELEASE) umma_commit(&empty[stage]);
          if (st.flags & TH_LAST) {
            umma_commit(&tfull[st.acc]);
            acc_uses[st.acc]++;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp > TH_MMA_WARP) {
    // ===================== epilogue: TMEM -> registers -> (+add) -> bf16 -> concat slot =====================
    const int quad = warp & 3;
    const int half = (warp - (TH_MMA_WARP + 1)) >> 2;   // which 4 slabs of the accumulator this warp drains
    const int w = quad * 32 + lane;
    uint32_t acc_uses[2] = {0u, 0u};
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const long long slab0 = (long long)tile * 8 + half * 4;
      int o = 0;
      for (int si = 0; si < p.n_steps; ++si) {
        const HopStep st = p.steps[si];
        if (!(st.flags & TH_LAST)) continue;
        const HopOut ho = p.outs[o++];
        mbar_wait(&tfull[st.acc], acc_uses[st.acc] & 1u);
        acc_uses[st.acc]++;
        tc_fence_after();
        bf16* obase = p.out[ho.buf] + ho.slot * p.slot_stride[ho.buf];
        const int opitch = p.out_pitch[ho.buf];
        const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)st.acc * 256u +
                               (uint32_t)half * 128u;
#pragma unroll 1
        for (int pr = 0; pr < 2; ++pr) {
          uint32_t r[2][32];
          tmem_ld32_issue(tbase + (uint32_t)(2 * pr) * 32u, r[0]);
          tmem_ld32_issue(tbase + (uint32_t)(2 * pr + 1) * 32u, r[1]);
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const long long slab = slab0 + 2 * pr + h;
            if (w < V && slab < p.slabs) {
              const long long row = slab * V + w;
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[h][j]);
              if (ho.add_buf >= 0) {
                const bf16* ap = p.in[ho.add_buf] + row * (long long)p.in_pitch[ho.add_buf] +
                                 ho.add_slot * p.slot_stride[ho.add_buf];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float t[4];
                  load4(ap + 4 * j, t);
                  v[4 * j] += t[0]; v[4 * j + 1] += t[1]; v[4 * j + 2] += t[2]; v[4 * j + 3] += t[3];
                }
              }
              bf16* dp = obase + row * (long long)opitch;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 pk;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
                __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(dp + 8 * j) = pk;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[st.acc]);    // one arrival per warp
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TH_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ONE launch for the three support images (hop_prep.cuh): blocks [0, nb0) build the per-support images, [nb0, nb0 + nb1)
// the stacked forward image, the rest the stacked transposed-hop image of the fused backward (nb2 == 0: no such kernel
// instance for this shape).
__global__ void __launch_bounds__(256) hop_mats_all_prep_kernel(MatPrep mp, int KT, BtGeom G, bf16* __restrict__ out0,
                                                                bf16* __restrict__ out1, bf16* __restrict__ out2, int nb0, int nb1,
                                                                int nb2) {
  pdl_wait();      // programmatic launch: the launch latency overlaps the predecessor (common.cuh)
  pdl_trigger();
  const int b = (int)blockIdx.x;
  if (b < nb0) hop_mats_prep_body(mp, out0, (long long)b * 256 + threadIdx.x, (long long)nb0 * 256);
  else if (b < nb0 + nb1) hop_mats_t_prep_body(mp, KT, mp.Kp, out1, (long long)(b - nb0) * 256 + threadIdx.x, (long long)nb1 * 256);
  else hop_mats_bt_prep_body(mp, G, out2, (long long)(b - nb0 - nb1) * 256 + threadIdx.x, (long long)nb2 * 256);
}

static int g_sm_count = 0;

int hops_tc_supported(int V, int n_mats) {
  if (V > 128 || n_mats > TH_MAX_MATS) return 0;
  int Kp = ((V + 15) / 16) * 16;
  size_t need = (size_t)n_mats * (Kp / 8) * 2048 + 2 * (size_t)8 * Kp * 64 + 1024 + 128;
  return need <= 227 * 1024 ? 1 : 0;
}

int launch_hops_tc(HopParams& p, cudaStream_t st) {
  if (p.slabs <= 0) return 0;
  p.Kp = ((p.V + 15) / 16) * 16;
  p.n_tiles = (int)cdiv(p.slabs, 8);
  GWN_REQUIRE(p.n_steps >= 1 && p.n_steps <= TH_MAX_STEPS && p.n_outs <= TH_MAX_OUTS, "hops_tc: too many steps");
  GWN_REQUIRE(hops_tc_supported(p.V, p.n_mats), "hops_tc: V=%d with %d resident matrices does not fit in smem", p.V,
              p.n_mats);
  size_t smem = (size_t)p.n_mats * (p.Kp / 8) * 2048 + 2 * (size_t)8 * p.Kp * 64 + 1024 + 128;
  HopMaps maps;
  bool have = false;
  for (int si = 0; si < p.n_steps; ++si) {
    const HopStep& hs = p.steps[si];
    if (!(hs.flags & TH_LOAD)) { if (have) maps.m[si] = maps.m[0]; continue; }
    const bf16* base = p.in[hs.in_buf] + hs.in_slot * p.slot_stride[hs.in_buf];
    if (int rc = tg_map_rows3d_box(&maps.m[si], base, (uint64_t)p.V, (uint64_t)p.slabs, (uint64_t)p.in_pitch[hs.in_buf],
                                   (uint32_t)p.Kp, 8))
      return rc;
    if (!have) { for (int j = 0; j < si; ++j) maps.m[j] = maps.m[si]; have = true; }
  }
  for (int si = p.n_steps; si < TH_MAX_STEPS; ++si) maps.m[si] = maps.m[0];
  if (g_sm_count == 0) {
    int dev = 0;
    GWN_CUDA(cudaGetDevice(&dev));
    GWN_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    GWN_CUDA(cudaFuncSetAttribute(hops_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  int grid = p.n_tiles < g_sm_count ? p.n_tiles : g_sm_count;
  hops_tc_kernel<<<grid, TH_THREADS, smem, st>>>(maps, p);
  GWN_LAUNCHED();
  return 0;
}

}  // namespace gwn

using namespace gwn;

static int hop_mats_t_kt(int V, int n_supports) { return (((1 + 2 * n_supports) * V + 15) / 16) * 16; }

// the 4 n images [Kp/8][128][8], the stacked T-form image [KT/8][NP][8] of the fused forward, and the stacked
// transposed-hop image of the T-form fused backward (gcn_fused_bwd_t.cu: BtGeom)
extern "C" int gwn_hop_mats_bytes(int V, int n_supports) {
  int Kp = ((V + 15) / 16) * 16;
  return n_supports * 4 * (Kp / 8) * 2048 + hop_mats_t_kt(V, n_supports) * Kp * 2 +
         gcn_bwd_t_image_elems(V, 2 * n_supports) * 2;
}

extern "C" int gwn_hop_mats_prep(const float* const* supports, int n_supports, int V, void* out, void* stream) {
  GWN_REQUIRE(supports && out && n_supports >= 1 && n_supports <= GWN_MAX_SUPPORTS && V >= 1 && V <= 128,
              "hop_mats_prep: bad argument (V=%d must be <= 128)", V);
  MatPrep mp{};
  for (int i = 0; i < n_supports; ++i) mp.A[i] = supports[i];
  mp.n = n_supports; mp.V = V; mp.Kp = ((V + 15) / 16) * 16;
  const long long total = (long long)n_supports * 4 * (mp.Kp / 8) * 1024;
  const int KT = hop_mats_t_kt(V, n_supports);
  const BtGeom G(2 * n_supports, ((V + 7) / 8) * 8);
  const int nb0 = (int)cdiv(total, 256), nb1 = (int)cdiv((long long)KT * mp.Kp, 256);
  const int nb2 = (int)cdiv((long long)(G.KW / 8) * G.NTOT * 8, 256);
  bf16* o = reinterpret_cast<bf16*>(out);
  GWN_CUDA(launch_pdl(hop_mats_all_prep_kernel, dim3((unsigned)(nb0 + nb1 + nb2)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                      mp, KT, G, o, o + total, o + total + (long long)KT * mp.Kp, nb0, nb1, nb2));
  GWN_LAUNCHED();
  return 0;
}

// Test/bench entry: one tensor-core hop  y[slot_out] = Mop[mat] * x[slot_in]  over a pitched bf16 buffer.
extern "C" int gwn_hop_tc(const void* mats, int n_mats, int mat, void* buf, int pitch, int slot_in, int slot_out,
                          int slabs, int V, void* stream) {
  GWN_REQUIRE(mats && buf && mat >= 0 && mat < n_mats && pitch % 8 == 0, "hop_tc: bad argument");
  HopParams p{};
  p.in[0] = p.in[1] = reinterpret_cast<const bf16*>(buf);
  p.out[0] = p.out[1] = reinterpret_cast<bf16*>(buf);
  p.in_pitch[0] = p.in_pitch[1] = p.out_pitch[0] = p.out_pitch[1] = pitch;
  p.slot_stride[0] = p.slot_stride[1] = 32;   // row-major [rows, pitch] buffer: slots are column groups
  p.mats = reinterpret_cast<const bf16*>(mats);
  p.n_mats = 1; p.mat_src[0] = mat; p.V = V; p.slabs = slabs;
  p.n_steps = 1; p.n_outs = 1;
  p.steps[0] = HopStep{0, slot_in, 0, 0, TH_LOAD | TH_RELEASE | TH_FIRST | TH_LAST};
  p.outs[0] = HopOut{0, slot_out, -1, 0};
  return launch_hops_tc(p, reinterpret_cast<cudaStream_t>(stream));
}
