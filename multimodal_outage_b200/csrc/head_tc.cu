// bf16 head on the tensor cores: skip sum + relu + end_conv_1 + relu + end_conv_2, forward and backward
// (graph_wavenet.py:231-236 via the crop identity, :252-254), as nine TMA-fed tcgen05 GEMMs (tma_gemm.cuh).
//
//   forward   s1 = relu(zcat Ws + bs)         [P,32nl]x[32nl,S]      A K-major, B = Ws^T image K-major
//             e1 = relu(s1 W1 + b1)           [P,S]x[S,E]
//             out = e1 W2 + b2 -> NCHW fp32   [P,E]x[E,O]
//   The two GEMMs that feed a ReLU run in SPLIT bf16 precision (x = hi + lo, two/three MMAs per product:
//   z.(Ws_hi + Ws_lo);  s1_hi.W1_hi + s1_lo.W1_hi + s1_hi.W1_lo): a bf16-rounded pre-activation flips ~0.2%
//   of the relu masks, which moves every upstream gradient by ~sqrt(0.002) = 4% - above the 2e-2 bar.  The
//   split costs 2-3x of a GEMM that is < 1% of the step.  s1 is therefore saved as [P][2S] = [hi | lo].
//   backward  de1 = (do W2^T) . [e1>0]   ds1 = (de1 W1^T) . [s1>0]   dz_i = ds1 Ws_i^T      (data grads, K-major)
//             dW2 = e1^T do   dW1 = s1^T de1   dWs = zcat^T ds1     (reduction over positions: both operands
//             MN-major, split-K over the position axis, fp32 atomic accumulation); bias grads are the column
//             sums of do / de1 / ds1, reduced in the epilogue that produces them.
// P = N*Lf*V positions; zcat = the per-layer z[..., -Lf:] slices concatenated along channels.
#include "tc_gemm_impl.cuh"   // warp_column_sums
#include "tma_gemm.cuh"

namespace gwn {

__device__ __forceinline__ void store_bf16x32(bf16* dst, const float v[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 pk;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]);
    __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
    __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(dst + 8 * j) = pk;
  }
}

struct EpiBiasReluBf16 {   // out[m][n] = relu(acc + bias[n]) (bf16); split: also out[m][N + n] = bf16 residual
  const float* bias; bf16* out; int ldo, N, split;
  __device__ __forceinline__ void chunk(int m, bool m_ok, int n0, float v[32]) const {
    if (!m_ok || n0 >= N) return;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + __ldg(bias + n0 + j), 0.f);
    store_bf16x32(out + (long long)m * ldo + n0, v);
    if (split) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] -= __bfloat162float(__float2bfloat16_rn(v[j]));
      store_bf16x32(out + (long long)m * ldo + N + n0, v);
    }
  }  // coalesced form (tma_gemm.cuh: kWarpScratch)
  static constexpr bool kWarpScratch = true;
  __device__ __forceinline__ void chunk_ws(int m0, int M, bool, int n0, float v[32], uint8_t* scr, int lane) const {
    if (n0 >= N) return;
    const int rows = M - m0 < 32 ? M - m0 : 32;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + __ldg(bias + n0 + j), 0.f);
    uint4 q[4];
    tg::pack_bf16x32(v, q);
    tg::ws_store_rows64(scr, q, out + (long long)m0 * ldo + n0, ldo, rows, lane);
    if (split) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] -= __bfloat162float(__float2bfloat16_rn(v[j]));
      tg::pack_bf16x32(v, q);
      tg::ws_store_rows64(scr, q, out + (long long)m0 * ldo + N + n0, ldo, rows, lane);
    }
  }
};

struct EpiOutNCHW {        // out[n, o, v, l] = acc + bias[o]; row m = (n, l, v)
  const float* bias; float* out; int O, V, Lf;
  __device__ __forceinline__ void chunk(int m, bool m_ok, int n0, float v[32]) const {
    if (!m_ok) return;
    const int R = Lf * V;
    const int n = m / R, rem = m % R, l = rem / V, node = rem % V;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int o = n0 + j;
      if (o < O) out[(((long long)n * O + o) * V + node) * Lf + l] = v[j] + __ldg(bias + o);
    }
  }
};

struct EpiMaskColsum {     // out = acc . [mask > 0] (bf16); colsum[n] += sum_m out[m][n]
  const bf16* mask; int ldm; bf16* out; int ld, N; float* colsum;
  __device__ __forceinline__ void chunk(int m, bool m_ok, int n0, float v[32]) const {
    if (n0 >= N) return;
    if (m_ok) {
      const bf16* mp = mask + (long long)m * ldm + n0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t[4];
        load4(mp + 4 * j, t);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[4 * j + i] = t[i] > 0.f ? v[4 * j + i] : 0.f;
      }
      store_bf16x32(out + (long long)m * ld + n0, v);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    const int lane = threadIdx.x & 31;
    const float s = warp_column_sums(v, lane);
    atomicAdd(colsum + n0 + lane, s);
  }  static constexpr bool kWarpScratch = true;
  __device__ __forceinline__ void chunk_ws(int m0, int M, bool m_ok, int n0, float v[32], uint8_t* scr, int lane) const {
    if (n0 >= N) return;
    const int rows = M - m0 < 32 ? M - m0 : 32;
    uint4 mq[4];
    tg::ws_load_rows64(scr, mq, mask + (long long)m0 * ldm + n0, ldm, rows, lane);
    if (m_ok) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w[4] = {mq[j].x, mq[j].y, mq[j].z, mq[j].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {      // bf16 > 0  <=>  its 16 bits, read as a signed integer, are > 0
          if ((int16_t)(w[i] & 0xFFFFu) <= 0) v[8 * j + 2 * i] = 0.f;
          if ((int16_t)(w[i] >> 16) <= 0) v[8 * j + 2 * i + 1] = 0.f;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    uint4 q[4];
    tg::pack_bf16x32(v, q);
    tg::ws_store_rows64(scr, q, out + (long long)m0 * ld + n0, ld, rows, lane);
    const float s = warp_column_sums(v, lane);
    atomicAdd(colsum + n0 + lane, s);
  }
};

struct EpiGroupBf16 {      // column group n0/32 -> its own [P,32] bf16 tensor
  bf16* outs[GWN_MAX_LAYERS]; int N;
  __device__ __forceinline__ void chunk(int m, bool m_ok, int n0, float v[32]) const {
    if (!m_ok || n0 >= N) return;
    store_bf16x32(outs[n0 >> 5] + (long long)m * 32, v);
  }  static constexpr bool kWarpScratch = true;
  __device__ __forceinline__ void chunk_ws(int m0, int M, bool, int n0, float v[32], uint8_t* scr, int lane) const {
    if (n0 >= N) return;
    const int rows = M - m0 < 32 ? M - m0 : 32;
    uint4 q[4];
    tg::pack_bf16x32(v, q);
    tg::ws_store_rows64(scr, q, outs[n0 >> 5] + (long long)m0 * 32, 32, rows, lane);
  }
};

struct EpiAtomicF32 {      // dW[m][n] += acc (split-K partial)
  float* C; int ldc, M, N;
  __device__ __forceinline__ void chunk(int m, bool m_ok, int n0, float v[32]) const {
    if (!m_ok || m >= M) return;
    float* dst = C + (long long)m * ldc + n0;
    // 16-byte vector reductions (a quarter of the L2 transactions of scalar atomics) when the row chunk is whole and aligned
    if (n0 + 32 <= N && (ldc & 3) == 0 && (reinterpret_cast<unsigned long long>(C) & 15ull) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) red_add_v4(dst + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
      return;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n0 + j < N) atomicAdd(dst + j, v[j]);
  }
};

// fp32 [R][C] -> bf16 image(s); up to 3 matrices per launch.
//   transpose = 0: dst[r][c] (pitch ld)          transpose = 1: dst[c][r] (pitch ld)
//   lo_off  >= 0: the bf16 residual (x - bf16(x)) is also written at column offset lo_off
//   hi2_off >= 0: a second copy of the hi part at column offset hi2_off
struct CvtJob { const float* src; bf16* dst; int R, C, transpose, ld, lo_off, hi2_off; };
struct CvtJobs { CvtJob j[3]; int n; };
__global__ void cvt_weights_kernel(CvtJobs jobs) {
  pdl_wait();      // programmatic launch: the launch latency overlaps the predecessor (common.cuh)
  pdl_trigger();
  // one flat index space over the jobs: every thread has its (few) loads in flight at once - with one job after the other a
  // thread paid an L2 round trip per job (8.7 us for 213k elements)
  const int t0 = jobs.j[0].R * jobs.j[0].C;   // (weights: far below 2^31; 32-bit index math - a 64-bit division costs hundreds of cycles)
  const int t1 = t0 + (jobs.n > 1 ? jobs.j[1].R * jobs.j[1].C : 0);
  const int t2 = t1 + (jobs.n > 2 ? jobs.j[2].R * jobs.j[2].C : 0);
  constexpr int U = 2;
  const int stride = gridDim.x * blockDim.x;
  for (int e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < t2; e0 += U * stride) {
    float x[U]; int q[U], idx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * stride;
      q[u] = e < t0 ? 0 : (e < t1 ? 1 : 2);
      idx[u] = e - (q[u] == 0 ? 0 : (q[u] == 1 ? t0 : t1));
      x[u] = e < t2 ? __ldg(jobs.j[q[u]].src + idx[u]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (e0 + u * stride >= t2) break;
      const CvtJob& J = jobs.j[q[u]];
      const int i = idx[u];
      const int r = i / J.C, c = i - r * J.C;
      const bf16 hi = __float2bfloat16_rn(x[u]);
      const int o = J.transpose ? c * J.ld + r : r * J.ld + c;
      J.dst[o] = hi;
      if (J.lo_off >= 0) J.dst[o + J.lo_off] = __float2bfloat16_rn(x[u] - __bfloat162float(hi));
      if (J.hi2_off >= 0) J.dst[o + J.hi2_off] = hi;
    }
  }
}

// dout NCHW fp32 [N,O,V,Lf] -> do bf16 [P][Opad] (zero padded), db2[o] += sum_p do[p][o]
__global__ void __launch_bounds__(256) dout_to_cl_kernel(const float* __restrict__ src, bf16* __restrict__ dst,
                                                         float* __restrict__ db, long long P, int O, int V, int Lf,
                                                         int Opad) {
  pdl_wait();      // programmatic launch: the launch latency overlaps the predecessor (common.cuh)
  pdl_trigger();
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = Opad / 32;
  float acc = 0.f;
  int my_o = -1;
  const long long items = P * groups;
  for (long long it = (long long)blockIdx.x * 8 + warp; it < items; it += (long long)gridDim.x * 8) {
    const long long p = it / groups; const int o = (int)(it % groups) * 32 + lane;
    const int R = Lf * V;
    const long long n = p / R; const int rem = (int)(p % R), l = rem / V, node = rem % V;
    const float v = (o < O) ? src[((n * O + o) * V + node) * Lf + l] : 0.f;
    dst[p * Opad + o] = __float2bfloat16_rn(v);
    if (groups == 1) { acc += v; my_o = o; } else if (o < O) atomicAdd(db + o, v);
  }
  if (groups == 1) {
    red[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][lane];
      if (lane < O) atomicAdd(db + lane, s);
    }
  }
  (void)my_o;
}

// C[P][N] = A[P][K] . B[N][K]^T : both operands K-major
static int data_gemm(const bf16* A, int lda, long long P, int K, const bf16* B, int ldb, int N, int bn, TgParams& p,
                     CUtensorMap& ma, CUtensorMap& mb) {
  if (int rc = tg_map_rows(&ma, A, (uint64_t)P, (uint64_t)K, (uint64_t)lda, 128)) return rc;
  if (int rc = tg_map_rows(&mb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)bn)) return rc;
  p = TgParams{};
  p.M = (int)P; p.N = N; p.K = K; p.bn = bn; p.splits = 1;
  tg_operand(p.a, TG_K_SW128, 128);
  tg_operand(p.b, TG_K_SW128, bn);
  return 0;
}

static int pick_bn(int N) { return N >= 256 ? 256 : (N >= 128 ? 128 : (N >= 64 ? 64 : 32)); }

// dW[M][N] (+)= A^T B over positions: A [P][lda] (M cols used), B [P][ldb] (N cols used)
static int wgrad_gemm(const bf16* A, int lda, int M, const bf16* B, int ldb, int N, long long P, float* dW, int ldw,
                      cudaStream_t st) {
  CUtensorMap ma, mb;
  if (int rc = tg_map_2d(&ma, A, (uint64_t)M, (uint64_t)P, (uint64_t)lda * 2, 64, 64)) return rc;
  if (int rc = tg_map_2d(&mb, B, (uint64_t)N, (uint64_t)P, (uint64_t)ldb * 2, 64, 64)) return rc;
  TgParams p{};
  const int bn = N >= 256 ? 256 : (N >= 128 ? 128 : 64);
  p.M = M; p.N = N; p.K = (int)P; p.bn = bn;
  const long long tiles = cdiv(M, 128) * cdiv(N, bn);
  // split-K over positions: enough CTAs to fill the GPU, but >= 8 K blocks per split so the fp32 atomic tail
  // (128 x bn vector reductions per CTA) stays small next to the streaming (16 left the skip GEMM on 64 CTAs)
  long long splits = cdiv(tg_sm_count(), tiles);
  const long long kb = cdiv(P, TG_BK);
  if (splits > kb / 8) splits = kb / 8;
  if (splits < 1) splits = 1;
  p.splits = (int)splits;
  tg_operand(p.a, TG_MN_SW128, 128);
  tg_operand(p.b, TG_MN_SW128, bn);
  EpiAtomicF32 e{dW, ldw, M, N};
  return launch_tma_gemm(ma, mb, p, e, st);
}

}  // namespace gwn

using namespace gwn;

extern "C" long long gwn_head_tc_ws_bytes(int n_layers, int S, int E, int O) {
  const long long Opad = 32 * ((O + 31) / 32);
  const long long K0p = 64 * ((32 * n_layers + 63) / 64);
  return 2 * ((long long)S * 2 * K0p + (long long)E * 3 * S + Opad * E) + 1024;
}

extern "C" int gwn_head_fwd_tc(const gwn_head_cfg* c, const gwn_head_tc_fwd_args* a, void* stream) {
  GWN_REQUIRE(c && a && a->zcat && a->w_skip && a->b_skip && a->w_end1 && a->b_end1 && a->w_end2 && a->b_end2 && a->s1 &&
                  a->e1 && a->out && a->ws_w, "head_fwd_tc: NULL argument");
  GWN_REQUIRE(c->S % 32 == 0 && c->E % 32 == 0 && c->n_layers >= 1 && c->n_layers <= GWN_MAX_LAYERS && c->O >= 1,
              "head_fwd_tc: unsupported channel counts");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long P = (long long)c->N * c->Lf * c->V;
  GWN_REQUIRE(P < (1ll << 31), "head_fwd_tc: too many positions");
  const int K0 = 32 * c->n_layers, S = c->S, E = c->E, Opad = 32 * ((c->O + 31) / 32);
  const int K0p = 64 * ((K0 + 63) / 64);
  bf16* wsT = reinterpret_cast<bf16*>(a->ws_w);          // [S][2*K0p] = [hi(Ws^T) | lo(Ws^T)], zero padded
  bf16* w1T = wsT + (long long)S * 2 * K0p;              // [E][3*S]   = [hi(W1^T) | hi(W1^T) | lo(W1^T)]
  bf16* w2T = w1T + (long long)E * 3 * S;                // [Opad][E]  = W2^T
  if (K0p != K0) GWN_CUDA(cudaMemsetAsync(wsT, 0, sizeof(bf16) * (size_t)S * 2 * K0p, st));
  CvtJobs jobs{};
  jobs.n = 3;
  jobs.j[0] = CvtJob{a->w_skip, wsT, K0, S, 1, 2 * K0p, K0p, -1};
  jobs.j[1] = CvtJob{a->w_end1, w1T, S, E, 1, 3 * S, 2 * S, S};
  jobs.j[2] = CvtJob{a->w_end2, w2T, E, Opad, 1, E, -1, -1};
  GWN_CUDA(launch_pdl(cvt_weights_kernel, dim3(592), dim3(256), 0, st, jobs));
  GWN_LAUNCHED();
  TgParams p; CUtensorMap ma, mb;
  {   // x1 = zcat.(Ws_hi + Ws_lo): the A operand is read twice (k wraps at K0p)
    const int bn = pick_bn(S);
    if (int rc = data_gemm(reinterpret_cast<const bf16*>(a->zcat), K0, P, K0, wsT, 2 * K0p, S, bn, p, ma, mb)) return rc;
    if (int rc = tg_map_rows(&mb, wsT, (uint64_t)S, (uint64_t)(2 * K0p), (uint64_t)(2 * K0p), (uint32_t)bn)) return rc;
    p.K = K0p + K0; p.a_kwrap = K0p;
    EpiBiasReluBf16 e{a->b_skip, reinterpret_cast<bf16*>(a->s1), 2 * S, S, 1};
    if (int rc = launch_tma_gemm(ma, mb, p, e, st)) return rc;
  }
  {   // x2 = s1_hi.W1_hi + s1_lo.W1_hi + s1_hi.W1_lo: s1 = [hi | lo] wraps at 2S
    const int bn = pick_bn(E);
    if (int rc = data_gemm(reinterpret_cast<const bf16*>(a->s1), 2 * S, P, 2 * S, w1T, 3 * S, E, bn, p, ma, mb)) return rc;
    if (int rc = tg_map_rows(&mb, w1T, (uint64_t)E, (uint64_t)(3 * S), (uint64_t)(3 * S), (uint32_t)bn)) return rc;
    p.K = 3 * S; p.a_kwrap = 2 * S;
    EpiBiasReluBf16 e{a->b_end1, reinterpret_cast<bf16*>(a->e1), E, E, 0};
    if (int rc = launch_tma_gemm(ma, mb, p, e, st)) return rc;
  }
  {
    const int bn = pick_bn(Opad);
    if (int rc = data_gemm(reinterpret_cast<const bf16*>(a->e1), E, P, E, w2T, E, Opad, bn, p, ma, mb)) return rc;
    EpiOutNCHW e{a->b_end2, a->out, c->O, c->V, c->Lf};
    if (int rc = launch_tma_gemm(ma, mb, p, e, st)) return rc;
  }
  return 0;
}

extern "C" int gwn_head_bwd_tc(const gwn_head_cfg* c, const gwn_head_tc_bwd_args* a, void* stream) {
  GWN_REQUIRE(c && a && a->zcat && a->w_skip && a->w_end1 && a->w_end2 && a->s1 && a->e1 && a->dout && a->dw_skip &&
                  a->db_skip && a->dw_end1 && a->db_end1 && a->dw_end2 && a->db_end2 && a->ws_do && a->ws_de1 &&
                  a->ws_ds1 && a->ws_w, "head_bwd_tc: NULL argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long P = (long long)c->N * c->Lf * c->V;
  GWN_REQUIRE(P < (1ll << 31), "head_bwd_tc: too many positions");
  const int K0 = 32 * c->n_layers, S = c->S, E = c->E, Opad = 32 * ((c->O + 31) / 32);
  bf16* ws = reinterpret_cast<bf16*>(a->ws_w);           // Ws  [K0][S]
  bf16* w1 = ws + (long long)S * K0;                     // W1  [S][E]
  bf16* w2 = w1 + (long long)E * S;                      // W2  [E][Opad]
  CvtJobs jobs{};
  jobs.n = 3;
  jobs.j[0] = CvtJob{a->w_skip, ws, K0, S, 0, S, -1, -1};
  jobs.j[1] = CvtJob{a->w_end1, w1, S, E, 0, E, -1, -1};
  jobs.j[2] = CvtJob{a->w_end2, w2, E, Opad, 0, Opad, -1, -1};
  GWN_CUDA(launch_pdl(cvt_weights_kernel, dim3(592), dim3(256), 0, st, jobs));
  GWN_LAUNCHED();
  if (!a->outputs_zeroed) {
  GWN_CUDA(cudaMemsetAsync(a->dw_skip, 0, sizeof(float) * (size_t)K0 * S, st));
  GWN_CUDA(cudaMemsetAsync(a->db_skip, 0, sizeof(float) * S, st));
  GWN_CUDA(cudaMemsetAsync(a->dw_end1, 0, sizeof(float) * (size_t)S * E, st));
  GWN_CUDA(cudaMemsetAsync(a->db_end1, 0, sizeof(float) * E, st));
  GWN_CUDA(cudaMemsetAsync(a->dw_end2, 0, sizeof(float) * (size_t)E * Opad, st));
  GWN_CUDA(cudaMemsetAsync(a->db_end2, 0, sizeof(float) * Opad, st));
  }
  bf16* d_o = reinterpret_cast<bf16*>(a->ws_do);
  bf16* de1 = reinterpret_cast<bf16*>(a->ws_de1);
  bf16* ds1 = reinterpret_cast<bf16*>(a->ws_ds1);
  const bf16* s1 = reinterpret_cast<const bf16*>(a->s1);
  const bf16* e1 = reinterpret_cast<const bf16*>(a->e1);
  const bf16* zcat = reinterpret_cast<const bf16*>(a->zcat);
  {
    long long items = P * (Opad / 32);
    long long blocks = cdiv(items, 8 * 4);       // (every item is one dependent gather: 16 per warp took 10 us for 1.6 MB)
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    GWN_CUDA(launch_pdl(dout_to_cl_kernel, dim3((unsigned)blocks), dim3(256), 0, st, a->dout, d_o, a->db_end2, P, c->O, c->V, c->Lf, Opad));
    GWN_LAUNCHED();
  }
  // end_conv_2
  if (int rc = wgrad_gemm(e1, E, E, d_o, Opad, Opad, P, a->dw_end2, Opad, st)) return rc;
  TgParams p; CUtensorMap ma, mb;
  {
    const int bn = pick_bn(E);
    if (int rc = data_gemm(d_o, Opad, P, Opad, w2, Opad, E, bn, p, ma, mb)) return rc;
    EpiMaskColsum e{e1, E, de1, E, E, a->db_end1};
    if (int rc = launch_tma_gemm(ma, mb, p, e, st)) return rc;
  }
  // end_conv_1
  if (int rc = wgrad_gemm(s1, 2 * S, S, de1, E, E, P, a->dw_end1, E, st)) return rc;   // hi part of s1
  {
    const int bn = pick_bn(S);
    if (int rc = data_gemm(de1, E, P, E, w1, E, S, bn, p, ma, mb)) return rc;
    EpiMaskColsum e{s1, 2 * S, ds1, S, S, a->db_skip};
    if (int rc = launch_tma_gemm(ma, mb, p, e, st)) return rc;
  }
  // skip convs
  if (int rc = wgrad_gemm(zcat, K0, K0, ds1, S, S, P, a->dw_skip, S, st)) return rc;
  {
    const int bn = pick_bn(K0);
    if (int rc = data_gemm(ds1, S, P, S, ws, S, K0, bn, p, ma, mb)) return rc;
    EpiGroupBf16 e{};
    e.N = K0;
    for (int i = 0; i < c->n_layers; ++i) {
      GWN_REQUIRE(a->dz_last[i] != nullptr, "head_bwd_tc: NULL dz_last[%d]", i);
      e.outs[i] = reinterpret_cast<bf16*>(a->dz_last[i]);
    }
    if (int rc = launch_tma_gemm(ma, mb, p, e, st)) return rc;
  }
  return 0;
}
