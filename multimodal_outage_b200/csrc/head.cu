// Block edges: start_conv (+pad, NCHW -> channels-last)  graph_wavenet.py:191-196
//              skip sum + relu + end_conv_1 + relu + end_conv_2 (-> NCHW)  :231-236, :252-254
#include "gemm.cuh"

namespace gwn {

// ------------------------------------------------------------------------------------------ start conv
// One warp per 4 consecutive time steps of one (n, v); lane = output channel.
template <typename T>
__global__ void __launch_bounds__(256) start_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, T* __restrict__ u0, int N,
                                                        int Cin, int V, int Tn, int L0) {
  constexpr int CK = 64, PPW = 4;
  __shared__ float ws[CK][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lgroups = (L0 + PPW - 1) / PPW;
  const long long total = (long long)N * V * lgroups;
  const long long item = (long long)blockIdx.x * 8 + warp;
  const bool active = item < total;
  const int pad = L0 - Tn;
  long long n = 0; int v = 0, l0 = 0;
  if (active) { n = item / ((long long)V * lgroups); long long r = item % ((long long)V * lgroups); v = (int)(r / lgroups); l0 = (int)(r % lgroups) * PPW; }
  float acc[PPW];
  const float bias = b[lane];
#pragma unroll
  for (int i = 0; i < PPW; ++i) acc[i] = bias;
  for (int c0 = 0; c0 < Cin; c0 += CK) {
    __syncthreads();
    for (int i = threadIdx.x; i < CK * 32; i += 256) {
      int ci = i / 32, c = i % 32;
      ws[ci][c] = (c0 + ci < Cin) ? w[(long long)c * Cin + c0 + ci] : 0.f;
    }
    __syncthreads();
    if (active) {
      const int cend = min(CK, Cin - c0);
      for (int ci = 0; ci < cend; ++ci) {
        const float wv = ws[ci][lane];
        const float* xr = x + ((n * Cin + c0 + ci) * V + v) * (long long)Tn;
#pragma unroll
        for (int i = 0; i < PPW; ++i) {
          int l = l0 + i, t = l - pad;
          float xv = (l < L0 && t >= 0) ? __ldg(xr + t) : 0.f;
          acc[i] = fmaf(wv, xv, acc[i]);
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < PPW; ++i) {
      int l = l0 + i;
      if (l < L0) st1(u0 + ((n * L0 + l) * V + v) * 32 + lane, acc[i]);
    }
  }
}

// dw[c, ci] += sum_p du[p,c] x[p,ci];  db[c] += sum_p du[p,c].  grid (position chunks, ci chunks of 32)
template <typename T>
__global__ void __launch_bounds__(256) start_bwd_w_kernel(const float* __restrict__ x, const T* __restrict__ du,
                                                          float* __restrict__ dw, float* __restrict__ db, int N,
                                                          int Cin, int V, int Tn, int L0, long long pos_per_block) {
  __shared__ float red[8][32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.y * 32;
  const int pad = L0 - Tn;
  const long long P = (long long)N * L0 * V;
  long long pb = (long long)blockIdx.x * pos_per_block, pe = min(P, pb + pos_per_block);
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  float bacc = 0.f;
  for (long long p = pb + warp; p < pe; p += 8) {
    long long n = p / ((long long)L0 * V); long long r = p % ((long long)L0 * V);
    int l = (int)(r / V), v = (int)(r % V);
    float g = ld1(du + p * 32 + lane);
    bacc += g;
    int t = l - pad;
    float xv = 0.f;
    if (t >= 0 && c0 + lane < Cin) xv = __ldg(x + ((n * Cin + c0 + lane) * V + v) * (long long)Tn + t);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = fmaf(g, __shfl_sync(0xffffffffu, xv, j), acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) red[warp][j][lane] = acc[j];
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 32; i += 256) {
    int j = i / 32, c = i % 32;
    float s = 0.f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) s += red[wq][j][c];
    if (c0 + j < Cin) atomicAdd(dw + (long long)c * Cin + c0 + j, s);
  }
  if (blockIdx.y == 0) {
    __syncthreads();
    red[warp][0][lane] = bacc;
    __syncthreads();
    if (threadIdx.x < 32) {
      float s = 0.f;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) s += red[wq][0][threadIdx.x];
      atomicAdd(db + threadIdx.x, s);
    }
  }
}

// dx[n,ci,v,t] = sum_c w[c,ci] du[(n, t+pad, v), c].  One warp per position; lanes sweep ci.
template <typename T>
__global__ void __launch_bounds__(256) start_bwd_x_kernel(const float* __restrict__ w, const T* __restrict__ du,
                                                          float* __restrict__ dx, int N, int Cin, int V, int Tn,
                                                          int L0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pad = L0 - Tn;
  const long long total = (long long)N * V * Tn;
  const long long item = (long long)blockIdx.x * 8 + warp;
  if (item >= total) return;
  long long n = item / ((long long)V * Tn); long long r = item % ((long long)V * Tn);
  int v = (int)(r / Tn), t = (int)(r % Tn);
  float g = ld1(du + ((n * L0 + t + pad) * V + v) * 32 + lane);
  for (int c0 = 0; c0 < Cin; c0 += 32) {
    int ci = c0 + lane;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      float gc = __shfl_sync(0xffffffffu, g, c);
      if (ci < Cin) s = fmaf(__ldg(w + (long long)c * Cin + ci), gc, s);
    }
    if (ci < Cin) dx[((n * Cin + ci) * V + v) * (long long)Tn + t] = s;
  }
}

// ------------------------------------------------------------------------------------------ start conv, narrow input
// Cin <= 8 (config 1-3, 5: in_dim = 2): the contraction is a handful of FMAs per output.  A thread produces 8 output
// channels of one position (one 16-byte bf16 / two 16-byte fp32 stores): four lanes per position, 8 positions per
// warp instruction; the Cin inputs of a position are the same address for its four lanes.
template <typename T, int CIN>
__global__ void __launch_bounds__(256) start_fwd_small_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ b, T* __restrict__ u0, int N,
                                                              int V, int Tn, int L0) {
  pdl_wait();      // programmatic launch: the launch latency overlaps the predecessor (common.cuh)
  pdl_trigger();
  const int cg = threadIdx.x & 3;                           // channels [8 cg, 8 cg + 8)
  float wr[8][CIN], bias[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    bias[c] = b[8 * cg + c];
#pragma unroll
    for (int i = 0; i < CIN; ++i) wr[c][i] = w[(8 * cg + c) * CIN + i];
  }
  const int pad = L0 - Tn;
  const uint32_t P = (uint32_t)((long long)N * L0 * V);     // launcher: P < 2^31
  const uint32_t stride = gridDim.x * 64u;
  for (uint32_t p = blockIdx.x * 64u + (threadIdx.x >> 2); p < P; p += stride) {
    const uint32_t r = p / (uint32_t)V;
    const int v = (int)(p - r * (uint32_t)V);
    const uint32_t n = r / (uint32_t)L0;
    const int l = (int)(r - n * (uint32_t)L0);
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = bias[c];
    if (l >= pad) {
      const float* xr = x + (((long long)n * CIN) * V + v) * (long long)Tn + (l - pad);
#pragma unroll
      for (int i = 0; i < CIN; ++i) {
        const float xv = __ldg(xr + (long long)i * V * Tn);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = fmaf(wr[c][i], xv, acc[c]);
      }
    }
    store8(u0 + (size_t)p * 32 + 8 * cg, acc);
  }
}

// dw[c][ci] += sum_p du[p][c] x[p][ci], db[c] += sum_p du[p][c].  A thread owns 8 channels (one 16/32-byte load per
// position) of every fourth-lane-group position: 8 positions per warp instruction instead of one, the index split
// amortised over 8 channels; per-thread accumulators, one shuffle + shared-memory reduction at the end.
template <typename T, int CIN>
__global__ void __launch_bounds__(256) start_bwd_w_small_kernel(const float* __restrict__ x, const T* __restrict__ du,
                                                                float* __restrict__ dw, float* __restrict__ db, int N,
                                                                int V, int Tn, int L0) {
  pdl_wait();      // programmatic launch: the launch latency overlaps the predecessor (common.cuh)
  pdl_trigger();
  constexpr int NA = 8 * (CIN + 1);
  __shared__ float red[8][4][NA];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = lane & 3;                                  // channels [8 cg, 8 cg + 8)
  const int pad = L0 - Tn;
  const uint32_t P = (uint32_t)((long long)N * L0 * V);     // launcher: P < 2^31
  float acc[CIN + 1][8];
#pragma unroll
  for (int i = 0; i <= CIN; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
  const uint32_t stride = gridDim.x * 64u;
  for (uint32_t p = blockIdx.x * 64u + (threadIdx.x >> 2); p < P; p += stride) {
    const uint32_t r = p / (uint32_t)V;
    const int v = (int)(p - r * (uint32_t)V);
    const uint32_t n = r / (uint32_t)L0;
    const int l = (int)(r - n * (uint32_t)L0);
    float g[8];
    load8(du + (size_t)p * 32 + 8 * cg, g);
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[CIN][c] += g[c];
    if (l >= pad) {
      const float* xr = x + (((long long)n * CIN) * V + v) * (long long)Tn + (l - pad);
#pragma unroll
      for (int i = 0; i < CIN; ++i) {
        const float xv = __ldg(xr + (long long)i * V * Tn);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[i][c] = fmaf(g[c], xv, acc[i][c]);
      }
    }
  }
  // lanes with the same channel group: xor 4, 8, 16
#pragma unroll
  for (int i = 0; i <= CIN; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float t = acc[i][c];
      t += __shfl_xor_sync(0xffffffffu, t, 4);
      t += __shfl_xor_sync(0xffffffffu, t, 8);
      t += __shfl_xor_sync(0xffffffffu, t, 16);
      acc[i][c] = t;
    }
  if (lane < 4) {
#pragma unroll
    for (int i = 0; i <= CIN; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c) red[warp][lane][i * 8 + c] = acc[i][c];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 4 * NA; t += 256) {
    const int g4 = t / NA, k = t % NA;                      // channel group, (i, c)
    const int i = k >> 3, c = 8 * g4 + (k & 7);
    float s = 0.f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) s += red[wq][g4][k];
    if (i < CIN) atomicAdd(dw + c * CIN + i, s); else atomicAdd(db + c, s);
  }
}

// ------------------------------------------------------------------------------------------ head epilogues
struct EpiBiasAct {  // out[p, col] = act(acc + bias)
  static constexpr bool kStats = false;
  double* stats;
  const float* bias; float* out; int pitch; int relu;
  __device__ __forceinline__ void apply(long long p, long long, long long, int col, float v[4], float*,
                                        float*) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] += __ldg(bias + col + j);
      if (relu) v[j] = fmaxf(v[j], 0.f);
    }
    store4(out + p * pitch + col, v);
  }
};
struct EpiMask {  // out = acc * [mask > 0]
  static constexpr bool kStats = false;
  double* stats;
  const float* mask; float* out; int pitch;
  __device__ __forceinline__ void apply(long long p, long long, long long, int col, float v[4], float*,
                                        float*) const {
    float m[4]; load4(mask + p * pitch + col, m);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = m[j] > 0.f ? v[j] : 0.f;
    store4(out + p * pitch + col, v);
  }
};
template <typename T>
struct EpiGroupStore {  // column group g = col/32 goes to its own [P,32] tensor
  static constexpr bool kStats = false;
  double* stats;
  T* outs[GWN_MAX_LAYERS];
  __device__ __forceinline__ void apply(long long p, long long, long long, int col, float v[4], float*,
                                        float*) const {
    store4(outs[col >> 5] + p * 32 + (col & 31), v);
  }
};

// [P, Opad] rows (n,l,v)  <->  NCHW [N,O,V,L]
__global__ void cl_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, long long N, int O,
                                  int V, int L, int Opad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = N * O * V * L;
  if (i >= total) return;
  int l = (int)(i % L); long long r = i / L;
  int v = (int)(r % V); r /= V;
  int o = (int)(r % O); long long n = r / O;
  dst[i] = src[((n * L + l) * V + v) * Opad + o];
}
__global__ void nchw_to_cl_kernel(const float* __restrict__ src, float* __restrict__ dst, long long N, int O,
                                  int V, int L, int Opad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = N * L * V * Opad;
  if (i >= total) return;
  int o = (int)(i % Opad); long long r = i / Opad;
  int v = (int)(r % V); r /= V;
  int l = (int)(r % L); long long n = r / L;
  dst[i] = (o < O) ? src[((n * O + o) * V + v) * L + l] : 0.f;
}

static void chunks_from_matrix(GemmA& A, const float* base, int cols, long long rows_per_n, long long P) {
  A.n_chunks = cols / 32; A.rows_per_n_out = rows_per_n; A.P = P;
  for (int q = 0; q < A.n_chunks; ++q) {
    AChunk& ch = A.ch[q];
    ch = AChunk{};
    ch.base = base; ch.rows_per_n = rows_per_n; ch.row_off = 0; ch.pitch = cols; ch.col_off = q * 32;
    ch.w_off = (long long)q * 32;
  }
}

template <typename TA, typename Epi>
static int gemm_any(const GemmA& A, const float* W, int ldw, const Epi& e, cudaStream_t st) {
  return (ldw % 64 == 0) ? launch_pos_gemm<TA, 64>(A, W, ldw, e, st) : launch_pos_gemm<TA, 32>(A, W, ldw, e, st);
}
template <typename TA, typename Epi>
static int gemm_any_wt(const GemmA& A, const float* W, int ldw, int ldk, const Epi& e, cudaStream_t st) {
  return (ldw % 64 == 0) ? launch_pos_gemm_wt<TA, 64>(A, W, ldw, ldk, e, st)
                         : launch_pos_gemm_wt<TA, 32>(A, W, ldw, ldk, e, st);
}

static int check_head(const gwn_head_cfg* c) {
  GWN_REQUIRE(c != nullptr, "head cfg NULL");
  GWN_REQUIRE(c->S % 32 == 0 && c->E % 32 == 0 && c->S / 32 <= GEMM_MAX_CHUNKS && c->E / 32 <= GEMM_MAX_CHUNKS,
              "skip/end channels must be multiples of 32 and <= %d (got %d, %d)", 32 * GEMM_MAX_CHUNKS, c->S, c->E);
  GWN_REQUIRE(c->n_layers >= 1 && c->n_layers <= GWN_MAX_LAYERS, "n_layers %d unsupported", c->n_layers);
  GWN_REQUIRE(c->O >= 1 && (c->O + 31) / 32 <= GEMM_MAX_CHUNKS, "out_dim %d unsupported", c->O);
  GWN_REQUIRE(c->dtype == GWN_F32 || c->dtype == GWN_BF16, "bad dtype");
  return 0;
}

template <typename T>
static int head_fwd_t(const gwn_head_cfg* c, const gwn_head_fwd_args* a, cudaStream_t st) {
  const long long R = (long long)c->Lf * c->V, P = c->N * R;
  const int Opad = 32 * ((c->O + 31) / 32);
  GemmA Z{};
  Z.n_chunks = c->n_layers; Z.rows_per_n_out = R; Z.P = P;
  for (int i = 0; i < c->n_layers; ++i) {
    Z.ch[i].base = a->z_last[i]; Z.ch[i].rows_per_n = R; Z.ch[i].pitch = 32;
  }
  EpiBiasAct e1{}; e1.bias = a->b_skip; e1.out = a->s1; e1.pitch = c->S; e1.relu = 1;
  if (int rc = gemm_any<T>(Z, a->w_skip, c->S, e1, st)) return rc;
  GemmA S{}; chunks_from_matrix(S, a->s1, c->S, R, P);
  EpiBiasAct e2{}; e2.bias = a->b_end1; e2.out = a->e1; e2.pitch = c->E; e2.relu = 1;
  if (int rc = gemm_any<float>(S, a->w_end1, c->E, e2, st)) return rc;
  GemmA E{}; chunks_from_matrix(E, a->e1, c->E, R, P);
  EpiBiasAct e3{}; e3.bias = a->b_end2; e3.out = a->ws; e3.pitch = Opad; e3.relu = 0;
  if (int rc = gemm_any<float>(E, a->w_end2, Opad, e3, st)) return rc;
  long long total = (long long)c->N * c->O * R;
  cl_to_nchw_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(a->ws, a->out, c->N, c->O, c->V, c->Lf, Opad);
  GWN_LAUNCHED();
  return 0;
}

template <typename T>
static int head_bwd_t(const gwn_head_cfg* c, const gwn_head_bwd_args* a, cudaStream_t st) {
  const long long R = (long long)c->Lf * c->V, P = c->N * R;
  const int Opad = 32 * ((c->O + 31) / 32);
  long long total = P * Opad;
  nchw_to_cl_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(a->dout, a->ws_do, c->N, c->O, c->V, c->Lf, Opad);
  GWN_LAUNCHED();
  // end_conv_2
  GemmA E{}; chunks_from_matrix(E, a->e1, c->E, R, P);
  if (int rc = launch_wgrad<float, float>(E, a->ws_do, Opad, 0, a->dw_end2, Opad, a->db_end2, st)) return rc;
  GemmA DO{}; chunks_from_matrix(DO, a->ws_do, Opad, R, P);
  EpiMask m1{}; m1.mask = a->e1; m1.out = a->ws_de1; m1.pitch = c->E;
  if (int rc = gemm_any_wt<float>(DO, a->w_end2, c->E, Opad, m1, st)) return rc;
  // end_conv_1
  GemmA S{}; chunks_from_matrix(S, a->s1, c->S, R, P);
  if (int rc = launch_wgrad<float, float>(S, a->ws_de1, c->E, 0, a->dw_end1, c->E, a->db_end1, st)) return rc;
  GemmA DE{}; chunks_from_matrix(DE, a->ws_de1, c->E, R, P);
  EpiMask m2{}; m2.mask = a->s1; m2.out = a->ws_ds1; m2.pitch = c->S;
  if (int rc = gemm_any_wt<float>(DE, a->w_end1, c->S, c->E, m2, st)) return rc;
  // skip convs
  GemmA Z{};
  Z.n_chunks = c->n_layers; Z.rows_per_n_out = R; Z.P = P;
  for (int i = 0; i < c->n_layers; ++i) {
    Z.ch[i].base = a->z_last[i]; Z.ch[i].rows_per_n = R; Z.ch[i].pitch = 32;
  }
  if (int rc = launch_wgrad<T, float>(Z, a->ws_ds1, c->S, 0, a->dw_skip, c->S, a->db_skip, st)) return rc;
  GemmA DS{}; chunks_from_matrix(DS, a->ws_ds1, c->S, R, P);
  EpiGroupStore<T> gs{};
  for (int i = 0; i < c->n_layers; ++i) gs.outs[i] = reinterpret_cast<T*>(a->dz_last[i]);
  return launch_pos_gemm_wt<float, 32>(DS, a->w_skip, 32 * c->n_layers, c->S, gs, st);
}

}  // namespace gwn

using namespace gwn;

extern "C" int gwn_start_fwd(const float* x, const float* w, const float* b, void* u0, int dtype, int N, int Cin,
                             int V, int T, int L0, void* stream) {
  GWN_REQUIRE(x && w && b && u0 && N >= 1 && Cin >= 1 && V >= 1 && T >= 1 && L0 >= T, "start_fwd: bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((Cin == 2 || Cin == 1 || Cin == 4) && (long long)N * L0 * V < (1ll << 31)) {   // narrow inputs: warp-per-position kernel (64-bit index math only, no smem)
    const unsigned nb = 148 * 8;
#define GWN_SF(CI)                                                                                                   \
  if (dtype == GWN_F32) GWN_CUDA(launch_pdl(start_fwd_small_kernel<float, CI>, dim3(nb), dim3(256), 0, st, x, w, b, (float*)u0, N, V, T, L0));      \
  else if (dtype == GWN_BF16) GWN_CUDA(launch_pdl(start_fwd_small_kernel<bf16, CI>, dim3(nb), dim3(256), 0, st, x, w, b, (bf16*)u0, N, V, T, L0));  \
  else GWN_REQUIRE(false, "bad dtype %d", dtype);
    if (Cin == 1) { GWN_SF(1) } else if (Cin == 2) { GWN_SF(2) } else { GWN_SF(4) }
#undef GWN_SF
    GWN_LAUNCHED();
    return 0;
  }
  long long items = (long long)N * V * ((L0 + 3) / 4);
  unsigned blocks = (unsigned)cdiv(items, 8);
  if (dtype == GWN_F32) start_fwd_kernel<float><<<blocks, 256, 0, st>>>(x, w, b, (float*)u0, N, Cin, V, T, L0);
  else if (dtype == GWN_BF16) start_fwd_kernel<bf16><<<blocks, 256, 0, st>>>(x, w, b, (bf16*)u0, N, Cin, V, T, L0);
  else GWN_REQUIRE(false, "bad dtype %d", dtype);
  GWN_LAUNCHED();
  return 0;
}

extern "C" int gwn_start_bwd(const float* x, const float* w, const void* du0, int dtype, float* dw, float* db,
                             float* dx, int N, int Cin, int V, int T, int L0, void* stream) {
  GWN_REQUIRE(x && w && du0 && dw && db && L0 >= T, "start_bwd: bad argument");
  GWN_REQUIRE(dtype == GWN_F32 || dtype == GWN_BF16, "bad dtype %d", dtype);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GWN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 32 * (size_t)Cin, st));
  GWN_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * 32, st));
  const long long P = (long long)N * L0 * V;
  if ((Cin == 2 || Cin == 1 || Cin == 4) && P < (1ll << 31)) {
    const unsigned nb = 148 * 4;
#define GWN_SB(CI)                                                                                                       \
  if (dtype == GWN_F32) GWN_CUDA(launch_pdl(start_bwd_w_small_kernel<float, CI>, dim3(nb), dim3(256), 0, st, x, (const float*)du0, dw, db, N, V, T, L0)); \
  else GWN_CUDA(launch_pdl(start_bwd_w_small_kernel<bf16, CI>, dim3(nb), dim3(256), 0, st, x, (const bf16*)du0, dw, db, N, V, T, L0));
    if (Cin == 1) { GWN_SB(1) } else if (Cin == 2) { GWN_SB(2) } else { GWN_SB(4) }
#undef GWN_SB
    GWN_LAUNCHED();
  } else {
  int cchunks = (Cin + 31) / 32;
  long long want_blocks = cdiv(148 * 8, cchunks);
  long long per = cdiv(P, want_blocks);
  if (per < 64) per = 64;
  dim3 grid((unsigned)cdiv(P, per), (unsigned)cchunks);
  if (dtype == GWN_F32) start_bwd_w_kernel<float><<<grid, 256, 0, st>>>(x, (const float*)du0, dw, db, N, Cin, V, T, L0, per);
  else start_bwd_w_kernel<bf16><<<grid, 256, 0, st>>>(x, (const bf16*)du0, dw, db, N, Cin, V, T, L0, per);
  GWN_LAUNCHED();
  }
  if (dx) {
    long long items = (long long)N * V * T;
    unsigned blocks = (unsigned)cdiv(items, 8);
    if (dtype == GWN_F32) start_bwd_x_kernel<float><<<blocks, 256, 0, st>>>(w, (const float*)du0, dx, N, Cin, V, T, L0);
    else start_bwd_x_kernel<bf16><<<blocks, 256, 0, st>>>(w, (const bf16*)du0, dx, N, Cin, V, T, L0);
    GWN_LAUNCHED();
  }
  return 0;
}

extern "C" int gwn_head_fwd(const gwn_head_cfg* cfg, const gwn_head_fwd_args* a, void* stream) {
  if (int rc = check_head(cfg)) return rc;
  GWN_REQUIRE(a && a->w_skip && a->b_skip && a->w_end1 && a->b_end1 && a->w_end2 && a->b_end2 && a->s1 && a->e1 &&
                  a->out && a->ws, "head_fwd: NULL argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return cfg->dtype == GWN_F32 ? head_fwd_t<float>(cfg, a, st) : head_fwd_t<bf16>(cfg, a, st);
}

extern "C" int gwn_head_bwd(const gwn_head_cfg* cfg, const gwn_head_bwd_args* a, void* stream) {
  if (int rc = check_head(cfg)) return rc;
  GWN_REQUIRE(a && a->w_skip && a->w_end1 && a->w_end2 && a->s1 && a->e1 && a->dout && a->dw_skip && a->db_skip &&
                  a->dw_end1 && a->db_end1 && a->dw_end2 && a->db_end2 && a->ws_do && a->ws_de1 && a->ws_ds1,
              "head_bwd: NULL argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return cfg->dtype == GWN_F32 ? head_bwd_t<float>(cfg, a, st) : head_bwd_t<bf16>(cfg, a, st);
}
