// start_conv (1x1, Cin -> 32) + left zero pad + NCHW -> channels-last, on the tensor cores, for wide inputs
// (config 4: Cin = 256 UNet features + 64 date2vec; graph_wavenet.py:191-196).
//   forward : x NCHW fp32 --tiled transpose--> xcl [N, L0, V, Cin] bf16 (time left-padded with zeros, saved for the
//             backward) ; u0 = xcl W^T + b          one TMA-fed tcgen05 GEMM (tma_gemm.cuh), bf16 out [P0, 32]
//   backward: dW [32, Cin] = du0^T xcl  (split-K over positions, fp32 atomics), db = column sums of du0,
//             dx = du0 W  -> bf16 [P0, Cin] -> transpose back to NCHW fp32 (only when the input needs a gradient)
#include "tma_gemm.cuh"

namespace gwn {

// x [N][Cin][V*T] fp32 -> xcl [N][L0][V][Cin] bf16, row (l, v) <- column (v, t = l - pad); l < pad rows are zero
__global__ void __launch_bounds__(256) nchw_to_cl_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ xcl, int Cin,
                                                              int V, int T, int L0) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, c0 = blockIdx.y * 32, q0 = blockIdx.x * 32;     // q = v*T + t
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;                   // 32 x 8
  const int VT = V * T, pad = L0 - T;
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, q = q0 + tx;
    tile[j][tx] = (c < Cin && q < VT) ? x[((long long)n * Cin + c) * VT + q] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int q = q0 + j, c = c0 + tx;
    if (q < VT && c < Cin) {
      const int v = q / T, t = q - v * T;
      xcl[(((long long)n * L0 + t + pad) * V + v) * Cin + c] = __float2bfloat16_rn(tile[tx][j]);
    }
  }
}
__global__ void zero_pad_rows_kernel(bf16* __restrict__ xcl, long long rows_per_n, long long pad_elems, int N) {
  // the first pad*V rows of every sample
  const long long total = (long long)N * pad_elems;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / pad_elems, r = i - n * pad_elems;
    xcl[n * rows_per_n + r] = __float2bfloat16_rn(0.f);
  }
}
// dxcl [N][L0][V][Cin] bf16 -> dx [N][Cin][V][T] fp32 (drops the padded time steps)
__global__ void __launch_bounds__(256) cl_bf16_to_nchw_kernel(const bf16* __restrict__ dxcl, float* __restrict__ dx, int Cin,
                                                              int V, int T, int L0) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, c0 = blockIdx.y * 32, q0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int VT = V * T, pad = L0 - T;
  for (int j = ty; j < 32; j += 8) {
    const int q = q0 + j, c = c0 + tx;
    float v = 0.f;
    if (q < VT && c < Cin) {
      const int node = q / T, t = q - node * T;
      v = __bfloat162float(dxcl[(((long long)n * L0 + t + pad) * V + node) * Cin + c]);
    }
    tile[j][tx] = v;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, q = q0 + tx;
    if (c < Cin && q < VT) dx[((long long)n * Cin + c) * VT + q] = tile[tx][j];
  }
}
// w [32][Cin] fp32 -> bf16 [32][Cin] (forward B operand, K-major) and its transpose [Cin][32] (dx B operand)
__global__ void start_wprep_kernel(const float* __restrict__ w, bf16* __restrict__ wk, bf16* __restrict__ wt, int Cin) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 32 * Cin; i += gridDim.x * blockDim.x) {
    const int c = i / Cin, ci = i - c * Cin;
    const bf16 v = __float2bfloat16_rn(w[i]);
    wk[i] = v;
    if (wt) wt[ci * 32 + c] = v;
  }
}
// db[c] = sum_p g[p][c], g bf16 [P][32]
__global__ void __launch_bounds__(256) colsum32_kernel(const bf16* __restrict__ g, float* __restrict__ db, long long P) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc = 0.f;
  for (long long p = (long long)blockIdx.x * 8 + warp; p < P; p += (long long)gridDim.x * 8) acc += __bfloat162float(g[p * 32 + lane]);
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][lane];
    atomicAdd(db + lane, s);
  }
}

struct EpiBiasBf16 {   // out[m][n] = acc + bias[n] (bf16), N = 32
  const float* bias; bf16* out; int ldo, N;
  __device__ __forceinline__ void chunk(int m, bool m_ok, int n0, float v[32]) const {
    if (!m_ok || n0 >= N) return;
    uint4* dst = reinterpret_cast<uint4*>(out + (long long)m * ldo + n0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 pk;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * j] + __ldg(bias + n0 + 8 * j), v[8 * j + 1] + __ldg(bias + n0 + 8 * j + 1));
      __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * j + 2] + __ldg(bias + n0 + 8 * j + 2), v[8 * j + 3] + __ldg(bias + n0 + 8 * j + 3));
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * j + 4] + __ldg(bias + n0 + 8 * j + 4), v[8 * j + 5] + __ldg(bias + n0 + 8 * j + 5));
      __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * j + 6] + __ldg(bias + n0 + 8 * j + 6), v[8 * j + 7] + __ldg(bias + n0 + 8 * j + 7));
      pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
      dst[j] = pk;
    }
  }
};
struct EpiStoreBf16 {  // out[m][n] = acc (bf16)
  bf16* out; int ldo, N;
  __device__ __forceinline__ void chunk(int m, bool m_ok, int n0, float v[32]) const {
    if (!m_ok || n0 >= N) return;
    uint4* dst = reinterpret_cast<uint4*>(out + (long long)m * ldo + n0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 pk;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]);
      __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
      __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
      pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
      dst[j] = pk;
    }
  }
};
struct EpiAtomicDw {   // dW[m][n] += acc
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

}  // namespace gwn

using namespace gwn;

extern "C" int gwn_start_tc_supported(int Cin) { return (Cin >= 64 && Cin % 8 == 0) ? 1 : 0; }

// x [N,Cin,V,T] fp32; w [32,Cin]; b [32]; xcl out [N,L0,V,Cin] bf16; u0 out [N,L0,V,32] bf16; ws_w >= 2*32*Cin*2 bytes
extern "C" int gwn_start_fwd_tc(const float* x, const float* w, const float* b, void* xcl, void* u0, void* ws_w, int N,
                                int Cin, int V, int T, int L0, void* stream) {
  GWN_REQUIRE(x && w && b && xcl && u0 && ws_w && gwn_start_tc_supported(Cin) && L0 >= T, "start_fwd_tc: bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long P0 = (long long)N * L0 * V;
  GWN_REQUIRE(P0 < (1ll << 31), "start_fwd_tc: too many positions");
  bf16* xc = reinterpret_cast<bf16*>(xcl);
  bf16* wk = reinterpret_cast<bf16*>(ws_w);
  start_wprep_kernel<<<16, 256, 0, st>>>(w, wk, wk + 32 * (size_t)Cin, Cin);
  GWN_LAUNCHED();
  if (L0 > T) {
    zero_pad_rows_kernel<<<148, 256, 0, st>>>(xc, (long long)L0 * V * Cin, (long long)(L0 - T) * V * Cin, N);
    GWN_LAUNCHED();
  }
  dim3 grid((unsigned)cdiv((long long)V * T, 32), (unsigned)cdiv(Cin, 32), (unsigned)N);
  nchw_to_cl_bf16_kernel<<<grid, 256, 0, st>>>(x, xc, Cin, V, T, L0);
  GWN_LAUNCHED();
  CUtensorMap ma, mb;
  if (int rc = tg_map_rows(&ma, xc, (uint64_t)P0, (uint64_t)Cin, (uint64_t)Cin, 128)) return rc;
  if (int rc = tg_map_rows(&mb, wk, 32, (uint64_t)Cin, (uint64_t)Cin, 32)) return rc;
  TgParams p{};
  p.M = (int)P0; p.N = 32; p.K = Cin; p.bn = 32; p.splits = 1;
  tg_operand(p.a, TG_K_SW128, 128);
  tg_operand(p.b, TG_K_SW128, 32);
  EpiBiasBf16 e{b, reinterpret_cast<bf16*>(u0), 32, 32};
  return launch_tma_gemm(ma, mb, p, e, st);
}

// du0 [P0,32] bf16 -> dw [32,Cin], db [32] (overwritten), dx [N,Cin,V,T] fp32 (optional); ws_dx >= P0*Cin*2 bytes when dx
extern "C" int gwn_start_bwd_tc(const void* xcl, const void* du0, void* ws_w, float* dw, float* db, float* dx, void* ws_dx,
                                int N, int Cin, int V, int T, int L0, void* stream) {
  GWN_REQUIRE(xcl && du0 && ws_w && dw && db && gwn_start_tc_supported(Cin) && (!dx || ws_dx), "start_bwd_tc: bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long P0 = (long long)N * L0 * V;
  const bf16* xc = reinterpret_cast<const bf16*>(xcl);
  const bf16* g = reinterpret_cast<const bf16*>(du0);
  GWN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 32 * (size_t)Cin, st));
  GWN_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * 32, st));
  colsum32_kernel<<<148 * 4, 256, 0, st>>>(g, db, P0);
  GWN_LAUNCHED();
  {  // dW[c][ci] = sum_p du0[p][c] xcl[p][ci]: both operands MN-major, split-K over positions
    CUtensorMap ma, mb;
    if (int rc = tg_map_2d(&ma, g, 32, (uint64_t)P0, 64, 64, 64)) return rc;
    if (int rc = tg_map_2d(&mb, xc, (uint64_t)Cin, (uint64_t)P0, (uint64_t)Cin * 2, 64, 64)) return rc;
    TgParams p{};
    const int bn = Cin >= 256 ? 256 : (Cin >= 128 ? 128 : 64);
    p.M = 32; p.N = Cin; p.K = (int)P0; p.bn = bn;
    const long long tiles = cdiv(Cin, bn), kb = cdiv(P0, TG_BK);
    long long splits = cdiv(tg_sm_count(), tiles);
    if (splits > kb / 16) splits = kb / 16;
    if (splits < 1) splits = 1;
    p.splits = (int)splits;
    tg_operand(p.a, TG_MN_SW128, 128);
    tg_operand(p.b, TG_MN_SW128, bn);
    EpiAtomicDw e{dw, Cin, 32, Cin};
    if (int rc = launch_tma_gemm(ma, mb, p, e, st)) return rc;
  }
  if (dx) {
    bf16* dxc = reinterpret_cast<bf16*>(ws_dx);
    const bf16* wt = reinterpret_cast<const bf16*>(ws_w) + 32 * (size_t)Cin;      // W^T [Cin][32] from the forward prep
    CUtensorMap ma, mb;
    const int bn = Cin >= 256 ? 256 : (Cin >= 128 ? 128 : 64);
    if (int rc = tg_map_rows(&ma, g, (uint64_t)P0, 32, 32, 128)) return rc;
    if (int rc = tg_map_rows(&mb, wt, (uint64_t)Cin, 32, 32, (uint32_t)bn)) return rc;
    TgParams p{};
    p.M = (int)P0; p.N = Cin; p.K = 32; p.bn = bn; p.splits = 1;
    tg_operand(p.a, TG_K_SW128, 128);
    tg_operand(p.b, TG_K_SW128, bn);
    EpiStoreBf16 e{dxc, Cin, Cin};
    if (int rc = launch_tma_gemm(ma, mb, p, e, st)) return rc;
    dim3 grid((unsigned)cdiv((long long)V * T, 32), (unsigned)cdiv(Cin, 32), (unsigned)N);
    cl_bf16_to_nchw_kernel<<<grid, 256, 0, st>>>(dxc, dx, Cin, V, T, L0);
    GWN_LAUNCHED();
  }
  return 0;
}
