// Adaptive adjacency  adp = softmax(relu(E1 @ E2), dim=1)   (graph_wavenet.py:202), fp32 always.
// Warp-per-row: the rank-R product, relu, row max, exp, row sum and normalisation never leave
// registers/shuffles.  Backward:
//   dR = P * (dP - rowsum(dP*P)),  dM = dR * [M > 0],  dE1 = dM E2^T,  dE2 = E1^T dM.
#include "common.cuh"

namespace gwn {

constexpr int ADP_MAX_R = 16;

__global__ void __launch_bounds__(256) adp_fwd_kernel(const float* __restrict__ e1, const float* __restrict__ e2,
                                                      float* __restrict__ adp, float* __restrict__ adp_t,
                                                      float* __restrict__ adp2, int V, int R) {
  pdl_wait();      // programmatic launch: the launch latency overlaps the predecessor (common.cuh)
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= V) return;
  float a[ADP_MAX_R];
#pragma unroll
  for (int r = 0; r < ADP_MAX_R; ++r) a[r] = r < R ? __ldg(e1 + (long long)row * R + r) : 0.f;
  float mx = 0.f;  // relu output is >= 0, so 0 is a valid lower bound only if a 0 exists; track true max
  mx = -INFINITY;
  // (the three column loops are unrolled by four: a row of a 3,100-node graph is 97 trips of R dependent L1 loads each - with
  //  one trip in flight per warp and ~20 warps per SM the kernel was latency bound: 254 -> 113 us)
#pragma unroll 4
  for (int j = lane; j < V; j += 32) {
    float m = 0.f;
#pragma unroll
    for (int r = 0; r < ADP_MAX_R; ++r)
      if (r < R) m = fmaf(a[r], __ldg(e2 + (long long)r * V + j), m);
    m = fmaxf(m, 0.f);
    adp[(long long)row * V + j] = m;  // stash relu(M); overwritten below
    mx = fmaxf(mx, m);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll 4
  for (int j = lane; j < V; j += 32) {
    float e = expf(adp[(long long)row * V + j] - mx);
    adp[(long long)row * V + j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
#pragma unroll 4
  for (int j = lane; j < V; j += 32) {
    float pv = adp[(long long)row * V + j] * inv;
    adp[(long long)row * V + j] = pv;
    if (adp_t) adp_t[(long long)j * V + row] = pv;
    if (adp2) adp2[(long long)row * V + j] = pv;
  }
}

// Completes a support gradient delivered in factored form (include/gwn.h: d_supports_sq):
//   out[p, q] = d0[p, q] + sum_w Q[p, w] A[q, w] + sum_v A[v, p] Q[v, q]          (dA = Q5 + Q6 A^T + A^T Q6, fp32)
// One block per row p; A and Q ([V,V], a few tens of KB) stay in L1/L2.
__global__ void __launch_bounds__(128) dadj_finish_kernel(const float* __restrict__ A, const float* __restrict__ d0,
                                                          const float* __restrict__ Q, float* __restrict__ out, int V) {
  pdl_wait();      // programmatic launch: the launch latency overlaps the predecessor (common.cuh)
  pdl_trigger();
  extern __shared__ float sh[];
  float* qrow = sh;          // Q[p, :]
  float* acol = sh + V;      // A[:, p]
  const int p = blockIdx.x;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    qrow[i] = Q[(long long)p * V + i];
    acol[i] = A[(long long)i * V + p];
  }
  __syncthreads();
  for (int q = threadIdx.x; q < V; q += blockDim.x) {
    float acc = d0[(long long)p * V + q];
    const float* arow = A + (long long)q * V;
    float t1 = 0.f, t2 = 0.f;
    for (int w = 0; w < V; ++w) {
      t1 = fmaf(qrow[w], __ldg(arow + w), t1);
      t2 = fmaf(acol[w], __ldg(Q + (long long)w * V + q), t2);
    }
    out[(long long)p * V + q] = acc + t1 + t2;
  }
}

// Small graphs (V <= 96): the same completion with A and Q staged in shared memory by every CTA (coalesced, 2 x 36 KB at
// V = 96) - the per-element loop above walks A row-wise per thread (a different line per thread and iteration).
// CTA = 8 output rows p, thread = (row p, column q) pairs strided over the CTA.
__global__ void __launch_bounds__(256) dadj_finish_small_kernel(const float* __restrict__ A, const float* __restrict__ d0,
                                                                const float* __restrict__ Q, float* __restrict__ out, int V) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sh[];
  float* As = sh;            // [V][V]
  float* Qs = sh + V * V;    // [V][V]
  for (int i = threadIdx.x; i < V * V; i += 256) { As[i] = __ldg(A + i); Qs[i] = __ldg(Q + i); }
  __syncthreads();
  const int p0 = blockIdx.x * 8;
  for (int o = threadIdx.x; o < 8 * V; o += 256) {
    const int p = p0 + o / V, q = o % V;
    if (p >= V) break;
    float t1 = 0.f, t2 = 0.f;
#pragma unroll 4
    for (int w = 0; w < V; ++w) {
      t1 = fmaf(Qs[p * V + w], As[q * V + w], t1);      // broadcast x stride-V (conflict-free for odd V; 2-way at worst)
      t2 = fmaf(As[w * V + p], Qs[w * V + q], t2);      // broadcast x consecutive
    }
    out[(long long)p * V + q] = __ldg(d0 + (long long)p * V + q) + t1 + t2;
  }
}

// dE2[r, j] = sum_i E1[i, r] dM[i, j] for small graphs: CTA = rank index r, thread = column j, eight independent loads in
// flight per thread; plain stores (no memset node, no atomics)
__global__ void __launch_bounds__(128) adp_bwd_cols_small_kernel(const float* __restrict__ e1, const float* __restrict__ dm,
                                                                 float* __restrict__ de2, int V, int R) {
  pdl_wait();
  pdl_trigger();
  __shared__ float ecol[128];
  const int r = blockIdx.x, j = threadIdx.x;
  for (int i = threadIdx.x; i < V; i += 128) ecol[i] = __ldg(e1 + (long long)i * R + r);
  __syncthreads();
  if (j >= V) return;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  int i = 0;
  for (; i + 8 <= V; i += 8) {
    float g[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = __ldg(dm + (long long)(i + k) * V + j);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = fmaf(ecol[i + k], g[k], acc[k]);
  }
  for (; i < V; ++i) acc[0] = fmaf(ecol[i], __ldg(dm + (long long)i * V + j), acc[0]);
  de2[(long long)r * V + j] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
}

// one warp per row: dM row -> ws, dE1 row
__global__ void __launch_bounds__(256) adp_bwd_rows_kernel(const float* __restrict__ e1,
                                                           const float* __restrict__ e2,
                                                           const float* __restrict__ adp,
                                                           const float* __restrict__ dadp,
                                                           float* __restrict__ de1, float* __restrict__ dm, int V,
                                                           int R) {
  pdl_wait();      // programmatic launch: the launch latency overlaps the predecessor (common.cuh)
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= V) return;
  float a[ADP_MAX_R], acc[ADP_MAX_R];
#pragma unroll
  for (int r = 0; r < ADP_MAX_R; ++r) { a[r] = r < R ? __ldg(e1 + (long long)row * R + r) : 0.f; acc[r] = 0.f; }
  float dot = 0.f;
  for (int j = lane; j < V; j += 32) dot = fmaf(dadp[(long long)row * V + j], adp[(long long)row * V + j], dot);
  dot = warp_sum(dot);
  // (NOT unrolled like the forward: measured at V = 3100, two trips per iteration took this kernel 263 -> 345 us, four
  //  294 us - the R accumulators and the row's a[] already fill the registers)
  for (int j = lane; j < V; j += 32) {
    float m = 0.f;
#pragma unroll
    for (int r = 0; r < ADP_MAX_R; ++r)
      if (r < R) m = fmaf(a[r], __ldg(e2 + (long long)r * V + j), m);
    float pv = adp[(long long)row * V + j];
    float g = (m > 0.f) ? pv * (dadp[(long long)row * V + j] - dot) : 0.f;
    dm[(long long)row * V + j] = g;
#pragma unroll
    for (int r = 0; r < ADP_MAX_R; ++r)
      if (r < R) acc[r] = fmaf(g, __ldg(e2 + (long long)r * V + j), acc[r]);
  }
#pragma unroll
  for (int r = 0; r < ADP_MAX_R; ++r) {
    float s = warp_sum(acc[r]);
    if (lane == 0 && r < R) de1[(long long)row * R + r] = s;
  }
}

// dE2[r, j] = sum_i E1[i, r] dM[i, j]; grid (col tiles of 256, row splits), atomics into zeroed dE2
__global__ void __launch_bounds__(256) adp_bwd_cols_kernel(const float* __restrict__ e1,
                                                           const float* __restrict__ dm,
                                                           float* __restrict__ de2, int V, int R,
                                                           int rows_per_split) {
  pdl_wait();      // programmatic launch: the launch latency overlaps the predecessor (common.cuh)
  pdl_trigger();
  __shared__ float es[64][ADP_MAX_R];
  const int j = blockIdx.x * 256 + threadIdx.x;
  const int ib = blockIdx.y * rows_per_split, ie = min(V, ib + rows_per_split);
  float acc[ADP_MAX_R];
#pragma unroll
  for (int r = 0; r < ADP_MAX_R; ++r) acc[r] = 0.f;
  for (int i0 = ib; i0 < ie; i0 += 64) {
    __syncthreads();
    for (int t = threadIdx.x; t < 64 * ADP_MAX_R; t += 256) {
      int ii = t / ADP_MAX_R, r = t % ADP_MAX_R;
      es[ii][r] = (i0 + ii < ie && r < R) ? e1[(long long)(i0 + ii) * R + r] : 0.f;
    }
    __syncthreads();
    if (j < V) {
      const int lim = min(64, ie - i0);
      for (int ii = 0; ii < lim; ++ii) {
        float g = dm[(long long)(i0 + ii) * V + j];
#pragma unroll
        for (int r = 0; r < ADP_MAX_R; ++r) acc[r] = fmaf(es[ii][r], g, acc[r]);
      }
    }
  }
  if (j < V) {
#pragma unroll
    for (int r = 0; r < ADP_MAX_R; ++r)
      if (r < R) atomicAdd(de2 + (long long)r * V + j, acc[r]);
  }
}

}  // namespace gwn

using namespace gwn;

extern "C" int gwn_adp_fwd(const float* e1, const float* e2, float* adp, float* adp_t, int V, int R, void* stream) {
  GWN_REQUIRE(e1 && e2 && adp && V >= 1 && R >= 1 && R <= ADP_MAX_R, "adp_fwd: bad argument (R=%d, max %d)", R,
              ADP_MAX_R);
  GWN_CUDA(launch_pdl(adp_fwd_kernel, dim3((unsigned)cdiv(V, 8)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), e1, e2, adp, adp_t, nullptr, V, R));
  GWN_LAUNCHED();
  return 0;
}

// The adaptive adjacency as a PAIR [2][V][V] (two identical copies): the layer kernels read copy 0 as the support; the
// gradient of the pair comes back as (d0, Q) = (first-order part, factored second-order part - gwn.h: d_supports_sq)
// and gwn_adp_pair_bwd completes it (d_adp = d0 + A^T Q + Q A^T) before the softmax / relu / rank-R backward.
extern "C" int gwn_adp_fwd_pair(const float* e1, const float* e2, float* pair, int V, int R, void* stream) {
  GWN_REQUIRE(e1 && e2 && pair && V >= 1 && R >= 1 && R <= ADP_MAX_R, "adp_fwd_pair: bad argument (R=%d, max %d)", R,
              ADP_MAX_R);
  GWN_CUDA(launch_pdl(adp_fwd_kernel, dim3((unsigned)cdiv(V, 8)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), e1, e2, pair, nullptr,
                                                                                          pair + (long long)V * V, V, R));
  GWN_LAUNCHED();
  return 0;
}

extern "C" int gwn_adp_bwd(const float* e1, const float* e2, const float* adp, const float* d_adp, float* d_e1,
                           float* d_e2, float* ws, int V, int R, void* stream);

// ws: [2][V][V] scratch
extern "C" int gwn_adp_pair_bwd(const float* e1, const float* e2, const float* adp, const float* d_pair, float* d_e1,
                                float* d_e2, float* ws, int V, int R, void* stream) {
  GWN_REQUIRE(e1 && e2 && adp && d_pair && d_e1 && d_e2 && ws && V >= 1 && R >= 1 && R <= ADP_MAX_R,
              "adp_pair_bwd: bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (V <= 96) {
    static bool attr = false;
    if (!attr) {
      GWN_CUDA(cudaFuncSetAttribute(dadj_finish_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
      attr = true;
    }
    GWN_CUDA(launch_pdl(dadj_finish_small_kernel, dim3((unsigned)cdiv(V, 8)), dim3(256), 2 * (size_t)V * V * sizeof(float), st, adp, d_pair,
                        d_pair + (long long)V * V, ws, V));
  } else {
    GWN_CUDA(launch_pdl(dadj_finish_kernel, dim3(V), dim3(128), 2 * V * sizeof(float), st, adp, d_pair, d_pair + (long long)V * V, ws, V));
  }
  GWN_LAUNCHED();
  return gwn_adp_bwd(e1, e2, adp, ws, d_e1, d_e2, ws + (long long)V * V, V, R, stream);
}

extern "C" int gwn_adp_bwd(const float* e1, const float* e2, const float* adp, const float* d_adp, float* d_e1,
                           float* d_e2, float* ws, int V, int R, void* stream) {
  GWN_REQUIRE(e1 && e2 && adp && d_adp && d_e1 && d_e2 && ws && V >= 1 && R >= 1 && R <= ADP_MAX_R,
              "adp_bwd: bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GWN_CUDA(launch_pdl(adp_bwd_rows_kernel, dim3((unsigned)cdiv(V, 8)), dim3(256), 0, st, e1, e2, adp, d_adp, d_e1, ws, V, R));
  GWN_LAUNCHED();
  if (V <= 128) {
    GWN_CUDA(launch_pdl(adp_bwd_cols_small_kernel, dim3(R), dim3(128), 0, st, e1, ws, d_e2, V, R));
    GWN_LAUNCHED();
    return 0;
  }
  GWN_CUDA(cudaMemsetAsync(d_e2, 0, sizeof(float) * (size_t)R * V, st));
  int col_tiles = (int)cdiv(V, 256);
  int splits = (int)cdiv(148 * 2, col_tiles);
  int per = (int)cdiv(V, splits);
  per = (int)cdiv(per, 64) * 64;
  splits = (int)cdiv(V, per);
  dim3 grid(col_tiles, splits);
  GWN_CUDA(launch_pdl(adp_bwd_cols_kernel, dim3(grid), dim3(256), 0, st, e1, ws, d_e2, V, R, per));
  GWN_LAUNCHED();
  return 0;
}
