// CUDA-core tiled contractions over channels-last "position rows".
//
//   pos_gemm : C[p, n]  = epi( sum_k A[p, k] * W[k, n] )      (forward / data-grad form)
//   wgrad    : dW[k, n] = sum_p A[p, k] * G[p, n]  (+ column sums of G for the bias)
//
// A is never materialised as one matrix: it is gathered from up to GEMM_MAX_CHUNKS
// 32-channel "chunks", each a channels-last source with its own base pointer, row
// offset (temporal tap / crop), pitch, column offset, optional per-channel affine
// (the previous layer's folded BatchNorm) and optional relu.  This is how the
// dilated temporal taps (graph_wavenet.py:150-156), the concat of diffusion hops
// (graph_wavenet.py:95) and the per-layer skip slices (graph_wavenet.py:231-236)
// feed one contraction without a gather/concat pass over HBM.
//
// fp32 accumulate everywhere; storage type TA is float or bf16.
#pragma once
#include "common.cuh"

namespace gwn {

constexpr int GEMM_MAX_CHUNKS = 32;

struct AChunk {
  const void* base;
  long long rows_per_n;  // source rows per sample
  long long row_off;     // source row-in-sample = output row-in-sample + row_off (masked if outside)
  int pitch;             // elements per source row
  int col_off;           // first of the 32 channels
  const float* scale;    // optional per-channel affine (32), applied to valid rows only
  const float* shift;
  int relu;
  int pad_;
  long long w_off;       // element offset of this chunk's 32 K-rows inside W (see launch helpers)
};

struct GemmA {
  AChunk ch[GEMM_MAX_CHUNKS];
  int n_chunks;
  long long rows_per_n_out;
  long long P;  // total output rows
};

constexpr int PG_BM = 128;
constexpr int PG_BK = 32;

// Loads this thread's 16 channels of one source row of chunk `c` (after affine / relu).
template <typename TA>
__device__ __forceinline__ void load_chunk_row(const AChunk& c, long long n, long long rem, bool pvalid,
                                               int lk, float v[16]) {
  long long sr = rem + c.row_off;
  bool valid = pvalid && sr >= 0 && sr < c.rows_per_n;
  if (valid) {
    const TA* src = reinterpret_cast<const TA*>(c.base) + (n * c.rows_per_n + sr) * (long long)c.pitch +
                    c.col_off + lk;
#pragma unroll
    for (int i = 0; i < 4; ++i) load4(src + 4 * i, v + 4 * i);
    if (c.scale) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], __ldg(c.scale + lk + i), __ldg(c.shift + lk + i));
    }
    if (c.relu) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
  }
}

// Epilogue contract:
//   static constexpr bool kStats;        // per-column (sum1,sum2) reduction to double atomics
//   __device__ void apply(long long p, long long n, long long rem, int col, float v[4],
//                         float s1[4], float s2[4]) const;   // 4 consecutive columns of row p
//   double* stats;                       // [2, ldw] when kStats
// WT=false: W is [K, ldw] row-major, element (k,n) of chunk q at W[w_off[q] + k*ldw + n].
// WT=true : W is stored transposed, element (k,n) of chunk q at W[w_off[q] + n*ldk + k]
//           (used by the data-grad passes so weights are never re-packed).
template <typename TA, int BN, bool WT, typename Epi>
__global__ void __launch_bounds__(256) pos_gemm_kernel(GemmA A, const float* __restrict__ W, int ldw,
                                                       int ldk, Epi epi) {
  constexpr int TN = 4, TX = BN / TN, TY = 256 / TX, TM = PG_BM / TY;
  __shared__ __align__(16) float As[PG_BK][PG_BM + 4];
  __shared__ __align__(16) float Ws[PG_BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const long long p0 = (long long)blockIdx.x * PG_BM;
  const int n0 = blockIdx.y * BN;

  const int lm = tid % PG_BM;
  const int lk = (tid / PG_BM) * 16;
  const long long lp = p0 + lm;
  const bool lvalid = lp < A.P;
  const long long ln = lvalid ? lp / A.rows_per_n_out : 0;
  const long long lrem = lvalid ? lp % A.rows_per_n_out : 0;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int q = 0; q < A.n_chunks; ++q) {
    float v[16];
    load_chunk_row<TA>(A.ch[q], ln, lrem, lvalid, lk, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) As[lk + i][lm] = v[i];
    if constexpr (!WT) {
      for (int i = tid * 4; i < PG_BK * BN; i += 1024) {
        int r = i / BN, cc = i % BN;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + cc < ldw)
          w = *reinterpret_cast<const float4*>(W + A.ch[q].w_off + (long long)r * ldw + n0 + cc);
        *reinterpret_cast<float4*>(&Ws[r][cc]) = w;
      }
    } else {
      for (int i = tid; i < PG_BK * BN; i += 256) {
        int r = i % PG_BK, cc = i / PG_BK;
        Ws[r][cc] = (n0 + cc < ldw) ? __ldg(W + A.ch[q].w_off + (long long)(n0 + cc) * ldk + r) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < PG_BK; ++k) {
      float a[TM];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 t = *reinterpret_cast<const float4*>(&As[k][ty * TM + i]);
        a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
      }
      float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * TN]);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        acc[i][0] = fmaf(a[i], w.x, acc[i][0]);
        acc[i][1] = fmaf(a[i], w.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], w.z, acc[i][2]);
        acc[i][3] = fmaf(a[i], w.w, acc[i][3]);
      }
    }
    __syncthreads();
  }

  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  const int col = n0 + tx * TN;
  if (col < ldw) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      long long p = p0 + ty * TM + i;
      if (p < A.P) {
        long long n = p / A.rows_per_n_out, rem = p % A.rows_per_n_out;
        epi.apply(p, n, rem, col, acc[i], s1, s2);
      }
    }
  }
  if constexpr (Epi::kStats) {
    // block reduce the per-thread column partials over ty, one double atomic per column per block
    float* red = &As[0][0];  // reuse (>= 2*TY*BN floats): safe after the final __syncthreads above
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[(ty * BN) + tx * 4 + j] = s1[j];
      red[(TY * BN) + (ty * BN) + tx * 4 + j] = s2[j];
    }
    __syncthreads();
    if (tid < 2 * BN) {
      int which = tid / BN, c = tid % BN;
      double s = 0.0;
      for (int r = 0; r < TY; ++r) s += (double)red[which * TY * BN + r * BN + c];
      if (n0 + c < ldw) atomicAdd(epi.stats + (long long)which * ldw + n0 + c, s);
    }
  }
}

// ldw = number of output columns.  Non-transposed W: chunk q defaults to rows [32q, 32q+32).
template <typename TA, int BN, typename Epi>
int launch_pos_gemm(GemmA A, const float* W, int ldw, const Epi& epi, cudaStream_t st) {
  if (A.P <= 0) return 0;
  static_assert(2 * (256 / (BN / 4)) * BN <= PG_BK * (PG_BM + 4), "stats scratch must fit in As");
  for (int q = 0; q < A.n_chunks; ++q) A.ch[q].w_off = (long long)q * 32 * ldw;
  dim3 grid((unsigned)cdiv(A.P, PG_BM), (unsigned)cdiv(ldw, BN));
  pos_gemm_kernel<TA, BN, false, Epi><<<grid, 256, 0, st>>>(A, W, ldw, 0, epi);
  GWN_LAUNCHED();
  return 0;
}
// Transposed W (element (k,n) at W[w_off[q] + n*ldk + k]); caller fills A.ch[q].w_off.
template <typename TA, int BN, typename Epi>
int launch_pos_gemm_wt(const GemmA& A, const float* W, int ldw, int ldk, const Epi& epi, cudaStream_t st) {
  if (A.P <= 0) return 0;
  dim3 grid((unsigned)cdiv(A.P, PG_BM), (unsigned)cdiv(ldw, BN));
  pos_gemm_kernel<TA, BN, true, Epi><<<grid, 256, 0, st>>>(A, W, ldw, ldk, epi);
  GWN_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------
// wgrad: dW[q*32 + k, n] += sum_p A_q[p, k] * G[p, n];  db[n] += sum_p G[p, n]
// grid (n_chunks, ceil(ldw/BN), splits); atomics into pre-zeroed fp32 outputs.
template <typename TA, typename TG, int BN>
__global__ void __launch_bounds__(256) wgrad_kernel(GemmA A, const TG* __restrict__ G, int ldg, int g_off,
                                                    float* __restrict__ dW, int ldw, float* __restrict__ db,
                                                    long long rows_per_split) {
  constexpr int BP = 64;
  constexpr int TX = BN / 4, KR = 256 / TX, NK = 32 / KR;  // BN=64: TX=16, KR=16, NK=2; BN=32: 8,32,1
  __shared__ __align__(16) float As[BP][33];
  __shared__ __align__(16) float Gs[BP][BN];
  const int tid = threadIdx.x, tx = tid % TX, tk = tid / TX;
  const int q = blockIdx.x, n0 = blockIdx.y * BN;
  const AChunk c = A.ch[q];
  long long pb = (long long)blockIdx.z * rows_per_split;
  long long pe = pb + rows_per_split;
  if (pe > A.P) pe = A.P;

  float acc[NK][4];
#pragma unroll
  for (int i = 0; i < NK; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bs[4] = {0.f, 0.f, 0.f, 0.f};

  const int lrow = tid / 4;         // 64 rows
  const int lk = (tid % 4) * 8;     // 8 channels each
  for (long long pt = pb; pt < pe; pt += BP) {
    {
      long long p = pt + lrow;
      bool pv = p < pe;
      long long n = pv ? p / A.rows_per_n_out : 0, rem = pv ? p % A.rows_per_n_out : 0;
      long long sr = rem + c.row_off;
      bool valid = pv && sr >= 0 && sr < c.rows_per_n;
      float v[8];
      if (valid) {
        const TA* src = reinterpret_cast<const TA*>(c.base) + (n * c.rows_per_n + sr) * (long long)c.pitch +
                        c.col_off + lk;
        load4(src, v); load4(src + 4, v + 4);
        if (c.scale) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], __ldg(c.scale + lk + i), __ldg(c.shift + lk + i));
        }
        if (c.relu) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) As[lrow][lk + i] = v[i];
    }
    for (int i = tid * 4; i < BP * BN; i += 1024) {
      int r = i / BN, cc = i % BN;
      long long p = pt + r;
      float g[4] = {0.f, 0.f, 0.f, 0.f};
      if (p < pe && n0 + cc < ldw) load4(G + p * (long long)ldg + g_off + n0 + cc, g);
      *reinterpret_cast<float4*>(&Gs[r][cc]) = make_float4(g[0], g[1], g[2], g[3]);
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < BP; ++r) {
      float4 g = *reinterpret_cast<const float4*>(&Gs[r][tx * 4]);
#pragma unroll
      for (int i = 0; i < NK; ++i) {
        float a = As[r][tk + i * KR];
        acc[i][0] = fmaf(a, g.x, acc[i][0]);
        acc[i][1] = fmaf(a, g.y, acc[i][1]);
        acc[i][2] = fmaf(a, g.z, acc[i][2]);
        acc[i][3] = fmaf(a, g.w, acc[i][3]);
      }
      if (tk == 0) { bs[0] += g.x; bs[1] += g.y; bs[2] += g.z; bs[3] += g.w; }
    }
    __syncthreads();
  }
  const int col = n0 + tx * 4;
  if (col < ldw) {
#pragma unroll
    for (int i = 0; i < NK; ++i) {
      float* dst = dW + (long long)(q * 32 + tk + i * KR) * ldw + col;
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(dst + j, acc[i][j]);
    }
    if (db != nullptr && q == 0 && tk == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(db + col + j, bs[j]);
    }
  }
}

// dW [n_chunks*32, ldw] and db [ldw] are zeroed here, then accumulated.
template <typename TA, typename TG>
int launch_wgrad(const GemmA& A, const TG* G, int ldg, int g_off, float* dW, int ldw, float* db,
                 cudaStream_t st) {
  GWN_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)A.n_chunks * 32 * ldw, st));
  if (db) GWN_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)ldw, st));
  if (A.P <= 0) return 0;
  const int BN = (ldw % 64 == 0) ? 64 : 32;
  long long col_tiles = cdiv(ldw, BN);
  long long base_blocks = (long long)A.n_chunks * col_tiles;
  long long want = cdiv(148 * 4, base_blocks);               // ~4 waves of 148 SMs
  long long max_splits = cdiv(A.P, 256);                      // >= 256 rows per split
  long long splits = want < 1 ? 1 : (want > max_splits ? max_splits : want);
  if (splits < 1) splits = 1;
  long long rows = cdiv(cdiv(A.P, splits), 64) * 64;
  splits = cdiv(A.P, rows);
  dim3 grid((unsigned)A.n_chunks, (unsigned)col_tiles, (unsigned)splits);
  if (BN == 64)
    wgrad_kernel<TA, TG, 64><<<grid, 256, 0, st>>>(A, G, ldg, g_off, dW, ldw, db, rows);
  else
    wgrad_kernel<TA, TG, 32><<<grid, 256, 0, st>>>(A, G, ldg, g_off, dW, ldw, db, rows);
  GWN_LAUNCHED();
  return 0;
}

}  // namespace gwn
