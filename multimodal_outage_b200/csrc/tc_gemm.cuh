// tcgen05 position GEMM:  C[p, n] = epi( sum_k A[p, k] * W[k, n] ),  M = 128 positions per tile.
//   A: gathered from 32-channel chunks of channels-last bf16 sources (taps / slots), cp.async'ed into the
//      K-major no-swizzle canonical layout; W: bf16 image resident in shared memory for the whole kernel
//      (weight-stationary, persistent CTAs); D: fp32 in TMEM, double buffered; epilogue functors fuse the
//      gate non-linearity / bias+dropout+residual+BN statistics / slot scatter / residual-grad add.
#pragma once
#include "common.cuh"

namespace gwn {

constexpr int PG_TC_MAX_CHUNKS = 16;

struct PgChunk {
  const bf16* base;
  long long rows_per_n;
  long long row_off;
  int pitch;
  int col_off;
};

constexpr int PG_TC_MAX_MAPS = 8;

// Weight source for building the bf16 UMMA weight image INSIDE the GEMM kernel's prologue (no separate weight-prep /
// BatchNorm-fold launches between a layer's kernels): fp32 weights, optional bias, optional fold of the previous
// layer's BatchNorm (scale into the weights, shift into the bias).  CTA 0 also publishes scale/shift/mean/rstd,
// updates the running statistics and zeroes the next statistics buffer.
struct PgWsrc {
  const float* W;               // NULL: the kernel copies PgParams::w_img instead
  int w_off[PG_TC_MAX_CHUNKS];  // element offset of chunk q
  int ld;                       // stride of the non-contiguous index
  int transposed;               // 0: (k,n) at w_off[q] + kk*ld + n ; 1: w_off[q] + n*ld + kk
  int half_odd;                 // odd output columns (and their bias) scaled by 0.5
  const float* bias;            // optional [N]
  int bn;                       // 1: fold BatchNorm(gamma, beta) of the A operand's channels; 2: fold the given scale_in / shift_in
  const float* scale_in; const float* shift_in;
  int bn_training;              // batch statistics from bn_stats (else running statistics)
  const double* bn_stats;       // [2][32] sum, sum of squares
  double bn_count;
  const float* gamma; const float* beta;
  float* running_mean; float* running_var;
  float eps, momentum;
  float* scale_out; float* shift_out; float* mean_out; float* rstd_out;   // [32] each (written by CTA 0)
  double* zero64;               // optional: 64 doubles zeroed by CTA 0
};

struct PgParams {
  PgChunk ch[PG_TC_MAX_CHUNKS];
  PgWsrc wsrc;
  int n_chunks;                 // K = 32 * n_chunks
  int n_extra;                  // extra (non-MMA) source tiles ch[n_chunks .. n_chunks+n_extra) brought in by TMA for the epilogue
  long long rows_per_n_out, P;
  int N;                        // output columns (multiple of 16, <= 256)
  const bf16* w_img;            // [K/8 (+2 with has_bias)][N][8]
  int has_bias;                 // the image carries a bias chunk (K = 16: row 0 = bf16 hi, row 1 = bf16 lo of the fp32 bias):
                                // one extra MMA per sub-tile against a resident "ones" tile plants it in the accumulator
  int n_tiles;
  // filled by the launcher: TMA tiling.  A tile is 128 consecutive output rows of ONE sample, so that a chunk
  // (temporal tap / concat slot) is one 3-D box {32 ch, 128 rows, 1 sample} whose out-of-range rows TMA zero-fills.
  int tiles_per_n, n_samples;
  int sub;                      // 128-row sub-tiles per pipeline step (macro tile = 128*sub rows): amortises hand-offs
  int rows_out;                 // output rows per (virtual) sample
  int n_out;                    // staged outputs (Epi::kTmaOut epilogues): bulk tensor stores from shared memory; 0 = direct stores
  int map_of[PG_TC_MAX_CHUNKS]; // chunk -> tensor map
  int row_off[PG_TC_MAX_CHUNKS];
  long long* trace;             // optional debug timeline of CTA 0 (GWN_PG_TRACE)
};

// builds w_img (+ folded bias) from fp32 weights; see tc_gemm.cu
struct WPrepParams {
  const float* W;
  long long w_off[PG_TC_MAX_CHUNKS];   // element offset of chunk q
  int ld;                       // stride of the non-contiguous index
  int transposed;               // 0: (k,n) at w_off[q] + kk*ld + n ; 1: w_off[q] + n*ld + kk
  int K, N;
  const float* scale;           // optional [32]: W'(k,n) = scale[k%32] * W(k,n)
  const float* shift;           // optional [32]: bias'(n) = bias(n) + sum_k shift[k%32] * W(k,n)
  const float* bias;            // optional [N]
  bf16* img;                    // out [K/8][N][8]
  float* bias_out;              // out [N] (may be NULL)
  int bias_chunk;               // != 0: append the (folded) bias as two K pieces [2][N][8] (k=0: bf16 hi, k=1: bf16 lo)
  int half_odd;                 // != 0: odd output columns (and their bias) are scaled by 0.5 (sigmoid(g) = 0.5 tanh(g/2) + 0.5)
};
int launch_wprep(const WPrepParams& w, cudaStream_t st);

}  // namespace gwn
