// Parameter blocks of the tcgen05 "reduce over positions" kernels (tc_wgrad.cu):
//   wgrad_tc : dW[k, n] += sum_p A[p, k] * G[p, n]  (+ db[n] = sum_p G[p, n] through a resident ones-row)
//   dadj_tc  : dA[v, w] += sum_{s, c} X[s, v, c] * G[s, w, c]
#pragma once
#include "common.cuh"

namespace gwn {

constexpr int WG_MAX_CHUNKS = 8;

struct WgChunk {
  const bf16* base;       // channels-last source, 32 channels at [row*pitch, +32)
  long long rows_per_n;   // source rows per sample
  long long row_off;      // source row-in-sample = output row-in-sample + row_off (temporal tap)
  int pitch;
  int pad_;
};

struct WgParams {
  WgChunk ch[WG_MAX_CHUNKS];
  int n_chunks;             // K rows of dW = 32*n_chunks
  long long rows_per_n_out, P;
  const bf16* G; int g_pitch; int N;   // G [P, g_pitch] bf16, columns [0, N), N in {32, 64}
  float* dW; int ldw;       // fp32 [32*n_chunks, ldw], accumulated with atomics (caller zeroes)
  float* db;                // fp32 [N] or NULL
  const float* scale;       // optional per-input-channel affine of A: dW = scale*D + shift*db
  const float* shift;
  int n_tiles;
  long long* trace;         // optional debug timeline of CTA 0 (GWN_WG_TRACE)
  // filled by the launcher (TMA tiling: a K tile is 64 consecutive rows of one sample)
  int tiles_per_n, n_samples;
  int map_of[WG_MAX_CHUNKS];
  int row_off[WG_MAX_CHUNKS];
};

struct DadjTerm { const bf16* X; const bf16* G; };   // both [slabs, V, 32] contiguous slots
struct DadjParams {
  DadjTerm t[4]; int n_terms;
  int V, slabs, n_tiles;
  float* dA;                // fp32 [V, V], atomics
};

int wgrad_tc_supported(int n_chunks, int N);
int launch_wgrad_tc(WgParams& p, cudaStream_t st);
int launch_dadj_tc(DadjParams& p, cudaStream_t st);

}  // namespace gwn
