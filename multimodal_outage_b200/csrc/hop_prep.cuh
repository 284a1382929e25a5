// Support-image builders of the on-chip hop kernels (V <= 128), as device functions over a virtual 1-D grid so that ONE launch
// builds all three images (gwn_hop_mats_prep): the four per-support UMMA images of hops_tc_kernel, the stacked image of the
// T-form fused forward (gcn_fused_t.cu) and the stacked transposed-hop image of the T-form fused backward
// (gcn_fused_bwd_t.cu).  They are the first kernels of every forward, each a few microseconds of dependent latency
// (an A^2 element is a V-long dot product): three launches in a row became one.
#pragma once
#include "gcn_fused.cuh"

namespace gwn {

// image[m][kc][r][e] = Mop[r][kc*8+e] in bf16; Mop = X or X^T, X = A or A*A (fp32 product), zero padded.
struct MatPrep {
  const float* A[GWN_MAX_SUPPORTS];
  int n, V, Kp;
};

__device__ __forceinline__ float prep_sq(const float* A, int V, int row, int col) {      // (A A)[row][col], fp32
  float acc = 0.f;
  for (int t = 0; t < V; ++t) acc = fmaf(A[(long long)row * V + t], A[(long long)t * V + col], acc);
  return acc;
}

// matrix index m = 4*s + variant; variant: 0 = A^T, 1 = (A^2)^T (forward), 2 = A, 3 = A^2 (backward)
__device__ __forceinline__ void hop_mats_prep_body(const MatPrep& mp, bf16* __restrict__ out, long long t0, long long stride) {
  const int per_mat = (mp.Kp / 8) * 128 * 8;
  const long long total = (long long)mp.n * 4 * per_mat;
  for (long long i = t0; i < total; i += stride) {
    int m = (int)(i / per_mat), rem = (int)(i % per_mat);
    int kc = rem / 1024, r = (rem / 8) % 128, e = rem % 8;
    int k = kc * 8 + e;
    int s = m / 4, variant = m % 4;
    float val = 0.f;
    if (r < mp.V && k < mp.V) {
      const float* A = mp.A[s];
      const bool transpose = variant < 2, square = variant & 1;
      int row = transpose ? k : r, col = transpose ? r : k;  // X[row][col]
      val = square ? prep_sq(A, mp.V, row, col) : A[(long long)row * mp.V + col];
    }
    out[i] = __float2bfloat16_rn(val);
  }
}

// Stacked support image of the transposed ("T-form") fused forward (gcn_fused_t.cu): the B operand of
//   h^T[(s,c), w] = sum_{j,v} U_j[(s,v), c] * Mt_j[v, w],   k = j*V + v,   Mt_0 = I, Mt_{2s+1} = A_s, Mt_{2s+2} = A_s A_s
// K-major no-swizzle canonical layout [KT/8][NP][8] bf16 (rows = output node w), zero padded.
__device__ __forceinline__ void hop_mats_t_prep_body(const MatPrep& mp, int KT, int NP, bf16* __restrict__ out, long long t0,
                                                     long long stride) {
  const int total = KT * NP;
  for (int i = (int)t0; i < total; i += (int)stride) {
    const int e = i & 7, w = (i >> 3) % NP, k = (i >> 3) / NP * 8 + e;
    const int j = k / mp.V, v = k - j * mp.V;
    float val = 0.f;
    if (w < mp.V && j <= 2 * mp.n) {
      if (j == 0) {
        val = (v == w) ? 1.f : 0.f;
      } else {
        const float* A = mp.A[(j - 1) >> 1];
        val = ((j - 1) & 1) == 0 ? A[(long long)v * mp.V + w] : prep_sq(A, mp.V, v, w);
      }
    }
    out[i] = __float2bfloat16_rn(val);
  }
}

// Stacked transposed-hop image: B operand of GEMM H of the T-form fused backward, K-major no-swizzle canonical layout
// [KW/8][NTOT][8] bf16:  element (k = w, n = column of item (r, h), slot j_l, node v_l) = Mt_j[v, w],  j = 4h + j_l, v = 32 r + v_l
__device__ __forceinline__ void hop_mats_bt_prep_body(const MatPrep& mp, const BtGeom& G, bf16* __restrict__ out, long long t0,
                                                      long long stride) {
  const int V = mp.V;
  const int total = (G.KW / 8) * G.NTOT * 8;
  for (int i = (int)t0; i < total; i += (int)stride) {
    const int e = i & 7, n = (i >> 3) % G.NTOT, w = (i >> 3) / G.NTOT * 8 + e;
    // decode n -> (r, h, j_l, v_l)
    int r = -1, h = 0, jl = 0, vl = 0, n0 = 0;
    for (int rr = 0; rr < G.NR && r < 0; ++rr)
      for (int hh = 0; hh < G.NHALF; ++hh) {
        const int cnt = G.nh(hh) * G.rs(rr);
        if (n < n0 + cnt) { r = rr; h = hh; jl = (n - n0) / G.rs(rr); vl = (n - n0) - jl * G.rs(rr); break; }
        n0 += cnt;
      }
    float val = 0.f;
    if (r >= 0) {
      const int j = 4 * h + jl, v = 32 * r + vl;
      if (v < V && w < V && j <= 2 * mp.n) {
        if (j == 0) val = (v == w) ? 1.f : 0.f;
        else {
          const float* A = mp.A[(j - 1) >> 1];
          val = ((j - 1) & 1) == 0 ? A[(long long)v * V + w] : prep_sq(A, V, v, w);
        }
      }
    }
    out[i] = __float2bfloat16_rn(val);
  }
}

}  // namespace gwn
