// Fused diffusion graph convolution for graphs whose supports fit on chip (V <= 80), bf16, forward:
//
//   u = dropout( mlp( concat[z, z A_0, z A_0^2, ..., z A_adp^2] ) + b ) + bn_prev(u_prev)[crop]     (+ BN statistics)
//   graph_wavenet.py:76-98 (gcn: nconv hops :87-93, concat :95, 1x1 mlp :96, dropout :97), :247 (residual), :250 (BN)
//
// in ONE kernel: the K-hop diffusion, the support concatenation and the mlp are one chained contraction on the
// tcgen05 tensor cores, with every support image and the mlp weights resident in shared memory, the
// intermediate in TMEM / shared memory, and nothing but z (read once) and u (written once) touching HBM.
//
// The 1x1 mlp commutes with node mixing (exact in fp64), so the contraction is evaluated in
// Horner form per slab (one (n,t) pair = V nodes x 32 channels):
//     U      = z W            [V,32] x [32, 32(1+H)]         GEMM 1 (M = node, K = channel), D in TMEM
//     h      = U_0 + b + sum_j M_j U_j                       GEMM 2 (M = node w, K = node v), H = 2*supports
// GEMM 1's accumulator is drained by "stage" warps that convert U_1..U_H to bf16 into the MN-major B-operand
// layout of GEMM 2 in shared memory and plant U_0 + b directly into GEMM 2's TMEM accumulator (tcgen05.st),
// so GEMM 2 only accumulates.  GEMM 2's accumulator is drained by the epilogue warps: Philox dropout, residual
// with the folded BatchNorm affine of the previous layer, bf16 store of u, per-channel (sum, sum^2).
#pragma once
#include "common.cuh"

namespace gwn {

constexpr int GF_MAX_MATS = 6;

struct GcnFwdParams {
  const bf16* z;               // [slabs*V, 32]
  const bf16* u_prev;          // [N, Lin, V, 32] (residual source)
  long long RI, RO, crop;      // rows per sample of u_prev / of the output, first residual row
  const float* scale;          // folded BN of the previous layer (NULL = identity)
  const float* shift;
  const bf16* mats;            // gwn_hop_mats_prep images
  int mat_src[GF_MAX_MATS];    // image index of hop j+1
  int n_mats;
  const bf16* w_img;           // [4][32*(1+n_mats)][8] bf16: (k = c, n = (j, c')) = W_mlp[j*32 + c][c']
  const float* bias;           // [32]
  const bf16* mask;            // optional explicit dropout mask [slabs*V, 32]
  float drop_p;
  uint64_t seed, offset;
  const uint64_t* rng;
  bf16* u;                     // out [slabs*V, 32]
  double* stats;               // [2][32], accumulated with atomics (caller zeroes)
  int V, Kp, slabs;
  const float* w_src;          // optional: packed fp32 mlp weight [32*(1+n_mats)][32]; the kernel builds its bf16 image itself (w_img unused)
  long long* trace;            // optional debug: [64 slabs][8] clock64 timestamps of CTA 0 (NULL = off)
  const bf16* mats_t;          // T-form kernel (gcn_fused_t.cu): stacked support image [KT/8][NP][8]; KT, NP filled by the launcher
  int KT, NP;
};

int gcn_fused_supported(int V, int n_mats);
// transposed contraction over groups of four slabs (gcn_fused_t.cu); needs w_src and mats_t
int gcn_fused_t_supported(int V, int n_mats);
int launch_gcn_fwd_t(GcnFwdParams& p, cudaStream_t st);
int launch_gcn_fwd(GcnFwdParams& p, cudaStream_t st);

}  // namespace gwn

// ------------------------------------------------------------------------------------------------------------------
// Fused diffusion graph convolution, BACKWARD (V <= 80, bf16), supports without gradient (gcn_fused_bwd.cu):
//   dh   = du . dropout-mask                                                   (graph_wavenet.py:97)
//   dU_j = M_j^T dh  (j = 1..H),  dU_0 = dh                                    (transposed hops, :60-66)
//   dz   = sum_j dU_j W_j^T  (+ dz_last on the tail rows)                      (mlp backward, :96)
//   dW_j = z^T dU_j,  db = sum dh        with z = a . b recomputed on chip     (weight gradients)
//   dfg  = gate backward of dz:  df = dz b (1 - a^2), dg = dz a b (1 - b)      (:222-226)
// One kernel per layer replaces zfill + 3 hop launches + drop_bwd + mlp wgrad + dcat GEMM + gate_bwd.
namespace gwn {
struct GcnBwdParams {
  const bf16* du;              // [slabs*V, 32]
  const bf16* a;               // [slabs*V, 32] tanh(f)
  const bf16* b;               // [slabs*V, 32] sigmoid(g)
  const bf16* dz_last;         // [N, Lf, V, 32] or NULL
  long long RO, last_begin, last_rows;
  const bf16* mats;            // gwn_hop_mats_prep images
  int mat_src[GF_MAX_MATS];    // image index of the TRANSPOSED hop j+1 (variants 2, 3)
  int n_mats;
  const bf16* wt_img;          // [4*(1+n_mats)][32][8]: (n = c, k = (j, c')) = W_mlp[j*32 + c][c']
  const float* w_src;          // optional: packed fp32 mlp weight; the kernel builds wt_img / w56_img itself (both unused)
  const bf16* mask;
  float drop_p;
  uint64_t seed, offset;
  const uint64_t* rng;
  bf16* dfg;                   // out [slabs*V, 64], interleaved (f, g)
  float* dw_mlp;               // [32*(1+n_mats), 32] fp32, accumulated with atomics (caller zeroes)
  float* db_mlp;               // [32]
  int V, Kp, slabs;
  // optional gradient wrt ONE support (the adaptive adjacency), sa = its index (-1: none):
  //   dA[v,w] += sum_{slab,c} T1[v,c] dh[w,c] + U6[v,c] dU5[w,c],   U5 = z W_{2sa+1}, U6 = z W_{2sa+2}, T1 = U5 + A^T-hop(U6)
  int sa;
  int mat_fwd;                 // image index of the FORWARD hop of support sa (variant 0)
  const bf16* w56_img;         // [4][64][8]: (n = (5|6, c'), k = c) = W_mlp[(2sa+1 | 2sa+2)*32 + c][c']
  float* dA;                   // [V, V] fp32, accumulated with atomics
  long long* trace;            // optional debug timeline of CTA 0 (GWN_GCN_TRACE)
  // transposed ("T-form") kernel (gcn_fused_bwd_t.cu) only:
  const bf16* mats_bt;         // stacked TRANSPOSED-hop image (gwn_hop_mats_prep, third region): B operand of GEMM H
  int debug;                   // experiment switches (GWN_BT_DEBUG), 0 in production
  float* dQ6;                  // [V, V] fp32, accumulated: Q6 = sum U6[v] . dh[w]; the caller owes dA += A^T Q6 + Q6 A^T
};
int gcn_bwd_fused_supported(int V, int n_mats);
int launch_gcn_bwd(GcnBwdParams& p, cudaStream_t st);

// ---- transposed fused backward: geometry shared by the kernel, its launcher and the image-prep kernel ----
// A group of four slabs is contracted at once.  GEMM H runs with M = (slab, channel) = 128 on the accumulator rows and
// N = (hop, node) on the columns, in ITEMS of (node range r of <= 32 nodes) x (hop half h of <= 4 hop slots, slot 0 =
// identity): item (r, h) owns columns n0(r, h) .. + nh(h) * rs(r) of the stacked image.
struct BtGeom {
  int NM, NH, NH0, NH1, NHALF, NPD, NR, RSL, KW, NTOT;
  __host__ __device__ constexpr BtGeom(int nm, int npd)
      : NM(nm), NH(1 + nm), NH0(1 + nm < 4 ? 1 + nm : 4), NH1(1 + nm - (1 + nm < 4 ? 1 + nm : 4)),
        NHALF(1 + nm > 4 ? 2 : 1), NPD(npd), NR((npd + 31) / 32), RSL(npd - 32 * ((npd + 31) / 32 - 1)),
        KW(((npd + 15) / 16) * 16), NTOT(0) {
    int n = 0, last = 0;
    for (int r = 0; r < NR; ++r)
      for (int h = 0; h < NHALF; ++h) { last = n + ((nh(h) * rs(r) + 15) / 16) * 16; n += nh(h) * rs(r); }
    NTOT = ((last + 7) / 8) * 8;
  }
  __host__ __device__ constexpr int rs(int r) const { return r < NR - 1 ? 32 : RSL; }       // nodes of range r
  __host__ __device__ constexpr int nh(int h) const { return h == 0 ? NH0 : NH1; }          // hop slots of half h
  __host__ __device__ constexpr int n0(int r, int h) const {                                // first column of item (r, h)
    int n = 0;
    for (int rr = 0; rr < NR; ++rr)
      for (int hh = 0; hh < NHALF; ++hh) {
        if (rr == r && hh == h) return n;
        n += nh(hh) * rs(rr);
      }
    return n;
  }
  __host__ __device__ constexpr int nmma(int r, int h) const { return ((nh(h) * rs(r) + 15) / 16) * 16; }
};
// bytes of the stacked transposed-hop image [KW/8][NTOT][8] bf16 for a graph of V nodes, n_mats = 2 * supports
int gcn_bwd_t_image_elems(int V, int n_mats);
int gcn_bwd_t_supported(int V, int n_mats, bool has_da);
int launch_gcn_bwd_t(GcnBwdParams& p, cudaStream_t st);
}  // namespace gwn
