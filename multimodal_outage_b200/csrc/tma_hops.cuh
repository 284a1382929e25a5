// Diffusion hop for supports that do not fit on chip (V > 128): TMA-tiled tcgen05 GEMM, see tma_gemm.cu.
#pragma once
#include "common.cuh"

namespace gwn {
// y[s,w,c] = sum_v img[w][v] * x[s,v,c] (+ add[s,w,c]);  img: bf16 [V][Vp] row-major; x, y, add: slot-major [slabs*V, 32]
int launch_hop_big(const bf16* img, int Vp, const bf16* X, bf16* Y, const bf16* add, long long slabs, int V,
                   cudaStream_t st);
// sparse form of the same hop: y[s,w,:] = sum_k val[w][k] * x[s, idx[w][k], :] (+ add; add may alias y), ELL rows of width W
int launch_hop_ell(const int* idx, const float* val, int W, const bf16* X, bf16* Y, const bf16* add, long long slabs, int V,
                   cudaStream_t st);
// dA[v,w] += sum_{s,c} X[s,v,c] * G[s,w,c]  (fp32 [V][V], accumulated in place)
int launch_dadj_big(const bf16* X, const bf16* G, float* dA, long long slabs, int V, cudaStream_t st);
}  // namespace gwn
