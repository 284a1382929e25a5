// Parameter block of the tcgen05 diffusion-hop kernel (tc_hops.cu).
#pragma once
#include "common.cuh"

namespace gwn {

constexpr int TH_MAX_STEPS = 16, TH_MAX_OUTS = 8, TH_MAX_MATS = 8;
enum { TH_LOAD = 1, TH_RELEASE = 2, TH_FIRST = 4, TH_LAST = 8 };
constexpr int TH_PRODUCERS = 128;                  // warps 0-3
constexpr int TH_MMA_WARP = 4;                     // warp 4 (lane 0 issues)
constexpr int TH_EPI_WARPS = 8;                     // warps 5-12: two per TMEM lane quadrant (warp%4)
constexpr int TH_THREADS = 32 * (5 + TH_EPI_WARPS);

struct HopStep { int in_buf, in_slot, mat, acc, flags; };
struct HopOut { int buf, slot, add_buf, add_slot; };  // add_buf < 0: no add-in
struct HopParams {
  const bf16* in[2]; int in_pitch[2];
  bf16* out[2]; int out_pitch[2];
  long long slot_stride[2];    // elements between consecutive 32-channel slots of buffer i
  const bf16* mats;            // all images: [*][Kp/8][128][8]
  int mat_src[TH_MAX_MATS];    // resident slot -> image index inside `mats`
  int n_mats, V, Kp, slabs, n_tiles;
  int n_steps, n_outs;
  HopStep steps[TH_MAX_STEPS];
  HopOut outs[TH_MAX_OUTS];
};

int hops_tc_supported(int V, int n_mats);
int launch_hops_tc(HopParams& p, cudaStream_t st);

}  // namespace gwn
