// Parameter re-layout between the reference's nn.Module shapes (graph_wavenet.py:150-183: Conv2d weights
// [out, in, 1, k], biases) and the kernels' packed layouts, and the reverse scatter of the gradients.  One launch each
// (the PyTorch stack / permute / contiguous glue this replaces was ~50 tiny launches per training step).
//
// packed buffer (fp32, back to back; offsets from gwn_pack_offsets):
//   w_fg  [nl][k*32][64]   (row = tap*32 + c_in, col = 2*c_out + {0: filter, 1: gate})
//   b_fg  [nl][64]
//   w_mlp [nl][mlp_in][32] (transpose of gconv.mlp.mlp.weight)
//   w_skip[nl*32][S]       (row = layer*32 + c)
//   b_skip[S]              (sum over layers: skip crop identity)
//   w_end1[S][E], w_end2[E][Opad] (zero padded), b_end2[Opad]
#include "common.cuh"
#include "../../include/gwn.h"

namespace gwn {

struct PackSeg { long long off[9]; };   // start of each of the 8 segments + total

__host__ __device__ inline PackSeg pack_segments(const gwn_pack_cfg& c) {
  PackSeg s;
  const long long nl = c.n_layers, k = c.taps, mi = c.mlp_in, S = c.S, E = c.E, Op = c.Opad;
  s.off[0] = 0;
  s.off[1] = s.off[0] + nl * k * 32 * 64;
  s.off[2] = s.off[1] + nl * 64;
  s.off[3] = s.off[2] + nl * mi * 32;
  s.off[4] = s.off[3] + nl * 32 * S;
  s.off[5] = s.off[4] + S;
  s.off[6] = s.off[5] + S * E;
  s.off[7] = s.off[6] + E * Op;
  s.off[8] = s.off[7] + Op;
  return s;
}

__global__ void __launch_bounds__(256) pack_params_kernel(const __grid_constant__ gwn_pack_cfg c,
                                                          const __grid_constant__ gwn_pack_ptrs p, float* __restrict__ out) {
  pdl_wait();      // programmatic launch: the launch latency overlaps the predecessor (common.cuh)
  pdl_trigger();
  const PackSeg s = pack_segments(c);
  const int k = c.taps, mi = c.mlp_in, S = c.S, E = c.E, Op = c.Opad, O = c.O;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < s.off[8]; i += (long long)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < s.off[1]) {                       // w_fg
      const int e = (int)i;
      const int col = e & 63, row = (e >> 6) % (k * 32), l = (e >> 6) / (k * 32);
      const int o = col >> 1, h = col & 1, j = row >> 5, ci = row & 31;
      const float* w = reinterpret_cast<const float*>(h ? p.w_gate[l] : p.w_filter[l]);
      v = w[(o * 32 + ci) * k + j];
    } else if (i < s.off[2]) {                // b_fg
      const int e = (int)(i - s.off[1]);
      const int col = e & 63, l = e >> 6;
      const float* b = reinterpret_cast<const float*>((col & 1) ? p.b_gate[l] : p.b_filter[l]);
      v = b[col >> 1];
    } else if (i < s.off[3]) {                // w_mlp
      const int e = (int)(i - s.off[2]);
      const int o = e & 31, ii = (e >> 5) % mi, l = (e >> 5) / mi;
      const float* w = reinterpret_cast<const float*>(p.w_mlp[l]);
      v = w ? w[o * mi + ii] : 0.f;
    } else if (i < s.off[4]) {                // w_skip
      const int e = (int)(i - s.off[3]);
      const int sc = e % S, row = e / S, ci = row & 31, l = row >> 5;
      v = reinterpret_cast<const float*>(p.w_skip[l])[sc * 32 + ci];
    } else if (i < s.off[5]) {                // b_skip
      const int sc = (int)(i - s.off[4]);
      for (int l = 0; l < c.n_layers; ++l) v += reinterpret_cast<const float*>(p.b_skip[l])[sc];
    } else if (i < s.off[6]) {                // w_end1
      const int e = (int)(i - s.off[5]);
      const int ec = e % E, sc = e / E;
      v = reinterpret_cast<const float*>(p.w_end1)[ec * S + sc];
    } else if (i < s.off[7]) {                // w_end2
      const int e = (int)(i - s.off[6]);
      const int o = e % Op, ec = e / Op;
      v = o < O ? reinterpret_cast<const float*>(p.w_end2)[o * E + ec] : 0.f;
    } else {                                  // b_end2
      const int o = (int)(i - s.off[7]);
      v = o < O ? reinterpret_cast<const float*>(p.b_end2)[o] : 0.f;
    }
    out[i] = v;
  }
}

// Gradient scatter.  g.* are the gradients of the packed tensors (NULL = no gradient: the matching parameter gradients
// are not written and the caller reports them as absent); out = one flat buffer in parameter order:
//   per layer l: dWf [32][32][k] | dbf [32] | dWg | dbg | dWm [32][mlp_in] | dWs [S][32] | dbs [S]   (layer-major)
//   then dW1 [E][S] | dW2 [O][E] | db2 [O]
__global__ void __launch_bounds__(256) unpack_grads_kernel(const __grid_constant__ gwn_pack_cfg c,
                                                           const __grid_constant__ gwn_unpack_ptrs g, float* __restrict__ out) {
  pdl_wait();      // programmatic launch: the launch latency overlaps the predecessor (common.cuh)
  pdl_trigger();
  const int k = c.taps, mi = c.mlp_in, S = c.S, E = c.E, Op = c.Opad, O = c.O;
  const long long per_layer = 2ll * (32 * 32 * k + 32) + 32ll * mi + 32ll * S + S;
  const long long total = per_layer * c.n_layers + (long long)E * S + (long long)O * E + O;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < per_layer * c.n_layers) {
      const int l = (int)(i / per_layer);
      int e = (int)(i - (long long)l * per_layer);
      const int nw = 32 * 32 * k;
      const float* wfg = reinterpret_cast<const float*>(g.w_fg[l]);
      const float* bfg = reinterpret_cast<const float*>(g.b_fg[l]);
      if (e < 2 * (nw + 32)) {                // filter / gate conv
        const int h = e >= nw + 32;
        if (h) e -= nw + 32;
        if (e < nw) {
          const int j = e % k, ci = (e / k) & 31, o = e / (k * 32);
          v = wfg ? wfg[(j * 32 + ci) * 64 + 2 * o + h] : 0.f;
        } else {
          v = bfg ? bfg[2 * (e - nw) + h] : 0.f;
        }
      } else {
        e -= 2 * (nw + 32);
        if (e < 32 * mi) {                    // gcn mlp weight
          const float* wm = reinterpret_cast<const float*>(g.w_mlp[l]);
          const int ii = e % mi, o = e / mi;
          v = wm ? wm[ii * 32 + o] : 0.f;
        } else {
          e -= 32 * mi;
          if (e < 32 * S) {                   // skip conv weight
            const float* ws = reinterpret_cast<const float*>(g.w_skip);
            const int ci = e & 31, sc = e >> 5;
            v = ws ? ws[(long long)(l * 32 + ci) * S + sc] : 0.f;
          } else {
            const float* bs = reinterpret_cast<const float*>(g.b_skip);
            v = bs ? bs[e - 32 * S] : 0.f;
          }
        }
      }
    } else {
      long long e = i - per_layer * c.n_layers;
      if (e < (long long)E * S) {
        const float* w1 = reinterpret_cast<const float*>(g.w_end1);
        const int sc = (int)(e % S), ec = (int)(e / S);
        v = w1 ? w1[(long long)sc * E + ec] : 0.f;
      } else if (e < (long long)E * S + (long long)O * E) {
        e -= (long long)E * S;
        const float* w2 = reinterpret_cast<const float*>(g.w_end2);
        const int ec = (int)(e % E), o = (int)(e / E);
        v = w2 ? w2[(long long)ec * Op + o] : 0.f;
      } else {
        const float* b2 = reinterpret_cast<const float*>(g.b_end2);
        v = b2 ? b2[e - (long long)E * S - (long long)O * E] : 0.f;
      }
    }
    out[i] = v;
  }
}

}  // namespace gwn

using namespace gwn;

static int check_pack_cfg(const gwn_pack_cfg* c) {
  GWN_REQUIRE(c && c->n_layers >= 1 && c->n_layers <= GWN_MAX_LAYERS && c->taps >= 1 && c->taps <= GWN_MAX_TAPS &&
                  c->mlp_in >= 32 && c->mlp_in % 32 == 0 && c->S >= 1 && c->E >= 1 && c->O >= 1 && c->Opad >= c->O,
              "pack: bad configuration");
  return 0;
}

extern "C" long long gwn_pack_offsets(const gwn_pack_cfg* cfg, long long* off9) {
  if (check_pack_cfg(cfg)) return -1;
  const PackSeg s = pack_segments(*cfg);
  if (off9) for (int i = 0; i < 9; ++i) off9[i] = s.off[i];
  return s.off[8];
}

extern "C" int gwn_pack_params(const gwn_pack_cfg* cfg, const gwn_pack_ptrs* ptrs, float* out, void* stream) {
  if (int rc = check_pack_cfg(cfg)) return rc;
  GWN_REQUIRE(ptrs && out, "pack_params: null argument");
  if (int rc = gwn_check_device()) return rc;
  const long long total = pack_segments(*cfg).off[8];
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  GWN_CUDA(launch_pdl(pack_params_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), *cfg, *ptrs, out));
  GWN_LAUNCHED();
  return 0;
}

extern "C" long long gwn_unpack_total(const gwn_pack_cfg* cfg) {
  if (check_pack_cfg(cfg)) return -1;
  const long long per_layer = 2ll * (32 * 32 * cfg->taps + 32) + 32ll * cfg->mlp_in + 32ll * cfg->S + cfg->S;
  return per_layer * cfg->n_layers + (long long)cfg->E * cfg->S + (long long)cfg->O * cfg->E + cfg->O;
}

extern "C" int gwn_unpack_grads(const gwn_pack_cfg* cfg, const gwn_unpack_ptrs* grads, float* out, void* stream) {
  if (int rc = check_pack_cfg(cfg)) return rc;
  GWN_REQUIRE(grads && out, "unpack_grads: null argument");
  if (int rc = gwn_check_device()) return rc;
  const long long total = gwn_unpack_total(cfg);
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  GWN_CUDA(launch_pdl(unpack_grads_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), *cfg, *grads, out));
  GWN_LAUNCHED();
  return 0;
}
