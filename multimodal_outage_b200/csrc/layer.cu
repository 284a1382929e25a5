// One Graph WaveNet layer (graph_wavenet.py:206-250), forward and backward, channels-last.
//   gate  : fused BN-affine-on-load + dilated (1,k) filter & gate convs + tanh*sigmoid  (:222-226)
//   hops  : nconv  y[s,w,c] = sum_v x[s,v,c] A[v,w]                                     (:60-66,:87-93)
//   mlp   : 1x1 conv over the never-materialised-in-NCHW concat + bias + dropout        (:95-97)
//           + residual add with the cropped, BN-folded input (:247) + BN statistics     (:250)
#include <type_traits>

#include "gemm.cuh"
#include "tc_hops.cuh"
#include "tma_hops.cuh"
#include "gcn_fused.cuh"
#include <cstdlib>
#include "tc_wgrad.cuh"
#include "tc_gemm_impl.cuh"

namespace gwn {

// ------------------------------------------------------------------------------------------ epilogues
template <typename T>
struct EpiGate {
  static constexpr bool kStats = false;
  double* stats;
  const float* bias;  // [64] interleaved (f0,g0,f1,g1,...)
  T* zcat; int zpitch;
  T* a; T* b;
  T* z_last; long long last_begin, last_rows;
  __device__ __forceinline__ void apply(long long p, long long n, long long rem, int col, float v[4],
                                        float*, float*) const {
    const int o = col >> 1;
    float f0 = v[0] + __ldg(bias + col), g0 = v[1] + __ldg(bias + col + 1);
    float f1 = v[2] + __ldg(bias + col + 2), g1 = v[3] + __ldg(bias + col + 3);
    float a0 = tanhf(f0), a1 = tanhf(f1);
    float b0 = 1.f / (1.f + expf(-g0)), b1 = 1.f / (1.f + expf(-g1));
    float z0 = a0 * b0, z1 = a1 * b1;
    store2(zcat + p * zpitch + o, z0, z1);
    if (a) { store2(a + p * 32 + o, a0, a1); store2(b + p * 32 + o, b0, b1); }
    if (z_last && rem >= last_begin) store2(z_last + (n * last_rows + rem - last_begin) * 32 + o, z0, z1);
  }
};

template <typename T>
struct EpiMlp {
  static constexpr bool kStats = true;
  double* stats;
  const float* bias;
  const T* u_prev; long long prev_rows_per_n, crop; const float* scale; const float* shift;
  const T* mask; float drop_p; uint64_t seed, offset; const uint64_t* rng;
  T* u;
  __device__ __forceinline__ void apply(long long p, long long n, long long rem, int col, float v[4],
                                        float s1[4], float s2[4]) const {
    float h[4], r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = v[j] + __ldg(bias + col + j);
    if (mask) {
      float m[4]; load4(mask + p * 32 + col, m);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] *= m[j];
    } else if (drop_p > 0.f) {
      const uint64_t sd = rng ? __ldg(rng) : seed;
      const uint64_t of = rng ? offset + __ldg(rng + 1) : offset;
      float m[4]; dropout4(sd, of, (uint64_t)p, col, drop_p, m);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] *= m[j];
    }
    load4(u_prev + (n * prev_rows_per_n + rem + crop) * 32 + col, r);
    if (scale) {
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = fmaf(r[j], __ldg(scale + col + j), __ldg(shift + col + j));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { h[j] += r[j]; s1[j] += h[j]; s2[j] += h[j] * h[j]; }
    store4(u + p * 32 + col, h);
  }
};

template <typename T>
struct EpiSlotStore {  // column group col/32 goes to slot col/32 of a slot-major buffer
  static constexpr bool kStats = false;
  double* stats;
  T* out; long long slot_stride;
  __device__ __forceinline__ void apply(long long p, long long, long long, int col, float v[4], float*,
                                        float*) const {
    store4(out + (col >> 5) * slot_stride + p * 32 + (col & 31), v);
  }
};

template <typename T>
struct EpiGateBwdData {  // dx = acc + du(cropped rows); stats = (sum dx, sum dx*u_prev)
  static constexpr bool kStats = true;
  double* stats;
  const T* du; long long du_rows_per_n, crop;
  const T* u_prev;
  float* dx;
  __device__ __forceinline__ void apply(long long p, long long n, long long rem, int col, float v[4],
                                        float s1[4], float s2[4]) const {
    if (du && rem >= crop) {
      float g[4]; load4(du + (n * du_rows_per_n + rem - crop) * 32 + col, g);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += g[j];
    }
    float up[4]; load4(u_prev + p * 32 + col, up);
#pragma unroll
    for (int j = 0; j < 4; ++j) { s1[j] += v[j]; s2[j] += v[j] * up[j]; }
    store4(dx + p * 32 + col, v);
  }
};

// ------------------------------------------------------------------------------------------ tcgen05 epilogues
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void store_bf16x16(bf16* dst, const float v[16]) {
  uint4 a, b;
  __nv_bfloat162 h;
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&h); }
  a = make_uint4(w[0], w[1], w[2], w[3]); b = make_uint4(w[4], w[5], w[6], w[7]);
  *reinterpret_cast<uint4*>(dst) = a;
  *reinterpret_cast<uint4*>(dst + 8) = b;
}

struct EpiGateTC {   // N = 64 interleaved (f,g); a chunk of 32 columns = 16 channels
  // the accumulator already holds f + bias and (g + bias) / 2: the bias rides in the GEMM (has_bias) and the weight
  // image's g columns are pre-halved (half_odd), so sigmoid(g) = 0.5 tanh(acc) + 0.5
  static constexpr bool kExtra = false;
  static constexpr bool kWgrad = false;
  static constexpr int kFast = 2;
  bf16* z; bf16* a; bf16* b; bf16* z_last; long long last_begin, last_rows;
  __device__ __forceinline__ void chunk(long long p, long long n, long long rem, bool valid, int c0, float v[32]) {
    if (!valid) return;
    float zz[16], aa[16], bb[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      aa[i] = tanh_fast(v[2 * i]);
      bb[i] = fmaf(0.5f, tanh_fast(v[2 * i + 1]), 0.5f);
      zz[i] = aa[i] * bb[i];
    }
    const int ch0 = c0 >> 1;
    store_bf16x16(z + p * 32 + ch0, zz);
    if (a) { store_bf16x16(a + p * 32 + ch0, aa); store_bf16x16(b + p * 32 + ch0, bb); }
    if (z_last && rem >= last_begin) store_bf16x16(z_last + (n * last_rows + rem - last_begin) * 32 + ch0, zz);
  }
  // staged form (tc_gemm_impl.cuh, kTmaOut): the thread's 16 channels of z (a, b) go to 16-byte pieces 2*ci, 2*ci + 1 of its
  // row in the 64B-swizzled [128][64 B] tiles (piece ^ ((row >> 1) & 3): a quarter warp's stores hit 8 distinct bank groups)
  static constexpr int kTmaOut = 3;
  int staged;                                                   // host switch (GWN_GATE_TMA_STORE=0: direct stores)
  int n_out() const { return (staged && a) ? 3 : 0; }           // (one output: the direct stores are faster, 18.9 vs 22.2 us)
  bf16* out_ptr(int i) const { return i == 0 ? z : i == 1 ? a : b; }
  __device__ __forceinline__ void chunk_st(long long p, long long n, long long rem, bool valid, int c0, float v[32], uint8_t* slot_s,
                                           int trow, uint64_t* sempty, uint32_t parity) {
    uint32_t zw[8], aw[8], bw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a0 = tanh_fast(v[4 * i]), a1 = tanh_fast(v[4 * i + 2]);
      const float b0 = fmaf(0.5f, tanh_fast(v[4 * i + 1]), 0.5f), b1 = fmaf(0.5f, tanh_fast(v[4 * i + 3]), 0.5f);
      __nv_bfloat162 h = __floats2bfloat162_rn(a0 * b0, a1 * b1);
      zw[i] = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2bfloat162_rn(a0, a1); aw[i] = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2bfloat162_rn(b0, b1); bw[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    tc::mbar_wait_lazy(sempty, parity);                              // the slot's previous tile has been read by its bulk stores
    const uint32_t sw = ((uint32_t)trow >> 1) & 3u, cc = (uint32_t)(c0 >> 5) * 2u;
    uint8_t* row = slot_s + (size_t)trow * 64;
    uint8_t* p0 = row + ((cc ^ sw) << 4);
    uint8_t* p1 = row + (((cc + 1u) ^ sw) << 4);
    *reinterpret_cast<uint4*>(p0) = make_uint4(zw[0], zw[1], zw[2], zw[3]);
    *reinterpret_cast<uint4*>(p1) = make_uint4(zw[4], zw[5], zw[6], zw[7]);
    if (a) {
      *reinterpret_cast<uint4*>(p0 + 8192) = make_uint4(aw[0], aw[1], aw[2], aw[3]);
      *reinterpret_cast<uint4*>(p1 + 8192) = make_uint4(aw[4], aw[5], aw[6], aw[7]);
      *reinterpret_cast<uint4*>(p0 + 16384) = make_uint4(bw[0], bw[1], bw[2], bw[3]);
      *reinterpret_cast<uint4*>(p1 + 16384) = make_uint4(bw[4], bw[5], bw[6], bw[7]);
    }
    if (z_last && valid && rem >= last_begin) {
      uint4* zl = reinterpret_cast<uint4*>(z_last + (n * last_rows + rem - last_begin) * 32 + (c0 >> 1));
      zl[0] = make_uint4(zw[0], zw[1], zw[2], zw[3]);
      zl[1] = make_uint4(zw[4], zw[5], zw[6], zw[7]);
    }
  }
  __device__ __forceinline__ void finish(float*) {}
  __device__ __forceinline__ void flush(const float*, int) {}
};

struct EpiMlpTC {    // N = 32: (bias in the GEMM) dropout + residual(BN-folded input) -> u, per-channel (sum, sum^2)
  static constexpr bool kExtra = false;
  static constexpr bool kWgrad = false;
  static constexpr int kFast = 0;
  const bf16* u_prev; long long prev_rows_per_n, crop; const float* scale; const float* shift;
  const bf16* mask; float drop_p; uint64_t seed, offset; const uint64_t* rng;
  bf16* u; double* stats;
  float s1, s2;
  __device__ __forceinline__ void chunk(long long p, long long n, long long rem, bool valid, int c0, float v[32]) {
    const int lane = threadIdx.x & 31;
    float h[32];
    if (valid) {
      const bf16* rp = u_prev + (n * prev_rows_per_n + rem + crop) * 32;
      uint64_t sd = 0, of = 0;
      if (!mask && drop_p > 0.f) { sd = rng ? __ldg(rng) : seed; of = rng ? offset + __ldg(rng + 1) : offset; }
      float m32[32];
      if (!mask && drop_p > 0.f) {
        dropout16(sd, of, (uint64_t)(p * 2), drop_p, m32);
        dropout16(sd, of, (uint64_t)(p * 2 + 1), drop_p, m32 + 16);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float m[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f}, r[8];
        if (mask) { load4(mask + p * 32 + 8 * j, m); load4(mask + p * 32 + 8 * j + 4, m + 4); }
        else if (drop_p > 0.f) {
#pragma unroll
          for (int i = 0; i < 8; ++i) m[i] = m32[8 * j + i];
        }
        load4(rp + 8 * j, r); load4(rp + 8 * j + 4, r + 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = 8 * j + i;
          float rr = scale ? fmaf(r[i], __ldg(scale + c), __ldg(shift + c)) : r[i];
          h[c] = fmaf(v[c], m[i], rr);
        }
      }
      store_bf16x16(u + p * 32, h);
      store_bf16x16(u + p * 32 + 16, h + 16);
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c) h[c] = 0.f;
    }
    float t[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) t[c] = h[c];
    s1 += warp_column_sums(t, lane);
#pragma unroll
    for (int c = 0; c < 32; ++c) t[c] = h[c] * h[c];
    s2 += warp_column_sums(t, lane);
  }
  __device__ __forceinline__ void finish(float* red_s) {
    const int lane = threadIdx.x & 31;
    atomicAdd(red_s + lane, s1);
    atomicAdd(red_s + 32 + lane, s2);
  }
  __device__ __forceinline__ void flush(const float* red_s, int lane) {
    atomicAdd(stats + lane, (double)red_s[lane]);
    atomicAdd(stats + 32 + lane, (double)red_s[32 + lane]);
  }
};

struct EpiSlotTC {   // 32-column chunk c0 -> slot c0/32 of a slot-major buffer
  static constexpr bool kExtra = false;
  static constexpr bool kWgrad = false;
  static constexpr int kFast = 1;
  bf16* out; long long slot_stride;
  __device__ __forceinline__ void chunk(long long p, long long, long long, bool valid, int c0, float v[32]) {
    if (!valid) return;
    bf16* dst = out + (long long)(c0 >> 5) * slot_stride + p * 32;
    store_bf16x16(dst, v);
    store_bf16x16(dst + 16, v + 16);
  }
  __device__ __forceinline__ void finish(float*) {}
  __device__ __forceinline__ void flush(const float*, int) {}
};

template <bool WG>
struct EpiGateBwdTC {   // N = 32: dx = acc + du(cropped rows); (sum dx, sum dx*u_prev); WG: fused gate weight gradient
  // du and u_prev rows of the tile arrive by TMA as two extra 64B-swizzled [128 rows][64 B] tiles (the du tile is
  // zero-filled outside the cropped range), so the epilogue never waits on a global load.  The 32 columns are
  // processed as two halves of 16 (keeps the live registers under the 96 the 640-thread CTA allows).
  static constexpr bool kExtra = true;
  static constexpr bool kWgrad = WG;
  static constexpr int kFast = 4;
  float* dx; bf16* dx16;          // exactly one of the two: fp32 or bf16 storage of dx
  double* stats;
  float s1[2], s2[2];
  // fused weight gradient (tc_gemm_impl.cuh): accumulator row c of chunk q = (tap j, half h) -> dw_fg[(j*32 + c), 32h + n],
  // with the BatchNorm fold of the layer input (dW' = scale[c] dW + shift[c] db[n]); ones row -> db_fg
  const float* wg_scale; const float* wg_shift; float* dw_fg; float* db_fg;
  // row (chunk q = (tap j, half h), column n) of the transposed accumulator: v[c] = sum u_prev[.][c] dfg_q[.][n], db = sum dfg_q[.][n]
  // over the same rows (so the fold's shift term pairs with exactly the partial sum this CTA holds)
  __device__ __forceinline__ void wgrad_row_t(int q, int n, const float v[32], float db, float* db_s, float* stg) const {
    float* dst = stg + (size_t)((q >> 1) * 32) * 64 + 32 * (q & 1) + n;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const float sc = wg_scale ? __ldg(wg_scale + c) : 1.f, sh = wg_scale ? __ldg(wg_shift + c) : 0.f;
      dst[(size_t)c * 64] = fmaf(sc, v[c], sh * db);
    }
    if (q < 2) db_s[32 * q + n] = db;                           // tap 0 sees every output position once: the bias gradient
  }
  __device__ __forceinline__ void wgrad_flush(const float* stg, const float* db_s, int n_chunks, int t, int nthr) const {
    red_flush_2d(dw_fg, 64, stg, 64, (n_chunks >> 1) * 32, 64, t, nthr);
    red_flush_1d(db_fg, db_s, 64, t, nthr);
  }
  __device__ __forceinline__ void chunk_ex(long long p, long long, long long, bool valid, int, float v[32],
                                           const uint8_t* extra, int r) {
    const int lane = threadIdx.x & 31;
    const int sw = (r >> 1) & 3;                                // 64B swizzle: logical 16-byte chunk c sits at c ^ sw
    const uint8_t* drow = extra + (size_t)r * 64;
    const uint8_t* urow = extra + 8192 + (size_t)r * 64;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float up[16];
      float* vh = v + 16 * h;
      if (valid) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = 2 * h + cc;
          const uint4 qd = *reinterpret_cast<const uint4*>(drow + ((c ^ sw) << 4));
          const uint4 qu = *reinterpret_cast<const uint4*>(urow + ((c ^ sw) << 4));
          const uint32_t wd[4] = {qd.x, qd.y, qd.z, qd.w}, wu[4] = {qu.x, qu.y, qu.z, qu.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            vh[8 * cc + 2 * i] += __uint_as_float(wd[i] << 16);
            vh[8 * cc + 2 * i + 1] += __uint_as_float(wd[i] & 0xFFFF0000u);
            up[8 * cc + 2 * i] = __uint_as_float(wu[i] << 16);
            up[8 * cc + 2 * i + 1] = __uint_as_float(wu[i] & 0xFFFF0000u);
          }
        }
        if (dx16) store_bf16x16(dx16 + p * 32 + 16 * h, vh);
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j) store4(dx + p * 32 + 16 * h + 4 * j, vh + 4 * j);
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) up[c] *= vh[c];
      } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) { vh[c] = 0.f; up[c] = 0.f; }
      }
      s1[h] += warp_column_sums16(vh, lane);      // in place: vh and up are dead afterwards
      s2[h] += warp_column_sums16(up, lane);
    }
  }
  __device__ __forceinline__ void finish(float* red_s) {
    const int lane = threadIdx.x & 31;
    if ((lane & 1) == 0) {                         // lanes 2c, 2c+1 both hold column c of each half
      const int c = lane >> 1;
      atomicAdd(red_s + c, s1[0]);
      atomicAdd(red_s + 16 + c, s1[1]);
      atomicAdd(red_s + 32 + c, s2[0]);
      atomicAdd(red_s + 48 + c, s2[1]);
    }
  }
  __device__ __forceinline__ void flush(const float* red_s, int lane) {
    atomicAdd(stats + lane, (double)red_s[lane]);
    atomicAdd(stats + 32 + lane, (double)red_s[32 + lane]);
  }
};

// ------------------------------------------------------------------------------------------ nconv
// Y[s, w, yoff+c] (+)= sum_v Aop(v,w) * X[s, v, xoff+c];  Aop(v,w) = TR ? A[w*V+v] : A[v*V+w].
// Tile: (16*TM) w-rows x 64 cols (2 slabs x 32 ch), K-step 16 nodes; 256 threads, TMx4 per thread.
template <typename T, int TM, bool TR>
__global__ void __launch_bounds__(256) node_mix_kernel(const T* __restrict__ X, int xp, int xoff,
                                                       T* __restrict__ Y, int yp, int yoff, int accumulate,
                                                       const float* __restrict__ A, int slabs, int V) {
  constexpr int BMW = 16 * TM, BK = 16;
  __shared__ __align__(16) float As[BK][BMW + 4];
  __shared__ __align__(16) float Xs[BK][64];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int w0 = blockIdx.y * BMW;
  const long long s0 = (long long)blockIdx.x * 2;
  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int xk = tid / 16, xc = (tid % 16) * 4;      // X tile loader: row kk, 4 cols
  const long long xs = s0 + xc / 32;
  const int xcc = xc % 32;
  for (int v0 = 0; v0 < V; v0 += BK) {
    for (int i = tid; i < BK * BMW; i += 256) {
      int kk, m;
      if (TR) { kk = i % BK; m = i / BK; } else { m = i % BMW; kk = i / BMW; }
      int v = v0 + kk, w = w0 + m;
      float val = 0.f;
      if (v < V && w < V) val = TR ? __ldg(A + (long long)w * V + v) : __ldg(A + (long long)v * V + w);
      As[kk][m] = val;
    }
    {
      float xv[4] = {0.f, 0.f, 0.f, 0.f};
      int v = v0 + xk;
      if (v < V && xs < slabs) load4(X + (xs * V + v) * (long long)xp + xoff + xcc, xv);
      *reinterpret_cast<float4*>(&Xs[xk][xc]) = make_float4(xv[0], xv[1], xv[2], xv[3]);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
      float4 x = *reinterpret_cast<const float4*>(&Xs[kk][tx * 4]);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        acc[i][0] = fmaf(a[i], x.x, acc[i][0]);
        acc[i][1] = fmaf(a[i], x.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], x.z, acc[i][2]);
        acc[i][3] = fmaf(a[i], x.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
  const long long s = s0 + (tx * 4) / 32;
  const int c = (tx * 4) % 32;
  if (s < slabs) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      int w = w0 + ty * TM + i;
      if (w < V) {
        T* dst = Y + (s * V + w) * (long long)yp + yoff + c;
        if (accumulate) {
          float o[4]; load4(dst, o);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] += o[j];
        }
        store4(dst, acc[i]);
      }
    }
  }
}

template <typename T>
int launch_node_mix(const T* X, int xp, int xoff, T* Y, int yp, int yoff, int accumulate, const float* A,
                    bool transpose_a, long long slabs, int V, cudaStream_t st) {
  if (slabs <= 0) return 0;
  const bool small = V <= 80;
  const int bmw = small ? 80 : 64;
  dim3 grid((unsigned)cdiv(slabs, 2), (unsigned)cdiv(V, bmw));
  GWN_REQUIRE(grid.y <= 65535u, "node_mix: V too large");
#define GWN_NM(TM, TR) \
  node_mix_kernel<T, TM, TR><<<grid, 256, 0, st>>>(X, xp, xoff, Y, yp, yoff, accumulate, A, (int)slabs, V)
  if (small) { if (transpose_a) GWN_NM(5, true); else GWN_NM(5, false); }
  else       { if (transpose_a) GWN_NM(4, true); else GWN_NM(4, false); }
#undef GWN_NM
  GWN_LAUNCHED();
  return 0;
}

// dA[v,w] += sum_{s,c} X[s,v,xoff+c] * G[s,w,goff+c]     (nconv weight-grad, adaptive support only)
template <typename T>
__global__ void __launch_bounds__(256) dadj_kernel(const T* __restrict__ X, int xp, int xoff,
                                                   const T* __restrict__ G, int gp, int goff,
                                                   float* __restrict__ dA, int slabs, int V,
                                                   int slabs_per_split) {
  __shared__ __align__(16) float Xs[32][68];
  __shared__ __align__(16) float Gs[32][68];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int v0 = blockIdx.x * 64, w0 = blockIdx.y * 64;
  int sb = blockIdx.z * slabs_per_split, se = min(slabs, sb + slabs_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = tid % 64, lc = (tid / 64) * 8;
  for (int s = sb; s < se; ++s) {
    float xv[8], gv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { xv[i] = 0.f; gv[i] = 0.f; }
    if (v0 + lr < V) {
      const T* src = X + ((long long)s * V + v0 + lr) * xp + xoff + lc;
      load4(src, xv); load4(src + 4, xv + 4);
    }
    if (w0 + lr < V) {
      const T* src = G + ((long long)s * V + w0 + lr) * gp + goff + lc;
      load4(src, gv); load4(src + 4, gv + 4);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { Xs[lc + i][lr] = xv[i]; Gs[lc + i][lr] = gv[i]; }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < 32; ++c) {
      float4 x = *reinterpret_cast<const float4*>(&Xs[c][ty * 4]);
      float4 g = *reinterpret_cast<const float4*>(&Gs[c][tx * 4]);
      float xa[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(xa[i], g.x, acc[i][0]);
        acc[i][1] = fmaf(xa[i], g.y, acc[i][1]);
        acc[i][2] = fmaf(xa[i], g.z, acc[i][2]);
        acc[i][3] = fmaf(xa[i], g.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int v = v0 + ty * 4 + i;
    if (v >= V) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int w = w0 + tx * 4 + j;
      if (w < V) atomicAdd(dA + (long long)v * V + w, acc[i][j]);
    }
  }
}

template <typename T>
int launch_dadj(const T* X, int xp, int xoff, const T* G, int gp, int goff, float* dA, long long slabs, int V,
                cudaStream_t st) {
  if (slabs <= 0) return 0;
  long long tiles = cdiv(V, 64) * cdiv(V, 64);
  long long want = cdiv(148 * 4, tiles);
  long long splits = want > slabs ? slabs : (want < 1 ? 1 : want);
  long long per = cdiv(slabs, splits);
  splits = cdiv(slabs, per);
  dim3 grid((unsigned)cdiv(V, 64), (unsigned)cdiv(V, 64), (unsigned)splits);
  dadj_kernel<T><<<grid, 256, 0, st>>>(X, xp, xoff, G, gp, goff, dA, (int)slabs, V, (int)per);
  GWN_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------ elementwise
// slot 0 of the concat buffer <- z = a*b  (backward recompute)
template <typename T>
__global__ void zfill_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ cat, int pitch,
                             long long P) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * 8) return;
  long long p = i >> 3; int c = (int)(i & 7) * 4;
  float av[4], bv[4];
  load4(a + p * 32 + c, av); load4(b + p * 32 + c, bv);
#pragma unroll
  for (int j = 0; j < 4; ++j) av[j] *= bv[j];
  store4(cat + p * pitch + c, av);
}

// dh = du * dropout mask
template <typename T>
__global__ void drop_bwd_kernel(const T* __restrict__ du, const T* __restrict__ mask, float p_drop,
                                uint64_t seed, uint64_t offset, const uint64_t* __restrict__ rng,
                                T* __restrict__ dh, long long P) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * 8) return;
  long long p = i >> 3; int c = (int)(i & 7) * 4;
  float g[4], m[4];
  load4(du + p * 32 + c, g);
  if (mask) {
    load4(mask + p * 32 + c, m);
  } else {
    const uint64_t sd = rng ? __ldg(rng) : seed;
    const uint64_t of = rng ? offset + __ldg(rng + 1) : offset;
    dropout4(sd, of, (uint64_t)p, c, p_drop, m);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) g[j] *= m[j];
  store4(dh + p * 32 + c, g);
}

// dfg[p, 2c] = dz*b*(1-a^2) ; dfg[p, 2c+1] = dz*a*b*(1-b);  dz = dcat slot0 (+ dz_last on the tail rows)
template <typename T, typename TO>
__global__ void gate_bwd_kernel(const T* __restrict__ dz, int dz_pitch, const T* __restrict__ dz_last,
                                long long rows_per_n, long long last_begin, long long last_rows,
                                const T* __restrict__ a, const T* __restrict__ b, TO* __restrict__ dfg,
                                long long P) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * 8) return;
  long long p = i >> 3; int c = (int)(i & 7) * 4;
  float g[4] = {0.f, 0.f, 0.f, 0.f}, av[4], bv[4];
  if (dz) load4(dz + p * dz_pitch + c, g);
  if (dz_last) {
    long long n, rem;
    split_pos(p, rows_per_n, n, rem);
    if (rem >= last_begin) {
      float t[4]; load4(dz_last + (n * last_rows + rem - last_begin) * 32 + c, t);
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] += t[j];
    }
  }
  load4(a + p * 32 + c, av); load4(b + p * 32 + c, bv);
  float o[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    o[2 * j] = g[j] * bv[j] * (1.f - av[j] * av[j]);
    o[2 * j + 1] = g[j] * av[j] * bv[j] * (1.f - bv[j]);
  }
  TO* dst = dfg + p * 64 + 2 * c;
  store4(dst, o);
  store4(dst + 4, o + 4);
}

// ------------------------------------------------------------------------------------------ BatchNorm fold
__global__ void bn_fold_kernel(const double* stats, double count, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float momentum, float eps,
                               int training, float* scale, float* shift, float* mean_out, float* rstd_out) {
  int c = threadIdx.x;
  if (c >= 32) return;
  float mean, var;
  if (training) {
    double m = stats[c] / count;
    double v = stats[32 + c] / count - m * m;
    if (v < 0) v = 0;
    mean = (float)m; var = (float)v;
    double unbiased = count > 1 ? v * (count / (count - 1.0)) : v;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  } else {
    mean = running_mean[c]; var = running_var[c];
  }
  float rstd = rsqrtf(var + eps);
  if (training) rstd = (float)(1.0 / sqrt((double)var + (double)eps));
  else rstd = 1.f / sqrtf(var + eps);
  float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  if (mean_out) { mean_out[c] = mean; rstd_out[c] = rstd; }
}

template <typename T, typename DX>
__global__ void bn_bwd_kernel(const DX* __restrict__ dx, const T* __restrict__ u, const double* dx_stats,
                              double count, const float* gamma, const float* mean, const float* rstd,
                              int training, T* __restrict__ du, float* dgamma, float* dbeta, long long rows) {
  __shared__ float k0[32], k1[32], k2[32];  // du = k0*dx + k1*u + k2
  pdl_wait();         // (programmatic launch: only the launch latency overlaps - dx_stats come from the predecessor)
  pdl_trigger();
  if (threadIdx.x < 32) {
    int c = threadIdx.x;
    double sdx = dx_stats[c], sdxu = dx_stats[32 + c];
    double mu = mean[c], rs = rstd[c], g = gamma[c];
    double sdxh = rs * (sdxu - mu * sdx);  // sum dx * xhat
    if (blockIdx.x == 0) { dgamma[c] = (float)sdxh; dbeta[c] = (float)sdx; }
    if (training) {
      double m1 = sdx / count, m2 = sdxh / count;
      // du = g*rs*(dx - m1 - xhat*m2),  xhat = (u-mu)*rs
      k0[c] = (float)(g * rs);
      k1[c] = (float)(-g * rs * m2 * rs);
      k2[c] = (float)(-g * rs * (m1 - mu * rs * m2));
    } else {
      k0[c] = (float)(g * rs); k1[c] = 0.f; k2[c] = 0.f;
    }
  }
  __syncthreads();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 4) return;
  long long p = i >> 2; int c = (int)(i & 3) * 8;       // a thread owns 8 channels of a position
  float g[8], uv[8];
  load8(dx + p * 32 + c, g); load8(u + p * 32 + c, uv);
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = k0[c + j] * g[j] + k1[c + j] * uv[j] + k2[c + j];
  store8(du + p * 32 + c, g);
}

// ------------------------------------------------------------------------------------------ orchestration
static int check_cfg(const gwn_layer_cfg* c) {
  GWN_REQUIRE(c != nullptr, "layer cfg is NULL");
  GWN_REQUIRE(c->dtype == GWN_F32 || c->dtype == GWN_BF16, "bad dtype %d", c->dtype);
  GWN_REQUIRE(c->taps >= 1 && c->taps <= GWN_MAX_TAPS, "kernel_size %d unsupported (1..%d)", c->taps, GWN_MAX_TAPS);
  GWN_REQUIRE(c->n_supports >= 0 && c->n_supports <= GWN_MAX_SUPPORTS, "n_supports %d > %d", c->n_supports,
              GWN_MAX_SUPPORTS);
  GWN_REQUIRE(c->order >= 1 && 1 + c->order * c->n_supports <= GEMM_MAX_CHUNKS, "order %d unsupported", c->order);
  GWN_REQUIRE(c->Lout == c->Lin - c->dilation * (c->taps - 1) && c->Lout >= 1, "bad Lin/Lout %d/%d", c->Lin, c->Lout);
  GWN_REQUIRE(c->Lf >= 1 && c->Lf <= c->Lout, "bad Lf %d", c->Lf);
  GWN_REQUIRE(c->N >= 1 && c->V >= 1, "bad N/V");
  GWN_REQUIRE((long long)c->N * c->Lin * c->V < (1ll << 31), "layer: N*Lin*V must be < 2^31 positions");
  return 0;
}

// The ONE rule that picks the image format of `hop_mats` (include/gwn.h): the Python side asks this same function which
// images to build, so the two sides cannot disagree about what the buffer holds.
extern "C" int gwn_hop_mode(int V, int n_supports) {
  if (n_supports < 1 || n_supports > GWN_MAX_SUPPORTS || V < 1) return 0;
  return (V <= 80 && hops_tc_supported(V, 2 * n_supports) != 0) ? 1 : 2;
}

// tensor-core path available for this layer?  (bf16 storage, images prepared, supports fit on chip)
template <typename T>
static bool use_tc_hops(const gwn_layer_cfg* c, const void* hop_mats) {
  if constexpr (!std::is_same<T, bf16>::value) return false;
  return hop_mats != nullptr && c->order == 2 && c->n_supports >= 1 && gwn_hop_mode(c->V, c->n_supports) == 1;
}

// 0: CUDA-core hops; 1: supports resident on chip (tc_hops.cu); 2: TMA-tiled GEMM per hop (tma_gemm.cu, V > 80).
// `hop_mats` holds gwn_hop_mats_prep images in mode 1 and gwn_support_images_prep images in mode 2.
template <typename T>
static int tc_mode(const gwn_layer_cfg* c, const void* hop_mats) {
  if constexpr (!std::is_same<T, bf16>::value) return 0;
  if (hop_mats == nullptr) return 0;
  if (c->n_supports == 0) return 1;    // no hops at all: the position GEMMs still run on tensor cores
  return c->order == 2 ? gwn_hop_mode(c->V, c->n_supports) : 2;
}
static const bf16* big_image(const gwn_layer_cfg* c, const void* hop_mats, int s, int which) {
  const long long Vp = ((c->V + 7) / 8) * 8;
  return reinterpret_cast<const bf16*>(hop_mats) + ((long long)s * 2 + which) * c->V * Vp;
}

// one hop of support s at V > 80: the sparse gather when the caller supplied ELL rows for it, else the dense tensor-core GEMM
static int hop_any(const gwn_layer_cfg* c, const void* hop_mats, const gwn_ell* ell, int s, int which, const bf16* X, bf16* Y,
                   const bf16* add, long long slabs, cudaStream_t st) {
  if (ell && ell[s].width > 0 && ell[s].idx[which] && ell[s].val[which])
    return launch_hop_ell(ell[s].idx[which], ell[s].val[which], ell[s].width, X, Y, add, slabs, c->V, st);
  const int Vp = ((c->V + 7) / 8) * 8;
  return launch_hop_big(big_image(c, hop_mats, s, which), Vp, X, Y, add, slabs, c->V, st);
}

static bool fused_gate_wgrad_enabled() {   // GWN_FUSED_WGRAD=0 keeps the separate weight-gradient launch (A/B measurements)
  static int v = -1;
  if (v < 0) { const char* e = getenv("GWN_FUSED_WGRAD"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

static bool horner_bwd_enabled() {   // GWN_HORNER_BWD=0 keeps the recompute-based backward of the big-graph path (A/B measurements)
  static int v = -1;
  if (v < 0) { const char* e = getenv("GWN_HORNER_BWD"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

static bool gate_tma_store_enabled() {   // GWN_GATE_TMA_STORE=0: the gated conv writes z, a, b with per-thread stores (A/B measurements)
  static int v = -1;
  if (v < 0) { const char* e = getenv("GWN_GATE_TMA_STORE"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

static bool gcn_t_enabled() {       // GWN_GCN_T=0 keeps the node-major fused forward (A/B measurements)
  static int v = -1;
  if (v < 0) { const char* e = getenv("GWN_GCN_T"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

static bool gcn_bwd_t_enabled() {   // GWN_GCN_BWD_T=0 keeps the node-major fused backward (A/B measurements)
  static int v = -1;
  if (v < 0) { const char* e = getenv("GWN_GCN_BWD_T"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

// start (in bf16 elements) of the stacked transposed-hop image of the T-form fused backward inside gwn_hop_mats_prep's buffer
static size_t mats_bt_offset(int V, int n_supports) {
  const int Kp = ((V + 15) / 16) * 16;
  const int KT = (((1 + 2 * n_supports) * V + 15) / 16) * 16;
  return (size_t)n_supports * 4 * (Kp / 8) * 1024 + (size_t)KT * Kp;
}

static bool fused_gcn_enabled() {   // GWN_NO_FUSED_GCN=1 keeps the unfused kernels (A/B measurements only)
  static int v = -1;
  if (v < 0) { const char* e = getenv("GWN_NO_FUSED_GCN"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

// The concat buffers are SLOT-MAJOR: slot q (32 channels of hop q) is a contiguous [P, 32] tensor at
// buf + q*P*32, so a tile of a slot is one contiguous run of 64-byte rows (full 128-byte DRAM lines).
static void hop_params_base(HopParams& p, const gwn_layer_cfg* c, bf16* buf, const void* hop_mats) {
  const long long P = (long long)c->N * c->Lout * c->V;
  p.in[0] = p.in[1] = buf; p.out[0] = p.out[1] = buf;
  p.in_pitch[0] = p.in_pitch[1] = p.out_pitch[0] = p.out_pitch[1] = 32;
  p.slot_stride[0] = p.slot_stride[1] = P * 32;
  p.mats = reinterpret_cast<const bf16*>(hop_mats);
  p.V = c->V; p.slabs = c->N * c->Lout;
}

// forward hops on tcgen05: the z tile is loaded once, all 2*S outputs (A_s and A_s^2) come from it
static int hops_forward_tc(const gwn_layer_cfg* c, bf16* cat, const void* hop_mats, cudaStream_t st) {
  HopParams p{};
  hop_params_base(p, c, cat, hop_mats);
  const int nh = 2 * c->n_supports;
  p.n_mats = nh; p.n_steps = nh; p.n_outs = nh;
  for (int j = 0; j < nh; ++j) {
    p.mat_src[j] = 4 * (j / 2) + (j % 2);                 // A_s^T, (A_s^2)^T
    p.steps[j] = HopStep{0, 0, j, j & 1, TH_FIRST | TH_LAST | (j == 0 ? TH_LOAD : 0) | (j == nh - 1 ? TH_RELEASE : 0)};
    p.outs[j] = HopOut{0, 1 + j, -1, 0};
  }
  return launch_hops_tc(p, st);
}

template <typename T>
static int hops_forward(const gwn_layer_cfg* c, T* cat, const float* const* supports,
                        const void* hop_mats, const gwn_ell* ell, cudaStream_t st) {
  const long long slabs = (long long)c->N * c->Lout;
  const long long SS = slabs * c->V * 32;   // slot stride
  if constexpr (std::is_same<T, bf16>::value) {
    const int mode = tc_mode<T>(c, hop_mats);
    if (mode == 1 && c->n_supports > 0) return hops_forward_tc(c, cat, hop_mats, st);
    if (mode == 2) {
      for (int s = 0; s < c->n_supports; ++s)
        for (int k = 1; k <= c->order; ++k) {
          const int slot = 1 + s * c->order + (k - 1), src = (k == 1) ? 0 : slot - 1;
          if (int rc = hop_any(c, hop_mats, ell, s, 0, cat + src * SS, cat + slot * SS, nullptr, slabs, st)) return rc;
        }
      return 0;
    }
  }
  for (int s = 0; s < c->n_supports; ++s)
    for (int k = 1; k <= c->order; ++k) {
      int slot = 1 + s * c->order + (k - 1);
      int src = (k == 1) ? 0 : slot - 1;
      if (int rc = launch_node_mix<T>(cat + src * SS, 32, 0, cat + slot * SS, 32, 0, 0, supports[s], false, slabs,
                                      c->V, st))
        return rc;
    }
  return 0;
}

template <typename T>
static int layer_fwd_t(const gwn_layer_cfg* c, const gwn_layer_fwd_args* g, cudaStream_t st) {
  const long long RO = (long long)c->Lout * c->V, RI = (long long)c->Lin * c->V;
  const long long P = c->N * RO;
  const int nslots = 1 + c->order * c->n_supports, mlp_in = 32 * nslots;
  T* cat = reinterpret_cast<T*>(g->ws_cat);
  // gate
  GemmA A{};
  A.n_chunks = c->taps; A.rows_per_n_out = RO; A.P = P;
  for (int j = 0; j < c->taps; ++j) {
    AChunk& ch = A.ch[j];
    ch.base = g->u_prev; ch.rows_per_n = RI; ch.row_off = (long long)j * c->dilation * c->V;
    ch.pitch = 32; ch.col_off = 0; ch.scale = g->scale; ch.shift = g->shift; ch.relu = 0;
  }
  bool tc = false;
  if constexpr (std::is_same<T, bf16>::value) tc = tc_mode<T>(c, g->hop_mats) != 0 && g->ws_w != nullptr && c->taps <= 4;
  GWN_REQUIRE(tc || !g->bn_gamma, "layer_fwd: the in-kernel BatchNorm fold (bn_gamma) needs the bf16 tensor-core path");
  uint8_t* wsw = reinterpret_cast<uint8_t*>(g->ws_w);
  const long long last_begin = (long long)(c->Lout - c->Lf) * c->V, last_rows = (long long)c->Lf * c->V;
  bool fused_fwd = false;     // supports on chip: hops + concat + mlp + dropout + residual + statistics as ONE kernel
  if constexpr (std::is_same<T, bf16>::value)
    fused_fwd = tc && c->has_gconv && fused_gcn_enabled() && tc_mode<T>(c, g->hop_mats) == 1 && c->order == 2 &&
                c->n_supports >= 1 && gcn_fused_supported(c->V, 2 * c->n_supports);
  if (tc) {
    if constexpr (std::is_same<T, bf16>::value) {
      // The gate kernel builds its own bf16 UMMA weight image in its prologue (tc_gemm.cuh: PgWsrc): the previous
      // layer's BatchNorm affine is folded into the weights / bias there (from the raw batch statistics when the caller
      // passes bn_gamma, else from scale / shift computed by gwn_bn_fold), the next statistics buffer is zeroed by
      // CTA 0 - no weight-prep or fold launch sits between the layer's kernels.
      PgParams pg{};
      pg.n_chunks = c->taps; pg.rows_per_n_out = RO; pg.P = P; pg.N = 64; pg.has_bias = 1;
      PgWsrc& ws = pg.wsrc;
      ws.W = g->w_fg; ws.ld = 64; ws.transposed = 0; ws.half_odd = 1; ws.bias = g->b_fg;
      for (int j = 0; j < c->taps; ++j) ws.w_off[j] = j * 32 * 64;
      if (g->bn_gamma) {
        GWN_REQUIRE(g->scale && g->shift && g->bn_running_mean && g->bn_running_var && (g->bn_stats || !c->training),
                    "layer_fwd: fused BatchNorm fold needs scale/shift outputs, running statistics and batch statistics");
        ws.bn = 1; ws.bn_training = c->training; ws.bn_stats = g->bn_stats; ws.bn_count = g->bn_count;
        ws.gamma = g->bn_gamma; ws.beta = g->bn_beta; ws.running_mean = g->bn_running_mean;
        ws.running_var = g->bn_running_var; ws.eps = g->bn_eps; ws.momentum = g->bn_momentum;
        ws.scale_out = const_cast<float*>(g->scale); ws.shift_out = const_cast<float*>(g->shift);
        ws.mean_out = g->bn_mean; ws.rstd_out = g->bn_rstd;
      } else if (g->scale) {          // scale / shift already computed by the caller (gwn_bn_fold)
        ws.bn = 2; ws.scale_in = g->scale; ws.shift_in = g->shift;
      }
      if (c->has_gconv) ws.zero64 = g->stats;
      for (int j = 0; j < c->taps; ++j)
        pg.ch[j] = PgChunk{reinterpret_cast<const bf16*>(g->u_prev), RI, (long long)j * c->dilation * c->V, 32, 0};
      EpiGateTC eg{};
      eg.z = cat;
      eg.a = c->training ? reinterpret_cast<bf16*>(g->a) : nullptr;
      eg.b = c->training ? reinterpret_cast<bf16*>(g->b) : nullptr;
      eg.z_last = reinterpret_cast<bf16*>(g->z_last); eg.last_begin = last_begin; eg.last_rows = last_rows;
      eg.staged = gate_tma_store_enabled() ? 1 : 0;
      if (int rc = launch_pos_gemm_tc(pg, eg, st)) return rc;
    }
  } else {
  EpiGate<T> eg{};
  eg.bias = g->b_fg; eg.zcat = cat; eg.zpitch = 32;   // slot 0 of the slot-major concat buffer
  eg.a = c->training ? reinterpret_cast<T*>(g->a) : nullptr;
  eg.b = c->training ? reinterpret_cast<T*>(g->b) : nullptr;
  eg.z_last = reinterpret_cast<T*>(g->z_last);
  eg.last_begin = last_begin; eg.last_rows = last_rows;
  if (int rc = launch_pos_gemm<T, 64>(A, g->w_fg, 64, eg, st)) return rc;
  }
  if (!c->has_gconv) return 0;
  if (!tc) GWN_CUDA(cudaMemsetAsync(g->stats, 0, sizeof(double) * 64, st));      // (tc: zeroed by CTA 0 of the gate kernel)
  if constexpr (std::is_same<T, bf16>::value) {
    // supports on chip: hops + concat + mlp + dropout + residual + statistics as ONE kernel (gcn_fused.cu)
    if (fused_fwd) {
      GcnFwdParams fp{};
      fp.z = cat; fp.u_prev = reinterpret_cast<const bf16*>(g->u_prev); fp.RI = RI; fp.RO = RO;
      fp.crop = (long long)(c->Lin - c->Lout) * c->V; fp.scale = g->scale; fp.shift = g->shift;
      fp.mats = reinterpret_cast<const bf16*>(g->hop_mats); fp.n_mats = 2 * c->n_supports;
      for (int j = 0; j < fp.n_mats; ++j) fp.mat_src[j] = 4 * (j / 2) + (j % 2);      // A_s^T, (A_s^2)^T
      fp.w_img = nullptr; fp.w_src = g->w_mlp; fp.bias = g->b_mlp;      // image built in the kernel's prologue
      fp.mask = c->training ? reinterpret_cast<const bf16*>(g->drop_mask) : nullptr;
      fp.drop_p = c->training ? c->dropout_p : 0.f; fp.seed = c->seed; fp.offset = c->offset; fp.rng = g->rng;
      fp.u = reinterpret_cast<bf16*>(g->u); fp.stats = g->stats; fp.V = c->V; fp.slabs = c->N * c->Lout;
      if (gcn_t_enabled() && gcn_fused_t_supported(c->V, fp.n_mats)) {      // transposed contraction over groups of 4 slabs
        const int Kp = ((c->V + 15) / 16) * 16;
        fp.mats_t = reinterpret_cast<const bf16*>(g->hop_mats) + (size_t)c->n_supports * 4 * (Kp / 8) * 1024;
        return launch_gcn_fwd_t(fp, st);
      }
      return launch_gcn_fwd(fp, st);
    }
  }
  // diffusion hops into the concat slots, then mlp + dropout + residual + stats
  if (int rc = hops_forward<T>(c, cat, g->supports, g->hop_mats, g->ell, st)) return rc;
  GemmA M{};
  M.n_chunks = nslots; M.rows_per_n_out = RO; M.P = P;
  for (int q = 0; q < nslots; ++q) {
    AChunk& ch = M.ch[q];
    ch.base = cat + q * P * 32; ch.rows_per_n = RO; ch.row_off = 0; ch.pitch = 32; ch.col_off = 0;
  }
  if (tc && nslots <= 7) {
    if constexpr (std::is_same<T, bf16>::value) {
      WPrepParams wp{};
      wp.W = g->w_mlp; wp.ld = 32; wp.transposed = 0; wp.K = mlp_in; wp.N = 32;
      for (int q = 0; q < nslots; ++q) wp.w_off[q] = (long long)q * 32 * 32;
      wp.bias = g->b_mlp;
      wp.img = reinterpret_cast<bf16*>(wsw + 64 * 1024); wp.bias_out = nullptr; wp.bias_chunk = 1;
      if (int rc = launch_wprep(wp, st)) return rc;
      PgParams pg{};
      pg.n_chunks = nslots; pg.rows_per_n_out = RO; pg.P = P; pg.N = 32; pg.w_img = wp.img; pg.has_bias = 1;
      for (int q = 0; q < nslots; ++q) pg.ch[q] = PgChunk{cat + q * P * 32, RO, 0, 32, 0};
      EpiMlpTC em{};
      em.u_prev = reinterpret_cast<const bf16*>(g->u_prev); em.prev_rows_per_n = RI;
      em.crop = (long long)(c->Lin - c->Lout) * c->V; em.scale = g->scale; em.shift = g->shift;
      em.mask = c->training ? reinterpret_cast<const bf16*>(g->drop_mask) : nullptr;
      em.drop_p = c->training ? c->dropout_p : 0.f; em.seed = c->seed; em.offset = c->offset; em.rng = g->rng;
      em.u = reinterpret_cast<bf16*>(g->u); em.stats = g->stats; em.s1 = 0.f; em.s2 = 0.f;
      return launch_pos_gemm_tc(pg, em, st);
    }
  }
  EpiMlp<T> em{};
  em.stats = g->stats; em.bias = g->b_mlp;
  em.u_prev = reinterpret_cast<const T*>(g->u_prev); em.prev_rows_per_n = RI;
  em.crop = (long long)(c->Lin - c->Lout) * c->V; em.scale = g->scale; em.shift = g->shift;
  em.mask = reinterpret_cast<const T*>(g->drop_mask);
  em.drop_p = c->training ? c->dropout_p : 0.f; em.seed = c->seed; em.offset = c->offset; em.rng = g->rng;
  if (!c->training) em.mask = nullptr;
  em.u = reinterpret_cast<T*>(g->u);
  return launch_pos_gemm<T, 32>(M, g->w_mlp, 32, em, st);
}

template <typename T>
static int layer_bwd_t(const gwn_layer_cfg* c, const gwn_layer_bwd_args* g, cudaStream_t st) {
  const long long RO = (long long)c->Lout * c->V, RI = (long long)c->Lin * c->V;
  const long long P = c->N * RO, PI = c->N * RI;
  const long long slabs = (long long)c->N * c->Lout;
  const int nslots = 1 + c->order * c->n_supports, mlp_in = 32 * nslots;
  T* cat = reinterpret_cast<T*>(g->ws_cat);
  T* dcat = reinterpret_cast<T*>(g->ws_dcat);
  const T* a = reinterpret_cast<const T*>(g->a);
  const T* b = reinterpret_cast<const T*>(g->b);
  const T* du = reinterpret_cast<const T*>(g->du);
  const unsigned eb = (unsigned)cdiv(P * 8, 256);
  const T* dz = nullptr;
  bool fused_bwd = false;
  if constexpr (std::is_same<T, bf16>::value) {
    // supports on chip and none of them needs a gradient: the whole diffusion backward (mask, transposed hops, mlp
    // data + weight gradients, gate backward) is ONE kernel (gcn_fused_bwd.cu) that leaves dfg for the conv backward
    int n_dA = 0, sa = -1;
    for (int s = 0; s < c->n_supports; ++s)
      if (g->support_needs_grad[s] && g->d_supports[s]) { ++n_dA; sa = s; }
    if (du && n_dA <= 1 && fused_gcn_enabled() && tc_mode<T>(c, g->hop_mats) == 1 && c->order == 2 && c->n_supports >= 1 &&
        g->ws_w != nullptr && gcn_bwd_fused_supported(c->V, 2 * c->n_supports) && wgrad_tc_supported(c->taps, 64)) {
      if (!g->outputs_zeroed) GWN_CUDA(cudaMemsetAsync(g->dw_mlp, 0, sizeof(float) * 32 * mlp_in, st));
      if (!g->outputs_zeroed) GWN_CUDA(cudaMemsetAsync(g->db_mlp, 0, sizeof(float) * 32, st));
      // (weight images: built by the kernels' own prologues - gcn_bwd from w_mlp, the dx GEMM below from w_fg)
      GcnBwdParams bp{};
      bp.du = du; bp.a = a; bp.b = b; bp.dz_last = reinterpret_cast<const bf16*>(g->dz_last);
      bp.RO = RO; bp.last_begin = (long long)(c->Lout - c->Lf) * c->V; bp.last_rows = (long long)c->Lf * c->V;
      bp.mats = reinterpret_cast<const bf16*>(g->hop_mats); bp.n_mats = 2 * c->n_supports;
      for (int j = 0; j < bp.n_mats; ++j) bp.mat_src[j] = 4 * (j / 2) + 2 + (j % 2);     // A_s, A_s^2 (transposed hops)
      bp.wt_img = nullptr; bp.w_src = g->w_mlp;
      const bool drop = c->training && (g->drop_mask != nullptr || c->dropout_p > 0.f);
      bp.mask = drop ? reinterpret_cast<const bf16*>(g->drop_mask) : nullptr;
      bp.drop_p = drop ? c->dropout_p : 0.f; bp.seed = c->seed; bp.offset = c->offset; bp.rng = g->rng;
      bp.dfg = reinterpret_cast<bf16*>(g->ws_dfg); bp.dw_mlp = g->dw_mlp; bp.db_mlp = g->db_mlp;
      bp.V = c->V; bp.slabs = c->N * c->Lout;
      {
        bp.trace = trace_ptr("GWN_GCN_TRACE");
      }
      bp.sa = sa; bp.mat_fwd = sa >= 0 ? 4 * sa : 0; bp.w56_img = nullptr; bp.dA = sa >= 0 ? g->d_supports[sa] : nullptr;
      // transposed hops over groups of four slabs (gcn_fused_bwd_t.cu) when the shape has an instance and the caller
      // takes the second-order part of the support gradient in factored form (d_supports_sq)
      if (gcn_bwd_t_enabled() && gcn_bwd_t_supported(c->V, bp.n_mats, sa >= 0) && (sa < 0 || g->d_supports_sq[sa])) {
        bp.mats_bt = reinterpret_cast<const bf16*>(g->hop_mats) + mats_bt_offset(c->V, c->n_supports);
        bp.dQ6 = sa >= 0 ? g->d_supports_sq[sa] : nullptr;
        if (int rc = launch_gcn_bwd_t(bp, st)) return rc;
      } else {
        if (int rc = launch_gcn_bwd(bp, st)) return rc;
      }
      fused_bwd = true;
    }
  }
  bool horner_done = false;
  if constexpr (std::is_same<T, bf16>::value) {
    // ---- big graphs (V > 80): Horner-form backward - no forward recompute ----
    // dU_0 = dh, dU_{2s+1} = dh A_s^T, dU_{2s+2} = dU_{2s+1} A_s^T (2 transposed hops per support on the 32-channel dh),
    // dz = sum_j dU_j W_j^T (one position GEMM), dW_j = z^T dU_j (K = positions), and for the support with a gradient:
    // T1 = z W_{2s+1} + (z W_{2s+2}) A_s,  dA += T1^T dh + (z W_{2s+2})^T dU_{2s+1}.  9 big GEMMs per layer instead of 14
    // (6 recomputed forward hops + 6 backward hops + 2 support-gradient GEMMs).
    int n_dA = 0, sa = -1;
    for (int s = 0; s < c->n_supports; ++s)
      if (g->support_needs_grad[s] && g->d_supports[s]) { ++n_dA; sa = s; }
    // (the support-gradient scratch U5 / U6 / T1 lives in cat slots 2..4: with a single support the concat workspace has
    //  only 3 slots, so that shape takes the recompute path below)
    if (!fused_bwd && du && tc_mode<T>(c, g->hop_mats) == 2 && c->order == 2 && c->n_supports >= 1 && nslots <= 7 && n_dA <= 1 &&
        (sa < 0 || nslots >= 5) && g->ws_w != nullptr && horner_bwd_enabled() && wgrad_tc_supported(1, 32)) {
      const int Vp = ((c->V + 7) / 8) * 8;
      const long long SS = P * 32;
      // z = a . b -> cat slot 0;  dh = du . mask -> dcat slot 0
      zfill_kernel<T><<<eb, 256, 0, st>>>(a, b, cat, 32, P);
      GWN_LAUNCHED();
      const bool drop = c->training && (g->drop_mask != nullptr || c->dropout_p > 0.f);
      if (drop) {
        drop_bwd_kernel<T><<<eb, 256, 0, st>>>(du, reinterpret_cast<const T*>(g->drop_mask), c->dropout_p, c->seed,
                                               c->offset, g->rng, dcat, P);
        GWN_LAUNCHED();
      } else {
        GWN_CUDA(cudaMemcpyAsync(dcat, du, sizeof(T) * (size_t)SS, cudaMemcpyDeviceToDevice, st));
      }
      // transposed hops
      for (int s = 0; s < c->n_supports; ++s)
        for (int k = 1; k <= 2; ++k) {
          const int slot = 2 * s + k, src = (k == 1) ? 0 : slot - 1;
          if (int rc = hop_any(c, g->hop_mats, g->ell, s, 1, dcat + src * SS, dcat + slot * SS, nullptr, slabs, st)) return rc;
        }
      // dW_j = z^T dU_j (j = 0 also gives db = sum dh through the ones row)
      if (!g->outputs_zeroed) GWN_CUDA(cudaMemsetAsync(g->dw_mlp, 0, sizeof(float) * 32 * mlp_in, st));
      if (!g->outputs_zeroed) GWN_CUDA(cudaMemsetAsync(g->db_mlp, 0, sizeof(float) * 32, st));
      for (int j = 0; j < nslots; ++j) {
        WgParams w{};
        w.n_chunks = 1; w.rows_per_n_out = RO; w.P = P;
        w.ch[0] = WgChunk{cat, RO, 0, 32, 0};
        w.G = dcat + j * SS; w.g_pitch = 32; w.N = 32; w.dW = g->dw_mlp + (size_t)j * 32 * 32; w.ldw = 32;
        w.db = j == 0 ? g->db_mlp : nullptr;
        if (int rc = launch_wgrad_tc(w, st)) return rc;
      }
      // dz = sum_j dU_j W_j^T -> cat slot 1   (weight image (k = (j, c'), n = c) = W[j*32 + c][c'] built in the kernel)
      {
        PgParams pg{};
        pg.n_chunks = nslots; pg.rows_per_n_out = RO; pg.P = P; pg.N = 32;
        pg.wsrc.W = g->w_mlp; pg.wsrc.ld = 32; pg.wsrc.transposed = 1;
        for (int q = 0; q < nslots; ++q) {
          pg.wsrc.w_off[q] = q * 32 * 32;
          pg.ch[q] = PgChunk{dcat + q * SS, RO, 0, 32, 0};
        }
        EpiSlotTC es{}; es.out = cat + SS; es.slot_stride = SS;
        if (int rc = launch_pos_gemm_tc(pg, es, st)) return rc;
      }
      if (sa >= 0) {
        // U5 = z W_{2sa+1} -> cat slot 2, U6 = z W_{2sa+2} -> cat slot 3;  T1 = U5 + U6 A_sa -> cat slot 4
        for (int h = 0; h < 2; ++h) {
          PgParams pg{};
          pg.n_chunks = 1; pg.rows_per_n_out = RO; pg.P = P; pg.N = 32;
          pg.wsrc.W = g->w_mlp + (size_t)(2 * sa + 1 + h) * 32 * 32; pg.wsrc.ld = 32; pg.wsrc.transposed = 0; pg.wsrc.w_off[0] = 0;
          pg.ch[0] = PgChunk{cat, RO, 0, 32, 0};
          EpiSlotTC es{}; es.out = cat + (2 + h) * SS; es.slot_stride = SS;
          if (int rc = launch_pos_gemm_tc(pg, es, st)) return rc;
        }
        if (int rc = launch_hop_big(big_image(c, g->hop_mats, sa, 0), Vp, cat + 3 * SS, cat + 4 * SS, cat + 2 * SS, slabs, c->V, st))
          return rc;
        if (int rc = launch_dadj_big(cat + 4 * SS, dcat, g->d_supports[sa], slabs, c->V, st)) return rc;
        if (int rc = launch_dadj_big(cat + 3 * SS, dcat + (2 * sa + 1) * SS, g->d_supports[sa], slabs, c->V, st)) return rc;
      }
      dz = cat + SS;
      horner_done = true;
    }
  }
  if (fused_bwd) {
    // dfg is ready: fall through to the conv weight / data gradients
  } else if (horner_done) {
    // dz is ready: fall through to the gate backward
  } else if (du) {
    // recompute the concat (z and its hops)
    zfill_kernel<T><<<eb, 256, 0, st>>>(a, b, cat, 32, P);
    GWN_LAUNCHED();
    if (int rc = hops_forward<T>(c, cat, g->supports, g->hop_mats, g->ell, st)) return rc;
    // dh = du * mask
    const T* dh = du;
    const bool drop = c->training && (g->drop_mask != nullptr || c->dropout_p > 0.f);
    if (drop) {
      T* tmp = reinterpret_cast<T*>(g->ws_dfg);
      drop_bwd_kernel<T><<<eb, 256, 0, st>>>(du, reinterpret_cast<const T*>(g->drop_mask), c->dropout_p, c->seed,
                                             c->offset, g->rng, tmp, P);
      GWN_LAUNCHED();
      dh = tmp;
    }
    // dW_mlp [mlp_in, 32], db_mlp
    GemmA M{};
    M.n_chunks = nslots; M.rows_per_n_out = RO; M.P = P;
    for (int q = 0; q < nslots; ++q) {
      AChunk& ch = M.ch[q];
      ch.base = cat + q * P * 32; ch.rows_per_n = RO; ch.row_off = 0; ch.pitch = 32; ch.col_off = 0;
    }
    bool wg_done = false;
    if constexpr (std::is_same<T, bf16>::value) {
      if (tc_mode<T>(c, g->hop_mats) != 0 && wgrad_tc_supported(nslots, 32)) {
        if (!g->outputs_zeroed) GWN_CUDA(cudaMemsetAsync(g->dw_mlp, 0, sizeof(float) * 32 * mlp_in, st));
        if (!g->outputs_zeroed) GWN_CUDA(cudaMemsetAsync(g->db_mlp, 0, sizeof(float) * 32, st));
        WgParams w{};
        w.n_chunks = nslots; w.rows_per_n_out = RO; w.P = P;
        for (int q = 0; q < nslots; ++q) w.ch[q] = WgChunk{cat + q * P * 32, RO, 0, 32, 0};
        w.G = dh; w.g_pitch = 32; w.N = 32; w.dW = g->dw_mlp; w.ldw = 32; w.db = g->db_mlp;
        if (int rc = launch_wgrad_tc(w, st)) return rc;
        wg_done = true;
      }
    }
    if (!wg_done)
      if (int rc = launch_wgrad<T, T>(M, dh, 32, 0, g->dw_mlp, 32, g->db_mlp, st)) return rc;
    // dcat[p, (slot,c)] = sum_o dh[p,o] * w_mlp_t[(slot,c), o]
    GemmA D{};
    D.n_chunks = 1; D.rows_per_n_out = RO; D.P = P;
    D.ch[0].base = dh; D.ch[0].rows_per_n = RO; D.ch[0].row_off = 0; D.ch[0].pitch = 32; D.ch[0].col_off = 0;
    D.ch[0].w_off = 0;
    bool dcat_done = false;
    if constexpr (std::is_same<T, bf16>::value) {
      if (tc_mode<T>(c, g->hop_mats) != 0 && g->ws_w != nullptr && mlp_in <= 256) {
        uint8_t* wsw = reinterpret_cast<uint8_t*>(g->ws_w);
        WPrepParams wp{};
        wp.W = g->w_mlp; wp.ld = 32; wp.transposed = 1; wp.K = 32; wp.N = mlp_in; wp.w_off[0] = 0;
        wp.img = reinterpret_cast<bf16*>(wsw); wp.bias_out = nullptr;
        if (int rc = launch_wprep(wp, st)) return rc;
        PgParams pg{};
        pg.n_chunks = 1; pg.rows_per_n_out = RO; pg.P = P; pg.N = mlp_in; pg.w_img = wp.img;
        pg.ch[0] = PgChunk{dh, RO, 0, 32, 0};
        EpiSlotTC es{}; es.out = dcat; es.slot_stride = P * 32;
        if (int rc = launch_pos_gemm_tc(pg, es, st)) return rc;
        dcat_done = true;
      }
    }
    if (!dcat_done) {
      EpiSlotStore<T> es{}; es.out = dcat; es.slot_stride = P * 32;
      if (int rc = launch_pos_gemm_wt<T, 32>(D, g->w_mlp, mlp_in, 32, es, st)) return rc;
    }
    // hops backward
    bool tc_done = false;
    if constexpr (std::is_same<T, bf16>::value) {
      if (tc_mode<T>(c, g->hop_mats) == 2) {
        // big graphs: sequential hops, each one TMA-tiled tensor-core GEMM; dA likewise (K = slab x channel)
        const int Vp = ((c->V + 7) / 8) * 8;
        for (int s = 0; s < c->n_supports; ++s)
          for (int k = c->order; k >= 1; --k) {
            const int slot = 1 + s * c->order + (k - 1), src = (k == 1) ? 0 : slot - 1;
            if (g->support_needs_grad[s] && g->d_supports[s])
              if (int rc = launch_dadj_big(cat + src * P * 32, dcat + slot * P * 32, g->d_supports[s], slabs, c->V, st))
                return rc;
            if (int rc = hop_any(c, g->hop_mats, g->ell, s, 1, dcat + slot * P * 32, dcat + src * P * 32, dcat + src * P * 32,
                                 slabs, st))
              return rc;
          }
        tc_done = true;
      } else if (use_tc_hops<T>(c, g->hop_mats)) {
        // (1) supports that need dA: g1' = g1 + g2 * A^T (in place in the y1 slot), then dA from (z, g1'), (y1, g2)
        for (int s = 0; s < c->n_supports; ++s) {
          if (!(g->support_needs_grad[s] && g->d_supports[s])) continue;
          HopParams p{};
          hop_params_base(p, c, dcat, g->hop_mats);
          p.n_mats = 1; p.mat_src[0] = 4 * s + 2; p.n_steps = 1; p.n_outs = 1;
          p.steps[0] = HopStep{0, 2 * s + 2, 0, 0, TH_LOAD | TH_RELEASE | TH_FIRST | TH_LAST};
          p.outs[0] = HopOut{0, 2 * s + 1, 0, 2 * s + 1};
          if (int rc = launch_hops_tc(p, st)) return rc;
          DadjParams dj{};
          dj.n_terms = 2; dj.V = c->V; dj.slabs = (int)slabs; dj.dA = g->d_supports[s];
          dj.t[0] = DadjTerm{cat + (2 * s + 1) * P * 32, dcat + (2 * s + 2) * P * 32};   // y1^T g2
          dj.t[1] = DadjTerm{cat, dcat + (2 * s + 1) * P * 32};                           // z^T  g1'
          if (int rc = launch_dadj_tc(dj, st)) return rc;
        }
        // (2) dz = g0 + sum_s [ g1_s A_s^T + g2_s (A_s^2)^T ]  (one accumulator; g1' A^T where step (1) ran)
        HopParams p{};
        hop_params_base(p, c, dcat, g->hop_mats);
        int n = 0;
        for (int s = 0; s < c->n_supports; ++s) {
          const bool folded = g->support_needs_grad[s] && g->d_supports[s];
          p.mat_src[n] = 4 * s + 2;
          p.steps[n] = HopStep{0, 2 * s + 1, n, 0, TH_LOAD | TH_RELEASE};
          ++n;
          if (!folded) {
            p.mat_src[n] = 4 * s + 3;
            p.steps[n] = HopStep{0, 2 * s + 2, n, 0, TH_LOAD | TH_RELEASE};
            ++n;
          }
        }
        p.steps[0].flags |= TH_FIRST;
        p.steps[n - 1].flags |= TH_LAST;
        p.n_mats = n; p.n_steps = n; p.n_outs = 1;
        p.outs[0] = HopOut{0, 0, 0, 0};
        if (int rc = launch_hops_tc(p, st)) return rc;
        tc_done = true;
      }
    }
    if (!tc_done)
    for (int s = 0; s < c->n_supports; ++s)
      for (int k = c->order; k >= 1; --k) {
        int slot = 1 + s * c->order + (k - 1);
        int src = (k == 1) ? 0 : slot - 1;
        if (g->support_needs_grad[s] && g->d_supports[s])
          if (int rc = launch_dadj<T>(cat + src * P * 32, 32, 0, dcat + slot * P * 32, 32, 0, g->d_supports[s], slabs,
                                      c->V, st))
            return rc;
        if (int rc = launch_node_mix<T>(dcat + slot * P * 32, 32, 0, dcat + src * P * 32, 32, 0, 1, g->supports[s], true,
                                        slabs, c->V, st))
          return rc;
      }
    dz = dcat;
  } else {
    if (!g->outputs_zeroed) GWN_CUDA(cudaMemsetAsync(g->dw_mlp, 0, sizeof(float) * 32 * mlp_in, st));
    if (!g->outputs_zeroed) GWN_CUDA(cudaMemsetAsync(g->db_mlp, 0, sizeof(float) * 32, st));
  }
  // gate backward: (dz + dz_last) -> dfg  (bf16 in tensor-core mode so it can be an MMA operand)
  bool tc_gate = false;
  if constexpr (std::is_same<T, bf16>::value) tc_gate = tc_mode<T>(c, g->hop_mats) != 0 && wgrad_tc_supported(c->taps, 64);
  const long long last_begin = (long long)(c->Lout - c->Lf) * c->V, last_rows = (long long)c->Lf * c->V;
  bf16* dfg16 = reinterpret_cast<bf16*>(g->ws_dfg);
  if (fused_bwd) {
  } else if (tc_gate)
    gate_bwd_kernel<T, bf16><<<eb, 256, 0, st>>>(dz, 32, reinterpret_cast<const T*>(g->dz_last), RO, last_begin,
                                                  last_rows, a, b, dfg16, P);
  else
    gate_bwd_kernel<T, float><<<eb, 256, 0, st>>>(dz, 32, reinterpret_cast<const T*>(g->dz_last), RO, last_begin,
                                                   last_rows, a, b, g->ws_dfg, P);
  GWN_LAUNCHED();
  // dW_fg [taps*32, 64], db_fg
  GemmA A{};
  A.n_chunks = c->taps; A.rows_per_n_out = RO; A.P = P;
  for (int j = 0; j < c->taps; ++j) {
    AChunk& ch = A.ch[j];
    ch.base = g->u_prev; ch.rows_per_n = RI; ch.row_off = (long long)j * c->dilation * c->V;
    ch.pitch = 32; ch.col_off = 0; ch.scale = g->scale; ch.shift = g->shift;
  }
  // k = 2 taps: the gate weight gradient is fused into the data-gradient GEMM below (same dfg / u_prev tiles)
  const bool fuse_wg = tc_gate && g->ws_w != nullptr && c->taps == 2 && fused_gate_wgrad_enabled();
  if (tc_gate) {
    if constexpr (std::is_same<T, bf16>::value) {
      if (!g->outputs_zeroed) GWN_CUDA(cudaMemsetAsync(g->dw_fg, 0, sizeof(float) * 64 * 32 * c->taps, st));
      if (!g->outputs_zeroed) GWN_CUDA(cudaMemsetAsync(g->db_fg, 0, sizeof(float) * 64, st));
      WgParams w{};
      w.n_chunks = c->taps; w.rows_per_n_out = RO; w.P = P;
      for (int j = 0; j < c->taps; ++j)
        w.ch[j] = WgChunk{reinterpret_cast<const bf16*>(g->u_prev), RI, (long long)j * c->dilation * c->V, 32, 0};
      w.G = dfg16; w.g_pitch = 64; w.N = 64; w.dW = g->dw_fg; w.ldw = 64; w.db = g->db_fg;
      w.scale = g->scale; w.shift = g->shift;
      if (!fuse_wg)
        if (int rc = launch_wgrad_tc(w, st)) return rc;
    }
  } else {
    if (int rc = launch_wgrad<T, float>(A, g->ws_dfg, 64, 0, g->dw_fg, 64, g->db_fg, st)) return rc;
  }
  // dx_prev[p_in, c] = sum_{j,fg} dfg[p_in - j*d*V, fg] * w_fg[(j,c), fg]  (+ residual du on cropped rows)
  GemmA X{};
  X.n_chunks = 2 * c->taps; X.rows_per_n_out = RI; X.P = PI;
  for (int j = 0; j < c->taps; ++j)
    for (int h = 0; h < 2; ++h) {
      AChunk& ch = X.ch[2 * j + h];
      ch.base = g->ws_dfg; ch.rows_per_n = RO; ch.row_off = -(long long)j * c->dilation * c->V;
      ch.pitch = 64; ch.col_off = h * 32; ch.w_off = (long long)j * 32 * 64 + h * 32;
    }
  if (!g->outputs_zeroed) GWN_CUDA(cudaMemsetAsync(g->dx_stats, 0, sizeof(double) * 64, st));
  EpiGateBwdData<T> ex{};
  ex.stats = g->dx_stats; ex.du = du; ex.du_rows_per_n = RO; ex.crop = (long long)(c->Lin - c->Lout) * c->V;
  ex.u_prev = reinterpret_cast<const T*>(g->u_prev); ex.dx = reinterpret_cast<float*>(g->dx_prev);
  GWN_REQUIRE(!g->dx_prev_bf16 || (tc_gate && g->ws_w != nullptr && 2 * c->taps <= PG_TC_MAX_CHUNKS),
              "layer_bwd: bf16 dx_prev needs the tensor-core data-gradient GEMM");
  if (tc_gate && g->ws_w != nullptr && 2 * c->taps <= PG_TC_MAX_CHUNKS) {
    if constexpr (std::is_same<T, bf16>::value) {
      // the (transposed) gate weight image of this GEMM is built by the kernel's own prologue (tc_gemm.cuh: PgWsrc)
      PgParams pg{};
      pg.n_chunks = 2 * c->taps; pg.rows_per_n_out = RI; pg.P = PI; pg.N = 32;
      pg.wsrc.W = g->w_fg; pg.wsrc.ld = 64; pg.wsrc.transposed = 1;
      for (int j = 0; j < c->taps; ++j)
        for (int h = 0; h < 2; ++h) pg.wsrc.w_off[2 * j + h] = j * 32 * 64 + h * 32;
      for (int j = 0; j < c->taps; ++j)
        for (int h = 0; h < 2; ++h)
          pg.ch[2 * j + h] = PgChunk{dfg16, RO, -(long long)j * c->dilation * c->V, 64, h * 32};
      // extra tiles for the epilogue: du shifted by the crop (zero outside), u_prev
      pg.n_extra = 2;
      // (no du for the last layer: a row offset far past the sample makes TMA zero-fill the whole tile)
      pg.ch[2 * c->taps] = du ? PgChunk{du, RO, -(long long)(c->Lin - c->Lout) * c->V, 32, 0}
                              : PgChunk{reinterpret_cast<const bf16*>(g->u_prev), RI, (long long)1 << 28, 32, 0};
      pg.ch[2 * c->taps + 1] = PgChunk{reinterpret_cast<const bf16*>(g->u_prev), RI, 0, 32, 0};
      if (fuse_wg) {
        EpiGateBwdTC<true> eb2{};
        eb2.dx = g->dx_prev_bf16 ? nullptr : reinterpret_cast<float*>(g->dx_prev);
        eb2.dx16 = g->dx_prev_bf16 ? reinterpret_cast<bf16*>(g->dx_prev) : nullptr;
        eb2.stats = g->dx_stats; eb2.s1[0] = eb2.s1[1] = eb2.s2[0] = eb2.s2[1] = 0.f;
        eb2.wg_scale = g->scale; eb2.wg_shift = g->shift; eb2.dw_fg = g->dw_fg; eb2.db_fg = g->db_fg;
        return launch_pos_gemm_tc(pg, eb2, st);
      }
      EpiGateBwdTC<false> eb2{};
      eb2.dx = g->dx_prev_bf16 ? nullptr : reinterpret_cast<float*>(g->dx_prev);
      eb2.dx16 = g->dx_prev_bf16 ? reinterpret_cast<bf16*>(g->dx_prev) : nullptr;
      eb2.stats = g->dx_stats; eb2.s1[0] = eb2.s1[1] = eb2.s2[0] = eb2.s2[1] = 0.f;
      eb2.wg_scale = nullptr; eb2.wg_shift = nullptr; eb2.dw_fg = nullptr; eb2.db_fg = nullptr;
      return launch_pos_gemm_tc(pg, eb2, st);
    }
  }
  if (tc_gate) return launch_pos_gemm_wt<bf16, 32>(X, g->w_fg, 32, 64, ex, st);
  return launch_pos_gemm_wt<float, 32>(X, g->w_fg, 32, 64, ex, st);
}

}  // namespace gwn

using namespace gwn;

extern "C" int gwn_layer_fwd(const gwn_layer_cfg* cfg, const gwn_layer_fwd_args* args, void* stream) {
  if (int rc = check_cfg(cfg)) return rc;
  GWN_REQUIRE(args && args->u_prev && args->w_fg && args->b_fg && args->z_last && args->ws_cat, "layer_fwd: NULL argument");
  GWN_REQUIRE(!cfg->has_gconv || (args->w_mlp && args->b_mlp && args->u && args->stats), "layer_fwd: NULL gconv argument");
  GWN_REQUIRE(!cfg->training || (args->a && args->b), "layer_fwd: training needs a,b buffers");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return cfg->dtype == GWN_F32 ? layer_fwd_t<float>(cfg, args, st) : layer_fwd_t<bf16>(cfg, args, st);
}

extern "C" int gwn_layer_bwd(const gwn_layer_cfg* cfg, const gwn_layer_bwd_args* args, void* stream) {
  if (int rc = check_cfg(cfg)) return rc;
  GWN_REQUIRE(args && args->u_prev && args->w_fg && args->a && args->b && args->dx_prev && args->dx_stats &&
                  args->dw_fg && args->db_fg && args->dw_mlp && args->db_mlp && args->ws_dfg,
              "layer_bwd: NULL argument");
  GWN_REQUIRE(args->du == nullptr || (args->ws_cat && args->ws_dcat && args->w_mlp), "layer_bwd: NULL workspace");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return cfg->dtype == GWN_F32 ? layer_bwd_t<float>(cfg, args, st) : layer_bwd_t<bf16>(cfg, args, st);
}

// The fused diffusion graph convolution alone (unit tests / roofline microbenchmark): z -> u (+ stats).
extern "C" int gwn_gcn_fwd(const void* z, const void* u_prev, const float* scale, const float* shift,
                           const void* hop_mats, int n_supports, const float* w_mlp, const float* b_mlp, void* ws_w,
                           float drop_p, unsigned long long seed, unsigned long long offset, void* u, double* stats,
                           int N, int V, int Lin, int Lout, void* stream) {
  GWN_REQUIRE(z && u_prev && hop_mats && w_mlp && b_mlp && ws_w && u && stats && n_supports >= 1 && Lout <= Lin,
              "gcn_fwd: bad argument");
  GWN_REQUIRE(gcn_fused_supported(V, 2 * n_supports), "gcn_fwd: V=%d with %d supports does not fit on chip", V, n_supports);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GWN_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 64, st));
  GcnFwdParams fp{};
  fp.z = reinterpret_cast<const bf16*>(z); fp.u_prev = reinterpret_cast<const bf16*>(u_prev);
  fp.RI = (long long)Lin * V; fp.RO = (long long)Lout * V; fp.crop = (long long)(Lin - Lout) * V;
  fp.scale = scale; fp.shift = shift;
  fp.mats = reinterpret_cast<const bf16*>(hop_mats); fp.n_mats = 2 * n_supports;
  for (int j = 0; j < fp.n_mats; ++j) fp.mat_src[j] = 4 * (j / 2) + (j % 2);
  fp.w_img = nullptr; fp.w_src = w_mlp; fp.bias = b_mlp; fp.mask = nullptr; fp.drop_p = drop_p; fp.seed = seed; fp.offset = offset;
  fp.rng = nullptr; fp.u = reinterpret_cast<bf16*>(u); fp.stats = stats; fp.V = V; fp.slabs = N * Lout;
  {  // debug timeline: GWN_GCN_TRACE=<device pointer of 64*8 int64> (scripts/gpu_gcn_trace.py)
    fp.trace = trace_ptr("GWN_GCN_TRACE");
  }
  if (gcn_t_enabled() && gcn_fused_t_supported(V, fp.n_mats)) {
    const int Kp = ((V + 15) / 16) * 16;
    fp.mats_t = reinterpret_cast<const bf16*>(hop_mats) + (size_t)n_supports * 4 * (Kp / 8) * 1024;
    return launch_gcn_fwd_t(fp, st);
  }
  return launch_gcn_fwd(fp, st);
}

// The fused diffusion backward alone (roofline microbenchmark): du, a, b -> dfg, dW_mlp, db_mlp, dA.
extern "C" int gwn_gcn_bwd(const void* du, const void* a, const void* b, const void* dz_last, const void* hop_mats,
                           int n_supports, const float* w_mlp, void* ws_w, float drop_p, unsigned long long seed,
                           unsigned long long offset, int sa, void* dfg, float* dw_mlp, float* db_mlp, float* dA,
                           int N, int V, int Lout, int Lf, void* stream) {
  GWN_REQUIRE(du && a && b && hop_mats && w_mlp && ws_w && dfg && dw_mlp && db_mlp && n_supports >= 1 && Lf >= 1 && Lf <= Lout,
              "gcn_bwd: bad argument");
  GWN_REQUIRE(sa < n_supports && (sa < 0 || dA != nullptr), "gcn_bwd: bad support-gradient argument");
  GWN_REQUIRE(gcn_bwd_fused_supported(V, 2 * n_supports), "gcn_bwd: V=%d with %d supports does not fit on chip", V, n_supports);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int mlp_in = 32 * (1 + 2 * n_supports);
  GWN_CUDA(cudaMemsetAsync(dw_mlp, 0, sizeof(float) * 32 * mlp_in, st));
  GWN_CUDA(cudaMemsetAsync(db_mlp, 0, sizeof(float) * 32, st));
  GcnBwdParams bp{};
  bp.du = reinterpret_cast<const bf16*>(du); bp.a = reinterpret_cast<const bf16*>(a); bp.b = reinterpret_cast<const bf16*>(b);
  bp.dz_last = reinterpret_cast<const bf16*>(dz_last);
  bp.RO = (long long)Lout * V; bp.last_begin = (long long)(Lout - Lf) * V; bp.last_rows = (long long)Lf * V;
  bp.mats = reinterpret_cast<const bf16*>(hop_mats); bp.n_mats = 2 * n_supports;
  for (int j = 0; j < bp.n_mats; ++j) bp.mat_src[j] = 4 * (j / 2) + 2 + (j % 2);
  bp.wt_img = nullptr; bp.w_src = w_mlp; bp.mask = nullptr; bp.drop_p = drop_p; bp.seed = seed; bp.offset = offset; bp.rng = nullptr;
  bp.dfg = reinterpret_cast<bf16*>(dfg); bp.dw_mlp = dw_mlp; bp.db_mlp = db_mlp; bp.V = V; bp.slabs = N * Lout;
  bp.sa = sa; bp.mat_fwd = sa >= 0 ? 4 * sa : 0; bp.w56_img = nullptr; bp.dA = dA; bp.trace = nullptr;
  return launch_gcn_bwd(bp, st);
}

extern "C" int gwn_dropout_apply(const void* x, void* out, long long rows, float p, unsigned long long seed,
                                 unsigned long long offset, void* stream) {
  GWN_REQUIRE(x && out && rows >= 0 && p >= 0.f && p < 1.f, "dropout_apply: bad argument");
  if (rows == 0) return 0;
  drop_bwd_kernel<bf16><<<(unsigned)cdiv(rows * 8, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(x), nullptr, p, seed, offset, nullptr, reinterpret_cast<bf16*>(out), rows);
  GWN_LAUNCHED();
  return 0;
}

extern "C" int gwn_gcn_bwd_t_supported(int V, int n_supports, int has_da) {
  return gcn_bwd_t_supported(V, 2 * n_supports, has_da != 0);
}

extern "C" int gwn_gcn_bwd_t(const void* du, const void* a, const void* b, const void* dz_last, const void* hop_mats,
                             int n_supports, const float* w_mlp, float drop_p, unsigned long long seed,
                             unsigned long long offset, int sa, void* dfg, float* dw_mlp, float* db_mlp, float* dA,
                             float* dQ6, int N, int V, int Lout, int Lf, void* stream) {
  GWN_REQUIRE(du && a && b && hop_mats && w_mlp && dfg && dw_mlp && db_mlp && n_supports >= 1 && Lf >= 1 && Lf <= Lout,
              "gcn_bwd_t: bad argument");
  GWN_REQUIRE(sa < n_supports && (sa < 0 || (dA != nullptr && dQ6 != nullptr)), "gcn_bwd_t: bad support-gradient argument");
  GWN_REQUIRE(gcn_bwd_t_supported(V, 2 * n_supports, sa >= 0), "gcn_bwd_t: no T-form instance for V=%d with %d supports", V,
              n_supports);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int mlp_in = 32 * (1 + 2 * n_supports);
  GWN_CUDA(cudaMemsetAsync(dw_mlp, 0, sizeof(float) * 32 * mlp_in, st));
  GWN_CUDA(cudaMemsetAsync(db_mlp, 0, sizeof(float) * 32, st));
  GcnBwdParams bp{};
  bp.du = reinterpret_cast<const bf16*>(du); bp.a = reinterpret_cast<const bf16*>(a); bp.b = reinterpret_cast<const bf16*>(b);
  bp.dz_last = reinterpret_cast<const bf16*>(dz_last);
  bp.RO = (long long)Lout * V; bp.last_begin = (long long)(Lout - Lf) * V; bp.last_rows = (long long)Lf * V;
  bp.mats = reinterpret_cast<const bf16*>(hop_mats); bp.n_mats = 2 * n_supports;
  bp.mats_bt = bp.mats + mats_bt_offset(V, n_supports);
  bp.w_src = w_mlp; bp.mask = nullptr; bp.drop_p = drop_p; bp.seed = seed; bp.offset = offset; bp.rng = nullptr;
  bp.dfg = reinterpret_cast<bf16*>(dfg); bp.dw_mlp = dw_mlp; bp.db_mlp = db_mlp; bp.V = V; bp.slabs = N * Lout;
  bp.sa = sa; bp.dA = dA; bp.dQ6 = dQ6; bp.trace = trace_ptr("GWN_GCN_TRACE");
  { const char* e = getenv("GWN_BT_DEBUG"); bp.debug = e ? atoi(e) : 0; }
  return launch_gcn_bwd_t(bp, st);
}

extern "C" int gwn_node_mix(const void* x, int x_pitch, int x_off, void* y, int y_pitch, int y_off, int accumulate,
                            const float* A, int transpose_a, int slabs, int V, int dtype, void* stream) {
  GWN_REQUIRE(x && y && A && slabs >= 0 && V >= 1, "node_mix: bad argument");
  GWN_REQUIRE(x_pitch % 4 == 0 && y_pitch % 4 == 0 && x_off % 4 == 0 && y_off % 4 == 0, "node_mix: unaligned pitch");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == GWN_F32)
    return launch_node_mix<float>((const float*)x, x_pitch, x_off, (float*)y, y_pitch, y_off, accumulate, A,
                                  transpose_a != 0, slabs, V, st);
  GWN_REQUIRE(dtype == GWN_BF16, "bad dtype %d", dtype);
  return launch_node_mix<bf16>((const bf16*)x, x_pitch, x_off, (bf16*)y, y_pitch, y_off, accumulate, A,
                               transpose_a != 0, slabs, V, st);
}

extern "C" int gwn_bn_fold(const double* stats, double count, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, float momentum, float eps, int training,
                           float* scale, float* shift, float* mean, float* rstd, void* stream) {
  GWN_REQUIRE(gamma && beta && running_mean && running_var && scale && shift, "bn_fold: NULL argument");
  GWN_REQUIRE(!training || stats, "bn_fold: training needs stats");
  bn_fold_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(stats, count, gamma, beta, running_mean,
                                                                       running_var, momentum, eps, training, scale,
                                                                       shift, mean, rstd);
  GWN_LAUNCHED();
  return 0;
}

extern "C" int gwn_bn_bwd(const void* dx, int dx_dtype, const void* u, int dtype, const double* dx_stats, double count,
                          const float* gamma, const float* mean, const float* rstd, int training, void* du,
                          float* dgamma, float* dbeta, long long rows, void* stream) {
  GWN_REQUIRE(dx && u && dx_stats && gamma && mean && rstd && du && dgamma && dbeta, "bn_bwd: NULL argument");
  GWN_REQUIRE((dx_dtype == GWN_F32 || dx_dtype == GWN_BF16) && (dtype == GWN_F32 || dtype == GWN_BF16) &&
                  !(dx_dtype == GWN_BF16 && dtype == GWN_F32), "bn_bwd: bad dtype combination");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned blocks = (unsigned)cdiv(rows * 4, 256);
  if (blocks == 0) blocks = 1;
  if (dtype == GWN_F32)
    GWN_CUDA(launch_pdl(bn_bwd_kernel<float, float>, dim3(blocks), dim3(256), 0, st, (const float*)dx, (const float*)u, dx_stats, count, gamma, mean, rstd,
                                                        training, (float*)du, dgamma, dbeta, rows));
  else if (dx_dtype == GWN_F32)
    GWN_CUDA(launch_pdl(bn_bwd_kernel<bf16, float>, dim3(blocks), dim3(256), 0, st, (const float*)dx, (const bf16*)u, dx_stats, count, gamma, mean, rstd,
                                                       training, (bf16*)du, dgamma, dbeta, rows));
  else
    GWN_CUDA(launch_pdl(bn_bwd_kernel<bf16, bf16>, dim3(blocks), dim3(256), 0, st, (const bf16*)dx, (const bf16*)u, dx_stats, count, gamma, mean, rstd,
                                                      training, (bf16*)du, dgamma, dbeta, rows));
  GWN_LAUNCHED();
  return 0;
}
