// The data-parallel exchange step fused with its consumer (lit.py:59-61 = Adam(lr=1e-3)):
//
//   one kernel = all-reduce (average) of the flat fp32 gradient over NVLink 5 / NVSwitch PEER MEMORY + the Adam update
//
// The reference has no distributed code at all; batch data parallelism needs exactly one exchange per
// step: the 1.2-2.1 MB flat gradient.  At that size a ring / tree collective is pure latency (NCCL: ~50 us at 8 GPUs,
// followed by three multi-tensor Adam launches over 108 small tensors); here every rank keeps its gradient in an
// exchange block that all peers have mapped (CUDA IPC), and ONE kernel per rank
//   1. signals "my gradient is complete" into every peer's flag row and waits for all peers' signals   (start barrier),
//   2. reads the same slice of all `world` gradients straight over NVSwitch (one-shot all-reduce: every rank sums
//      all ranks' values in the SAME order -> bit-identical replicas), averages,
//   3. signals "done reading" (so a peer may overwrite its gradient next step) and
//   4. applies Adam to its parameters (flat fp32 p / m / v, bias corrections from a device-side step counter:
//      CUDA-graph replays advance it),  5. waits for the peers' "done reading" signals                    (end barrier).
// world == 1 runs steps 2 (local) and 4 only: one launch instead of torch's three multi-tensor launches.
//
// Exchange block of a rank (one cudaMalloc, exported with cudaIpcGetMemHandle):
//   [0, 64)    start flags, one uint64 per source rank      [128, 192)  end flags, one uint64 per source rank
//   [256, ...) the flat gradient (fp32)
// Flags carry the step number (monotonic), so they are never reset.  Every poll is bounded: a missing peer traps the
// kernel instead of hanging the GPU.
#include "common.cuh"

namespace gwn {

constexpr int PEER_MAX = 8;
constexpr int PEER_HDR = 256;            // bytes in front of the gradient
constexpr int ADAM_THREADS = 256;
constexpr int ADAM_MAXIT = 8;            // float4 per thread

struct PeerTable {
  const float* grads[PEER_MAX];
  unsigned long long* flags[PEER_MAX];   // base of every rank's exchange block
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_flag(const unsigned long long* p, unsigned long long want) {
  for (unsigned long long i = 0; i < (1ull << 25); ++i) {          // ~30 s of polling, then trap
    if (ld_acquire_sys(p) >= want) return;
    if (i > 1024) __nanosleep(64);
  }
  __trap();    // a peer never arrived (dead rank / mismatched step count): fail loudly, never hang
}

// state[0] = completed steps, state[1] = CTA completion counter of the running launch
template <bool MULTI>
__global__ void __launch_bounds__(ADAM_THREADS) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g_local,
                                                                 float* __restrict__ m, float* __restrict__ v, long long n,
                                                                 float lr, float b1, float b2, float eps,
                                                                 unsigned long long* __restrict__ state, PeerTable pt,
                                                                 int rank, int world) {
  pdl_wait();      // programmatic launch: the launch latency overlaps the predecessor (common.cuh)
  pdl_trigger();
  __shared__ unsigned long long s_step;
  __shared__ int s_last;
  __shared__ float s_coef[2];
  const int tid = threadIdx.x;
  if (tid == 0) {
    s_step = *reinterpret_cast<volatile unsigned long long*>(state) + 1ull;
    const double bc1 = 1.0 - pow((double)b1, (double)s_step), bc2 = 1.0 - pow((double)b2, (double)s_step);
    s_coef[0] = (float)((double)lr / bc1);          // step size
    s_coef[1] = (float)(1.0 / sqrt(bc2));           // 1 / sqrt(bias correction 2)
  }
  __syncthreads();
  const unsigned long long step = s_step;
  if (MULTI) {
    // ---- start barrier: every rank's gradient is complete (written by earlier kernels of its stream)
    if (blockIdx.x == 0 && tid < world) {
      __threadfence_system();
      st_release_sys(pt.flags[tid] + rank, step);
    }
    if (tid < world) wait_flag(pt.flags[rank] + tid, step);
    __syncthreads();
  }
  // ---- gather + average (every rank adds the ranks' values in rank order: replicas stay bit-identical)
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * ADAM_THREADS;
  float4 acc[ADAM_MAXIT];
#pragma unroll
  for (int it = 0; it < ADAM_MAXIT; ++it) {
    const long long i = (long long)blockIdx.x * ADAM_THREADS + tid + it * stride;
    acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
      if (MULTI) {
        float4 part[PEER_MAX];
#pragma unroll
        for (int r = 0; r < PEER_MAX; ++r)
          if (r < world) part[r] = __ldcv(reinterpret_cast<const float4*>(pt.grads[r]) + i);   // never through a stale L1 line
#pragma unroll
        for (int r = 0; r < PEER_MAX; ++r)
          if (r < world) { acc[it].x += part[r].x; acc[it].y += part[r].y; acc[it].z += part[r].z; acc[it].w += part[r].w; }
        const float inv = 1.f / (float)world;
        acc[it].x *= inv; acc[it].y *= inv; acc[it].z *= inv; acc[it].w *= inv;
      } else {
        acc[it] = *(reinterpret_cast<const float4*>(g_local) + i);
      }
    }
  }
  float tail[3] = {0.f, 0.f, 0.f};
  const bool tail_owner = blockIdx.x == 0 && tid < (int)(n & 3);
  if (tail_owner) {
    const long long i = (n4 << 2) + tid;
    if (MULTI) {
      float s = 0.f;
      for (int r = 0; r < world; ++r) s += __ldcv(pt.grads[r] + i);
      tail[0] = s / (float)world;
    } else {
      tail[0] = g_local[i];
    }
  }
  // ---- this CTA is done reading; the last CTA of the launch tells the peers and publishes the step count
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned long long old = atomicAdd(state + 1, 1ull);
    s_last = (old == (unsigned long long)gridDim.x - 1ull) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    if (tid == 0) { state[1] = 0ull; state[0] = step; }
    if (MULTI && tid < world) st_release_sys(pt.flags[tid] + 16 + rank, step);
  }
  // ---- Adam (torch.optim.Adam, amsgrad off, no weight decay): bias corrections from the step count
  const float step_size = s_coef[0], rs2 = s_coef[1];
  auto upd = [&](float gv, float& pv, float& mv, float& vv) {
    mv = mv + (gv - mv) * (1.f - b1);
    vv = vv * b2 + (1.f - b2) * gv * gv;
    pv = pv - step_size * (mv / (sqrtf(vv) * rs2 + eps));
  };
#pragma unroll
  for (int it = 0; it < ADAM_MAXIT; ++it) {
    const long long i = (long long)blockIdx.x * ADAM_THREADS + tid + it * stride;
    if (i < n4) {
      float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      upd(acc[it].x, pv.x, mv.x, vv.x); upd(acc[it].y, pv.y, mv.y, vv.y);
      upd(acc[it].z, pv.z, mv.z, vv.z); upd(acc[it].w, pv.w, mv.w, vv.w);
      reinterpret_cast<float4*>(p)[i] = pv; reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv;
    }
  }
  if (tail_owner) {
    const long long i = (n4 << 2) + tid;
    upd(tail[0], p[i], m[i], v[i]);
  }
  // ---- end barrier: nobody overwrites its gradient before every peer has read it
  if (MULTI) {
    if (tid < world) wait_flag(pt.flags[rank] + 16 + tid, step);
  }
}

// ---- gather of the per-parameter gradients into the flat gradient: ONE launch, pointer table passed by value
constexpr int GATHER_MAX = 160;
struct GatherTable {
  const float* src[GATHER_MAX];
  int off[GATHER_MAX + 1];       // element offsets in the flat buffer (prefix sums), off[n] = total
  int n;
};
__global__ void __launch_bounds__(256) gather_flat_kernel(const __grid_constant__ GatherTable t, float* __restrict__ out) {
  pdl_wait();      // programmatic launch: the launch latency overlaps the predecessor (common.cuh)
  pdl_trigger();
  // element-parallel over the whole flat buffer (tensors range from 32 to 131072 elements: a per-tensor split would
  // leave most CTAs idle); the owning tensor of an element is found by bisection in the offset table (constant bank)
  const int total = t.off[t.n];
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    int lo = 0, hi = t.n;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (i >= t.off[mid]) lo = mid; else hi = mid;
    }
    const float* s = t.src[lo];
    const int j = i - t.off[lo];
    if (s == nullptr) out[i] = 0.f;
    else if (s + j != out + i) out[i] = __ldg(s + j);      // (src == dst: the gradient is already in place)
  }
}

}  // namespace gwn

using namespace gwn;

extern "C" int gwn_gather_flat(const void* const* srcs, const long long* counts, int n, float* out, void* stream) {
  GWN_REQUIRE(srcs && counts && out && n >= 1 && n <= GATHER_MAX, "gather_flat: 1..%d tensors", GATHER_MAX);
  GatherTable t{};
  long long off = 0;
  for (int k = 0; k < n; ++k) {
    GWN_REQUIRE(counts[k] >= 0 && off + counts[k] < (1ll << 31), "gather_flat: sizes out of range");
    t.src[k] = reinterpret_cast<const float*>(srcs[k]);
    t.off[k] = (int)off;
    off += counts[k];
  }
  t.off[n] = (int)off;
  t.n = n;
  const long long blocks = cdiv(off, 256 * 4);
  GWN_CUDA(launch_pdl(gather_flat_kernel, dim3((unsigned)(blocks < 1 ? 1 : blocks > 592 ? 592 : blocks)), dim3(256), 0, (cudaStream_t)stream, t, out));
  GWN_LAUNCHED();
  return 0;
}

extern "C" long long gwn_peer_header_bytes(void) { return PEER_HDR; }

extern "C" int gwn_peer_alloc(long long bytes, void** out) {
  GWN_REQUIRE(out && bytes > 0, "peer_alloc: bad argument");
  if (int rc = gwn_check_device()) return rc;
  void* p = nullptr;
  GWN_CUDA(cudaMalloc(&p, (size_t)bytes + PEER_HDR));
  GWN_CUDA(cudaMemset(p, 0, (size_t)bytes + PEER_HDR));
  GWN_CUDA(cudaDeviceSynchronize());
  *out = p;
  return 0;
}

extern "C" int gwn_peer_free(void* block) {
  if (block) GWN_CUDA(cudaFree(block));
  return 0;
}

extern "C" int gwn_peer_export(const void* block, void* handle64) {
  GWN_REQUIRE(block && handle64, "peer_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  GWN_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), const_cast<void*>(block)));
  return 0;
}

extern "C" int gwn_peer_open(const void* handle64, void** out) {
  GWN_REQUIRE(handle64 && out, "peer_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  GWN_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int gwn_peer_close(void* mapped) {
  if (mapped) GWN_CUDA(cudaIpcCloseMemHandle(mapped));
  return 0;
}

static int adam_grid(long long n) {
  long long g = cdiv(cdiv(n, 4), (long long)ADAM_THREADS * 4);          // ~4 float4 per thread
  if (g < 1) g = 1;
  if (g > 128) g = 128;                                                  // co-resident with room to spare (the CTAs spin)
  return (int)g;
}

extern "C" int gwn_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
                             float eps, void* state, void* stream) {
  GWN_REQUIRE(p && g && m && v && state && n > 0, "adam_flat: bad argument");
  const int grid = adam_grid(n);
  GWN_REQUIRE(cdiv(n, 4) <= (long long)grid * ADAM_THREADS * ADAM_MAXIT, "adam_flat: %lld parameters exceed one launch", n);
  PeerTable pt{};
  GWN_CUDA(launch_pdl(adam_flat_kernel<false>, dim3(grid), dim3(ADAM_THREADS), 0, (cudaStream_t)stream, p, g, m, v, n, lr, b1, b2, eps,
                                                                          reinterpret_cast<unsigned long long*>(state), pt, 0, 1));
  GWN_LAUNCHED();
  return 0;
}

extern "C" int gwn_allreduce_adam(float* p, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                                  void* state, const void* const* blocks, int rank, int world, void* stream) {
  GWN_REQUIRE(p && m && v && state && blocks && n > 0 && world >= 2 && world <= PEER_MAX && rank >= 0 && rank < world,
              "allreduce_adam: bad argument");
  const int grid = adam_grid(n);
  GWN_REQUIRE(cdiv(n, 4) <= (long long)grid * ADAM_THREADS * ADAM_MAXIT, "allreduce_adam: %lld parameters exceed one launch", n);
  PeerTable pt{};
  for (int r = 0; r < world; ++r) {
    GWN_REQUIRE(blocks[r] != nullptr, "allreduce_adam: exchange block of rank %d missing", r);
    pt.flags[r] = reinterpret_cast<unsigned long long*>(const_cast<void*>(blocks[r]));
    pt.grads[r] = reinterpret_cast<const float*>(reinterpret_cast<const char*>(blocks[r]) + PEER_HDR);
  }
  GWN_CUDA(launch_pdl(adam_flat_kernel<true>, dim3(grid), dim3(ADAM_THREADS), 0, (cudaStream_t)stream, p, pt.grads[rank], m, v, n, lr, b1, b2, eps,
                                                                         reinterpret_cast<unsigned long long*>(state), pt, rank,
                                                                         world));
  GWN_LAUNCHED();
  return 0;
}
