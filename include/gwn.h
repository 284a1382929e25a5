/*
 * libgwn — C ABI of the B200-native (sm_100a) Graph WaveNet block.
 *
 * Drop-in boundary: the reference has no FFI layer; the unit being replaced is the
 * `gwnet` nn.Module of models/graph_wavenet.py:100-256.  Each entry
 * point below names the reference lines whose arithmetic it replaces.  The Python
 * host (multimodal_outage_b200/ops.py) binds these with ctypes and registers them as
 * torch.library custom ops; INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless noted;
 *  - the caller owns all memory (outputs, saved tensors, workspaces);
 *  - `stream` is a cudaStream_t passed as void*; entry points are re-entrant (the
 *    backward runs on an autograd worker thread);
 *  - return 0 on success, <0 on error; gwn_last_error() returns a thread-local message;
 *  - sm_100 only: any other device is an error, there is no fallback;
 *  - activations inside the block are "channels-last": [N, L, V, 32] (batch, time,
 *    node, channel), dtype GWN_F32 or GWN_BF16; parameters, statistics and gradients
 *    of parameters are always fp32.  The channel width (residual = dilation channels)
 *    is fixed at 32, the value every reference/BASELINE configuration uses
 *    (graph_wavenet.py:101).
 */
#ifndef GWN_H
#define GWN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GWN_C 32            /* residual_channels == dilation_channels */
#define GWN_MAX_SUPPORTS 4  /* fixed supports + adaptive */
#define GWN_MAX_TAPS 8      /* temporal kernel_size */
#define GWN_MAX_LAYERS 32

enum { GWN_F32 = 0, GWN_BF16 = 1 };

const char* gwn_last_error(void);
int gwn_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches evidence) */
long long gwn_launch_count(void);
/* 0 if the current device is sm_100 (B200); <0 otherwise. */
int gwn_check_device(void);

/* ---- adaptive adjacency: softmax(relu(E1 @ E2), dim=1)   graph_wavenet.py:202 ---- */
/* e1 [V,R], e2 [R,V], adp [V,V] fp32 row-major (adp[v,w]); adp_t optional transpose copy. */
int gwn_adp_fwd(const float* e1, const float* e2, float* adp, float* adp_t, int V, int R, void* stream);
/* d_adp [V,V] -> d_e1 [V,R], d_e2 [R,V] (overwritten). ws: V*V floats. */
int gwn_adp_bwd(const float* e1, const float* e2, const float* adp, const float* d_adp,
                float* d_e1, float* d_e2, float* ws, int V, int R, void* stream);
/* The adaptive adjacency as a PAIR [2][V][V] (two identical copies; copy 0 is what the layer kernels read as the
 * support).  Its gradient comes back from gwn_layer_bwd as (d_supports, d_supports_sq) = (d0, Q), laid out [2][V][V];
 * gwn_adp_pair_bwd completes it in fp32 (d_adp = d0 + A^T Q + Q A^T, the A*A hop of graph_wavenet.py:91-93 is
 * linear in Q) and then runs gwn_adp_bwd.  ws: [2][V][V] scratch. */
int gwn_adp_fwd_pair(const float* e1, const float* e2, float* pair, int V, int R, void* stream);
int gwn_adp_pair_bwd(const float* e1, const float* e2, const float* adp, const float* d_pair, float* d_e1, float* d_e2,
                     float* ws, int V, int R, void* stream);

/* ---- start_conv (1x1, Cin->32) + left zero pad + NCHW -> channels-last
 *      graph_wavenet.py:191-196 ---- */
/* x [N,Cin,V,T] fp32 NCHW; w [32,Cin]; b [32]; u0 [N,L0,V,32] (dtype), L0 >= T. */
int gwn_start_fwd(const float* x, const float* w, const float* b, void* u0, int dtype,
                  int N, int Cin, int V, int T, int L0, void* stream);
/* du0 [N,L0,V,32] -> dw [32,Cin], db [32] (overwritten), dx [N,Cin,V,T] (optional, may be NULL). */
int gwn_start_bwd(const float* x, const float* w, const void* du0, int dtype, float* dw, float* db,
                  float* dx, int N, int Cin, int V, int T, int L0, void* stream);

/* start_conv on the tensor cores for wide inputs (Cin >= 64, Cin % 8 == 0: config 4's 256 UNet features + 64
 * date2vec; csrc/start_tc.cu).  Forward also writes xcl = the input in channels-last bf16 [N,L0,V,Cin] (time
 * left-padded), which the backward consumes.  ws_w: >= 128*Cin bytes (bf16 W and W^T), the SAME buffer in fwd and bwd.
 * ws_dx: >= N*L0*V*Cin*2 bytes, only when dx != NULL. */
int gwn_start_tc_supported(int Cin);
int gwn_start_fwd_tc(const float* x, const float* w, const float* b, void* xcl, void* u0, void* ws_w, int N,
                     int Cin, int V, int T, int L0, void* stream);
int gwn_start_bwd_tc(const void* xcl, const void* du0, void* ws_w, float* dw, float* db, float* dx, void* ws_dx,
                     int N, int Cin, int V, int T, int L0, void* stream);

/* ---- one WaveNet layer  graph_wavenet.py:206-250 ----
 * forward:  r = bn_prev(u_prev)  (folded affine: r = u_prev*scale + shift, NULL = identity)
 *           (f,g) = dilated conv k taps  (:222,224)      z = tanh(f)*sigmoid(g)  (:223-226)
 *           z_last = z[:, -Lf:]  (the only part of z the skip path keeps, :231-236)
 *           h = mlp(concat[z, zA0, zA0^2, ...]) + bm     (gcn, :85-96; nconv :65)
 *           h = dropout(h)                               (:97)
 *           u = h + r[:, -Lout:]                         (:247)
 *           stats = per-channel (sum u, sum u^2)         (training BatchNorm, :250)
 */
typedef struct {
  int N, V, Lin, Lout, Lf;      /* Lout = Lin - dilation*(taps-1); Lf = final time length */
  int taps, dilation;
  int n_supports, order;        /* n_supports = 0 with mlp_in = 32: gcn_bool=False path (:245) */
  int dtype;                    /* GWN_F32 / GWN_BF16: storage of activations */
  int training;                 /* save a,b for backward */
  int has_gconv;                /* 0: stop after z/z_last (eval-mode last layer) */
  float dropout_p;              /* 0 disables */
  uint64_t seed, offset;        /* Philox key/offset for the fused dropout mask */
} gwn_layer_cfg;

/* Optional SPARSE form of a fixed support.  County adjacency graphs have a handful of neighbours per node and asym_adj
 * (utils.py:152-158) keeps that pattern, so at V = 3100 a dense V x V hop multiplies 99.7 % exact zeros.  With width > 0 the
 * bf16 big-graph path (V > 80) applies the support with a gather kernel over ELL rows instead of a dense tensor-core GEMM:
 * the same sums with the zero terms skipped (values in fp32).  The adaptive adjacency (a dense softmax) stays a GEMM. */
typedef struct {
  const int* idx[2];            /* [V][width] int32, -1 pads.  which = 0: neighbours v of column w (y[w] = sum_v A[v,w] x[v], the
                                   forward hop); which = 1: neighbours v of row w (y[w] = sum_v A[w,v] x[v], its transpose) */
  const float* val[2];          /* [V][width] fp32 */
  int width;                    /* 0 = dense */
} gwn_ell;

typedef struct {
  const void* u_prev;           /* [N,Lin,V,32] */
  const float* scale;           /* [32] or NULL */
  const float* shift;           /* [32] or NULL */
  const float* w_fg;            /* packed [taps*32, 64]: row (j,c), col 2o = filter, 2o+1 = gate */
  const float* b_fg;            /* [64] interleaved */
  const float* w_mlp;           /* packed [mlp_in, 32] = mlp.weight[:, :, 0, 0]^T, mlp_in = 32*(1+order*n_supports) */
  const float* b_mlp;           /* [32] */
  const float* supports[GWN_MAX_SUPPORTS];   /* each [V,V] fp32, A[v,w] */
  const void* drop_mask;        /* optional explicit mask [N,Lout,V,32] (dtype); overrides Philox */
  const uint64_t* rng;          /* optional DEVICE {seed, offset}: overrides cfg.seed, added to cfg.offset
                                   (keeps the mask fresh across CUDA-graph replays) */
  const void* hop_mats;         /* optional gwn_hop_mats_prep images: bf16 hops run on tcgen05 */
  void* ws_w;                   /* optional 128 KiB scratch for the bf16 weight images of the tcgen05 GEMMs */
  void* a;                      /* out (training) tanh(f)   [N,Lout,V,32] */
  void* b;                      /* out (training) sigmoid(g) */
  void* z_last;                 /* out [N,Lf,V,32] */
  void* u;                      /* out [N,Lout,V,32] */
  double* stats;                /* out [2,32]: sum, sum of squares (zeroed by callee) */
  void* ws_cat;                 /* workspace [N*Lout*V, 32*(1+order*n_supports)] (dtype) */
  /* Optional (bf16 tensor-core path only): BatchNorm of the previous layer (graph_wavenet.py:250) folded INSIDE the
   * gate kernel's prologue - no separate fold / weight-prep launches.  When bn_gamma != NULL, `scale` and `shift`
   * above are OUTPUTS ([32] each, written by the kernel) together with bn_mean / bn_rstd; with cfg.training the batch
   * statistics come from bn_stats ([2,32] sum, sum of squares over bn_count elements per channel) and the running
   * statistics are updated (momentum, unbiased variance); otherwise the running statistics are used. */
  const double* bn_stats; const float* bn_gamma; const float* bn_beta;
  float* bn_running_mean; float* bn_running_var;
  float* bn_mean; float* bn_rstd;
  double bn_count;
  float bn_momentum, bn_eps;
  const gwn_ell* ell;           /* optional HOST array [GWN_MAX_SUPPORTS] (see gwn_ell), NULL = all supports dense */
} gwn_layer_fwd_args;

int gwn_layer_fwd(const gwn_layer_cfg* cfg, const gwn_layer_fwd_args* args, void* stream);

typedef struct {
  /* saved from forward */
  const void* u_prev; const float* scale; const float* shift;
  const float* w_fg; const float* w_mlp;
  const float* supports[GWN_MAX_SUPPORTS];
  int support_needs_grad[GWN_MAX_SUPPORTS];
  const void* drop_mask;
  const uint64_t* rng;
  const void* hop_mats;
  void* ws_w;
  const void* a; const void* b;
  /* incoming gradients */
  const void* du;               /* [N,Lout,V,32] (dtype) or NULL (dead gconv: last layer) */
  const void* dz_last;          /* [N,Lf,V,32] (dtype) or NULL */
  /* outputs */
  void* dx_prev;                /* [N,Lin,V,32] grad wrt r = bn_prev(u_prev): fp32, or bf16 when dx_prev_bf16 != 0 */
  double* dx_stats;             /* [2,32]: sum dx, sum dx*u_prev (zeroed by callee) */
  float* dw_fg; float* db_fg;   /* [taps*32,64], [64] (overwritten) */
  float* dw_mlp; float* db_mlp; /* packed [mlp_in, 32], [32] (overwritten) */
  float* d_supports[GWN_MAX_SUPPORTS];   /* [V,V] fp32, ACCUMULATED into (caller zeroes) */
  /* workspaces */
  void* ws_cat;                 /* [P, mlp_in] (dtype): recomputed hops */
  void* ws_dcat;                /* [P, mlp_in] (dtype): grads of the concat */
  float* ws_dfg;                /* [P, 64] fp32 */
  int outputs_zeroed;           /* != 0: the caller already zeroed dx_stats, dw_*, db_* (one fill instead of six memsets) */
  int dx_prev_bf16;             /* != 0 (bf16 tensor-core path only): dx_prev is stored as bf16 - it only lives until the
                                   BatchNorm backward reads it; its statistics are taken from the fp32 accumulator */
  float* d_supports_sq[GWN_MAX_SUPPORTS];
                                /* optional [V,V] fp32, ACCUMULATED into (caller zeroes): when non-NULL for a support
                                   with support_needs_grad, a kernel path MAY leave the gradient that reaches the
                                   support through its SECOND-order hop A*A in factored form: with Q = d_supports_sq[i]
                                   the full gradient is d_supports[i] + A^T Q + Q A^T (applied once per step by
                                   gwn_adp_pair_bwd, in fp32).  Paths that do not use it leave Q = 0. */
  const gwn_ell* ell;           /* optional HOST array [GWN_MAX_SUPPORTS] (see gwn_ell), NULL = all supports dense */
} gwn_layer_bwd_args;

int gwn_layer_bwd(const gwn_layer_cfg* cfg, const gwn_layer_bwd_args* args, void* stream);

/* The fused diffusion graph convolution of gwn_layer_fwd alone (bf16, supports resident on chip; csrc/gcn_fused.cu):
 *   u = dropout(mlp(concat[z, z A_s, z A_s^2 ...]) + b) + (u_prev*scale + shift)[crop],  stats = (sum u, sum u^2).
 * z [N,Lout,V,32], u_prev [N,Lin,V,32], u [N,Lout,V,32] bf16; hop_mats from gwn_hop_mats_prep; w_mlp packed
 * [32*(1+2*n_supports), 32]; ws_w >= 16 KiB scratch.  graph_wavenet.py:76-98, :247, :250. */
int gwn_gcn_fwd(const void* z, const void* u_prev, const float* scale, const float* shift,
                const void* hop_mats, int n_supports, const float* w_mlp, const float* b_mlp, void* ws_w,
                float drop_p, unsigned long long seed, unsigned long long offset, void* u, double* stats,
                int N, int V, int Lin, int Lout, void* stream);

/* The fused diffusion BACKWARD of gwn_layer_bwd alone (bf16, supports on chip; csrc/gcn_fused_bwd.cu):
 *   dh = du.mask;  dU_j = M_j^T dh;  dz = sum_j dU_j W_j^T (+ dz_last);  dW_j = z^T dU_j (z = a.b);  db = sum dh;
 *   dfg = gate backward of dz (df = dz b (1-a^2), dg = dz a b (1-b), interleaved, bf16 [N*Lout*V, 64]);
 *   sa >= 0: dA[V,V] += gradient wrt support `sa` (the adaptive adjacency).  dw_mlp/db_mlp are overwritten, dA accumulated.
 * ws_w >= 24 KiB scratch.  graph_wavenet.py:76-98 backward + :222-226 backward. */
/* as gwn_gcn_bwd, through the transposed ("T-form") kernel (csrc/gcn_fused_bwd_t.cu); dQ6 as d_supports_sq above
 * (required when sa >= 0).  Returns -1 when the shape has no T-form instance (gwn_gcn_bwd_t_supported). */
int gwn_gcn_bwd_t_supported(int V, int n_supports, int has_da);
int gwn_gcn_bwd_t(const void* du, const void* a, const void* b, const void* dz_last, const void* hop_mats,
                  int n_supports, const float* w_mlp, float drop_p, unsigned long long seed, unsigned long long offset,
                  int sa, void* dfg, float* dw_mlp, float* db_mlp, float* dA, float* dQ6, int N, int V, int Lout, int Lf,
                  void* stream);
int gwn_gcn_bwd(const void* du, const void* a, const void* b, const void* dz_last, const void* hop_mats,
                int n_supports, const float* w_mlp, void* ws_w, float drop_p, unsigned long long seed,
                unsigned long long offset, int sa, void* dfg, float* dw_mlp, float* db_mlp, float* dA,
                int N, int V, int Lout, int Lf, void* stream);

/* The fused dropout stream on its own (graph_wavenet.py:97, F.dropout(h, p, training)): out = x * mask over `rows` rows of
 * 32 bf16 channels, mask = the Philox4x32-7 keep-mask {seed, offset, element} every fused kernel draws (0 or
 * 256 / (256 - round(256 p))).  Used by the statistical tests of the stream; the layer kernels apply it in their epilogues. */
int gwn_dropout_apply(const void* x, void* out, long long rows, float p, unsigned long long seed,
                      unsigned long long offset, void* stream);

/* One sparse hop on its own (tests / microbenchmarks): y[s,w,:] = sum_k val[w][k] x[s, idx[w][k], :] (+ add), bf16 slabs of
 * V rows x 32 channels. */
int gwn_hop_ell(const int* idx, const float* val, int width, const void* x, void* y, const void* add, long long slabs,
                int V, void* stream);

/* ---- BatchNorm2d(32) folded to an affine  graph_wavenet.py:167,250 ----
 * training: mean/var from stats (count = N*L*V), scale = gamma*rstd, shift = beta - mean*scale,
 *           running <- (1-m)*running + m*(mean, unbiased var); saves mean,rstd.
 * eval:     scale/shift from the running statistics. */
int gwn_bn_fold(const double* stats, double count, const float* gamma, const float* beta,
                float* running_mean, float* running_var, float momentum, float eps, int training,
                float* scale, float* shift, float* mean, float* rstd, void* stream);
/* du = BN backward of dx (grad wrt BN output) given u, mean, rstd and dx_stats=(sum dx, sum dx*u).
 * training=0: du = dx*scale.  dgamma,dbeta [32] overwritten.  du has dtype `dtype`. */
int gwn_bn_bwd(const void* dx, int dx_dtype, const void* u, int dtype, const double* dx_stats, double count,
               const float* gamma, const float* mean, const float* rstd, int training,
               void* du, float* dgamma, float* dbeta, long long rows, void* stream);

/* ---- head: relu(sum_i Ws_i z_last_i + bs) -> relu(end_conv_1) -> end_conv_2 -> NCHW
 *      graph_wavenet.py:231-236 (skip identity), :252-254 ---- */
typedef struct {
  int N, V, Lf, n_layers, S, E, O;   /* skip, end, out channels; S,E multiples of 32 */
  int dtype;
} gwn_head_cfg;
typedef struct {
  const void* z_last[GWN_MAX_LAYERS];  /* each [N,Lf,V,32] */
  const float* w_skip;    /* packed [32*n_layers, S]  (row (i,c)) */
  const float* b_skip;    /* [S] = sum_i bs_i */
  const float* w_end1;    /* packed [S, E] (= end_conv_1.weight^T) */
  const float* b_end1;    /* [E] */
  const float* w_end2;    /* packed [E, Opad] zero-padded, Opad = 32*ceil(O/32) */
  const float* b_end2;    /* [Opad] */
  float* s1;              /* out/saved [P,S] fp32  relu(skip) */
  float* e1;              /* out/saved [P,E] fp32  relu(end_conv_1) */
  float* out;             /* out [N,O,V,Lf] fp32 NCHW */
  float* ws;              /* workspace [P,Opad] fp32 */
} gwn_head_fwd_args;
int gwn_head_fwd(const gwn_head_cfg* cfg, const gwn_head_fwd_args* a, void* stream);
typedef struct {
  const void* z_last[GWN_MAX_LAYERS];
  const float* w_skip; const float* w_end1; const float* w_end2;
  const float* s1; const float* e1;
  const float* dout;      /* [N,O,V,Lf] fp32 NCHW */
  float* dw_skip; float* db_skip; float* dw_end1; float* db_end1; float* dw_end2; float* db_end2;
  void* dz_last[GWN_MAX_LAYERS];       /* out, each [N,Lf,V,32] (dtype) */
  float* ws_do;           /* [P,Opad] */
  float* ws_de1;          /* [P,E] */
  float* ws_ds1;          /* [P,S] */
} gwn_head_bwd_args;
int gwn_head_bwd(const gwn_head_cfg* cfg, const gwn_head_bwd_args* a, void* stream);

/* bf16 head on tensor cores (TMA-fed tcgen05 GEMMs, csrc/head_tc.cu).  zcat = the per-layer z_last slices
 * concatenated along channels: [P, 32*n_layers] bf16, P = N*Lf*V.  s1/e1 are saved in bf16.
 * ws_w: gwn_head_tc_ws_bytes() bytes of scratch for the bf16 weight images. */
typedef struct {
  const void* zcat;
  const float* w_skip; const float* b_skip; const float* w_end1; const float* b_end1;
  const float* w_end2; const float* b_end2;      /* packed as in gwn_head_fwd_args */
  void* s1;               /* out/saved [P,2S] bf16: [hi | lo] split of relu(skip) */
  void* e1;               /* out/saved [P,E] bf16 */
  float* out;             /* out [N,O,V,Lf] fp32 NCHW */
  void* ws_w;
} gwn_head_tc_fwd_args;
typedef struct {
  const void* zcat;
  const float* w_skip; const float* w_end1; const float* w_end2;
  const void* s1; const void* e1;
  const float* dout;      /* [N,O,V,Lf] fp32 NCHW */
  float* dw_skip; float* db_skip; float* dw_end1; float* db_end1; float* dw_end2; float* db_end2;  /* overwritten */
  void* dz_last[GWN_MAX_LAYERS];       /* out, each [P,32] bf16 */
  void* ws_do;            /* [P,Opad] bf16 */
  void* ws_de1;           /* [P,E] bf16 */
  void* ws_ds1;           /* [P,S] bf16 */
  void* ws_w;
  int outputs_zeroed;     /* != 0: the caller already zeroed dw_*, db_* */
} gwn_head_tc_bwd_args;
long long gwn_head_tc_ws_bytes(int n_layers, int S, int E, int O);
int gwn_head_fwd_tc(const gwn_head_cfg* cfg, const gwn_head_tc_fwd_args* a, void* stream);
int gwn_head_bwd_tc(const gwn_head_cfg* cfg, const gwn_head_tc_bwd_args* a, void* stream);

/* ---- tensor-core (tcgen05) diffusion hops for on-chip-resident supports (V <= 128), bf16 ----
 * gwn_hop_mats_prep builds, once per forward, the UMMA A-operand images of every support:
 * image 4*s+0 = A_s^T, 4*s+1 = (A_s^2)^T (forward hops), 4*s+2 = A_s, 4*s+3 = A_s^2 (backward hops),
 * bf16, K-major no-swizzle canonical layout, zero padded to [Kp/8][128][8], Kp = 16*ceil(V/16).
 * `supports` is a HOST array of device pointers.  gwn_hop_tc runs one hop (tests / microbench). */
/* Which image format gwn_layer_fwd / gwn_layer_bwd expect in `hop_mats` for a graph of V nodes with n_supports supports
 * (order 2): 1 = gwn_hop_mats_prep images (every support resident in shared memory: V <= 80 and the images fit),
 * 2 = gwn_support_images_prep images (one TMA-tiled GEMM per hop), 0 = bad arguments.  Callers MUST build the images
 * this function names - it is the same rule the layer entry points apply. */
int gwn_hop_mode(int V, int n_supports);
int gwn_hop_mats_bytes(int V, int n_supports);
int gwn_hop_mats_prep(const float* const* supports, int n_supports, int V, void* out, void* stream);
int gwn_hop_tc(const void* mats, int n_mats, int mat, void* buf, int pitch, int slot_in, int slot_out,
               int slabs, int V, void* stream);

/* ---- tensor-core diffusion hops for supports that do NOT fit on chip (V > 128: the 3,100-node configs), bf16 ----
 * TMA-tiled tcgen05 GEMM (csrc/tma_gemm.cuh).  gwn_support_images_prep builds per support the bf16 images
 * [2][V][Vp] (Vp = 8*ceil(V/8)): image 0 = A^T (operand of the forward hop y[w] = sum_v A[v,w] x[v]),
 * image 1 = A (operand of the backward hop).  `supports` is a HOST array of device pointers.
 * gwn_hop_big: y = image(support, which) * x (+ add) over slot-major [slabs*V, 32] bf16 buffers.
 * gwn_gemm_test: C[M][N] fp32 = A * B^T through every operand staging mode of the GEMM (unit tests). */
long long gwn_support_images_bytes(int V, int n_supports);
int gwn_support_images_prep(const float* const* supports, int n_supports, int V, void* out, void* stream);
int gwn_hop_big(const void* images, int n_supports, int support, int which, const void* x, void* y,
                const void* add, long long slabs, int V, void* stream);
/* dA[v,w] += sum_{s,c} x[s,v,c] * g[s,w,c]: gradient of nconv wrt the support (fp32 [V][V], accumulated). */
int gwn_dadj_big(const void* x, const void* g, float* dA, long long slabs, int V, void* stream);
int gwn_gemm_test(const void* A, const void* B, float* C, int M, int N, int K, int a_mode, int b_mode,
                  int lda, int ldb, int bn, int splits, void* stream);

/* ---- training loss: nn.MSELoss() of the reference's step (lit.py:24, applied at lit.py:41 and :50), fp32 ----
 * gwn_mse_loss_fwd: loss[0] = mean((a - b)^2) over n elements, ONE launch (a cluster of eight CTAs, partial sums added in
 * a fixed order through distributed shared memory: deterministic, no workspace, nothing to zero).
 * gwn_mse_loss_bwd: d_a[i] = (a[i] - b[i]) * 2 * grad_loss[0] / n (every element written; grad_loss on the device). */
int gwn_mse_loss_fwd(const float* a, const float* b, long long n, float* loss, void* stream);
int gwn_mse_loss_bwd(const float* a, const float* b, const float* grad_loss, long long n, float* d_a, void* stream);

/* ---- parameter re-layout (one launch instead of ~50 stack/permute/copy launches per step) ----
 * gwn_pack_params gathers the reference-shaped parameters (graph_wavenet.py:150-183: Conv2d weights [out,in,1,k] and
 * biases, fp32) into one flat fp32 buffer holding the kernels' packed layouts back to back (segment offsets from
 * gwn_pack_offsets: w_fg [nl][k*32][64] | b_fg [nl][64] | w_mlp [nl][mlp_in][32] | w_skip [nl*32][S] | b_skip [S] |
 * w_end1 [S][E] | w_end2 [E][Opad] | b_end2 [Opad]).  gwn_unpack_grads scatters the gradients of those packed tensors
 * (NULL = absent) back into one flat buffer in parameter order (per layer: dWf, dbf, dWg, dbg, dWm, dWs, dbs; then
 * dW_end1, dW_end2, db_end2).  Pointer tables are passed by value (no device-side table). */
typedef struct {
  int n_layers, taps, mlp_in, S, E, O, Opad;
} gwn_pack_cfg;
typedef struct {
  const void* w_filter[GWN_MAX_LAYERS]; const void* b_filter[GWN_MAX_LAYERS];
  const void* w_gate[GWN_MAX_LAYERS];   const void* b_gate[GWN_MAX_LAYERS];
  const void* w_mlp[GWN_MAX_LAYERS];    /* gconv.mlp.mlp.weight [32][mlp_in] (or residual_convs weight) */
  const void* w_skip[GWN_MAX_LAYERS];   const void* b_skip[GWN_MAX_LAYERS];
  const void* w_end1; const void* w_end2; const void* b_end2;
} gwn_pack_ptrs;
typedef struct {
  const void* w_fg[GWN_MAX_LAYERS]; const void* b_fg[GWN_MAX_LAYERS]; const void* w_mlp[GWN_MAX_LAYERS];
  const void* w_skip; const void* b_skip; const void* w_end1; const void* w_end2; const void* b_end2;
} gwn_unpack_ptrs;
long long gwn_pack_offsets(const gwn_pack_cfg* cfg, long long* off9);   /* returns the total element count */
int gwn_pack_params(const gwn_pack_cfg* cfg, const gwn_pack_ptrs* ptrs, float* out, void* stream);
long long gwn_unpack_total(const gwn_pack_cfg* cfg);
int gwn_unpack_grads(const gwn_pack_cfg* cfg, const gwn_unpack_ptrs* grads, float* out, void* stream);

/* ---- nconv primitive, exposed for unit tests  graph_wavenet.py:60-66 ----
 * y[s,w,c] = sum_v x[s,v,c] * A[v,w] (transpose_a=0) or A[w,v] (transpose_a=1);
 * x,y: [slabs, V, pitch] slots of 32 channels at column offsets xoff/yoff. */
int gwn_node_mix(const void* x, int x_pitch, int x_off, void* y, int y_pitch, int y_off, int accumulate,
                 const float* A, int transpose_a, int slabs, int V, int dtype, void* stream);

/* ---- optimizer step, and the data-parallel gradient exchange fused with it (csrc/peer.cu) ----
 * Replaces, for the train step of lit.py:29-43,59-61 (torch.optim.Adam(lr=1e-3), amsgrad off, no weight decay):
 *   1 GPU : `optimizer.step()` over ~110 small tensors (three multi-tensor launches) -> gwn_adam_flat, ONE launch over
 *           the flat fp32 parameter / gradient / moment buffers;
 *   N GPUs: gradient all-reduce (the reference has no distributed code) + `optimizer.step()` ->
 *           gwn_allreduce_adam, ONE launch per rank: one-shot all-reduce (average) over NVLink / NVSwitch peer memory
 *           fused with the Adam update.
 * Exchange block = gwn_peer_header_bytes() of flags followed by the flat gradient; allocate with gwn_peer_alloc (cudaMalloc,
 * zeroed), export with gwn_peer_export (64-byte CUDA IPC handle, host buffer), map the peers' blocks with gwn_peer_open.
 * `state` = 2 x uint64 on the device, zero-initialised: [completed steps, CTA counter]; the step count advances on the
 * device, so CUDA-graph replays work.  `blocks[r]` = exchange block of rank r as mapped in THIS process (blocks[rank] = own).
 * Every rank must launch once per step; a rank that never arrives makes the others trap (bounded polls), never hang. */
/* srcs[k] (device pointers, NULL = zeros) of counts[k] fp32 elements each -> `out` back to back: one launch (pointer table
 * by value, n <= 160).  The gradient packing step in front of gwn_adam_flat / gwn_allreduce_adam. */
int gwn_gather_flat(const void* const* srcs, const long long* counts, int n, float* out, void* stream);
long long gwn_peer_header_bytes(void);
int gwn_peer_alloc(long long grad_bytes, void** block);
int gwn_peer_free(void* block);
int gwn_peer_export(const void* block, void* handle64);
int gwn_peer_open(const void* handle64, void** mapped_block);
int gwn_peer_close(void* mapped_block);
int gwn_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                  void* state, void* stream);
int gwn_allreduce_adam(float* p, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                       void* state, const void* const* blocks, int rank, int world, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GWN_H */
