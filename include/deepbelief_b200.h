/* deepbelief_b200 — C ABI of the B200-native hot paths of zeis/deepbelief.
 *
 * The reference has no FFI layer: its hot path is the TensorFlow graph built by the
 * Python classes in deepbelief/layers. The boundary kept is that Python API (deepbelief_b200.layers);
 * this header is what its ctypes loader (deepbelief_b200/_lib.py) binds, one call per fused stage.
 * Every entry cites the reference lines whose TF ops it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - all matrices fp32, row-major, device pointers, leading dimension in elements (config.py:1);
 *   - the library never allocates or frees device memory: callers pass outputs and a workspace
 *     (size from the matching *_workspace_bytes); calls are asynchronous on `stream`. The GP entries
 *     (db_gplvm_lml_grad, db_gpdbn_factor) fork onto two auxiliary streams owned by the handle and join
 *     back before they return, so `stream` order is all a caller sees (also under stream capture);
 *   - return value: 0 ok; < 0 error (DB_ERR_*), message via db_last_error(); > 0 LAPACK-style info
 *     (1-based index of the failing Cholesky pivot / first negative predictive variance);
 *   - no entry point has a CPU fallback: without a CUDA device db_create fails.
 */
#ifndef DEEPBELIEF_B200_H
#define DEEPBELIEF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct db_context_s* db_handle_t;
typedef void* db_stream_t; /* cudaStream_t */

enum {
  DB_OK = 0,
  DB_ERR_BAD_ARG = -1,
  DB_ERR_CUDA = -2,
  DB_ERR_UNSUPPORTED = -3,
  DB_ERR_WORKSPACE = -4
};

/* activation of a unit layer */
enum { DB_ACT_LINEAR = 0, DB_ACT_SIGMOID = 1 };
/* sampling rule applied to the activation */
enum {
  DB_SAMPLE_NONE = 0,      /* return the probabilities / means (sample=False, rbm.py:209-210) */
  DB_SAMPLE_BERNOULLI = 1, /* float(p > u), strict (rbm.py:206-208, 234-236) */
  DB_SAMPLE_CONCRETE = 2,  /* relaxed Bernoulli with temperature (rbm.py:243-275) */
  DB_SAMPLE_GAUSSIAN = 3   /* mean + sigma * eps (bgrbm.py:106-110) */
};
/* broadcast mode of an optional scale operand */
enum { DB_BCAST_NONE = 0, DB_BCAST_ROW = 1 /* [1, n] */, DB_BCAST_FULL = 2 /* [m, n] */ };
/* RBM family member: decides activations and the free-energy / CD-gradient form */
enum {
  DB_LAYER_RBM = 0,   /* rbm.py   : Bernoulli visible, Bernoulli hidden */
  DB_LAYER_BGRBM = 1, /* bgrbm.py : Bernoulli visible, Gaussian hidden (sigma_h) */
  DB_LAYER_GBRBM = 2, /* gbrbm.py : Gaussian (unit variance, noise-free mean) visible, Bernoulli hidden */
  DB_LAYER_GGRBM = 3  /* ggrbm.py : Gaussian visible, Gaussian hidden */
};
/* optimizer rule (rbm.py:666-670 takes any TF optimizer; [TF, not vendored] update rules restated) */
enum {
  DB_OPT_SGD = 0,      /* theta -= lr * g                                   (GradientDescentOptimizer) */
  DB_OPT_MOMENTUM = 1, /* a = mu*a + g; theta -= lr * a                     (MomentumOptimizer)        */
  DB_OPT_ADAM = 2      /* m,v moments; theta -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps) (AdamOptimizer) */
};

/* Counter-based RNG state: Philox4x32-10 keyed by `seed`; `draw` identifies one sampling call (the
 * caller increments it per call); `row_offset` is the global index of local row 0 so that a
 * row-sharded batch draws the same numbers on any number of GPUs. */
typedef struct {
  uint64_t seed;
  uint64_t draw;
  uint32_t row_offset;
  uint32_t reserved;
  const uint64_t* clock; /* optional device clock (db_clock_advance): the draw id used is draw + clock[0], read by the
                            kernel — a captured CUDA graph replays with fresh draws. NULL = draw as given */
} db_rng_t;

typedef struct {
  int kind;            /* DB_OPT_* */
  float lr;
  float momentum;      /* DB_OPT_MOMENTUM */
  float beta1, beta2, eps; /* DB_OPT_ADAM */
  float weight_decay;  /* g += weight_decay * theta (builder-defined extension) */
  int step;            /* 1-based step number t (Adam bias correction) */
  const uint64_t* clock; /* optional device clock: the step number used is step + clock[1] (graph replay). NULL = step */
} db_opt_t;

/* ---- context ------------------------------------------------------------------------------- */
int db_create(db_handle_t* out, int device);
int db_destroy(db_handle_t h);
const char* db_last_error(db_handle_t h);
int db_version(void);
int db_num_sms(db_handle_t h);
/* Lets kernels of this context's device dereference pointers into `peer_device`'s memory (cudaDeviceEnablePeerAccess); needed
 * once per peer before db_cd_tc_step_push / db_cd_tc_apply_update_sharded are given peer-mapped pointers. */
int db_enable_peer_access(db_handle_t h, int peer_device);
/* CUDA IPC for the peer-push data-parallel step (all ranks on one node): db_ipc_export gives the 64-byte handle of the
 * allocation that contains ptr and ptr's offset inside it; the other ranks exchange these by any means and map the buffer
 * with db_ipc_import, which opens it in the calling rank's own device context (kernels of that device may then store
 * into it over NVLink). */
int db_ipc_export(db_handle_t h, const void* ptr, void* handle_out64, unsigned long long* offset_out);
int db_ipc_import(db_handle_t h, const void* handle64, unsigned long long offset, void** ptr_out);

/* ---- K1: fused unit layer  out = sample(act(col_scale * ((A / a_div) @ op(W)) + bias)) --------
 * Replaces, in one launch: rbm.py:119-131 h_net_input, rbm.py:133-147 v_net_input (transW=1, no
 * materialised transposes), rbm.py:149-175 sigmoid, rbm.py:187-275 sampling; bgrbm.py:69-82 (col_scale =
 * sigma_h), bgrbm.py:116-131 (a_div = sigma_h), bgrbm.py:96-114 (Gaussian sample), gbrbm.py:56-64.
 *   A [M,K]; W is [K,N] (transW=0) or [N,K] used transposed (transW=1); bias [N] or NULL.
 *   a_div / col_scale / noise_scale: NULL or vector [K]/[N] (DB_BCAST_ROW) or matrix (DB_BCAST_FULL,
 *   same leading dimension as the tensor they scale: lda for a_div, N for the others).
 *   injected: uniforms U (Bernoulli/Concrete) or standard normals (Gaussian) [M,N], or NULL -> Philox.
 *   out_act (value after activation) and out_sample may each be NULL. */
int db_layer_forward(db_handle_t h, db_stream_t stream,
                     const float* A, int lda, const float* a_div, int a_div_mode,
                     const float* W, int ldw, int transW,
                     const float* col_scale, int col_scale_mode, const float* bias,
                     int M, int N, int K, int act, int sample_kind, float temperature,
                     const float* noise_scale, int noise_scale_mode,
                     const float* injected, const db_rng_t* rng,
                     float* out_act, float* out_sample);

/* Elementwise sampling of given probabilities/means (rbm.py:177-275 sample_h / sample_v called on a
 * tensor; bgrbm.py:96-114). P, out: [M,N] contiguous. */
int db_sample(db_handle_t h, db_stream_t stream, const float* P, int M, int N, int sample_kind,
              float temperature, const float* noise_scale, int noise_scale_mode,
              const float* injected, const db_rng_t* rng, float* out);

/* Free energy per datapoint, out[N] (rbm.py:378-396; bgrbm.py:133-153 when layer_kind has Gaussian
 * hiddens, sigma_h [H] required then). workspace: N*H floats. */
int db_free_energy(db_handle_t h, db_stream_t stream, int layer_kind, const float* v, int N, int V, int H,
                   const float* W, const float* b, const float* c, const float* sigma_h,
                   float* workspace, float* out);

/* ---- K2: CD gradient + optimizer ------------------------------------------------------------- */
/* Packed parameter / gradient layout used by db_cd_gradients and db_optimizer_step:
 *   [ W (V*H) | b (V) | c (H) | opt_sigma_h (H, only BGRBM/GGRBM) ]                                    */
size_t db_cd_param_count(int layer_kind, int V, int H);

/* Gradient of cd_loss (rbm.py:398-415, closed form SURVEY.md §8 R9 / B2) from the two ends of the
 * chain: v0, vk [N,V]. Writes SUMS scaled by inv_global_n (= 1/N over all ranks) into grad (packed),
 * so a sum-allreduce over row shards yields the reference's mean gradient. workspace: 2*N*H floats. */
int db_cd_gradients(db_handle_t h, db_stream_t stream, int layer_kind,
                    const float* v0, const float* vk, int N, int V, int H,
                    const float* params, float inv_global_n, float* workspace, float* grad);

/* Same gradient when the caller already holds the hidden probabilities hp0 = sigmoid(v0 W + c) and hpk = sigmoid(vk W + c)
 * [N,H] — on Bernoulli-hidden layers (RBM, GBRBM) they are what the chain sampled h_0 and h_k from (rbm.py:628, 368-370),
 * so the two free-energy GEMMs of rbm.py:388-390 are not repeated. Gaussian-hidden kinds are refused. */
int db_cd_gradients_hidden(db_handle_t h, db_stream_t stream, int layer_kind,
                           const float* v0, const float* vk, const float* hp0, const float* hpk, int N, int V, int H,
                           const float* params, float inv_global_n, float* grad);

/* Device clock for CUDA-graph replay of a training iteration: clock is 2 uint64 in device memory, [0] = draw ids
 * consumed so far, [1] = optimizer steps taken so far (both relative to the values baked into the captured launches).
 * db_clock_advance adds (draws, steps) on the stream — captured as the last node of the iteration. */
int db_clock_advance(db_handle_t h, db_stream_t stream, uint64_t* clock, uint64_t draws, uint64_t steps);

/* One optimizer step over n packed elements; state: Momentum n floats, Adam 2n floats (m | v). */
int db_optimizer_step(db_handle_t h, db_stream_t stream, const db_opt_t* opt, float* params,
                      const float* grad, float* state, size_t n);

/* ---- fused CD-k chain for the binary RBM on tensor cores (tcgen05) ----------------------------
 * One call = positive phase + k Gibbs steps + dW/db/dc (rbm.py:628-646, 664 + autodiff), using fp16
 * hi/lo split operands so results hold fp32 tolerance. Falls back to nothing: returns
 * DB_ERR_UNSUPPORTED when the shape does not qualify (caller then uses db_layer_forward et al.). */
typedef struct {
  int N;              /* local batch rows (multiple of 8) */
  int V, H;           /* >= 8 (internal fp16 buffers pad their row pitch to a multiple of 8) */
  int k;              /* Gibbs steps >= 1 */
  int pcd;            /* 1: chain starts from (and updates) the persistent hidden state */
  float inv_global_n; /* 1 / (batch rows over all ranks) */
  int v0_u8;          /* 1: v0 is uint8 {0,1} instead of fp32 (binary data sets staged as bytes: 4x less host->device traffic) */
} db_cd_tc_desc_t;

size_t db_cd_tc_workspace_bytes(const db_cd_tc_desc_t* d);
/* bytes of the persistent split-weight cache (W and W^T as fp16 hi/lo + scale) */
size_t db_cd_tc_weight_cache_bytes(int V, int H);
/* refresh the cache after W changed (optimizer step / restore) */
int db_cd_tc_refresh_weights(db_handle_t h, db_stream_t stream, const float* W, int V, int H, void* cache);
/* v0: [N,V] fp32 {0,1} (uint8 when d->v0_u8). params packed (W|b|c). pcd_state: fp16 persistent chain [N, H rounded up to 8]
 * (pcd=1) or NULL.
 * injected: NULL (Philox) or uniforms laid out [U_h0 (N*H) | per step: U_v (N*V), U_h (N*H)].
 * grad: packed (W|b|c) sums * inv_global_n. Optional fp32 exports (may be NULL): vk_out [N,V],
 * hk_out [N,H] (last samples), p0_out / pk_out [N,H] (hidden probabilities at both chain ends). */
int db_cd_tc_step(db_handle_t h, db_stream_t stream, const db_cd_tc_desc_t* d, const void* v0,
                  const float* params, const void* weight_cache, void* pcd_state,
                  const float* injected, const db_rng_t* rng, void* workspace, float* grad,
                  float* vk_out, float* hk_out, float* p0_out, float* pk_out);

/* North-star K2, fused: db_cd_tc_step whose dW kernel also applies the optimizer rule (rbm.py:664-670: cd_loss ->
 * optimizer.minimize) to the packed parameters IN PLACE and rewrites the split-weight cache, so the following step needs no
 * db_optimizer_step / db_cd_tc_refresh_weights. opt_state: as db_optimizer_step (Momentum n floats, Adam 2n, n = V*H+V+H).
 * grad still receives the packed gradient. Results are bit-identical to db_cd_tc_step + db_optimizer_step +
 * db_cd_tc_refresh_weights. opt->clock must be NULL. DB_ERR_UNSUPPORTED when db_cd_tc_update_fusable(d) == 0. */
int db_cd_tc_step_update(db_handle_t h, db_stream_t stream, const db_cd_tc_desc_t* d, const void* v0,
                         float* params, void* weight_cache, void* pcd_state,
                         const float* injected, const db_rng_t* rng, void* workspace, float* grad,
                         const db_opt_t* opt, float* opt_state);
/* The data-parallel form of the same tail, run after the caller has summed `grad` over the ranks (NCCL): optimizer rule
 * on the packed parameters (W | b | c) and rewrite of the split-weight cache in ONE pass over W — replaces
 * db_optimizer_step + db_cd_tc_refresh_weights, bit-identical to them. Any V, H. */
int db_cd_tc_apply_update(db_handle_t h, db_stream_t stream, int V, int H, float* params, const float* grad,
                          const db_opt_t* opt, float* opt_state, void* weight_cache);
/* 0: the dW product of this shape runs K-split (small layers), db_cd_tc_step_update is not available. 1: available (one work
 * item per 128x256 tile). 2: available AND faster than db_cd_tc_step + db_cd_tc_apply_update — the epilogue's pass over W
 * hides under the tile's MMAs only when the batch (K = 2N of the dW product) is deep enough (>= 16384). */
int db_cd_tc_update_fusable(const db_cd_tc_desc_t* d);

/* Data-parallel CD step with the gradient reduce-scattered by PUSH from the dW kernel (no counterpart in the single-process
 * reference; replaces the NCCL all-reduce of the 4 V H byte gradient of db_cd_tc_step + db_cd_tc_apply_update):
 *  - db_cd_tc_step_push: db_cd_tc_step whose dW epilogue stores the rows of dW owned by rank r (V / world consecutive rows
 *    of W each) straight into rank r's slot buffer peer_slots[r] (peer-mapped device pointers, [world][V/world * H] floats,
 *    slot index = the writing rank) while the MMAs of the next tile run; grad receives db | dc only.
 *  - the caller then sums db | dc over the ranks and orders the ranks (any stream-ordered collective, e.g. NCCL);
 *  - db_cd_tc_apply_update_sharded: sums the world slots of this rank in rank order, applies the optimizer rule to its rows
 *    of W with its slice of opt_state, writes the reduced rows into grad and the NEW rows into every rank's parameter
 *    vector peer_params[r] (the all-gather, by push), and updates b | c identically on every rank;
 *  - after a second ordering of the ranks every rank refreshes its split-weight cache (db_cd_tc_refresh_weights).
 * db_cd_tc_push_ok: 1 when the shape qualifies (2..8 ranks, V a multiple of 256 x ranks, dW not K-split). */
int db_cd_tc_push_ok(const db_cd_tc_desc_t* d, int world);
int db_cd_tc_step_push(db_handle_t h, db_stream_t stream, const db_cd_tc_desc_t* d, const void* v0,
                       const float* params, const void* weight_cache, void* pcd_state,
                       const float* injected, const db_rng_t* rng, void* workspace, float* grad,
                       float* const* peer_slots, int world, int rank);
int db_cd_tc_apply_update_sharded(db_handle_t h, db_stream_t stream, int V, int H, int world, int rank,
                                  float* const* peer_params, const float* slots, float* grad, const db_opt_t* opt,
                                  float* opt_state);

/* ---- GP path (gplvm.py / gprbm.py) ----------------------------------------------------------- */
/* hyp (device, 3 floats + workspace use): effective kernel variance alpha, length-scale gamma, noise
 * variance (+jitter) AFTER softplus, i.e. the values gplvm.py:39,49,152,219-221 produce. */

/* K3: RBF Gram  K[i,j] = alpha * exp(-0.5 * |x_i/gamma - y_j/gamma|^2) (+ noise on the diagonal when
 * Y == NULL)  — gplvm.py:61-86 + 218-224. X [N,Q], Y [M,Q] or NULL (then M = N). out [N,M]. */
int db_gram_rbf(db_handle_t h, db_stream_t stream, const float* X, const float* Y, int N, int M, int Q,
                const float* hyp, int add_noise, float* out, int ldo);

/* K4: in-place lower Cholesky (tf.cholesky, gplvm.py:212 / gprbm.py:109). The strict upper triangle is
 * left untouched. info (device int): 0 or 1-based index of the first non-positive pivot. */
int db_potrf_lower(db_handle_t h, db_stream_t stream, float* A, int lda, int N, int* info);

/* K5: triangular solve with the Cholesky factor (tf.matrix_triangular_solve, gplvm.py:235,248-249,281).
 * trans=0: B <- L^-1 B; trans=1: B <- L^-T B. B [N,nrhs]. */
int db_trsm_lower(db_handle_t h, db_stream_t stream, const float* L, int ldl, int N, float* B, int ldb,
                  int nrhs, int trans);

/* K3-K6 fused pipeline: GPLVM negative log marginal likelihood and its gradient (gplvm.py:210-244 +
 * the autodiff of gplvm.py:425-428; closed forms SURVEY.md §8 GP6/GP7).
 *   X [N,Q], Y [N,D] (already centred), hyp_opt (device): [opt_alpha, opt_gamma, opt_noise] pre-softplus.
 *   flags bit0: alpha trainable, bit1: gamma trainable, bit2: noise trainable (=> + jitter);
 *   untrainable alpha/gamma are the constant 1.0 (gplvm.py:31-32,41-42); untrainable noise = fixed_noise.
 *   x_prior_weight: adds w*|X|^2 (gprbm.py:238-242); include_const: add D*N*log(2*pi) (gplvm.py:238).
 *   out_scalars (device, 8 floats): [loss, logdet, |L^-1 Y|^2, d/dopt_alpha, d/dopt_gamma, d/dopt_noise, alpha, noise_eff]
 *   dX [N,Q]. L_out (optional, [N,N]) receives the Cholesky factor for later prediction calls. */
size_t db_gplvm_workspace_bytes(int N, int Q, int D);
int db_gplvm_lml_grad(db_handle_t h, db_stream_t stream, const float* X, const float* Y, int N, int Q, int D,
                      const float* hyp_opt, int flags, float fixed_noise, float jitter,
                      float x_prior_weight, int include_const, int want_grad,
                      void* workspace, float* out_scalars, float* dX, float* L_out, int* info);

/* GP prediction at M test latents (gplvm.py:246-254, 278-288; gprbm.py:208-232):
 *   S = L^-1 Yc [N,D] cached by the caller (db_trsm_lower), L [N,N].
 *   mean [M,D] = (L^-1 K_Xx)^T S (+ y_mean [D] if not NULL); var [M] = alpha - colsum((L^-1 K_Xx)^2).
 *   info: 1-based index of the first negative variance (tf.assert_non_negative), else 0.
 *   workspace: N*M floats. */
int db_gp_predict(db_handle_t h, db_stream_t stream, const float* X, const float* L, const float* S,
                  const float* y_mean, const float* xstar, int N, int Q, int D, int M, const float* hyp,
                  float* workspace, float* mean, float* var, int* info);

/* effective hyper-parameters from optimisation variables: hyp_out[3] = [alpha, gamma, noise(+jitter)] */
int db_gp_effective_hyp(db_handle_t h, db_stream_t stream, const float* hyp_opt, int flags, float fixed_noise,
                        float jitter, float* hyp_out);

/* out[i] = softplus(x[i]) — the positivity constraint of sigma_h / alpha_h (bgrbm.py:51, gprbm.py:83). */
int db_softplus(db_handle_t h, db_stream_t stream, const float* x, int n, float* out);

/* ---- small fused helpers used by the layer API ------------------------------------------------ */
/* mean over rows of the sum over columns of (a-b)^2 — dist.py:5-21. out: device float. */
int db_mse(db_handle_t h, db_stream_t stream, const float* a, const float* b, int N, int D, float* out);
/* dist.py:24-35 cross entropy with clipping to [1e-6, 1-1e-6]; reduce_mean=1: mean of row sums, 0: total. */
int db_cross_entropy(db_handle_t h, db_stream_t stream, const float* probs, const float* targets, int N, int D,
                     int reduce_mean, float* out);

/* ---- 3xTF32 tensor-core GEMM (tcgen05 kind::tf32) for the GP chain ------------------------------------------
 * C[M,N] = alpha * X[M,K] Y[N,K]^T + beta * C, fp32, both operands row-major (K contiguous). X_lo / Y_lo are the
 * tf32 remainders x - tf32(x) of the operands (same shape and leading dimension), produced by db_tf32_split_lo.
 * Three MMA passes (hi*hi + lo*hi + hi*lo) keep fp32-level accuracy; K is accumulated in chunks of 1024 columns.
 * lower_only = 1 writes only elements with column <= row. Needs 16-byte aligned pointers and ldx, ldy % 4 == 0
 * (else DB_ERR_UNSUPPORTED). Replaces the dense products inside tf.cholesky / tf.matrix_triangular_solve and their
 * autodiff (gplvm.py:212, 235, 425-428). */
int db_tf32_split_lo(db_handle_t h, db_stream_t stream, const float* x, int rows, int cols, int ld, float* lo);
int db_gemm_tf32x3_nt(db_handle_t h, db_stream_t stream, int M, int N, int K, float alpha, const float* X,
                      const float* X_lo, int ldx, const float* Y, const float* Y_lo, int ldy, float beta, float* C,
                      int ldc, int lower_only);
/* The same product from power-of-two scaled fp16 pairs (x 2^e = hi + lo, three kind::f16 passes: twice the tensor rate
 * of the 3xTF32 form at the same accuracy): the operands are converted into `scratch` (db_gemm_f16x3_scratch_bytes) with
 * scales taken from their largest entries. The GP chain uses the kernel with operands it keeps in that form. */
size_t db_gemm_f16x3_scratch_bytes(int M, int N, int K);
int db_gemm_f16x3_nt(db_handle_t h, db_stream_t stream, int M, int N, int K, float alpha, const float* X, int ldx,
                     const float* Y, int ldy, float beta, float* C, int ldc, int lower_only, void* scratch);

/* ---- GPDBN top layer (gprbm.py) in training mode (x_ph = X, gprbm.py:106) ---------------------------
 * One optimisation step of GPRBM.optimize (gprbm.py:375-438) evaluates loss = CE(downprop(sample_mean), V)
 * + gp_loss (gprbm.py:247-256) and the gradient TF autodiff derives for every trainable. The calls below
 * are its GP block and top-layer stages; the RBM stack around them uses db_layer_forward / db_gemm /
 * db_concrete_backward. All share one workspace (db_gpdbn_workspace_bytes). */
size_t db_gpdbn_workspace_bytes(int N, int Q, int D);
/* K = K_f(X) + noise I, L = chol K, K^-1, logdet (gprbm.py:108-112, 193-206);
 * pred_var[i] = noise - noise^2 K^-1_ii  (= gprbm.py:208-218 at x_ph = X).
 * info (device int[2]): [0] failing Cholesky pivot, [1] first negative pred_var (assert_non_negative), 1-based.
 * L_out (optional [N,N]) receives the factor for later db_gp_predict calls. */
int db_gpdbn_factor(db_handle_t h, db_stream_t stream, const float* X, int N, int Q, int D, const float* hyp_opt,
                    int flags, float fixed_noise, float jitter, void* workspace, float* pred_var, float* L_out,
                    int* info);
/* Hm = (L^-1 K_Xx)^T (L^-1 H2) = H2 - noise K^-1 H2 (gprbm.py:220-224); scalars (device, 8 floats):
 * [gp_loss = 0.5(sum H2 o K^-1 H2 + D logdet) + |X|^2 (gprbm.py:234-245), logdet, sum H2 o K^-1 H2, |X|^2,
 *  alpha, gamma, noise, 0]. */
int db_gpdbn_apply(db_handle_t h, db_stream_t stream, const float* H2, int N, int Q, int D, void* workspace,
                   float* Hm, float* scalars);
/* g_H2 = dLoss/dH2 given g_Hm = dLoss/dHm (includes the gp_loss term K^-1 H2). */
int db_gpdbn_backward_h(db_handle_t h, db_stream_t stream, const float* g_Hm, int N, int Q, int D, void* workspace,
                        float* g_H2);
/* dX [N,Q] and dhyp[3] = d/d(opt_alpha, opt_gamma, opt_noise) given g_pv = dLoss/dpred_var [N]; must follow
 * db_gpdbn_backward_h of the same step. x_prior_weight: the |X|^2 term of gprbm.py:242 (1.0). */
int db_gpdbn_backward_x(db_handle_t h, db_stream_t stream, const float* X, const float* g_pv, int N, int Q, int D,
                        const float* hyp_opt, int flags, float x_prior_weight, void* workspace, float* dX,
                        float* dhyp);
/* H1 = sigma_h (x + eps1) + c, sigma_h = alpha_h sqrt(pred_var) (bgrbm.py:79-82,107-110; gprbm.py:114-115);
 * H1_mean [D]; H2 = (H1 - H1_mean) / alpha_h (gprbm.py:146-147). x = a W [N,D], eps1 standard normals. */
int db_gpdbn_top_forward_up(db_handle_t h, db_stream_t stream, const float* x, const float* pred_var,
                            const float* alpha_h, const float* c, const float* eps1, int N, int D, float* H1,
                            float* H1_mean, float* H2);
/* H = Hm alpha_h + H1_mean + sigma_h eps2 (gprbm.py:225-231); y = H / sigma_h (bgrbm.py:126). H may be NULL. */
int db_gpdbn_top_forward_mid(db_handle_t h, db_stream_t stream, const float* Hm, const float* pred_var,
                             const float* alpha_h, const float* H1_mean, const float* eps2, int N, int D, float* H,
                             float* y);
/* backward of the two stages above: (1) from g_y to g_Hm and partial sums, (2) from g_H2 to g_x, g_c,
 * g_opt_alpha_h and g_pv. gsig [N,D], ga/gmean/g_alpha_scratch [D] are scratch carried between the calls. */
int db_gpdbn_top_backward1(db_handle_t h, db_stream_t stream, const float* g_y, const float* y, const float* Hm,
                           const float* pred_var, const float* alpha_h, const float* eps2, int N, int D, float* g_Hm,
                           float* gsig, float* ga, float* gmean);
int db_gpdbn_top_backward2(db_handle_t h, db_stream_t stream, const float* g_H2, const float* H2, const float* x,
                           const float* eps1, const float* pred_var, const float* alpha_h, const float* opt_alpha_h,
                           const float* ga, const float* gmean, int N, int D, float* g_x, float* gsig, float* g_c,
                           float* g_alpha_scratch, float* g_opt_alpha_h, float* g_pv);

/* ---- test-time latent search (GPLVM.test_generalisation, gplvm.py:470-541), batched over R restarts ---------
 * loss[r] = CE(normalised(pm[r,:]), target) / D + log_var_scaling * log(pred_var[r]) with
 * normalised(v) = (v - min v) / (max v - min v) (gplvm.py:293-296), CE as dist.py:24-35; also d loss / d pm and
 * d loss / d pred_var. pm_norm / pm_clip (clip to [0,1], gplvm.py:480) may be NULL. */
int db_testgen_rowloss(db_handle_t h, db_stream_t stream, const float* pm, const float* target, const float* pred_var,
                       int R, int D, float log_var_scaling, float* loss, float* pm_norm, float* pm_clip, float* g_pm,
                       float* g_pv);
/* pred_var[r] = alpha - sum_i Kxs[i,r] (K^-1 Kxs)[i,r] (gplvm.py:278-288); info: first negative entry (1-based). */
int db_testgen_pred_var(db_handle_t h, db_stream_t stream, const float* Kxs, const float* Cv, int N, int R,
                        const float* hyp, float* pred_var, int* info);
/* g_x[r,:] = sum_i (gK[i,r] - 2 g_pv[r] Cv[i,r]) dKxs[i,r]/dx_r — the chain through the RBF kernel (gplvm.py:61-86). */
int db_testgen_grad_x(db_handle_t h, db_stream_t stream, const float* gK, const float* Cv, const float* Kxs,
                      const float* g_pv, const float* X, const float* x, int N, int R, int Q, const float* hyp,
                      float* g_x);

/* GPRBM.test_generalisation (gprbm.py:468-562), R restarts x S samples per restart (rows ordered r*S + s):
 * loss[r] = CE(mean_s Vs[(r,s),:], target) / Dim + log_var_scaling * log(pred_var[r]) (gprbm.py:497-499);
 * model_point [R,Dim] = the sample mean (may be NULL); g_Vs = d loss / d Vs. */
int db_testgen_meanloss(db_handle_t h, db_stream_t stream, const float* Vs, const float* target, const float* pred_var,
                        int R, int S, int Dim, float log_var_scaling, float* loss, float* model_point, float* g_Vs);
/* sums the per-row outputs of db_gpdbn_top_backward1 over the S samples of each restart:
 * g_Hm [R,D]; g_pv[r] = sum_{s,d} gsig alpha_h / (2 sqrt(pred_var[r])) + log_var_scaling / pred_var[r]. */
int db_testgen_reduce_samples(db_handle_t h, db_stream_t stream, const float* gHm_rows, const float* gsig_rows,
                              const float* alpha_h, const float* pred_var, int R, int S, int D, float log_var_scaling,
                              float* g_Hm, float* g_pv);

/* GPRBM.test_held_out_data (gprbm.py:440-466): loss[r] = mean_m sum_d (T[m,d] - Vs[r,d])^2 — dist.mse (dist.py:5-21)
 * of the [M,Dim] test set against the one down-propagated row of restart r; g_Vs [R,Dim] = d loss / d Vs. */
int db_heldout_mse(db_handle_t h, db_stream_t stream, const float* Vs, const float* T, int R, int M, int Dim,
                   float* loss, float* g_Vs);
/* GPRBM.interp_test (gprbm.py:651-660): objective = mean over the P+1 segments x_start -> Xi[0] -> ... -> Xi[P-1] -> x_end
 * of |segment|^2 + lambda * mean_p log pred_var[p]; g_x [P,Q] = gradient of the segment term with respect to Xi,
 * g_pv [P] = lambda / (P pred_var[p]) (chained through db_testgen_grad_x by the caller). */
int db_interp_objective(db_handle_t h, db_stream_t stream, const float* Xi, const float* x_start, const float* x_end,
                        const float* pred_var, int P, int Q, float lambda, float* objective, float* g_x, float* g_pv);

/* ---- backward building blocks (what TF autodiff derives from tf.matmul / the Concrete units / dist.py) -- */
/* C[M,N] = alpha * op(A) op(B) + beta * C, fp32 row-major; op = transpose when the flag is set. */
int db_gemm(db_handle_t h, db_stream_t stream, int transA, int transB, int M, int N, int K, float alpha,
            const float* A, int lda, const float* B, int ldb, float beta, float* C, int ldc);
/* dz = g * d concrete(p,u,T)/dp * p(1-p) with a = the saved Concrete sample (rbm.py:243-275). */
int db_concrete_backward(db_handle_t h, db_stream_t stream, const float* g, const float* p, const float* a, int M,
                         int N, float temperature, float* dz);
/* dz = g * p (1 - p): derivative of tf.sigmoid (rbm.py:160,174) where probabilities are propagated (sample=False). */
int db_sigmoid_backward(db_handle_t h, db_stream_t stream, const float* g, const float* p, int M, int N, float* dz);
/* out = scale * d cross_entropy / d probs (dist.py:24-35; zero where probs was clipped). */
int db_cross_entropy_grad(db_handle_t h, db_stream_t stream, const float* probs, const float* targets, int M, int N,
                          float scale, float* out);
/* out[c] = beta * out[c] + scale * sum_rows A[:,c]  (bias gradients). */
/* SBM_Lower (reference deepbelief/layers/sbm_lower.py:59-117): one weight matrix W_shared [P*P, num_h/4] shared by the
 * four overlapping P x P patches of the side x side image, P = side/2 + side_overlap/2. db_sbm_expand writes the dense
 * equivalent W_eff [side*side, num_h] (the tf.slice / tf.pad / tf.concat plumbing of h_net_input and v_net_input as one
 * scatter), so every RBM entry point runs unchanged on it; db_sbm_gather sums the four scattered blocks of a dense
 * gradient into the gradient of the shared matrix (what TF autodiff derives for the shared variable). */
int db_sbm_expand(db_handle_t h, db_stream_t stream, const float* W_shared, int side, int side_overlap, int num_h,
                  float* W_eff);
int db_sbm_gather(db_handle_t h, db_stream_t stream, const float* gW_eff, int side, int side_overlap, int num_h,
                  float* gW_shared);

int db_colsum(db_handle_t h, db_stream_t stream, const float* A, int M, int N, float scale, float beta, float* out);

/* ---- host-side data loader helper (no device work) --------------------------------------------- */
/* dst[i, :] = src[rows[i], :] for i < n_rows — the batch gather of Data.next_batch (data.py:92-94:
 * `datapoints[sorted_rows]`), multi-threaded, written straight into a (pinned) staging buffer.
 * rows are validated against n_src_rows; threads <= 0 picks min(16, hardware threads). */
int db_host_gather_rows(const void* src, size_t row_bytes, const int64_t* rows, size_t n_rows,
                        size_t n_src_rows, void* dst, int threads);

#ifdef __cplusplus
}
#endif
#endif /* DEEPBELIEF_B200_H */
