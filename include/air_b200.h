/* air_b200.h -- C ABI of the B200-native AIR hot path (libair_b200.so).
 *
 * The reference (aakhundov/tf-attend-infer-repeat) is pure TensorFlow-1.3 Python and has
 * no plugin / FFI interface: its boundary for this path is the Python call surface of
 * air/transformer.py, air/concrete.py, air/vae.py and air/air_model.py.  Each entry point
 * below names the reference lines whose arithmetic it replaces; INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer to a contiguous fp32 (or int32 where stated)
 *     buffer owned by the caller; nothing is allocated, cached or freed here;
 *   - launches are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream); no global state, thread-safe per stream, graph-capturable;
 *   - noise is always an input, never generated inside;
 *   - return 0 (AIR_OK) or a negative AIR_ERR_* code; air_last_error() gives the
 *     thread-local message.  There is NO CPU fallback: without a CUDA device every
 *     compute entry point fails with AIR_ERR_CUDA.
 */
#ifndef AIR_B200_H_
#define AIR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIR_OK 0
#define AIR_ERR_BAD_SHAPE (-1)
#define AIR_ERR_BAD_ALIGN (-2)
#define AIR_ERR_CUDA (-3)
#define AIR_ERR_UNSUPPORTED (-4)
#define AIR_ERR_NULL (-5)

typedef void *air_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------- */
int air_abi_version(void);            /* bumped on any signature change */
const char *air_last_error(void);     /* thread-local, never NULL */
/* sm count / compute capability of the current device; AIR_ERR_CUDA without a GPU */
int air_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* number of kernels this library has launched on the calling thread (bench evidence) */
int64_t air_launch_count(void);
/* host-side CRC-32C (Castagnoli) of a HOST buffer, continuing from `crc` (0 to start): the checksum of
 * TensorFlow bundle checkpoints (training.py:141, 203-207 tf.train.Saver); used by checkpoint.py */
uint32_t air_crc32c(const void *data, uint64_t nbytes, uint32_t crc);

/* ---- Spatial Transformer: air/transformer.py:18 transformer(U, theta, out_size) ----
 * U [B,H,W,C] NHWC, theta [B,6] (row-major 2x3), out [B,oh,ow,C].
 * Replaces _meshgrid (:119-136), _transform (:138-171) and _interpolate (:56-117) with
 * the same fp32 rounding sequence (no FMA contraction): bit-exact vs the oracle. */
int air_st_forward(const float *U, const float *theta, float *out, int64_t B, int H, int W, int C, int oh,
                   int ow, air_stream_t stream);

/* Backward of the above (what TF autodiff builds for transformer.py:56-171).
 * dout [B,oh,ow,C]; dtheta [B,6] always written; dU [B,H,W,C] written if non-NULL
 * (NULL for the crop ST, whose U is input data: air_model.py:330-333).
 * dU is deterministic (gather form, no atomics) for axis-aligned theta with C == 1. */
int air_st_backward(const float *U, const float *theta, const float *dout, float *dU, float *dtheta, int64_t B,
                    int H, int W, int C, int oh, int ow, air_stream_t stream);

/* ---- fused write-back + canvas: air_model.py:363-366 and :429-439 ------------------
 * canvas_out[b] = canvas_in[b] + (stop_new[b] < thr ? z[b] * ST(window[b], theta_inv[b]) : 0)
 * window [B,wh,ww] (C == 1), theta_inv [B,6], z [B], stop_new [B], canvas [B,ch,cw].
 * canvas_out may alias canvas_in (in place).  canvas_in == NULL stands for an all-zero canvas (the first loop
 * step, air_model.py:552: no memset, no read).  Bit-exact vs the oracle. */
int air_st_writeback_canvas_fwd(const float *window, const float *theta_inv, const float *z, const float *stop_new,
                                float thr, const float *canvas_in, float *canvas_out, int64_t B, int wh, int ww,
                                int ch, int cw, air_stream_t stream);

/* ---- all T loop steps of an op in one launch --------------------------------------------------------------
 * The per-step operands are stacked [T, B, ...]; the operand every step shares is passed once:
 *   air_st_forward_steps   out[t,b] = ST(U[b], theta[t,b])                 (the T attention crops of one canvas)
 *   air_st_backward_steps  dtheta[t,b] from dout[t,b]                      (U is data: no dU)
 *   air_st_writeback_canvas_bwd_steps  the fused backward of every step against the one dcanvas [B,ch,cw]; z / stop_new
 *                          point at step 0, consecutive steps are step_stride floats apart
 * Each is bit-identical to T calls of the single-step entry point, which is also the fallback when the batched
 * kernel does not apply (other sizes, B % 4 != 0, flags without AIR_WB_AXIS_ALIGNED_THETA). */
int air_st_forward_steps(const float *U, const float *theta, float *out, int64_t B, int T, int H, int W, int C, int oh, int ow,
                         air_stream_t stream);
int air_st_backward_steps(const float *U, const float *theta, const float *dout, float *dtheta, int64_t B, int T, int H, int W,
                          int C, int oh, int ow, air_stream_t stream);
int air_st_writeback_canvas_bwd_steps(const float *windows, const float *theta_inv, const float *z, const float *stop_new,
                                      int64_t step_stride, float thr, const float *dcanvas, float *dwindow, float *dtheta_inv,
                                      float *dz, int flags, int64_t B, int T, int wh, int ww, int ch, int cw, air_stream_t stream);

/* All T write-backs of the loop in one pass over the canvas:
 *   canvas_out = (((canvas_in + a_0) + a_1) + ...) + a_{T-1},  a_t = stop_new_t < thr ? z_t * ST(window_t, theta_inv_t) : +0
 * windows [T,B,wh,ww], theta_inv [T,B,6]; z / stop_new point at step 0 and consecutive steps are step_stride floats
 * apart (rows of the per-step field table).  Bit-identical to T consecutive air_st_writeback_canvas_fwd calls (which is
 * also the fallback for other sizes); canvas_in may be NULL (zeros) or alias canvas_out. */
int air_st_writeback_canvas_fwd_steps(const float *windows, const float *theta_inv, const float *z, const float *stop_new,
                                      int64_t step_stride, float thr, const float *canvas_in, float *canvas_out, int64_t B,
                                      int T, int wh, int ww, int ch, int cw, air_stream_t stream);

/* Backward: dcanvas [B,ch,cw] is d(loss)/d(canvas_out) (== d/d(canvas_in), not rewritten).
 * Writes dwindow [B,wh,ww], dtheta_inv [B,6], dz [B]; all zero for rows with stop_new >= thr.
 * flags (bit mask):
 *   AIR_WB_SIGMOID_WINDOW      the window is a sigmoid output w (vae.py:39-41) and dwindow receives the gradient
 *                              w.r.t. its PRE-sigmoid input, dwindow * w * (1 - w): SigmoidGrad fused into the store.
 *   AIR_WB_AXIS_ALIGNED_THETA  the caller builds theta_inv = [[1/s,0,-x/s],[0,1/s,-y/s]] (air_model.py:351-360) and
 *                              consumes only dtheta_inv[0], [2], [4], [5]; the gradients w.r.t. the structural zeros
 *                              ([1], [3]) are written as 0 for axis-aligned rows.  Enables the warp-specialised
 *                              kernel (28x28 window, 50x50 canvas); rows with shear/rotation still get all six.
 *   AIR_WB_REFERENCE_ROUNDING  sum what the reference's fp32 autodiff sums (gradient subgraph of transformer.py:84-116 in
 *                              model/air-model.meta): a canvas pixel outside the window meets one clipped border pixel
 *                              twice, with the weights (v - i) and (i - v); the reference multiplies the upstream
 *                              gradient into each corner term separately and the ~1e10 products of a lit, not yet
 *                              reconstructed pixel (BCE gradient -x / (0 + 1e-9)) cancel only to rounding error.  The
 *                              reference's training depends on these residues; without the flag
 *                              such pixels are skipped as the exact zeros they are on paper (the fp64 gradient).  Per
 *                              pixel the arithmetic is the graph's own (dz, dtheta_inv: all six entries); the order in
 *                              which dwindow accumulates over pixels (unspecified in the reference: UnsortedSegmentSum)
 *                              is fixed and deterministic for 28x28 windows on 50x50 canvases with an axis-aligned
 *                              theta_inv of positive scale (the model's case; dz and a 16-byte aligned dwindow
 *                              required for that kernel); any other size or theta takes the same per-pixel
 *                              arithmetic with shared-memory atomics for dwindow (order not fixed, like the
 *                              reference on a GPU). */
#define AIR_WB_SIGMOID_WINDOW 1
#define AIR_WB_AXIS_ALIGNED_THETA 2
#define AIR_WB_REFERENCE_ROUNDING 4
int air_st_writeback_canvas_bwd(const float *window, const float *theta_inv, const float *z, const float *stop_new,
                                float thr, const float *dcanvas, float *dwindow, float *dtheta_inv, float *dz,
                                int flags, int64_t B, int wh, int ww, int ch, int cw, air_stream_t stream);

/* ---- Concrete / ACT step: concrete.py:20-43 + air_model.py:380-427 -----------------
 * y = (log_odds + log(u+eps) - log(1-u+eps)) / temperature ; z = sigmoid(y) (rounded
 * half-to-even if !train) ; kl = log q(y) - log p(y) ; loss_new = loss_prev + (stop_prev<thr ? kl : 0) ;
 * stop_new = stop_prev + (1 - z) ; digits_new = digits_prev + (stop_new < thr).
 * prior_log_odds is a DEVICE scalar (it is an annealed tensor in the reference:
 * air_model.py:76-82).  Outputs may alias the matching *_prev inputs. */
int air_concrete_step_fwd(const float *log_odds, const float *u, const float *stop_prev, const float *loss_prev,
                          const int32_t *digits_prev, const float *prior_log_odds, float temperature, float thr,
                          int train, float *y, float *z, float *z_prob, float *kl, float *stop_new,
                          float *loss_new, int32_t *digits_new, int64_t B, air_stream_t stream);

/* dlog_odds from dz (gradient reaching z; ignored when !train: tf.round has no gradient)
 * and dkl (gradient reaching kl, i.e. dloss * [stop_prev < thr]). */
int air_concrete_step_bwd(const float *log_odds, const float *y, const float *z, const float *dz, const float *dkl,
                          const float *prior_log_odds, float temperature, int train, float *dlog_odds, int64_t B,
                          air_stream_t stream);

/* ---- dense contractions: MatMul + BiasAdd (+ activation) ---------------------------
 * Replaces the dense layers of the reference -- tf.contrib.layers.fully_connected
 * (air/vae.py:13-34, air/air_model.py:292-316, 372-376), BasicLSTMCell's linear map
 * (air/air_model.py:284-286, 539) -- and the MatMul gradients TF autodiff derives from them.
 *
 *   C[M,N] = epilogue( (Cinit[M,N] + op(A)[M,K] * op(B)[K,N]) + bias[N] )
 *
 * op(A) = A (row-major [M,K], leading dim lda) or A^T (A stored [K,M]) when transA;
 * op(B) = B (row-major [K,N], ldb)            or B^T (B stored [N,K]) when transB.
 * Cinit / bias / aux may be NULL; Cinit and aux use ldc; Cinit may alias C.
 * mode AIR_GEMM_FP32_EXACT: SIMT FFMA, one FMA per k in strictly increasing k order
 *   (bit-reproducible, equals oracle_gemm_seq_fma);
 * mode AIR_GEMM_TF32: tcgen05 tensor cores, operands cut to TF32 (top 19 bits), FP32 accumulation in TMEM
 *   (~1e-3 relative: the throughput mode);
 * mode AIR_GEMM_TF32X3: tcgen05 tensor cores at FP32 accuracy ("3xTF32"): every operand is split in-kernel into
 *   hi + lo TF32 parts, three MMAs per k-step (hi*hi, lo*hi, hi*lo), FP32 accumulation; error 1-3e-6 norm-wise,
 *   the same order as an FP32 FMA chain of that length.  This is the mode the model-level parity bars
 *   (<= 1e-5 outputs / ELBO, <= 1e-4 gradients) are checked in at tensor-core speed.  What remains over FP32 is the
 *   tensor core's accumulate step, which rounds toward zero (~0.5 ulp of bias per 8-deep MMA): the kernels keep the
 *   hi*hi accumulation chains short (three round-robin TMEM accumulators, or K split through the caller's workspace
 *   -- air_gemm_ws -- and summed in FP32 round-to-nearest); a K > ~8192 GEMM WITHOUT a workspace is ~1e-5. */
#define AIR_EPI_NONE 0
#define AIR_EPI_RELU 1
#define AIR_EPI_SOFTPLUS 2      /* tf.nn.softplus: x>13.94->x, x<-13.94->exp(x), else log(exp(x)+1) */
#define AIR_EPI_MUL_DRELU 3     /* out = v * (aux > 0)            (ReLU backward, aux = ReLU output) */
#define AIR_EPI_MUL_DSOFTPLUS 4 /* out = v * (1 - exp(-aux))      (softplus backward, aux = softplus output) */
#define AIR_EPI_SIGMOID_NOISE 5 /* out = sigmoid(v + aux * epi_param): vae.py:36-41 (aux = N(0,1) noise, epi_param = likelihood std) */
#define AIR_EPI_SIGMOID_RNG 6   /* the same with the noise GENERATED in the epilogue: aux points at the device RNG state
                                 * (air_rng_state_t, cast to const float *); element (m, n) gets the N(0,1) sample
                                 * m * N + n of stream AIR_RNG_LIKE for the state's current step counter.  No noise
                                 * tensor is written or read (38 MB per train step at B = 4096). */
#define AIR_GEMM_FP32_EXACT 0
#define AIR_GEMM_TF32 1
#define AIR_GEMM_TF32X3 2
/* The tensor-core modes need 16-byte aligned A / B with lda, ldb multiples of 4 (TMA). */
int air_gemm(const float *A, const float *B, float *C, const float *Cinit, const float *bias, const float *aux,
             int64_t M, int N, int K, int lda, int ldb, int ldc, int transA, int transB, int epilogue, int mode,
             air_stream_t stream);
/* same, with the scalar parameter some epilogues take (AIR_EPI_SIGMOID_NOISE: the likelihood std) */
int air_gemm_ex(const float *A, const float *B, float *C, const float *Cinit, const float *bias, const float *aux,
                int64_t M, int N, int K, int lda, int ldb, int ldc, int transA, int transB, int epilogue,
                float epi_param, int mode, air_stream_t stream);
/* same, with a caller-owned scratch buffer (device, 16-byte aligned, workspace_floats floats; NULL = none).
 * GEMMs with few output tiles and a long K (the weight gradients: reduction over the batch) then run split-K
 * through it -- as many splits as fit, M*N floats each, summed in a fixed order by a second pass -- instead of
 * leaving most SMs idle.  The library keeps no pointer to it after the call's kernels; calls sharing one
 * buffer must be stream-ordered.  M*N*32 floats of the largest such GEMM allow every split the heuristic wants. */
int air_gemm_ws(const float *A, const float *B, float *C, const float *Cinit, const float *bias, const float *aux,
                int64_t M, int N, int K, int lda, int ldb, int ldc, int transA, int transB, int epilogue,
                float epi_param, int mode, float *workspace, int64_t workspace_floats, air_stream_t stream);

/* ---- sampling noise generated on the device, where it is consumed -------------------------------------------------
 * The reference samples with tf.random_normal / tf.random_uniform inside the graph (air_model.py:123-128, vae.py:23, 37,
 * concrete.py:23).  Every compute entry point takes its noise as an input (parity runs inject it); for production
 * steps the noise is a pure function of (seed, step counter, stream, element index): Philox4x32-10 + a Box-Muller
 * transform built from correctly rounded fp32 operations only, restated bit for bit in oracle/rng_oracle.py.
 * State: 4 x uint64 on the DEVICE, caller-owned: [0] seed, [1] step counter, [2] scratch (zero-initialised), [3] unused.
 *
 * air_noise_fill: advances the step counter by one and fills the step's small noise tensors from the new counter:
 *   scale [T*B] and shift [T*B*2] and vae_latent [T*B*L] ~ N(0,1), concrete_u [T*B] ~ U[0,1) (each nullable).  The big
 *   one -- the VAE likelihood noise [T*B, window^2] -- is never materialised: AIR_EPI_SIGMOID_RNG generates it in the GEMM
 *   epilogue from the same state.  One launch; graph-capturable (the counter lives on the device).
 * air_rng_normals: out[i] = N(0,1) sample i of `stream` at the state's CURRENT counter (no advance): test / inspection
 *   helper, e.g. to reproduce the samples an AIR_EPI_SIGMOID_RNG epilogue used. */
typedef struct air_rng_state { uint64_t seed, counter, scratch, unused; } air_rng_state_t;
#define AIR_RNG_SCALE 1
#define AIR_RNG_SHIFT 2
#define AIR_RNG_LATENT 3
#define AIR_RNG_CONCRETE 4
#define AIR_RNG_LIKE 5
int air_noise_fill(air_rng_state_t *state, float *scale, float *shift, float *vae_latent, float *concrete_u, int64_t TB,
                   int L, air_stream_t stream);
int air_rng_normals(const air_rng_state_t *state, int rng_stream, float *out, int64_t n, air_stream_t stream);
int air_rng_uniforms(const air_rng_state_t *state, int rng_stream, float *out, int64_t n, air_stream_t stream);

/* ---- fused model-specific elementwise kernels (air/air_model.py loop body) ----------
 * Hyper-parameters that are plain Python floats in the reference constructor
 * (air_model.py:13-22).  Passed by pointer to a HOST struct. */
typedef struct air_hyper {
  float scale_prior_mean, scale_prior_variance;
  float shift_prior_mean, shift_prior_variance;
  float vae_prior_mean, vae_prior_variance;
  float vae_likelihood_std;
  float z_pres_temperature, stopping_threshold;
  int32_t train; /* 0: z_pres is rounded (air_model.py:389-390) */
} air_hyper_t;

/* Per-step table of per-image scalars, laid out [AIR_NF][B] (struct of arrays). */
enum air_field {
  AIR_F_SCALE_MEAN = 0, AIR_F_SCALE_LV, AIR_F_SHIFT_MEAN_X, AIR_F_SHIFT_MEAN_Y, AIR_F_SHIFT_LV_X, AIR_F_SHIFT_LV_Y,
  AIR_F_LOG_ODDS,                       /* the 7 head outputs */
  AIR_F_S, AIR_F_X, AIR_F_Y,            /* sigmoid / tanh of the Gaussian samples (air_model.py:301, 318) */
  AIR_F_YPRE, AIR_F_Z, AIR_F_ZPROB,     /* Concrete pre-sigmoid sample, z_pres, sigmoid(log_odds) */
  AIR_F_KL_Z, AIR_F_KL_SCALE, AIR_F_KL_SHIFT, AIR_F_KL_VAE,
  AIR_F_STOP_PREV, AIR_F_STOP_NEW,      /* stopping_sum before / after this step */
  AIR_NF = 20
};

/* BasicLSTMCell pointwise part (air_model.py:286, 539): gates [B,4H] = [x,h]K + b, order
 * i,j,f,o, forget bias 1.0.  c_prev NULL = zero state. */
int air_lstm_fwd(const float *gates, const float *c_prev, float *c_new, float *h_new, int64_t B, int H,
                 air_stream_t stream);
/* dgates [B,4H], dc_prev [B,H]; dc_new NULL = 0; dgates_sum (nullable) += dgates. */
int air_lstm_bwd(const float *gates, const float *c_prev, const float *c_new, const float *dh, const float *dc_new,
                 float *dgates, float *dc_prev, float *dgates_sum, int64_t B, int H, air_stream_t stream);

/* Everything between the hidden head layers and the VAE for one step
 * (air_model.py:294-327, 353-356, 376-427, 441-477): the 7 head outputs (k-sequential FMA
 * dot products over `hidden` [B,5*HU], post-ReLU, heads ordered scale/mean, scale/log_variance,
 * shift/mean, shift/log_variance, z_pres/log_odds; w_out [7,HU], b_out [7]), Gaussian sampling +
 * sigmoid/tanh, scale & shift KLs, theta / theta_inv, and the fused Concrete/ACT step.
 * stop / loss / digits are updated in place; fields [AIR_NF,B]; theta, theta_inv [B,6]. */
int air_heads_fwd(const float *hidden, const float *w_out, const float *b_out, const float *noise_scale,
                  const float *noise_shift, const float *u, const float *prior_log_odds, const air_hyper_t *hyper,
                  float *stop, float *loss, int32_t *digits, float *fields, float *theta, float *theta_inv, int64_t B,
                  int HU, air_stream_t stream);
/* Backward of air_heads_fwd.  dtheta / dtheta_inv [B,6], dz [B] come from the ST kernels;
 * dloss is d(total)/d(running_loss[b]) (1/batch).  Writes dhidden [B,5*HU] (ReLU mask applied)
 * and accumulates (if accumulate) or writes d(w_out) [7,HU] and d(b_out) [7]; deterministic.
 * workspace: at least air_heads_bwd_workspace(B, HU) floats = R rows of (7*HU + 7) per-CTA partial sums.
 * dw_out == db_out == NULL: only the partials are written; the caller sums the rows of all its steps once with
 * air_reduce_rows (stride 7*HU + 7: d(w_out) at column 0, d(b_out) at column 7*HU). */
int64_t air_heads_bwd_workspace(int64_t B, int HU);
int air_heads_bwd(const float *hidden, const float *w_out, const float *noise_scale, const float *noise_shift,
                  const float *fields, const float *dtheta, const float *dtheta_inv, const float *dz,
                  const float *prior_log_odds, const air_hyper_t *hyper, float dloss, float *dhidden, float *dw_out,
                  float *db_out, int accumulate, float *workspace, int64_t B, int HU, air_stream_t stream);

/* VAE latent (vae.py:22-24, air_model.py:479-493): ml [B,2L] = (mean | log_variance);
 * sample [B,L] with leading dimension ld_sample >= L (padded so it can feed the TMA GEMM)
 * = mean + noise*sqrt(exp(lv)); KL vs N(prior) -> fields[AIR_F_KL_VAE];
 * loss += (stop_new < thr ? kl : 0). */
int air_vae_latent_fwd(const float *ml, const float *noise, const air_hyper_t *hyper, float *sample, int ld_sample,
                       float *fields, float *loss, int64_t B, int L, air_stream_t stream);
int air_vae_latent_bwd(const float *ml, const float *noise, const float *dsample, const float *fields,
                       const air_hyper_t *hyper, float dloss, float *dml, int64_t B, int L, air_stream_t stream);

/* vae.py:36-41: out = sigmoid(gen + noise*std); backward dgen = dout*out*(1-out) (may be in place). */
int air_sigmoid_noise_fwd(const float *gen, const float *noise, float std, float *out, int64_t n, air_stream_t stream);
int air_sigmoid_bwd(const float *out, const float *dout, float *dgen, int64_t n, air_stream_t stream);

/* air_model.py:580-590: r = max(min(canvas,1),0); rec_loss[b] = -sum(x log(r+1e-9) + (1-x) log(1-r+1e-9)).
 * recon (nullable) receives r; dcanvas (nullable) receives dscale * d rec_loss / d canvas
 * (gradient passes at r == 0 and r == 1, TF minimum/maximum tie rules). */
int air_bce_loss(const float *canvas, const float *x, float *recon, float *rec_loss, float *dcanvas, float dscale,
                 int64_t B, int N, air_stream_t stream);

/* air_model.py:593-611: out[0] = mean(running_loss + rec_loss), out[1] = mean(target == digits);
 * loss_per_item (nullable) [B]. */
int air_finalize_loss(const float *running_loss, const float *rec_loss, const int32_t *digits, const int32_t *target,
                      float *out, float *loss_per_item, int64_t B, air_stream_t stream);

/* out[n] (+)= sum_b X[b, n]  (bias gradients); deterministic.  workspace >= air_colsum_workspace(B, N) floats,
 * zero-initialised once by the caller (the kernel leaves its counters zeroed). */
int64_t air_colsum_workspace(int64_t B, int N);
int air_colsum(const float *X, int ld, float *out, int accumulate, float *workspace, int64_t B, int N,
               air_stream_t stream);

/* Several column sums in one launch pair (all bias gradients of a train step).  items[i]: out[N] (+)= column sums
 * of X [rows, N] (leading dimension ld).  At most 16 items; workspace >= air_colsum_multi_workspace(items, n). */
typedef struct air_colsum_item {
  const float *X;
  float *out;
  int64_t rows;
  int ld, N, accumulate;
} air_colsum_item_t;
int64_t air_colsum_multi_workspace(const air_colsum_item_t *items, int n_items);
int air_colsum_multi(const air_colsum_item_t *items, int n_items, float *workspace, air_stream_t stream);

/* out[e] (+)= sum_{r < R} partials[r * stride + e] for e < n, fixed summation order (deterministic).  The second
 * stage of air_heads_bwd when that is called with dw_out == db_out == NULL (per-step partials reduced once per
 * train step instead of once per loop step). */
int air_reduce_rows(const float *partials, int R, int stride, int n, float *out, int accumulate, air_stream_t stream);

/* air_model.py:673, 692: tf.clip_by_global_norm + tf.train.AdamOptimizer.apply_gradients on a flat
 * parameter buffer of n floats.  state (device, 8 floats): [0] beta1^t, [1] beta2^t, [2] global_step,
 * [3] last global norm, [4] learning rate; updated in place (t -> t+1).  clip_norm <= 0 disables clipping.
 * grad_scale multiplies the gradients first (1/world_size after an allreduce-sum).
 * workspace >= air_adam_workspace(n) floats. */
int64_t air_adam_workspace(int64_t n);
int air_adam_step(float *params, const float *grads, float *m, float *v, float *state, float clip_norm, float beta1,
                  float beta2, float epsilon, float grad_scale, float *workspace, int64_t n, air_stream_t stream);
/* The same with flags.  AIR_ADAM_SKIP_NONFINITE: when the global gradient norm is not finite the step is a no-op
 * (parameters, Adam slots, beta powers and global_step untouched; state[3] = the norm, state[5] += 1 counts the skip).
 * The reference has no such guard: there an overflowed gradient (a window that collapsed to ~1e-14 of the canvas makes
 * the un-cancelled corner products of air/transformer.py:108-116 exceed fp32) turns every variable into NaN for good. */
#define AIR_ADAM_SKIP_NONFINITE 1
int air_adam_step_ex(float *params, const float *grads, float *m, float *v, float *state, float clip_norm, float beta1,
                     float beta2, float epsilon, float grad_scale, float *workspace, int64_t n, int flags,
                     air_stream_t stream);

/* air_model.py:94-121: value = init * factor^(step/iters) [floor if staircase], clamped to
 * [min,max] (NaN = no bound), optional log(value + 1e-9); step read from adam state[2].
 * Writes the device scalar *out. */
int air_anneal(const float *state, float init, float factor, float iters, int staircase, float vmin, float vmax,
               int take_log, float *out, air_stream_t stream);

/* ---- synthetic input canvases, generated on the device -------------------------------
 * Stand-in for the data set of multi_mnist.py (needs the MNIST download) with generate_multi_image's placement
 * semantics (multi_mnist.py:82-183, use_pixel_overlap, gap = margin = 0): 0..max_digits "digits" per canvas -- here
 * stroke-like blobs (14-24 x 10-24 pixel frames, values in (0, 1], background exactly 0.0) cropped to their non-empty
 * bounding box like crop_non_empty (:36-43) -- each placed at a uniform position, x drawn before y (:143-144); the
 * first always fits, later ones are re-drawn up to 100 times (:141) until no pixel overlaps the canvas (pixels_overlap,
 * :61-65); a digit that cannot be placed restarts the whole canvas with fresh digits (:95-171), so counts[b] always
 * equals the number of digits drawn.  positions [B, max_digits, 2] = (x, y) and boxes [B, max_digits, 2] = (w, h) of
 * the placed digits (:165-166), zero padded; both nullable.
 * Counter-based RNG: image first_index + b depends only on (seed, first_index + b), so ranks generate their own
 * shards of one global data set.  images [B, canvas_size^2], counts [B].  Bit-exact vs oracle/synth_oracle.py, which
 * is pinned to the reference's own generate_multi_image (tests/test_reference_source.py). */
int air_synth_canvases(uint64_t seed, int64_t first_index, float *images, int32_t *counts, int64_t B, int canvas_size,
                       int max_digits, air_stream_t stream);
int air_synth_canvases_ex(uint64_t seed, int64_t first_index, float *images, int32_t *counts, int32_t *positions,
                          int32_t *boxes, int64_t B, int canvas_size, int max_digits, air_stream_t stream);

/* Zero up to 8 device buffers (16-byte aligned, sizes multiples of 16 bytes) with ONE launch: the accumulators a step
 * starts from (stopping sums, running loss, digit counts of air_model.py:550-553; the summed gate gradients).  `buffers`
 * and `nbytes` are HOST arrays. */
int air_zero_buffers(void *const *buffers, const int64_t *nbytes, int n, air_stream_t stream);

/* uint8 canvases -> fp32 on the device: dst[i] = fl(float(src[i]) * fl(1/255)), the scaling of the MNIST loader behind
 * multi_mnist.py (tensorflow's input_data: numpy.multiply(images.astype(float32), 1.0 / 255.0)).  The reference's
 * default data set (digits pasted without overlap, no rescaling / rotation) only holds those 256 values, so its
 * canvases can cross PCIe as bytes -- a quarter of the fp32 traffic -- and be expanded bit for bit next to the model. */
int air_expand_u8(const uint8_t *src, float *dst, int64_t n, air_stream_t stream);

/* ---- host-side input path: the reference's multi-MNIST TFRecord files -----------------------------------------
 * multi_mnist.py:186-212 writes tf.train.Example records (features height, width, digits: int64; indices, positions,
 * boxes, labels: int32 bytes; image: canvas^2 float32 bytes); training reads them with TFRecordReader +
 * parse_single_example + shuffle_batch on 4 threads (multi_mnist.py:228-251, training.py:28, 76-81).  These three
 * entry points take HOST pointers (no CUDA); tfrecords.py builds the batched reader on them.
 *
 * air_tfrecord_index: one pass over the bytes of a .tfrecords file (e.g. a memory map): checks the framing
 *   (u64 length, masked CRC-32C of the length, payload, masked CRC-32C of the payload; CRCs only if verify_crc), parses
 *   each Example and stores the byte offset / length of its `image` payload and its `digits` value.  Returns the number
 *   of records (with all three output pointers NULL: only counts, for sizing) or a negative AIR_ERR_* code.
 * air_shuffle_order: the order in which a shuffle queue with min_after_dequeue = buffer, fed with record indices
 *   0..n-1 `epochs` times, emits them (order[n * epochs]); deterministic in seed.
 * air_gather_rows: dst[i, :] = src[offsets[i] : offsets[i] + row_bytes], i < n, on up to n_threads host threads
 *   (dst is typically a pinned batch buffer). */
int64_t air_tfrecord_index(const uint8_t *buf, uint64_t nbytes, int verify_crc, int64_t max_records, uint64_t *image_off,
                           uint32_t *image_len, int32_t *digits);
int air_shuffle_order(int64_t n, int64_t buffer, int epochs, uint64_t seed, int64_t *order);
int air_gather_rows(const uint8_t *src, const uint64_t *offsets, int64_t n, uint64_t row_bytes, uint8_t *dst, int n_threads);

/* ---- CNN front-end of AIRModel(cnn=True): air_model.py:510-535 -------------------------
 * tf.layers.conv2d(filters=8, kernel_size=5, padding="same", activation=relu) optionally followed by
 * tf.layers.max_pooling2d(pool_size=2, strides=2) ("valid": 25 -> 12), NHWC.
 * in [B,H,W,cin], w [5,5,cin,cout] (TF HWIO), bias [cout], out [B,H',W',cout] with H' = pool ? H/2 : H;
 * argmax [B,H',W',cout] uint8 (pool only): position 0..3 of the window maximum, first maximum wins.
 * Built for the reference's three layers: (cin,H,W,pool) = (1,50,50,1), (8,25,25,1), (8,12,12,0), cout = 8;
 * anything else returns AIR_ERR_UNSUPPORTED. */
int air_conv5x5_fwd(const float *in, const float *w, const float *bias, float *out, uint8_t *argmax, int64_t B, int H,
                    int W, int cin, int cout, int pool, air_stream_t stream);
/* Backward (ReluGrad, MaxPoolGrad and Conv2DBackpropFilter/Input of TF autodiff): dout [B,H',W',cout];
 * dw [5,5,cin,cout] and db [cout] are written (accumulate == 0) or added to; din [B,H,W,cin] is written if
 * non-NULL (NULL for the first layer, whose input is data).  Deterministic (fixed-order reductions).
 * workspace >= air_conv5x5_bwd_workspace(B, cin, cout) floats. */
int64_t air_conv5x5_bwd_workspace(int64_t B, int cin, int cout);
int air_conv5x5_bwd(const float *in, const float *w, const float *out, const uint8_t *argmax, const float *dout,
                    float *din, float *dw, float *db, int accumulate, float *workspace, int64_t B, int H, int W, int cin,
                    int cout, int pool, air_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AIR_B200_H_ */
