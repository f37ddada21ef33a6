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
 * canvas_out may alias canvas_in (in place).  Bit-exact vs the oracle. */
int air_st_writeback_canvas_fwd(const float *window, const float *theta_inv, const float *z, const float *stop_new,
                                float thr, const float *canvas_in, float *canvas_out, int64_t B, int wh, int ww,
                                int ch, int cw, air_stream_t stream);

/* Backward: dcanvas [B,ch,cw] is d(loss)/d(canvas_out) (== d/d(canvas_in), not rewritten).
 * Writes dwindow [B,wh,ww], dtheta_inv [B,6], dz [B]; all zero for rows with stop_new >= thr. */
int air_st_writeback_canvas_bwd(const float *window, const float *theta_inv, const float *z, const float *stop_new,
                                float thr, const float *dcanvas, float *dwindow, float *dtheta_inv, float *dz,
                                int64_t B, int wh, int ww, int ch, int cw, air_stream_t stream);

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

#ifdef __cplusplus
}
#endif
#endif /* AIR_B200_H_ */
