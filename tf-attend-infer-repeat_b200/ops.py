"""Tensor-level wrappers over the C ABI (no autograd): each function launches one
entry point of libair_b200.so on torch's current CUDA stream."""
from __future__ import annotations

import ctypes

import torch

from . import _cabi as C
from ._cabi import check, lib, ptr, stream


def _ld(t):
    if t.dim() != 2 or t.stride(1) != 1:
        raise C.AirError("GEMM operands must be 2-D with unit inner stride")
    return t.stride(0)


def _p2(t):
    """pointer of a 2-D row-strided tensor (view allowed)."""
    if t is None:
        return None
    if not t.is_cuda or t.dtype != torch.float32:
        raise C.AirError("air_b200 ops need float32 CUDA tensors (no CPU fallback)")
    return ctypes.c_void_p(t.data_ptr())


class GemmMode(int):
    """An air_gemm mode (AIR_GEMM_*) that carries the caller's split-K scratch tensor: ``GemmMode(mode, ws)`` can be
    passed wherever a plain mode integer is accepted and makes every gemm() call hand ``ws`` to the library."""

    def __new__(cls, mode, ws=None):
        self = super().__new__(cls, int(mode))
        self.ws = ws
        return self


def gemm(A, B, out, Cinit=None, bias=None, aux=None, tA=False, tB=False, epi=C.EPI_NONE, mode=0, epi_param=0.0, ws=None):
    """out[M,N] = epi((Cinit + op(A) op(B)) + bias); see include/air_b200.h (air_gemm_ws).
    ``ws``: optional float32 CUDA scratch tensor on the operands' device: lets long-K / few-tile GEMMs run split-K.
    ``aux``: the [M, N] second operand of an epilogue, or -- EPI_SIGMOID_RNG -- the int64 device RNG state (rng_state())."""
    M, N = out.shape
    K = A.shape[0] if tA else A.shape[1]
    kb = B.shape[1] if tB else B.shape[0]
    if kb != K or (A.shape[1] if tA else A.shape[0]) != M or (B.shape[0] if tB else B.shape[1]) != N:
        raise C.AirError(f"gemm shape mismatch: A{tuple(A.shape)} tA={tA} B{tuple(B.shape)} tB={tB} out{tuple(out.shape)}")
    ldc = _ld(out)
    if epi == C.EPI_SIGMOID_RNG:
        if aux is None or aux.dtype != torch.int64 or aux.numel() != 4 or not aux.is_cuda:
            raise C.AirError("EPI_SIGMOID_RNG takes the device RNG state (ops.rng_state) as aux")
        aux_p = ptr(aux)
    else:
        aux_p = _p2(aux)
    for t in (Cinit, None if epi == C.EPI_SIGMOID_RNG else aux):
        if t is not None and (_ld(t) != ldc or t.shape != out.shape):
            raise C.AirError("Cinit / aux must have the layout of out")
    if ws is None:
        ws = getattr(mode, "ws", None)
    if ws is not None and ws.device != out.device:
        raise C.AirError("the GEMM workspace must live on the device of the operands")
    check(lib().air_gemm_ws(_p2(A), _p2(B), _p2(out), _p2(Cinit), ptr(bias), aux_p, M, N, K, _ld(A), _ld(B), ldc,
                            int(tA), int(tB), epi, float(epi_param), int(mode), ptr(ws), 0 if ws is None else ws.numel(),
                            stream()), "air_gemm")
    return out


def lstm_fwd(gates, c_prev, c_new, h_new):
    B, H4 = gates.shape
    check(lib().air_lstm_fwd(ptr(gates), ptr(c_prev), ptr(c_new), ptr(h_new), B, H4 // 4, stream()), "air_lstm_fwd")


def lstm_bwd(gates, c_prev, c_new, dh, dc_new, dgates, dc_prev, dgates_sum):
    B, H4 = gates.shape
    check(lib().air_lstm_bwd(ptr(gates), ptr(c_prev), ptr(c_new), ptr(dh), ptr(dc_new), ptr(dgates), ptr(dc_prev),
                             ptr(dgates_sum), B, H4 // 4, stream()), "air_lstm_bwd")


def heads_fwd(hidden, w_out, b_out, n_scale, n_shift, u, prior, hyper, stop, loss, digits, fields, theta, theta_inv):
    B = hidden.shape[0]
    HU = w_out.shape[1]
    check(lib().air_heads_fwd(ptr(hidden), ptr(w_out), ptr(b_out), ptr(n_scale), ptr(n_shift), ptr(u), ptr(prior),
                              ctypes.byref(hyper), ptr(stop), ptr(loss), ptr(digits), ptr(fields), ptr(theta),
                              ptr(theta_inv), B, HU, stream()), "air_heads_fwd")


def heads_bwd(hidden, w_out, n_scale, n_shift, fields, dtheta, dtheta_inv, dz, prior, hyper, dloss, dhidden, dw_out,
              db_out, accumulate, workspace):
    B = hidden.shape[0]
    HU = w_out.shape[1]
    # dw_out = db_out = None: per-CTA partials only (summed later over all steps with reduce_rows)
    check(lib().air_heads_bwd(ptr(hidden), ptr(w_out), ptr(n_scale), ptr(n_shift), ptr(fields), ptr(dtheta),
                              ptr(dtheta_inv), ptr(dz), ptr(prior), ctypes.byref(hyper), float(dloss), ptr(dhidden),
                              ptr(dw_out), ptr(db_out), int(accumulate), ptr(workspace), B, HU, stream()),
          "air_heads_bwd")


def vae_latent_fwd(ml, noise, hyper, sample, fields, loss):
    B, L2 = ml.shape
    check(lib().air_vae_latent_fwd(ptr(ml), ptr(noise), ctypes.byref(hyper), _p2(sample), _ld(sample), ptr(fields),
                                   ptr(loss), B, L2 // 2, stream()), "air_vae_latent_fwd")


def vae_latent_bwd(ml, noise, dsample, fields, hyper, dloss, dml):
    B, L2 = ml.shape
    check(lib().air_vae_latent_bwd(ptr(ml), ptr(noise), ptr(dsample), ptr(fields), ctypes.byref(hyper), float(dloss),
                                   ptr(dml), B, L2 // 2, stream()), "air_vae_latent_bwd")


def sigmoid_noise_fwd(gen, noise, std, out):
    check(lib().air_sigmoid_noise_fwd(ptr(gen), ptr(noise), float(std), ptr(out), gen.numel(), stream()),
          "air_sigmoid_noise_fwd")


def sigmoid_bwd(out, dout, dgen):
    check(lib().air_sigmoid_bwd(ptr(out), ptr(dout), ptr(dgen), out.numel(), stream()), "air_sigmoid_bwd")


def bce_loss(canvas, x, recon, rec_loss, dcanvas, dscale):
    B, N = canvas.shape
    check(lib().air_bce_loss(ptr(canvas), ptr(x), ptr(recon), ptr(rec_loss), ptr(dcanvas), float(dscale), B, N,
                             stream()), "air_bce_loss")


def finalize_loss(running_loss, rec_loss, digits, target, out, loss_per_item=None):
    check(lib().air_finalize_loss(ptr(running_loss), ptr(rec_loss), ptr(digits), ptr(target), ptr(out),
                                  ptr(loss_per_item), running_loss.shape[0], stream()), "air_finalize_loss")


def colsum(X, out, accumulate, workspace):
    B, N = X.shape
    check(lib().air_colsum(_p2(X), _ld(X), ptr(out), int(accumulate), ptr(workspace), B, N, stream()), "air_colsum")


def colsum_items(pairs):
    """[(X [rows,N] (row-strided view allowed), out [N], accumulate), ...] -> ctypes array for colsum_multi."""
    arr = (C.ColsumItem * len(pairs))()
    for a, (X, out, acc) in zip(arr, pairs):
        a.X, a.out = X.data_ptr(), out.data_ptr()
        a.rows, a.ld, a.N, a.accumulate = X.shape[0], _ld(X), X.shape[1], int(acc)
        if not (X.is_cuda and out.is_cuda and X.dtype == out.dtype == torch.float32 and out.is_contiguous()):
            raise C.AirError("colsum_multi needs float32 CUDA tensors (no CPU fallback)")
    return arr


def colsum_multi_workspace(items):
    return int(lib().air_colsum_multi_workspace(items, len(items)))


def colsum_multi(items, workspace):
    """All column sums of ``items`` (from colsum_items) in one launch pair."""
    check(lib().air_colsum_multi(items, len(items), ptr(workspace), stream()), "air_colsum_multi")


def reduce_rows(partials, R, stride, n, out, accumulate=False):
    check(lib().air_reduce_rows(ptr(partials), R, stride, n, ptr(out), int(accumulate), stream()), "air_reduce_rows")


def adam_step(params, grads, m, v, state, clip_norm, beta1, beta2, eps, grad_scale, workspace, skip_nonfinite=False):
    """skip_nonfinite: AIR_ADAM_SKIP_NONFINITE -- a step whose global gradient norm is not finite changes nothing
    (state[5] counts such steps); the reference would write NaN into every variable."""
    check(lib().air_adam_step_ex(ptr(params), ptr(grads), ptr(m), ptr(v), ptr(state), float(clip_norm or 0.0), beta1,
                                 beta2, eps, float(grad_scale), ptr(workspace), params.numel(), 1 if skip_nonfinite else 0,
                                 stream()), "air_adam_step")


def anneal(state, schedule, out):
    nan = float("nan")
    check(lib().air_anneal(ptr(state), float(schedule["init"]), float(schedule["factor"]), float(schedule["iters"]),
                           int(bool(schedule.get("staircase", False))), float(schedule.get("min", nan)),
                           float(schedule.get("max", nan)), int(bool(schedule.get("log", False))), ptr(out), stream()),
          "air_anneal")


def st_forward(U, theta, out, H, W, Cc, oh, ow):
    check(lib().air_st_forward(ptr(U), ptr(theta), ptr(out), U.shape[0], H, W, Cc, oh, ow, stream()), "air_st_forward")


def st_backward(U, theta, dout, dU, dtheta, H, W, Cc, oh, ow):
    check(lib().air_st_backward(ptr(U), ptr(theta), ptr(dout), ptr(dU), ptr(dtheta), U.shape[0], H, W, Cc, oh, ow,
                                stream()), "air_st_backward")


def writeback_canvas_fwd(window, theta_inv, z, stop_new, thr, canvas_in, canvas_out, wh, ww, ch, cw):
    check(lib().air_st_writeback_canvas_fwd(ptr(window), ptr(theta_inv), ptr(z), ptr(stop_new), float(thr),
                                            ptr(canvas_in), ptr(canvas_out), window.shape[0], wh, ww, ch, cw, stream()),
          "air_st_writeback_canvas_fwd")


def st_forward_steps(U, theta, out, H, W, C_, oh, ow):
    """out[t,b] = ST(U[b], theta[t,b]) for all T steps in one launch (theta [T,B,6], out [T,B,oh*ow*C])."""
    T, B = theta.shape[0], theta.shape[1]
    check(lib().air_st_forward_steps(ptr(U), ptr(theta), ptr(out), B, T, H, W, C_, oh, ow, stream()), "air_st_forward_steps")


def st_backward_steps(U, theta, dout, dtheta, H, W, C_, oh, ow):
    T, B = theta.shape[0], theta.shape[1]
    check(lib().air_st_backward_steps(ptr(U), ptr(theta), ptr(dout), ptr(dtheta), B, T, H, W, C_, oh, ow, stream()),
          "air_st_backward_steps")


def writeback_canvas_bwd_steps(windows, theta_inv, z0, stop0, step_stride, thr, dcanvas, dwindow, dtheta_inv, dz, wh, ww, ch, cw,
                               window_is_sigmoid=False, axis_aligned_theta=False, reference_rounding=False):
    T, B = windows.shape[0], windows.shape[1]
    flags = (1 if window_is_sigmoid else 0) | (2 if axis_aligned_theta else 0) | (4 if reference_rounding else 0)
    check(lib().air_st_writeback_canvas_bwd_steps(ptr(windows), ptr(theta_inv), ptr(z0), ptr(stop0), int(step_stride), float(thr),
                                                  ptr(dcanvas), ptr(dwindow), ptr(dtheta_inv), ptr(dz), flags, B, T, wh, ww, ch, cw,
                                                  stream()), "air_st_writeback_canvas_bwd_steps")


def writeback_canvas_fwd_steps(windows, theta_inv, z0, stop0, step_stride, thr, canvas_in, canvas_out, wh, ww, ch, cw):
    """All T write-backs in one pass: windows [T,B,wh*ww], theta_inv [T,B,6]; z0 / stop0 = the step-0 rows, consecutive
    steps ``step_stride`` floats apart."""
    T, B = windows.shape[0], windows.shape[1]
    check(lib().air_st_writeback_canvas_fwd_steps(ptr(windows), ptr(theta_inv), ptr(z0), ptr(stop0), int(step_stride),
                                                  float(thr), ptr(canvas_in), ptr(canvas_out), B, T, wh, ww, ch, cw, stream()),
          "air_st_writeback_canvas_fwd_steps")


def writeback_canvas_bwd(window, theta_inv, z, stop_new, thr, dcanvas, dwindow, dtheta_inv, dz, wh, ww, ch, cw,
                         window_is_sigmoid=False, axis_aligned_theta=False, reference_rounding=False):
    """flags of include/air_b200.h: AIR_WB_SIGMOID_WINDOW = 1, AIR_WB_AXIS_ALIGNED_THETA = 2 (dtheta_inv[1], [3] := 0),
    AIR_WB_REFERENCE_ROUNDING = 4 (out-of-window pixels contribute their un-cancelled fp32 corner terms)."""
    flags = (1 if window_is_sigmoid else 0) | (2 if axis_aligned_theta else 0) | (4 if reference_rounding else 0)
    check(lib().air_st_writeback_canvas_bwd(ptr(window), ptr(theta_inv), ptr(z), ptr(stop_new), float(thr),
                                            ptr(dcanvas), ptr(dwindow), ptr(dtheta_inv), ptr(dz), flags,
                                            window.shape[0], wh, ww, ch, cw, stream()), "air_st_writeback_canvas_bwd")


def conv5x5_fwd(x, w, b, out, argmax, H, W, cin, cout, pool):
    """NHWC 5x5 'same' conv + ReLU (+ 2x2/2 max-pool): air_model.py:510-535.  x [B,H,W,cin] (any [B, H*W*cin] view)."""
    check(lib().air_conv5x5_fwd(ptr(x), ptr(w), ptr(b), ptr(out), ptr(argmax), x.shape[0], H, W, cin, cout, int(pool),
                                stream()), "air_conv5x5_fwd")


def conv5x5_bwd_workspace(B, cin, cout):
    return int(lib().air_conv5x5_bwd_workspace(B, cin, cout))


def conv5x5_bwd(x, w, out, argmax, dout, din, dw, db, accumulate, workspace, H, W, cin, cout, pool):
    check(lib().air_conv5x5_bwd(ptr(x), ptr(w), ptr(out), ptr(argmax), ptr(dout), ptr(din), ptr(dw), ptr(db),
                                int(accumulate), ptr(workspace), x.shape[0], H, W, cin, cout, int(pool), stream()),
          "air_conv5x5_bwd")


def rng_state(seed, device, counter=0):
    """Device RNG state of include/air_b200.h (air_rng_state_t): int64 [4] = seed, step counter, scratch, unused."""
    return torch.tensor([int(seed) & 0x7FFFFFFFFFFFFFFF, int(counter), 0, 0], dtype=torch.int64, device=device)


def noise_fill(state, scale, shift, vae_latent, concrete_u, TB, L):
    """Advance the step counter and fill the small noise tensors of a step (TB = T * B rows): scale [TB], shift [TB, 2],
    vae_latent [TB, L] ~ N(0,1), concrete_u [TB] ~ U[0,1); any of them may be None."""
    for t, n in ((scale, TB), (shift, 2 * TB), (vae_latent, TB * L), (concrete_u, TB)):
        if t is not None and (t.numel() != n or t.dtype != torch.float32):
            raise C.AirError("noise_fill: tensor sizes must be T*B, 2*T*B, T*B*L, T*B (float32)")
    check(lib().air_noise_fill(ptr(state), ptr(scale), ptr(shift), ptr(vae_latent), ptr(concrete_u), TB, L, stream()), "air_noise_fill")


def rng_normals(state, rng_stream, n):
    out = torch.empty(n, device=state.device)
    check(lib().air_rng_normals(ptr(state), rng_stream, ptr(out), n, stream()), "air_rng_normals")
    return out


def rng_uniforms(state, rng_stream, n):
    out = torch.empty(n, device=state.device)
    check(lib().air_rng_uniforms(ptr(state), rng_stream, ptr(out), n, stream()), "air_rng_uniforms")
    return out


def zero_items(tensors):
    """-> opaque handle for zero_many(): the (pointer, byte count) arrays of up to 8 contiguous CUDA tensors whose
    storage is 16-byte aligned and a multiple of 16 bytes long (keep the tensors alive as long as the handle)."""
    n = len(tensors)
    ptrs, sizes = (ctypes.c_void_p * n)(), (ctypes.c_int64 * n)()
    for i, t in enumerate(tensors):
        nb = t.numel() * t.element_size()
        if not (t.is_cuda and t.is_contiguous()) or nb % 16 or t.data_ptr() % 16:
            raise C.AirError("zero_many: contiguous CUDA tensors, 16-byte aligned, sizes multiples of 16 bytes")
        ptrs[i], sizes[i] = t.data_ptr(), nb
    return ptrs, sizes, n


def zero_many(handle):
    """Zero all tensors of a zero_items() handle with one launch."""
    ptrs, sizes, n = handle
    check(lib().air_zero_buffers(ptrs, sizes, n, stream()), "air_zero_buffers")


def expand_u8(src, dst):
    """dst (float32) = float(src) * fl(1/255) for a uint8 CUDA tensor of the same number of elements (air_expand_u8)."""
    if src.dtype != torch.uint8 or dst.dtype != torch.float32 or src.numel() != dst.numel():
        raise C.AirError("expand_u8: uint8 source and float32 destination of equal size required")
    check(lib().air_expand_u8(ptr(src), ptr(dst), src.numel(), stream()), "air_expand_u8")
    return dst


def synth_canvases(images, counts, seed=0, first_index=0, canvas_size=50, max_digits=2, positions=None, boxes=None):
    """Fill images [B, canvas_size**2] / counts [B] int32 with device-generated multi-digit canvases; optionally the
    (x, y) positions and (w, h) boxes of the placed digits, int32 [B, max_digits, 2] each (multi_mnist.py:165-166)."""
    for t in (counts, positions, boxes):
        if t is not None and t.dtype != torch.int32:
            raise C.AirError("counts / positions / boxes must be int32")
    for t in (positions, boxes):
        if t is not None and t.numel() != images.shape[0] * 2 * max_digits:
            raise C.AirError("positions / boxes must hold [B, max_digits, 2] entries")
    check(lib().air_synth_canvases_ex(int(seed), int(first_index), ptr(images), ptr(counts), ptr(positions), ptr(boxes),
                                      images.shape[0], canvas_size, max_digits, stream()), "air_synth_canvases")
    return images, counts
