"""Spatial transformer with the reference signature (air/transformer.py:18, :178).

``transformer(U, theta, out_size)`` launches the hand-written sm_100a kernels in
csrc/st.cu through the C ABI; gradients w.r.t. U and theta come from air_st_backward.
"""
from __future__ import annotations

import torch

from .. import _cabi


class _STFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, U, theta, oh, ow):
        U = _cabi.f32(U)
        theta6 = _cabi.f32(theta).reshape(-1, 6)
        B, H, W, C = U.shape
        if theta6.shape[0] != B:
            raise ValueError(f"theta has {theta6.shape[0]} rows, U has batch {B}")
        out = torch.empty((B, oh, ow, C), device=U.device, dtype=torch.float32)
        _cabi.check(_cabi.lib().air_st_forward(_cabi.ptr(U), _cabi.ptr(theta6), _cabi.ptr(out), B, H, W, C, oh, ow,
                                               _cabi.stream()), "air_st_forward")
        ctx.save_for_backward(U, theta6)
        ctx.theta_shape = theta.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        U, theta6 = ctx.saved_tensors
        B, H, W, C = U.shape
        dout = _cabi.f32(dout)
        oh, ow = dout.shape[1], dout.shape[2]
        dU = torch.empty_like(U) if ctx.needs_input_grad[0] else None
        dtheta = torch.empty_like(theta6)
        _cabi.check(_cabi.lib().air_st_backward(_cabi.ptr(U), _cabi.ptr(theta6), _cabi.ptr(dout), _cabi.ptr(dU),
                                                _cabi.ptr(dtheta), B, H, W, C, oh, ow, _cabi.stream()),
                    "air_st_backward")
        return dU, dtheta.reshape(ctx.theta_shape), None, None


def transformer(U, theta, out_size, name="SpatialTransformer", **kwargs):
    """Drop-in for air/transformer.py:18.  U [B,H,W,C] (NHWC), theta [B,6] or [B,2,3],
    out_size (oh, ow) -> [B,oh,ow,C].  ``name``/kwargs are accepted and ignored."""
    if U.dim() != 4:
        raise ValueError("U must be [num_batch, height, width, num_channels]")
    return _STFunction.apply(U, theta, int(out_size[0]), int(out_size[1]))


def batch_transformer(U, thetas, out_size, name="BatchSpatialTransformer"):
    """Drop-in for air/transformer.py:178-195: thetas [B,N,6] ->
    [B*N, oh, ow, C] (each input repeated N times)."""
    num_batch, num_transforms = int(thetas.shape[0]), int(thetas.shape[1])
    idx = torch.arange(num_batch, device=U.device).repeat_interleave(num_transforms)
    return transformer(U.index_select(0, idx), thetas.reshape(num_batch * num_transforms, -1), out_size)


class _WritebackCanvasFunction(torch.autograd.Function):
    """canvas + where(stop_new < thr, z * ST(window, theta_inv), 0)
    (air_model.py:363-366 fused with :429-439)."""

    @staticmethod
    def forward(ctx, window, theta_inv, z, stop_new, canvas, thr):
        window = _cabi.f32(window)
        th6 = _cabi.f32(theta_inv).reshape(-1, 6)
        z = _cabi.f32(z)
        stop_new = _cabi.f32(stop_new)
        canvas = _cabi.f32(canvas)
        B, wh, ww = window.shape
        ch, cw = canvas.shape[1], canvas.shape[2]
        out = torch.empty_like(canvas)
        _cabi.check(_cabi.lib().air_st_writeback_canvas_fwd(
            _cabi.ptr(window), _cabi.ptr(th6), _cabi.ptr(z), _cabi.ptr(stop_new), float(thr), _cabi.ptr(canvas),
            _cabi.ptr(out), B, wh, ww, ch, cw, _cabi.stream()), "air_st_writeback_canvas_fwd")
        ctx.save_for_backward(window, th6, z, stop_new)
        ctx.thr = float(thr)
        ctx.theta_shape = theta_inv.shape
        return out

    @staticmethod
    def backward(ctx, dcanvas):
        window, th6, z, stop_new = ctx.saved_tensors
        dcanvas = _cabi.f32(dcanvas)
        B, wh, ww = window.shape
        ch, cw = dcanvas.shape[1], dcanvas.shape[2]
        dwindow = torch.empty_like(window)
        dtheta = torch.empty_like(th6)
        dz = torch.empty_like(z)
        _cabi.check(_cabi.lib().air_st_writeback_canvas_bwd(
            _cabi.ptr(window), _cabi.ptr(th6), _cabi.ptr(z), _cabi.ptr(stop_new), ctx.thr, _cabi.ptr(dcanvas),
            _cabi.ptr(dwindow), _cabi.ptr(dtheta), _cabi.ptr(dz), 0, B, wh, ww, ch, cw, _cabi.stream()),
            "air_st_writeback_canvas_bwd")
        return dwindow, dtheta.reshape(ctx.theta_shape), dz, None, dcanvas, None


def writeback_canvas(window, theta_inv, z, stop_new, canvas, stopping_threshold):
    """Fused write-back ST + z_pres-scaled canvas accumulation.
    window [B,wh,ww], theta_inv [B,6]|[B,2,3], z [B], stop_new [B], canvas [B,ch,cw]."""
    return _WritebackCanvasFunction.apply(window, theta_inv, z, stop_new, canvas, stopping_threshold)
