"""Concrete / Gumbel-softmax ops with the reference signatures (air/concrete.py).

The reference draws ``u = tf.random_uniform`` inside each function (concrete.py:7, :23);
here the uniform noise is an explicit keyword (``u=``) so results are reproducible, and is
drawn with torch's CUDA generator when omitted.  The model path uses the fused
``concrete_step`` (csrc/concrete.cu); the three reference functions are thin compositions
kept for API compatibility.
"""
from __future__ import annotations

import torch

from .. import _cabi

EPS = 10e-10


def _scalar_dev(v, device):
    if torch.is_tensor(v):
        return v.to(device=device, dtype=torch.float32).reshape(1).contiguous()
    return torch.tensor([float(v)], device=device, dtype=torch.float32)


class _ConcreteStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_odds, u, stop_prev, loss_prev, digits_prev, prior_log_odds, temperature, thr, train):
        log_odds = _cabi.f32(log_odds)
        B = log_odds.shape[0]
        dev = log_odds.device
        u, stop_prev, loss_prev = _cabi.f32(u), _cabi.f32(stop_prev), _cabi.f32(loss_prev)
        digits_prev = digits_prev.to(torch.int32).contiguous()
        prior = _scalar_dev(prior_log_odds, dev)
        y, z, z_prob, kl, stop_new, loss_new = (torch.empty(B, device=dev, dtype=torch.float32) for _ in range(6))
        digits_new = torch.empty(B, device=dev, dtype=torch.int32)
        _cabi.check(_cabi.lib().air_concrete_step_fwd(
            _cabi.ptr(log_odds), _cabi.ptr(u), _cabi.ptr(stop_prev), _cabi.ptr(loss_prev), _cabi.ptr(digits_prev),
            _cabi.ptr(prior), float(temperature), float(thr), int(bool(train)), _cabi.ptr(y), _cabi.ptr(z),
            _cabi.ptr(z_prob), _cabi.ptr(kl), _cabi.ptr(stop_new), _cabi.ptr(loss_new), _cabi.ptr(digits_new), B,
            _cabi.stream()), "air_concrete_step_fwd")
        ctx.save_for_backward(log_odds, y, z, stop_prev, prior)
        ctx.cfg = (float(temperature), float(thr), bool(train))
        ctx.mark_non_differentiable(z_prob, stop_new, digits_new, y)
        return y, z, z_prob, kl, stop_new, loss_new, digits_new

    @staticmethod
    def backward(ctx, dy, dz, dzp, dkl, dstop, dloss, ddig):
        log_odds, y, z, stop_prev, prior = ctx.saved_tensors
        tau, thr, train = ctx.cfg
        B = log_odds.shape[0]
        # loss_new = loss_prev + where(stop_prev < thr, kl, 0)
        gk = torch.zeros_like(log_odds) if dkl is None else dkl.clone()
        if dloss is not None:
            gk = gk + torch.where(stop_prev < thr, dloss, torch.zeros_like(dloss))
        dz = None if dz is None else _cabi.f32(dz)
        dlo = torch.empty_like(log_odds)
        _cabi.check(_cabi.lib().air_concrete_step_bwd(
            _cabi.ptr(log_odds), _cabi.ptr(y), _cabi.ptr(z), _cabi.ptr(dz), _cabi.ptr(gk.contiguous()),
            _cabi.ptr(prior), tau, int(train), _cabi.ptr(dlo), B, _cabi.stream()), "air_concrete_step_bwd")
        return dlo, None, None, dloss, None, None, None, None, None


def concrete_step(log_odds, u, stop_prev, loss_prev, digits_prev, prior_log_odds, temperature,
                  stopping_threshold, train):
    """One fused z_pres / ACT step (air_model.py:380-427).  Returns
    (y_pre_sigmoid, z_pres, z_pres_prob, z_pres_kl, stop_new, loss_new, digits_new)."""
    return _ConcreteStep.apply(log_odds, u, stop_prev, loss_prev, digits_prev, prior_log_odds, temperature,
                               stopping_threshold, train)


def _uniform_like(log_odds, u):
    if u is None:
        u = torch.rand(log_odds.shape[0], device=log_odds.device, dtype=torch.float32)
    return u


def _step_on_zero_state(log_odds, temperature, u, prior=0.0, train=True):
    B = log_odds.shape[0]
    zf = torch.zeros(B, device=log_odds.device, dtype=torch.float32)
    zi = torch.zeros(B, device=log_odds.device, dtype=torch.int32)
    return concrete_step(log_odds, u, zf, zf, zi, prior, temperature, 2.0, train)


def concrete_binary_pre_sigmoid_sample(log_odds, temperature, eps=EPS, u=None):
    """Drop-in for air/concrete.py:20-27 -> y = (log_odds + logistic noise) / temperature."""
    _require_default_eps(eps)
    return _PreSigmoid.apply(log_odds, _uniform_like(log_odds, u), float(temperature))


def concrete_binary_sample(log_odds, temperature, hard=False, eps=EPS, u=None):
    """Drop-in for air/concrete.py:4-17 -> (y, sig_y); note y is NOT divided by the
    temperature here (concrete.py:10-11)."""
    _require_default_eps(eps)
    y_over_t = _PreSigmoid.apply(log_odds, _uniform_like(log_odds, u), float(temperature))
    y = y_over_t * float(temperature)
    sig_y = torch.sigmoid(y_over_t)
    if hard:
        sig_y = (torch.round(sig_y) - sig_y).detach() + sig_y
    return y, sig_y


def concrete_binary_kl_mc_sample(y, prior_log_odds, prior_temperature, posterior_log_odds, posterior_temperature,
                                 eps=EPS):
    """Drop-in for air/concrete.py:30-43 (log q(y) - log p(y)).  Not on the model's hot
    path (the fused concrete_step computes the same quantity); composed from torch CUDA
    elementwise ops with the reference's op order."""
    def t(v):
        return v if torch.is_tensor(v) else torch.tensor(float(v), device=y.device, dtype=y.dtype)
    pl, pt, ql, qt = t(prior_log_odds), t(prior_temperature), t(posterior_log_odds), t(posterior_temperature)
    ytp = y * pt
    log_prior = torch.log(pt + eps) - ytp + pl - 2.0 * torch.log(1.0 + torch.exp(-ytp + pl) + eps)
    ytq = y * qt
    log_post = torch.log(qt + eps) - ytq + ql - 2.0 * torch.log(1.0 + torch.exp(-ytq + ql) + eps)
    return log_post - log_prior


def _require_default_eps(eps):
    if abs(eps - EPS) > 1e-18:
        raise ValueError("the CUDA Concrete kernels are compiled for the reference's eps=10e-10")


class _PreSigmoid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_odds, u, temperature):
        y = _step_on_zero_state(log_odds.detach(), temperature, u)[0]
        ctx.temperature = temperature
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy / ctx.temperature, None, None
