"""ctypes binding of oracle/st_oracle.c -- TEST INFRASTRUCTURE ONLY (see its header)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return os.path.join(_HERE, "_build", "liboracle.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        src = os.path.join(_HERE, "st_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def st_forward(U, theta, out_size):
    U, pU = _f(U)
    B, H, W, C = U.shape
    theta, pT = _f(np.reshape(theta, (B, 6)))
    oh, ow = out_size
    out = np.empty((B, oh, ow, C), np.float32)
    lib().oracle_st_forward(pU, pT, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(B), H, W, C, oh, ow)
    return out


def st_backward(U, theta, dout, need_dU=True):
    U, pU = _f(U)
    B, H, W, C = U.shape
    theta, pT = _f(np.reshape(theta, (B, 6)))
    dout, pD = _f(dout)
    oh, ow = dout.shape[1:3]
    dU = np.empty_like(U) if need_dU else None
    dtheta = np.empty((B, 6), np.float32)
    lib().oracle_st_backward(pU, pT, pD, dU.ctypes.data_as(ctypes.c_void_p) if need_dU else None,
                             dtheta.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(B), H, W, C, oh, ow)
    return dU, dtheta.reshape(B, 2, 3)


def canvas_update(canvas_in, window_recon, z, stop_new, thr):
    canvas_in, pC = _f(canvas_in)
    window_recon, pW = _f(window_recon)
    z, pZ = _f(z)
    stop_new, pS = _f(stop_new)
    B, N = canvas_in.shape
    out = np.empty_like(canvas_in)
    lib().oracle_canvas_update(pC, pW, pZ, pS, ctypes.c_float(thr), out.ctypes.data_as(ctypes.c_void_p),
                               ctypes.c_int64(B), N)
    return out


def concrete_step(log_odds, u, stop_prev, loss_prev, digits_prev, prior_log_odds, temperature, thr, train):
    log_odds, pL = _f(log_odds)
    u, pU = _f(u)
    stop_prev, pS = _f(stop_prev)
    loss_prev, pR = _f(loss_prev)
    digits_prev, pD = _i(digits_prev)
    B = log_odds.shape[0]
    outs = {k: np.empty(B, np.float32) for k in ("y", "z", "z_prob", "kl", "stop_new", "loss_new")}
    outs["digits_new"] = np.empty(B, np.int32)
    ptr = {k: v.ctypes.data_as(ctypes.c_void_p) for k, v in outs.items()}
    lib().oracle_concrete_step(pL, pU, pS, pR, pD, ctypes.c_float(prior_log_odds), ctypes.c_float(temperature),
                               ctypes.c_float(thr), int(bool(train)), ptr["y"], ptr["z"], ptr["z_prob"], ptr["kl"],
                               ptr["stop_new"], ptr["loss_new"], ptr["digits_new"], ctypes.c_int64(B))
    return outs


def gemm_seq_fma(A, Bm, Cinit=None, bias=None):
    A, pA = _f(A)
    Bm, pB = _f(Bm)
    M, K = A.shape
    N = Bm.shape[1]
    out = np.empty((M, N), np.float32)
    pC = pb = None
    if Cinit is not None:
        Cinit, pC = _f(Cinit)
    if bias is not None:
        bias, pb = _f(bias)
    lib().oracle_gemm_seq_fma(pA, pB, pC, pb, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(M), N, K, K, N, N)
    return out


def conv5x5_relu_pool(x, w, b, pool):
    """oracle_conv5x5_relu_pool: x [B,H,W,cin], w [5,5,cin,cout] (HWIO), b [cout] -> [B,H',W',cout]."""
    x, px = _f(x)
    w, pw = _f(w)
    b, pb = _f(b)
    B, H, W, cin = x.shape
    cout = w.shape[3]
    PH, PW = (H // 2, W // 2) if pool else (H, W)
    out = np.empty((B, PH, PW, cout), np.float32)
    lib().oracle_conv5x5_relu_pool(px, pw, pb, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(B), H, W, cin, cout, int(pool))
    return out


def conv5x5_relu_pool_bwd(x, w, b, dout, pool, need_dx=True):
    x, px = _f(x)
    w, pw = _f(w)
    b, pb = _f(b)
    dout, pd = _f(dout)
    B, H, W, cin = x.shape
    cout = w.shape[3]
    dw, db = np.empty_like(w), np.empty_like(b)
    dx = np.empty_like(x) if need_dx else None
    lib().oracle_conv5x5_relu_pool_bwd(px, pw, pb, pd, dx.ctypes.data_as(ctypes.c_void_p) if need_dx else None,
                                       dw.ctypes.data_as(ctypes.c_void_p), db.ctypes.data_as(ctypes.c_void_p),
                                       ctypes.c_int64(B), H, W, cin, cout, int(pool))
    return dx, dw, db
