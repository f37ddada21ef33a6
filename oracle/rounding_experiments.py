"""Which part of the reference's fp32 gradient does its training depend on?  (test infrastructure / evidence, CPU only)

The oracle (oracle/air_oracle.py, pinned to the reference's own graph) is trained for a few hundred iterations with the
reference's training configuration while one aspect of the WRITE-BACK spatial transformer's arithmetic
(air_model.py:351-366 -> transformer.py:56-117) is changed:

  base          unchanged fp32 oracle (torch autograd: four gather buffers, each accumulated in pixel order)
  fp64          the whole model in double (the out-of-window residues sink below the loss's 1e-9 epsilon)
  clean_all     write-back evaluated in fp64 and rounded: no residues in the canvas, exactly cancelled gradient
  dz_only       canvas keeps its fp32 residues (so dz sees them), d(window) and d(theta_inv) exactly cancelled
  dU_noisy      d(window) through the fp32 graph, d(theta_inv) exactly cancelled
  dtheta_noisy  d(theta_inv) through the fp32 graph, d(window) exactly cancelled
  tf_order      the four gathers of _interpolate as ONE gather of the concatenated index list: its gradient is then
                accumulated in the order of the reference graph's gradients/concat -> UnsortedSegmentSum (all a-corner
                terms in pixel order, then b, c, d) -- what TensorFlow's CPU kernel does
  perm          the same list in a fixed random order (what a GPU's atomics amount to)

    python oracle/rounding_experiments.py VARIANT ITERS [THREADS]

Results of round 2 (batch 64, mean loss per 50 iterations; profiles/r2_reference_rounding.md): every variant that keeps
un-cancelled fp32 terms in d(window) or d(theta_inv) learns, whatever the order; clean_all / dz_only / fp64 stay at
loss ~1900 with a gradient norm of ~50 instead of 1e4 ... 1e7."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import air_oracle as O   # noqa: E402

variant = sys.argv[1]
iters = int(sys.argv[2])
torch.set_num_threads(int(sys.argv[3]) if len(sys.argv) > 3 else 4)
dt = torch.float64 if variant == "fp64" else torch.float32
orig_transformer, orig_interpolate = O.transformer, O._interpolate
_perm = {}


def transformer_variant(U, theta, out_size, **kw):
    if int(out_size[0]) != 50 or variant in ("base", "fp64", "tf_order", "perm"):   # only the write-back is touched
        return orig_transformer(U, theta, out_size, **kw)
    w32 = orig_transformer(U, theta, out_size, **kw)
    w64 = orig_transformer(U.double(), theta.double(), out_size, **kw).float()
    if variant == "clean_all":
        return w64
    if variant == "dz_only":
        return w64 + (w32.detach() - w64.detach())
    if variant == "dU_noisy":
        wU = orig_transformer(U, theta.detach(), out_size, **kw)
        wt = orig_transformer(U.detach().double(), theta.double(), out_size, **kw).float()
        return wU + (wt - wt.detach())
    if variant == "dtheta_noisy":
        wt = orig_transformer(U.detach(), theta, out_size, **kw)
        wU = orig_transformer(U.double(), theta.detach().double(), out_size, **kw).float()
        return wt + (wU - wU.detach())
    raise SystemExit(f"unknown variant {variant}")


def interpolate_one_gather(im, x, y, out_size):
    """transformer.py:56-117 with the four gathers expressed as one, so that autograd accumulates their gradients as one
    list (tf_order) or in a fixed random order (perm).  Forward values are bit-identical to O._interpolate."""
    num_batch, height, width, channels = im.shape
    dtype = im.dtype
    oh, ow = out_size
    x = (x + 1.0) * (torch.tensor(float(width), dtype=dtype) - 1.001) / 2.0
    y = (y + 1.0) * (torch.tensor(float(height), dtype=dtype) - 1.001) / 2.0
    x0 = torch.floor(x).to(torch.int64); x1 = x0 + 1
    y0 = torch.floor(y).to(torch.int64); y1 = y0 + 1
    x0, x1 = torch.clamp(x0, 0, width - 1), torch.clamp(x1, 0, width - 1)
    y0, y1 = torch.clamp(y0, 0, height - 1), torch.clamp(y1, 0, height - 1)
    base = (torch.arange(num_batch, dtype=torch.int64) * (width * height)).repeat_interleave(oh * ow)
    idx = torch.cat([base + y0 * width + x0, base + y1 * width + x0, base + y0 * width + x1, base + y1 * width + x1])
    im_flat = im.reshape(-1, channels)
    n = x.numel()
    if variant == "perm":
        if idx.numel() not in _perm:
            _perm[idx.numel()] = torch.randperm(idx.numel(), generator=torch.Generator().manual_seed(5))
        pm = _perm[idx.numel()]
        inv = torch.empty_like(pm)
        inv[pm] = torch.arange(pm.numel())
        vals = im_flat[idx[pm]][inv]
    else:
        vals = im_flat[idx]
    Ia, Ib, Ic, Id = vals[:n], vals[n:2 * n], vals[2 * n:3 * n], vals[3 * n:]
    x0_f, x1_f, y0_f, y1_f = x0.to(dtype), x1.to(dtype), y0.to(dtype), y1.to(dtype)
    wa = ((x1_f - x) * (y1_f - y)).unsqueeze(1); wb = ((x1_f - x) * (y - y0_f)).unsqueeze(1)
    wc = ((x - x0_f) * (y1_f - y)).unsqueeze(1); wd = ((x - x0_f) * (y - y0_f)).unsqueeze(1)
    return ((wa * Ia + wb * Ib) + wc * Ic) + wd * Id


O.transformer = transformer_variant
if variant in ("tf_order", "perm"):
    O._interpolate = interpolate_one_gather
train, cnt = O.synthetic_canvases(20000, seed=0)
m = O.AIROracle(params={k: v.to(dt) for k, v in O.init_params(seed=0).items()}, annealing_schedules=O.DEFAULT_ANNEALING,
                train=True, dtype=dt)
m.adam_m = {k: v.to(dt) for k, v in m.adam_m.items()}
m.adam_v = {k: v.to(dt) for k, v in m.adam_v.items()}
g = torch.Generator().manual_seed(1)
t0 = time.time()
rl = ra = rn = 0.0
for it in range(iters):
    idx = torch.randint(0, 20000, (64,), generator=g)
    noise = {k: v.to(dt) for k, v in O.make_noise(it, 3, 64).items()}
    out, _ = m.train_step(train[idx].to(dt), cnt[idx], noise)
    rl += float(out["loss"]); ra += float(out["accuracy"]); rn += float(out["grad_global_norm"])
    if it % 50 == 49:
        print(f"{variant} it {it + 1} loss {rl / 50:.2f} acc {ra / 50:.3f} |g| {rn / 50:.3e} t {time.time() - t0:.0f}", flush=True)
        rl = ra = rn = 0.0
