"""VAE with the reference signature (air/vae.py:5-43), running on the C-ABI kernels.

``vae(inputs, input_dim, rec_hidden_units, latent_dim, gen_hidden_units, likelihood_std,
activation)`` returns ``(reconstruction, rec_mean, rec_log_variance, rec_mean)`` exactly
like the reference (the 4th item is the MEAN: vae.py:43).  The two noise tensors the
reference draws internally (vae.py:23, :37) are explicit keyword arguments.

``vae_forward`` / ``vae_backward`` are the kernel schedules shared with AIRModel.
"""
from __future__ import annotations

import torch

from .. import _cabi as C
from .. import ops


class VAEWeights:
    """Views into a ParamStore-like dict, fused mean|log_variance layer included."""

    def __init__(self, p, g, n_rec, n_gen):
        self.rec = [(p[f"vae/recognition_{i + 1}/weights"], p[f"vae/recognition_{i + 1}/biases"]) for i in range(n_rec)]
        self.ml = (p["vae/rec_ml/weights"], p["vae/rec_ml/biases"])
        self.gen = [(p[f"vae/generative_{i + 1}/weights"], p[f"vae/generative_{i + 1}/biases"]) for i in range(n_gen)]
        self.gm = (p["vae/gen_mean/weights"], p["vae/gen_mean/biases"])
        if g is not None:
            self.g_rec = [(g[f"vae/recognition_{i + 1}/weights"], g[f"vae/recognition_{i + 1}/biases"])
                          for i in range(n_rec)]
            self.g_ml = (g["vae/rec_ml/weights"], g["vae/rec_ml/biases"])
            self.g_gen = [(g[f"vae/generative_{i + 1}/weights"], g[f"vae/generative_{i + 1}/biases"])
                          for i in range(n_gen)]
            self.g_gm = (g["vae/gen_mean/weights"], g["vae/gen_mean/biases"])


def alloc_vae_buffers(B, win, rec_units, L, gen_units, device, lead=()):
    """Activation buffers of one (or ``lead``-many) VAE evaluations."""
    z = lambda *s: torch.empty(*lead, *s, device=device, dtype=torch.float32)
    Lp = (L + 3) // 4 * 4  # the sample feeds a TMA GEMM: leading dimension must be a multiple of 4
    return dict(enc=[z(B, u) for u in rec_units], ml=z(B, 2 * L), zs=torch.zeros(*lead, B, Lp, device=device)[..., :L],
                dec=[z(B, u) for u in gen_units], recon=z(B, win))


def vae_forward(x, w: VAEWeights, noise_latent, noise_like, likelihood_std, hyper, buf, gen_tmp, fields, loss, mode,
                latent_fn=None):
    """x [rows,win] -> buf['recon'].  Softplus MLP 784->512->256 -> (mean|logvar) -> sample
    -> 256->512->784 -> sigmoid(gen + noise*std)   (vae.py:9-41).  Also accumulates the VAE
    KL into ``loss`` / ``fields`` (air_model.py:479-493).  AIRModel runs all T loop steps as ONE evaluation on
    T*B rows (the VAE of step t depends on nothing but theta_t) and passes ``latent_fn`` to do the latent step --
    the only place that touches per-step state -- step by step."""
    a = x
    for (W, b), out in zip(w.rec, buf["enc"]):
        ops.gemm(a, W, out, bias=b, epi=C.EPI_SOFTPLUS, mode=mode)
        a = out
    ops.gemm(a, w.ml[0], buf["ml"], bias=w.ml[1], mode=mode)
    if latent_fn is not None:
        latent_fn()
    else:
        ops.vae_latent_fwd(buf["ml"], noise_latent, hyper, buf["zs"], fields, loss)
    a = buf["zs"]
    for (W, b), out in zip(w.gen, buf["dec"]):
        ops.gemm(a, W, out, bias=b, epi=C.EPI_SOFTPLUS, mode=mode)
        a = out
    # gen_mean layer with the noisy-sigmoid output fused into the GEMM epilogue (vae.py:33-41); noise_like is either
    # the [rows, win] noise tensor (injected) or the int64 device RNG state: the samples are then generated in the epilogue
    rng = noise_like.dtype == torch.int64
    ops.gemm(a, w.gm[0], buf["recon"], bias=w.gm[1], aux=noise_like, epi=C.EPI_SIGMOID_RNG if rng else C.EPI_SIGMOID_NOISE,
             epi_param=likelihood_std, mode=mode)
    return buf["recon"]


def dense_dx(x_in, W, dY, dX, mode, act_of_x=False):
    """dX = (dY W^T) * softplus'(pre-activation of x_in) when x_in is a softplus output."""
    ops.gemm(dY, W, dX, tB=True, aux=x_in if act_of_x else None,
             epi=C.EPI_MUL_DSOFTPLUS if act_of_x else C.EPI_NONE, mode=mode)
    return dX


def dense_dw(x_in, dY, gW, gb, colsum_ws, mode, accumulate=False, gemm_ws=None):
    """gW (+)= x_in^T dY, gb (+)= colsum(dY).  x_in / dY may be time-batched [T*B, .] views.
    ``gemm_ws``: split-K scratch of the tensor-core GEMM (the reduction runs over the batch: few tiles, long K)."""
    ops.gemm(x_in, dY, gW, Cinit=gW if accumulate else None, tA=True, mode=mode, ws=gemm_ws)
    if colsum_ws is not None:  # None: the caller sums all its bias gradients in one launch (vae_bias_items)
        ops.colsum(dY, gb, accumulate, colsum_ws)


def vae_backward_dx(x, w: VAEWeights, noise_latent, hyper, buf, dbuf, dloss, fields, mode, dx_out=None,
                    dgen_is_presigmoid=False, dml_extra=None, latent_bwd_fn=None):
    """Backward through one VAE evaluation, activations only (air/vae.py:9-41 reversed).
    dbuf['dgen'] holds d(loss)/d(reconstruction) on entry; on exit dbuf holds the gradient
    w.r.t. every layer's pre-activation output (dgen, ddec[i], dml, denc[i]) -- the dY
    operands of the deferred weight-gradient GEMMs (vae_weight_grads)."""
    if not dgen_is_presigmoid:  # (AIRModel's write-back backward already applied recon * (1 - recon))
        ops.sigmoid_bwd(buf["recon"], dbuf["dgen"], dbuf["dgen"])  # d gen_mean, in place
    dY = dbuf["dgen"]
    acts = [buf["zs"]] + list(buf["dec"])
    layers = list(w.gen) + [w.gm]
    douts = [dbuf["dzs"]] + list(dbuf["ddec"])
    for i in range(len(layers) - 1, -1, -1):
        dY = dense_dx(acts[i], layers[i][0], dY, douts[i], mode, act_of_x=(i > 0))
    if latent_bwd_fn is not None:  # AIRModel: T*B rows, the latent step per loop step (see vae_forward)
        latent_bwd_fn()
    else:
        ops.vae_latent_bwd(buf["ml"], noise_latent, dY, fields, hyper, dloss, dbuf["dml"])
    if dml_extra is not None:  # gradients arriving directly at the (mean | log_variance) outputs of vae()
        dbuf["dml"].add_(dml_extra)
    dY = dbuf["dml"]
    acts = [x] + list(buf["enc"])
    layers = list(w.rec) + [w.ml]
    douts = [dx_out] + list(dbuf["denc"])
    for i in range(len(layers) - 1, -1, -1):
        if douts[i] is None:
            break
        dY = dense_dx(acts[i], layers[i][0], dY, douts[i], mode, act_of_x=(i > 0))
    return dx_out


def vae_weight_grads(x, w: VAEWeights, buf, dbuf, colsum_ws, mode, accumulate=False, gemm_ws=None):
    """All VAE parameter gradients from (time-batched) activations ``buf`` and the matching
    pre-activation gradients ``dbuf``: one long-K GEMM + one column sum per layer."""
    flat = lambda t: t.reshape(-1, t.shape[-1]) if t.is_contiguous() else t.flatten(0, -2)
    acts = [buf["zs"]] + list(buf["dec"])
    dys = list(dbuf["ddec"]) + [dbuf["dgen"]]
    grads = list(w.g_gen) + [w.g_gm]
    for a, dy, (gW, gb) in zip(acts, dys, grads):
        dense_dw(_flat2(a), _flat2(dy), gW, gb, colsum_ws, mode, accumulate, gemm_ws)
    acts = [x] + list(buf["enc"])
    dys = list(dbuf["denc"]) + [dbuf["dml"]]
    grads = list(w.g_rec) + [w.g_ml]
    for a, dy, (gW, gb) in zip(acts, dys, grads):
        dense_dw(_flat2(a), _flat2(dy), gW, gb, colsum_ws, mode, accumulate, gemm_ws)


def vae_bias_items(w: VAEWeights, dbuf, accumulate=False):
    """(dY, d(bias), accumulate) of every VAE layer, for ops.colsum_multi (same pairing as vae_weight_grads)."""
    dys = list(dbuf["ddec"]) + [dbuf["dgen"]] + list(dbuf["denc"]) + [dbuf["dml"]]
    gbs = [gb for _, gb in list(w.g_gen) + [w.g_gm] + list(w.g_rec) + [w.g_ml]]
    return [(_flat2(dy), gb, accumulate) for dy, gb in zip(dys, gbs)]


def _flat2(t):
    """[T,B,n] (possibly a padded view with unit inner stride) -> [T*B, n] view."""
    if t.dim() == 2:
        return t
    T, B, n = t.shape
    assert t.stride(2) == 1 and t.stride(0) == B * t.stride(1)
    return t.as_strided((T * B, n), (t.stride(1), 1), t.storage_offset())


def alloc_vae_scratch(B, win, rec_units, L, gen_units, device, lead=()):
    z = lambda *s: torch.empty(*lead, *s, device=device, dtype=torch.float32)
    return dict(denc=[z(B, u) for u in rec_units], dml=z(B, 2 * L), dzs=z(B, L), ddec=[z(B, u) for u in gen_units],
                dgen=z(B, win))


# ------------------------------------------------------------------------------------------
# public function with the reference signature
# ------------------------------------------------------------------------------------------
_SCOPES = {}


def _softplus_marker(x):  # placeholder so that ``activation=softplus`` can be passed explicitly
    raise RuntimeError("marker only")


softplus = _softplus_marker


class _VAEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, flat, store, n_rec, n_gen, noise_latent, noise_like, likelihood_std, mode):
        ctx.set_materialize_grads(False)
        B = inputs.shape[0]
        d = store.dims
        dev = inputs.device
        buf = alloc_vae_buffers(B, d["win"], d["rec_units"], d["L"], d["gen_units"], dev)
        w = VAEWeights(store.p, store.g, n_rec, n_gen)
        hyper = C.Hyper(0, 1, 0, 1, 0.0, 1.0, likelihood_std, 1.0, 2.0, 1)
        fields = torch.zeros(C.NF, B, device=dev)
        loss = torch.zeros(B, device=dev)
        gen_tmp = torch.empty(B, d["win"], device=dev)
        x = C.f32(inputs)
        vae_forward(x, w, C.f32(noise_latent), C.f32(noise_like), likelihood_std, hyper, buf, gen_tmp, fields, loss, mode)
        ctx.stuff = (x, buf, w, hyper, fields, C.f32(noise_latent), store, mode)
        L = d["L"]
        mean, logvar = buf["ml"][:, :L], buf["ml"][:, L:]
        return buf["recon"], mean, logvar

    @staticmethod
    def backward(ctx, drecon, dmean, dlogvar):
        x, buf, w, hyper, fields, noise_latent, store, mode = ctx.stuff
        B = x.shape[0]
        d = store.dims
        dev = x.device
        dbuf = alloc_vae_scratch(B, d["win"], d["rec_units"], d["L"], d["gen_units"], dev)
        ws = torch.zeros(int(C.lib().air_colsum_workspace(B, max(d["win"], 2 * d["L"], *d["rec_units"], *d["gen_units"]))),
                         device=dev)
        dx = torch.empty_like(x)
        store.grad.zero_()
        # dloss = 0: the KL term belongs to the caller (air_model.py:479-493), who differentiates through the
        # rec_mean / rec_log_variance outputs; those gradients are added to d(mean | log_variance) here
        L = d["L"]
        extra = None
        if dmean is not None or dlogvar is not None:
            extra = torch.zeros(B, 2 * L, device=dev)
            if dmean is not None:
                extra[:, :L] = dmean
            if dlogvar is not None:
                extra[:, L:] = dlogvar
        if drecon is None:
            dbuf["dgen"].zero_()
        else:
            dbuf["dgen"].copy_(drecon)
        vae_backward_dx(x, w, noise_latent, hyper, buf, dbuf, 0.0, fields, mode, dx_out=dx, dml_extra=extra)
        vae_weight_grads(x, w, buf, dbuf, ws, mode)
        return dx, store.grad.clone(), None, None, None, None, None, None, None


def vae(inputs, input_dim, rec_hidden_units, latent_dim, gen_hidden_units, likelihood_std=0.0, activation=softplus,
        *, scope="vae", reuse=False, noise_latent=None, noise_like=None, seed=0, gemm_mode="fp32"):
    """Drop-in for air/vae.py:5-43.  Variables are created on first use under ``scope``
    (Xavier-uniform, zero biases) and re-used with ``reuse=True``; the flat parameter
    buffer is ``vae.variables(scope).flat`` (requires_grad leaf returned by variables())."""
    if activation is not softplus:
        raise NotImplementedError("the CUDA VAE is compiled for the reference's softplus activation")
    from .params import ParamStore
    key = (scope, input_dim, tuple(rec_hidden_units), latent_dim, tuple(gen_hidden_units), inputs.device)
    if key not in _SCOPES:
        if reuse:
            raise ValueError(f"vae scope {scope!r} does not exist (reuse=True)")
        store = ParamStore(inputs.device, in_dim=1, win=input_dim, R=1, HU=1, L=latent_dim,
                           rec_units=rec_hidden_units, gen_units=gen_hidden_units, seed=seed)
        store.flat.requires_grad_(True)
        _SCOPES[key] = store
    store = _SCOPES[key]
    B = inputs.shape[0]
    if noise_latent is None:
        noise_latent = torch.randn(B, latent_dim, device=inputs.device)
    if noise_like is None:
        noise_like = torch.randn(B, input_dim, device=inputs.device)
    recon, mean, logvar = _VAEFunction.apply(inputs, store.flat, store, len(rec_hidden_units), len(gen_hidden_units),
                                             noise_latent, noise_like, float(likelihood_std),
                                             C.GEMM_MODES[gemm_mode])
    return recon, mean, logvar, mean


def variables(scope="vae"):
    """ParamStores created by vae() under ``scope`` (named_views() gives the TF names)."""
    return [s for k, s in _SCOPES.items() if k[0] == scope]
