"""Flat parameter / gradient / Adam-slot storage with checkpoint-named views.

The reference keeps 36 separate TF variables under ``air/rnn/...`` (model/air-model.index;
created at air_model.py:284-376, vae.py:11-34).  Here they live in ONE flat fp32 buffer
(one NCCL allreduce, one fused clip+Adam launch), and tensors that feed the same GEMM are
stored fused: the five 256->64 hidden head layers as one [256, 320] matrix, the seven
head outputs as one [7, 64] matrix, rec_mean / rec_log_variance as one [256, 100] matrix.
``named_views()`` exposes exactly the reference's variable names and shapes as (strided)
views, so checkpoints and the oracle's parameter dict load/save by name.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

HEADS = (("scale", "mean"), ("scale", "log_variance"), ("shift", "mean"), ("shift", "log_variance"),
         ("z_pres", "log_odds"))
# rows of heads/out_w: (head block, first output column, number of outputs)
HEAD_OUT_ROWS = ((0, 0, 1), (1, 1, 1), (2, 2, 2), (3, 4, 2), (4, 6, 1))


def head_sizes(HU):
    """hidden units of the five head layers (HEADS order) from an int or a (scale, shift, z_pres) triple."""
    if isinstance(HU, int):
        return (HU,) * 5
    hs, hf, hz = (int(v) for v in HU)
    return (hs, hs, hf, hf, hz)


def reference_shapes(in_dim, win, R, HU, L, rec_units, gen_units, cnn_filters=None):
    """name -> shape exactly as in model/air-model.index (minus the 'air/rnn/' prefix),
    ordered [cnn,] rnn, scale, shift, z_pres, vae (the order only matters for the init RNG stream).
    The CNN front-end's variables (air_model.py:510-535, graph names air/cnn/convN/{kernel,bias}) are keyed
    ``cnn/...``; ``in_dim`` is then its flattened output size."""
    s = OrderedDict()
    if cnn_filters:
        prev = 1
        for name in ("conv1", "conv2", "conv3"):
            s[f"cnn/{name}/kernel"] = (5, 5, prev, cnn_filters)
            s[f"cnn/{name}/bias"] = (cnn_filters,)
            prev = cnn_filters
    s["rnn/kernel"] = (in_dim + R, 4 * R)
    s["rnn/bias"] = (4 * R,)
    for (head, stat), hu in zip(HEADS, head_sizes(HU)):
        out = 2 if head == "shift" else 1
        s[f"{head}/{stat}/hidden/weights"] = (R, hu)
        s[f"{head}/{stat}/hidden/biases"] = (hu,)
        s[f"{head}/{stat}/output/weights"] = (hu, out)
        s[f"{head}/{stat}/output/biases"] = (out,)
    prev = win
    for i, u in enumerate(rec_units):
        s[f"vae/recognition_{i + 1}/weights"] = (prev, u)
        s[f"vae/recognition_{i + 1}/biases"] = (u,)
        prev = u
    for nm in ("rec_mean", "rec_log_variance"):
        s[f"vae/{nm}/weights"] = (prev, L)
        s[f"vae/{nm}/biases"] = (L,)
    prev = L
    for i, u in enumerate(gen_units):
        s[f"vae/generative_{i + 1}/weights"] = (prev, u)
        s[f"vae/generative_{i + 1}/biases"] = (u,)
        prev = u
    s["vae/gen_mean/weights"] = (prev, win)
    s["vae/gen_mean/biases"] = (win,)
    return s


class ParamStore:
    def __init__(self, device, in_dim, win, R, HU, L, rec_units, gen_units, seed=0, cnn_filters=None):
        self.device = torch.device(device)
        # Unequal scale / shift / z_pres hidden sizes (air_model.py:288-316, 372-376 allow them): every head block of the
        # fused [R, 5*HU] hidden layer and [7, HU] output matrix is padded to the widest head with zeros.  A padded unit
        # is relu(0 . h + 0) = 0, meets a zero output weight and receives a zero gradient (so it stays zero under Adam):
        # the arithmetic of the real units is unchanged, bit for bit; named_views() expose the reference-shaped slices.
        self.head_units = head_sizes(HU)
        HU_spec, HU = HU, max(self.head_units)
        self.dims = dict(in_dim=in_dim, win=win, R=R, HU=HU, HU_spec=HU_spec, L=L, rec_units=tuple(rec_units),
                         gen_units=tuple(gen_units), cnn_filters=cnn_filters)
        fused = OrderedDict()
        if cnn_filters:
            prev = 1
            for name in ("conv1", "conv2", "conv3"):
                fused[f"cnn/{name}/kernel"] = (5, 5, prev, cnn_filters)
                fused[f"cnn/{name}/bias"] = (cnn_filters,)
                prev = cnn_filters
        fused["rnn/kernel"] = (in_dim + R, 4 * R)
        fused["rnn/bias"] = (4 * R,)
        fused["heads/hidden_w"] = (R, 5 * HU)
        fused["heads/hidden_b"] = (5 * HU,)
        fused["heads/out_w"] = (7, HU)
        fused["heads/out_b"] = (7,)
        prev = win
        for i, u in enumerate(rec_units):
            fused[f"vae/recognition_{i + 1}/weights"] = (prev, u)
            fused[f"vae/recognition_{i + 1}/biases"] = (u,)
            prev = u
        fused["vae/rec_ml/weights"] = (prev, 2 * L)
        fused["vae/rec_ml/biases"] = (2 * L,)
        prev = L
        for i, u in enumerate(gen_units):
            fused[f"vae/generative_{i + 1}/weights"] = (prev, u)
            fused[f"vae/generative_{i + 1}/biases"] = (u,)
            prev = u
        fused["vae/gen_mean/weights"] = (prev, win)
        fused["vae/gen_mean/biases"] = (win,)
        self.fused_shapes = fused
        # Order in the flat buffer = the order in which the backward PRODUCES the gradients, so that the data-parallel
        # all-reduce can run as contiguous buckets that leave while the remaining weight-gradient GEMMs still run
        # (AIRModel._reduce_bucket): LSTM kernel (image rows first, 64 % of all bytes) | CNN front-end | head hidden
        # weights | VAE weights, decoder first (the order of vae_weight_grads) | an arena with everything that is
        # produced last by the batched reductions -- head outputs and ALL biases, 6 K floats: the only exposed message.
        vae_w = [f"vae/generative_{i + 1}/weights" for i in range(len(gen_units))] + ["vae/gen_mean/weights"] + \
                [f"vae/recognition_{i + 1}/weights" for i in range(len(rec_units))] + ["vae/rec_ml/weights"]
        order = ["rnn/kernel"] + [k for k in fused if k.startswith("cnn/")] + ["heads/hidden_w"] + vae_w
        arena = [k for k in fused if k not in order]
        self.bucket_names = dict(rnn=["rnn/kernel"], cnn=[k for k in fused if k.startswith("cnn/")], heads=["heads/hidden_w"],
                                 vae=vae_w, arena=arena)
        self.offsets = OrderedDict()
        off = 0
        for k in order + arena:
            self.offsets[k] = off
            off += (int(math.prod(fused[k])) + 3) & ~3  # 16-byte aligned tensors
        self.n = off
        self.n_params = sum(int(math.prod(s)) for s in fused.values())
        self.flat = torch.zeros(self.n, device=self.device, dtype=torch.float32)
        self.grad = torch.zeros_like(self.flat)
        self.adam_m = torch.zeros_like(self.flat)
        self.adam_v = torch.zeros_like(self.flat)
        # [beta1^t, beta2^t, global_step, last global grad norm, learning rate, -, -, -]
        self.state = torch.zeros(8, device=self.device, dtype=torch.float32)
        self.reset_optimizer()
        self.p = self._views(self.flat)
        self.g = self._views(self.grad)
        self.init_xavier(seed)

    # ---- views -------------------------------------------------------------------------
    def _views(self, flat):
        return OrderedDict((k, flat[self.offsets[k]:self.offsets[k] + int(math.prod(shp))].view(shp))
                           for k, shp in self.fused_shapes.items())

    def bucket_bounds(self):
        """Flat-buffer boundaries of the gradient buckets, in production order:
        [0, kernel image rows | kernel h rows + cnn | head hidden weights | vae weights | arena, n]."""
        in_rows = self.dims["in_dim"] * 4 * self.dims["R"]
        b = [0, in_rows, self.offsets["heads/hidden_w"], self.offsets[self.bucket_names["vae"][0]],
             self.offsets[self.bucket_names["arena"][0]], self.n]
        assert b == sorted(b) and in_rows % 4 == 0
        return b

    def _named(self, v):
        d, HU, L = OrderedDict(), self.dims["HU"], self.dims["L"]
        for k in v:
            if k.startswith("cnn/"):
                d[k] = v[k]
        d["rnn/kernel"], d["rnn/bias"] = v["rnn/kernel"], v["rnn/bias"]
        for (head, stat), (blk, row, nout), hu in zip(HEADS, HEAD_OUT_ROWS, self.head_units):
            d[f"{head}/{stat}/hidden/weights"] = v["heads/hidden_w"][:, blk * HU:blk * HU + hu]
            d[f"{head}/{stat}/hidden/biases"] = v["heads/hidden_b"][blk * HU:blk * HU + hu]
            d[f"{head}/{stat}/output/weights"] = v["heads/out_w"][row:row + nout, :hu].t()
            d[f"{head}/{stat}/output/biases"] = v["heads/out_b"][row:row + nout]
        for k in v:
            if k.startswith("vae/") and not k.startswith("vae/rec_ml/"):
                d[k] = v[k]
        d["vae/rec_mean/weights"] = v["vae/rec_ml/weights"][:, :L]
        d["vae/rec_log_variance/weights"] = v["vae/rec_ml/weights"][:, L:]
        d["vae/rec_mean/biases"] = v["vae/rec_ml/biases"][:L]
        d["vae/rec_log_variance/biases"] = v["vae/rec_ml/biases"][L:]
        return d

    def named_views(self):
        """reference variable name -> view of the flat parameter buffer."""
        return self._named(self.p)

    def named_grads(self):
        return self._named(self.g)

    def named_adam(self):
        return self._named(self._views(self.adam_m)), self._named(self._views(self.adam_v))

    # ---- init / load / save ------------------------------------------------------------------
    def init_xavier(self, seed=0):
        """glorot-uniform weights (limit sqrt(6/(fan_in+fan_out)) of the REFERENCE-shaped
        tensor), zero biases: the TF defaults of fully_connected / BasicLSTMCell."""
        g = torch.Generator().manual_seed(seed)
        d = self.dims
        named = self.named_views()
        for name, shape in reference_shapes(d["in_dim"], d["win"], d["R"], d["HU_spec"], d["L"], d["rec_units"],
                                            d["gen_units"], d["cnn_filters"]).items():
            if len(shape) >= 2:  # conv kernels [kh,kw,in,out]: fan_in = kh*kw*in, fan_out = kh*kw*out (TF glorot)
                rf = math.prod(shape[:-2])
                lim = math.sqrt(6.0 / (rf * shape[-2] + rf * shape[-1]))
                w = (torch.rand(shape, generator=g, dtype=torch.float64) * 2.0 - 1.0) * lim
                named[name].copy_(w.to(torch.float32))
            else:
                named[name].zero_()

    def load_named(self, tensors, strict=True):
        named = self.named_views()
        missing = [k for k in named if k not in tensors]
        if strict and missing:
            raise KeyError(f"missing parameters: {missing}")
        for k, t in tensors.items():
            if k in named:
                if tuple(named[k].shape) != tuple(t.shape):
                    raise ValueError(f"{k}: shape {tuple(t.shape)} != {tuple(named[k].shape)}")
                named[k].copy_(torch.as_tensor(t).to(torch.float32))

    def state_dict(self):
        return OrderedDict((k, v.detach().clone()) for k, v in self.named_views().items())

    def reset_optimizer(self, learning_rate=None):
        self.adam_m.zero_()
        self.adam_v.zero_()
        st = torch.zeros(8, dtype=torch.float32)
        st[0], st[1] = 0.9, 0.999  # beta powers start at beta (TF creates them initialised to beta)
        if learning_rate is not None:
            st[4] = learning_rate
        else:
            st[4] = float(self.state[4].item()) if self.state[4].item() != 0 else 0.0
        self.state.copy_(st)

    @property
    def global_step(self):
        return int(self.state[2].item())

    @global_step.setter
    def global_step(self, v):
        self.state[2] = float(v)
