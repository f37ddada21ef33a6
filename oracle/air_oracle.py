"""CPU oracle for the AIR hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``tf-attend-infer-repeat_b200``) never imports it and has no CPU fallback.

PARITY PIN: the reference (aakhundov/tf-attend-infer-repeat) is TensorFlow-1.3 graph code;
TensorFlow is not installable here and the reference ships no tests, golden vectors or
fixtures.  This file is an op-for-op restatement: one torch CPU
op per TF op, fp32, so every intermediate is rounded exactly where TF's op-at-a-time
executor rounds it (no FMA contraction).  It is pinned against the reference's OWN
serialized graph (model/air-model.meta: forward loop, autodiff gradient graph, clip,
ApplyAdam, test model) executed by the numpy graph interpreter in oracle/tfgraph --
tests/test_reference_graph.py: ST bit-exact, loss identical, 36 gradients <= 7e-7 on the
covered fixture, fp64 agreement on realistic poses.  TensorFlow's own kernels never ran:
the per-op arithmetic of that interpreter is itself restated.  The
checked-in SOURCE (air/*.py, whole AIRModel in test mode incl. the cnn=True branch that no
saved graph contains) is also executed, through the numpy tf shim oracle/tfgraph/tf_shim.py,
and compared with this file in tests/test_reference_source.py.  TF-library semantics that cannot be read from the reference tree
(LSTM gate order, softplus thresholds, Adam epsilon placement ...) are restated from the
published TF 1.3 behaviour and are listed in SURVEY.md section 8(c).

Every function cites the reference file:line it follows (paths relative to the
reference checkout).  All sampling noise is an explicit argument.

Passing ``dtype=torch.float64`` tensors runs the same op sequence in fp64; that is the
"truth" used to show that oracle-fp32 and the CUDA path are equally close to it.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

EPS = 10e-10  # == 1e-9, the reference's spelling (concrete.py:4, air_model.py:587)


# --------------------------------------------------------------------------------------
# TF-library primitives restated
# --------------------------------------------------------------------------------------
def tf_linspace(start: float, stop: float, num: int, dtype=torch.float32) -> torch.Tensor:
    """tf.linspace: ``start + i * step`` with ``step = (stop-start)/(num-1)`` computed in
    the output dtype (TF LinSpaceOp).  Used by transformer.py:127-130."""
    start_t = torch.tensor(start, dtype=dtype)
    if num == 1:
        return start_t.reshape(1)
    step = (torch.tensor(stop, dtype=dtype) - start_t) / torch.tensor(num - 1, dtype=dtype)
    i = torch.arange(num, dtype=dtype)
    return start_t + step * i


def tf_softplus(x: torch.Tensor) -> torch.Tensor:
    """tf.nn.softplus (Eigen functor): x > -thr -> x ; x < thr -> exp(x) ; else
    log(exp(x) + 1), thr = log(eps_fp) + 2.  Its gradient is g / (exp(-x) + 1)
    (TF SoftplusGrad), which is what autograd of the custom Function below returns.
    vae.py:6 (activation default)."""
    return _TFSoftplus.apply(x)


class _TFSoftplus(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        thr = math.log(torch.finfo(x.dtype).eps) + 2.0
        e = torch.exp(x)
        mid = torch.log(e + 1.0)
        return torch.where(x > -thr, x, torch.where(x < thr, e, mid))

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g / (torch.exp(-x) + 1.0)


def fully_connected(x, w, b, activation=None):
    """tf.contrib.layers.fully_connected: act(x @ W + b) (MatMul, BiasAdd, act)."""
    y = torch.matmul(x, w) + b
    if activation is not None:
        y = activation(y)
    return y


# --------------------------------------------------------------------------------------
# Spatial transformer  (air/transformer.py)
# --------------------------------------------------------------------------------------
def _meshgrid(height: int, width: int, dtype=torch.float32) -> torch.Tensor:
    """transformer.py:119-136.  ones[h,1] @ linspace[1,w] is exact (x*1), so the matmuls
    reduce to broadcasts."""
    x_lin = tf_linspace(-1.0, 1.0, width, dtype)
    y_lin = tf_linspace(-1.0, 1.0, height, dtype)
    x_t = x_lin.reshape(1, width).expand(height, width)
    y_t = y_lin.reshape(height, 1).expand(height, width)
    ones = torch.ones(1, height * width, dtype=dtype)
    return torch.cat([x_t.reshape(1, -1), y_t.reshape(1, -1), ones], dim=0)  # [3, h*w]


def _interpolate(im, x, y, out_size):
    """transformer.py:56-117, op for op.  ``im`` is NHWC, ``x``/``y`` are flat [B*oh*ow]."""
    num_batch, height, width, channels = im.shape
    dtype = im.dtype
    height_f = torch.tensor(float(height), dtype=dtype)
    width_f = torch.tensor(float(width), dtype=dtype)
    out_height, out_width = out_size
    max_y = height - 1
    max_x = width - 1

    # :75-76  (x + 1.0)*(width_f-1.001) / 2.0   -- three separately rounded ops
    x = (x + 1.0) * (width_f - 1.001) / 2.0
    y = (y + 1.0) * (height_f - 1.001) / 2.0

    # :79-82
    x0 = torch.floor(x).to(torch.int64)
    x1 = x0 + 1
    y0 = torch.floor(y).to(torch.int64)
    y1 = y0 + 1

    # :84-87
    x0 = torch.clamp(x0, 0, max_x)
    x1 = torch.clamp(x1, 0, max_x)
    y0 = torch.clamp(y0, 0, max_y)
    y1 = torch.clamp(y1, 0, max_y)
    dim2 = width
    dim1 = width * height
    # :90  _repeat(range(B)*dim1, oh*ow) (:48-54) is an int matmul that tiles the base
    base = (torch.arange(num_batch, dtype=torch.int64) * dim1).repeat_interleave(out_height * out_width)
    base_y0 = base + y0 * dim2
    base_y1 = base + y1 * dim2
    idx_a = base_y0 + x0
    idx_b = base_y1 + x0
    idx_c = base_y0 + x1
    idx_d = base_y1 + x1

    # :100-105
    im_flat = im.reshape(-1, channels)
    Ia = im_flat[idx_a]
    Ib = im_flat[idx_b]
    Ic = im_flat[idx_c]
    Id = im_flat[idx_d]

    # :108-115  weights from CLIPPED corners and UNCLIPPED coordinates
    x0_f = x0.to(dtype)
    x1_f = x1.to(dtype)
    y0_f = y0.to(dtype)
    y1_f = y1.to(dtype)
    wa = ((x1_f - x) * (y1_f - y)).unsqueeze(1)
    wb = ((x1_f - x) * (y - y0_f)).unsqueeze(1)
    wc = ((x - x0_f) * (y1_f - y)).unsqueeze(1)
    wd = ((x - x0_f) * (y - y0_f)).unsqueeze(1)
    # :116  add_n is left-to-right
    return ((wa * Ia + wb * Ib) + wc * Ic) + wd * Id


def transformer(U, theta, out_size, name="SpatialTransformer", **kwargs):
    """transformer.py:18 / _transform :138-171.  U [B,H,W,C], theta [B,6] or [B,2,3]."""
    num_batch, height, width, num_channels = U.shape
    dtype = U.dtype
    theta = theta.reshape(-1, 2, 3).to(dtype)
    out_height, out_width = int(out_size[0]), int(out_size[1])
    grid = _meshgrid(out_height, out_width, dtype)  # [3, n]
    # :159 BatchMatMul with k=3, no FMA: fl(fl(fl(a0 b0)+fl(a1 b1))+fl(a2 b2))
    g0, g1, g2 = grid[0][None, :], grid[1][None, :], grid[2][None, :]
    x_s = (theta[:, 0, 0:1] * g0 + theta[:, 0, 1:2] * g1) + theta[:, 0, 2:3] * g2
    y_s = (theta[:, 1, 0:1] * g0 + theta[:, 1, 1:2] * g1) + theta[:, 1, 2:3] * g2
    out = _interpolate(U, x_s.reshape(-1), y_s.reshape(-1), (out_height, out_width))
    return out.reshape(num_batch, out_height, out_width, num_channels)


def batch_transformer(U, thetas, out_size, name="BatchSpatialTransformer"):
    """transformer.py:178-195."""
    num_batch, num_transforms = thetas.shape[:2]
    idx = torch.arange(num_batch).repeat_interleave(num_transforms)
    return transformer(U[idx], thetas, out_size)


# --------------------------------------------------------------------------------------
# Concrete / Gumbel-softmax  (air/concrete.py)
# --------------------------------------------------------------------------------------
def concrete_binary_pre_sigmoid_sample(log_odds, temperature, u, eps=EPS):
    """concrete.py:20-27; ``u`` replaces tf.random_uniform (:23)."""
    noise = torch.log(u + eps) - torch.log(1.0 - u + eps)
    return (log_odds + noise) / temperature


def concrete_binary_sample(log_odds, temperature, u, hard=False, eps=EPS):
    """concrete.py:4-17."""
    noise = torch.log(u + eps) - torch.log(1.0 - u + eps)
    y = log_odds + noise
    sig_y = torch.sigmoid(y / temperature)
    if hard:
        sig_y = (torch.round(sig_y) - sig_y).detach() + sig_y
    return y, sig_y


def concrete_binary_kl_mc_sample(y, prior_log_odds, prior_temperature,
                                 posterior_log_odds, posterior_temperature, eps=EPS):
    """concrete.py:30-43.  Scalars are promoted to tensors of y's dtype first, as TF does
    with Python constants."""
    def t(v):
        return v if torch.is_tensor(v) else torch.tensor(v, dtype=y.dtype)

    prior_log_odds, prior_temperature = t(prior_log_odds), t(prior_temperature)
    posterior_log_odds, posterior_temperature = t(posterior_log_odds), t(posterior_temperature)

    y_times_prior_temp = y * prior_temperature
    log_prior = torch.log(prior_temperature + eps) - y_times_prior_temp + prior_log_odds - \
        2.0 * torch.log(1.0 + torch.exp(-y_times_prior_temp + prior_log_odds) + eps)

    y_times_posterior_temp = y * posterior_temperature
    log_posterior = torch.log(posterior_temperature + eps) - y_times_posterior_temp + posterior_log_odds - \
        2.0 * torch.log(1.0 + torch.exp(-y_times_posterior_temp + posterior_log_odds) + eps)

    return log_posterior - log_prior


# --------------------------------------------------------------------------------------
# VAE  (air/vae.py)
# --------------------------------------------------------------------------------------
def vae(inputs, params, prefix, n_rec, n_gen, noise_latent, noise_like,
        likelihood_std=0.0, activation=tf_softplus):
    """vae.py:5-43.  ``params[prefix + 'recognition_1/weights']`` ... follow the
    checkpoint names.  Returns (reconstruction, mean, log_variance, mean)
    -- the 4th item is the MEAN, not the sample (vae.py:43)."""
    h = inputs
    for i in range(n_rec):
        s = f"{prefix}recognition_{i + 1}/"
        h = fully_connected(h, params[s + "weights"], params[s + "biases"], activation)
    mean = fully_connected(h, params[prefix + "rec_mean/weights"], params[prefix + "rec_mean/biases"])
    log_var = fully_connected(h, params[prefix + "rec_log_variance/weights"],
                              params[prefix + "rec_log_variance/biases"])
    sample = mean + noise_latent * torch.sqrt(torch.exp(log_var))  # :23-24
    h = sample
    for i in range(n_gen):
        s = f"{prefix}generative_{i + 1}/"
        h = fully_connected(h, params[s + "weights"], params[s + "biases"], activation)
    gen_mean = fully_connected(h, params[prefix + "gen_mean/weights"], params[prefix + "gen_mean/biases"])
    gen_sample = gen_mean + noise_like * likelihood_std  # :37-38
    return torch.sigmoid(gen_sample), mean, log_var, mean


# --------------------------------------------------------------------------------------
# Parameters
# --------------------------------------------------------------------------------------
CNN_OUT = 12 * 12  # spatial size after the two valid 2x2/2 max-pools of a 50x50 canvas (air_model.py:533)


def cnn_frontend(x_img, p, filters=8):
    """air_model.py:510-535: tf.layers.conv2d(5x5, 'same', relu) x3 with max_pooling2d(2, 2) ('valid') after the
    first two, on the canvas reshaped to [-1, 50, 50, 1] (hard-coded in the reference), flattened in NHWC order.
    Kernels are stored as TF stores them (HWIO); TF's conv2d is a cross-correlation, like torch's."""
    import torch.nn.functional as F
    x = x_img.reshape(-1, 1, 50, 50)  # one channel: NCHW and NHWC coincide

    def conv(t, name):
        w = p[f"cnn/{name}/kernel"].permute(3, 2, 0, 1)  # HWIO -> OIHW
        return torch.relu(F.conv2d(t, w, p[f"cnn/{name}/bias"], padding=2))

    t = F.max_pool2d(conv(x, "conv1"), 2, 2)    # 50 -> 25
    t = F.max_pool2d(conv(t, "conv2"), 2, 2)    # 25 -> 12 (valid)
    t = conv(t, "conv3")
    return t.permute(0, 2, 3, 1).reshape(-1, CNN_OUT * filters)  # NHWC flatten (:533)


def param_shapes(canvas_size=50, windows_size=28, rnn_units=256, vae_latent_dimensions=50,
                 vae_recognition_units=(512, 256), vae_generative_units=(256, 512),
                 scale_hidden_units=64, shift_hidden_units=64, z_pres_hidden_units=64,
                 rnn_input_dim=None, cnn=False, cnn_filters=8):
    """Trainable tensors in checkpoint order-independent form: name -> shape.  Names are
    the checkpoint names (model/air-model.index) minus the ``air/rnn/`` prefix; the CNN front-end's
    variables live under ``air/cnn/`` in the reference graph and are keyed ``cnn/...`` here."""
    in_dim = canvas_size * canvas_size if rnn_input_dim is None else rnn_input_dim
    win = windows_size * windows_size
    s = OrderedDict()
    if cnn:
        in_dim = CNN_OUT * cnn_filters
        prev = 1
        for name in ("conv1", "conv2", "conv3"):
            s[f"cnn/{name}/kernel"] = (5, 5, prev, cnn_filters)
            s[f"cnn/{name}/bias"] = (cnn_filters,)
            prev = cnn_filters
    s["rnn/kernel"] = (in_dim + rnn_units, 4 * rnn_units)
    s["rnn/bias"] = (4 * rnn_units,)
    for head, hid, out in (("scale", scale_hidden_units, 1), ("shift", shift_hidden_units, 2)):
        for stat in ("mean", "log_variance"):
            s[f"{head}/{stat}/hidden/weights"] = (rnn_units, hid)
            s[f"{head}/{stat}/hidden/biases"] = (hid,)
            s[f"{head}/{stat}/output/weights"] = (hid, out)
            s[f"{head}/{stat}/output/biases"] = (out,)
    s["z_pres/log_odds/hidden/weights"] = (rnn_units, z_pres_hidden_units)
    s["z_pres/log_odds/hidden/biases"] = (z_pres_hidden_units,)
    s["z_pres/log_odds/output/weights"] = (z_pres_hidden_units, 1)
    s["z_pres/log_odds/output/biases"] = (1,)
    prev = win
    for i, u in enumerate(vae_recognition_units):
        s[f"vae/recognition_{i + 1}/weights"] = (prev, u)
        s[f"vae/recognition_{i + 1}/biases"] = (u,)
        prev = u
    for nm in ("rec_mean", "rec_log_variance"):
        s[f"vae/{nm}/weights"] = (prev, vae_latent_dimensions)
        s[f"vae/{nm}/biases"] = (vae_latent_dimensions,)
    prev = vae_latent_dimensions
    for i, u in enumerate(vae_generative_units):
        s[f"vae/generative_{i + 1}/weights"] = (prev, u)
        s[f"vae/generative_{i + 1}/biases"] = (u,)
        prev = u
    s["vae/gen_mean/weights"] = (prev, win)
    s["vae/gen_mean/biases"] = (win,)
    return s


def init_params(seed=0, dtype=torch.float32, **shape_kwargs):
    """Xavier/glorot-uniform weights (+-sqrt(6/(fan_in+fan_out))), zero biases -- the TF
    defaults of fully_connected and of BasicLSTMCell's get_variable."""
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, shape in param_shapes(**shape_kwargs).items():
        if len(shape) >= 2:  # glorot-uniform; conv kernels [kh,kw,in,out]: fan = kh*kw*in / kh*kw*out (TF)
            rf = math.prod(shape[:-2])
            lim = math.sqrt(6.0 / (rf * shape[-2] + rf * shape[-1]))
            w = (torch.rand(shape, generator=g, dtype=torch.float64) * 2.0 - 1.0) * lim
            p[name] = w.to(dtype)
        else:
            p[name] = torch.zeros(shape, dtype=dtype)
    return p


def make_noise(seed, T, B, latent=50, win=784, dtype=torch.float32):
    """The five injected noise tensors, shaped [T,B,...]."""
    g = torch.Generator().manual_seed(seed)
    return {
        "scale": torch.randn(T, B, 1, generator=g).to(dtype),        # air_model.py:301
        "shift": torch.randn(T, B, 2, generator=g).to(dtype),        # air_model.py:318
        "vae_latent": torch.randn(T, B, latent, generator=g).to(dtype),   # vae.py:23
        "vae_like": torch.randn(T, B, win, generator=g).to(dtype),        # vae.py:37
        "concrete_u": torch.rand(T, B, generator=g).to(dtype),       # concrete.py:23
    }


def annealed_value(schedule, global_step, eps=EPS):
    """air_model.py:94-121 -- tf.train.exponential_decay in fp32, then min/max/log."""
    f32 = torch.float32
    p = torch.tensor(float(global_step), dtype=f32) / torch.tensor(float(schedule["iters"]), dtype=f32)
    if schedule.get("staircase", False):
        p = torch.floor(p)
    value = torch.tensor(float(schedule["init"]), dtype=f32) * torch.pow(
        torch.tensor(float(schedule["factor"]), dtype=f32), p)
    if "min" in schedule:
        value = torch.maximum(value, torch.tensor(float(schedule["min"]), dtype=f32))
    if "max" in schedule:
        value = torch.minimum(value, torch.tensor(float(schedule["max"]), dtype=f32))
    if schedule.get("log", False):
        value = torch.log(value + eps)
    return value


# --------------------------------------------------------------------------------------
# AIR model  (air/air_model.py)
# --------------------------------------------------------------------------------------
DEFAULT_HYPER = dict(  # training.py:100-122 ("default" of the checkpoint / README)
    max_steps=3, max_digits=2, rnn_units=256, canvas_size=50, windows_size=28,
    vae_latent_dimensions=50, vae_recognition_units=(512, 256), vae_generative_units=(256, 512),
    scale_prior_mean=-1.0, scale_prior_variance=0.05, shift_prior_mean=0.0, shift_prior_variance=1.0,
    vae_prior_mean=0.0, vae_prior_variance=1.0, vae_likelihood_std=0.3,
    scale_hidden_units=64, shift_hidden_units=64, z_pres_hidden_units=64,
    z_pres_prior_log_odds=-0.01, z_pres_temperature=1.0, stopping_threshold=0.99,
    learning_rate=1e-4, gradient_clipping_norm=1.0, cnn=False,
)
DEFAULT_ANNEALING = {  # training.py:110-115
    "z_pres_prior_log_odds": {"init": 10000.0, "min": 0.000000001, "factor": 0.1,
                              "iters": 3000, "staircase": False, "log": True},
}


class AIROracle:
    """Functional restatement of AIRModel._create_model (air_model.py:269-611) and the
    optimizer tail (:651-694).  Holds parameters, Adam slots and global_step."""

    def __init__(self, params=None, annealing_schedules=None, train=True, seed=0,
                 dtype=torch.float32, **hyper):
        self.h = dict(DEFAULT_HYPER)
        self.h.update(hyper)
        self.h.setdefault("cnn_filters", 8)
        if self.h["cnn"]:
            assert self.h["canvas_size"] == 50, "the reference's CNN front-end hard-codes 50x50 canvases (:512, :533)"
        self.train = train
        self.dtype = dtype
        self.annealing = annealing_schedules
        h = self.h
        self.params = params if params is not None else init_params(
            seed=seed, dtype=dtype, canvas_size=h["canvas_size"], windows_size=h["windows_size"],
            rnn_units=h["rnn_units"], vae_latent_dimensions=h["vae_latent_dimensions"],
            vae_recognition_units=h["vae_recognition_units"], vae_generative_units=h["vae_generative_units"],
            scale_hidden_units=h["scale_hidden_units"], shift_hidden_units=h["shift_hidden_units"],
            z_pres_hidden_units=h["z_pres_hidden_units"], cnn=h["cnn"], cnn_filters=h["cnn_filters"])
        self.global_step = 0
        self.adam_m = {k: torch.zeros_like(v) for k, v in self.params.items()}
        self.adam_v = {k: torch.zeros_like(v) for k, v in self.params.items()}
        self.beta1_power = torch.tensor(0.9, dtype=torch.float32)
        self.beta2_power = torch.tensor(0.999, dtype=torch.float32)

    # -- hyper-parameters that may be annealed (air_model.py:76-82)
    def hyper(self, name):
        if self.annealing and name in self.annealing:
            return annealed_value(self.annealing[name], self.global_step).to(self.dtype)
        return self.h[name]

    def _head(self, outputs, head, stat):
        p = self.params
        hidden = fully_connected(outputs, p[f"{head}/{stat}/hidden/weights"],
                                 p[f"{head}/{stat}/hidden/biases"], torch.relu)
        return fully_connected(hidden, p[f"{head}/{stat}/output/weights"], p[f"{head}/{stat}/output/biases"])

    def forward(self, input_images, target_num_digits, noise, keep=None, taps=None):
        """Runs exactly ``max_steps`` iterations of body() (air_model.py:278-508); items
        that already stopped contribute exact +0.0 (tf.where), so loss / canvas / digit
        counts equal the reference's early-exit loop.  ``executed_steps``
        is the trip count the reference's cond() (:271-275) would have produced.

        ``keep`` (optional dict) receives intermediates for stage-wise parity checks; ``taps``
        (optional dict) receives the raw per-step graph tensors with retain_grad() set, so a
        later backward() exposes d(loss)/d(intermediate) for localising gradient mismatches."""
        h, p, dt = self.h, self.params, self.dtype
        B = input_images.shape[0]
        cs, ws = h["canvas_size"], h["windows_size"]
        thr = h["stopping_threshold"]
        x_img = input_images.to(dt)
        U_canvas = x_img.reshape(-1, cs, cs).unsqueeze(3)  # :331
        rnn_input = cnn_frontend(x_img, p, h["cnn_filters"]) if h["cnn"] else x_img  # :510-535

        stopping_sum = torch.zeros(B, dtype=dt)
        c_state = torch.zeros(B, h["rnn_units"], dtype=dt)
        h_state = torch.zeros(B, h["rnn_units"], dtype=dt)
        running_recon = torch.zeros_like(x_img)
        running_loss = torch.zeros(B, dtype=dt)
        running_digits = torch.zeros(B, dtype=torch.int32)
        ta = {k: [] for k in ("scales", "shifts", "z_pres_probs", "z_pres_kls", "scale_kls", "shift_kls",
                              "vae_kls", "st_backward", "windows", "latents", "z_pres", "stop_masks",
                              "windows_in", "thetas")}
        executed_steps = 0
        z_temp = self.hyper("z_pres_temperature")
        z_prior = self.hyper("z_pres_prior_log_odds")

        def log_c(v):  # tf.log(python constant) evaluated in fp32 (air_model.py:72-74)
            return torch.log(torch.tensor(float(v), dtype=dt))

        scale_prior_lv = log_c(h["scale_prior_variance"])
        shift_prior_lv = log_c(h["shift_prior_variance"])
        vae_prior_lv = log_c(h["vae_prior_variance"])

        for step in range(h["max_steps"]):
            if bool((stopping_sum < thr).any()) and executed_steps == step:
                executed_steps = step + 1  # cond() true on entry to this iteration

            # ---- LSTM step, BasicLSTMCell: gates i, j, f, o; forget bias 1.0 (:284-286)
            concat = torch.matmul(torch.cat([rnn_input, h_state], dim=1), p["rnn/kernel"]) + p["rnn/bias"]
            gi, gj, gf, go = torch.split(concat, h["rnn_units"], dim=1)
            c_state = c_state * torch.sigmoid(gf + 1.0) + torch.sigmoid(gi) * torch.tanh(gj)
            h_state = torch.tanh(c_state) * torch.sigmoid(go)
            outputs = h_state

            # ---- scale (:288-303)
            scale_mean = self._head(outputs, "scale", "mean")
            scale_log_variance = self._head(outputs, "scale", "log_variance")
            scale_variance = torch.exp(scale_log_variance)
            scale = torch.sigmoid(scale_mean + noise["scale"][step] * torch.sqrt(scale_variance))  # :123-128
            ta["scales"].append(scale)
            s = scale[:, 0]

            # ---- shift (:305-320)
            shift_mean = self._head(outputs, "shift", "mean")
            shift_log_variance = self._head(outputs, "shift", "log_variance")
            shift_variance = torch.exp(shift_log_variance)
            shift = torch.tanh(shift_mean + noise["shift"][step] * torch.sqrt(shift_variance))
            ta["shifts"].append(shift)
            x, y = shift[:, 0], shift[:, 1]

            # ---- ST forward: canvas -> window (:322-333)
            zeros = torch.zeros_like(s)
            theta = torch.stack([torch.stack([s, zeros, x], dim=1),
                                 torch.stack([zeros, s, y], dim=1)], dim=1)
            window = transformer(U_canvas, theta, [ws, ws])[:, :, :, 0]
            ta["thetas"].append(theta)
            ta["windows_in"].append(window.reshape(-1, ws * ws))

            # ---- VAE (:335-349)
            vae_recon, vae_mean, vae_log_variance, vae_latent = vae(
                window.reshape(-1, ws * ws), p, "vae/", len(h["vae_recognition_units"]),
                len(h["vae_generative_units"]), noise["vae_latent"][step], noise["vae_like"][step],
                h["vae_likelihood_std"])
            ta["windows"].append(vae_recon)
            ta["latents"].append(vae_latent)

            # ---- ST backward: window -> canvas (:351-366); three separate divisions
            theta_recon = torch.stack([torch.stack([1.0 / s, zeros, -x / s], dim=1),
                                       torch.stack([zeros, 1.0 / s, -y / s], dim=1)], dim=1)
            ta["st_backward"].append(theta_recon)
            window_recon = transformer(vae_recon.reshape(-1, ws, ws).unsqueeze(3), theta_recon, [cs, cs])[:, :, :, 0]

            # ---- z_pres (:368-396)
            z_pres_log_odds = self._head(outputs, "z_pres", "log_odds")[:, 0]
            z_pres_pre_sigmoid = concrete_binary_pre_sigmoid_sample(
                z_pres_log_odds, z_temp, noise["concrete_u"][step])
            z_pres = torch.sigmoid(z_pres_pre_sigmoid)
            if not self.train:
                z_pres = torch.round(z_pres)  # tf.round: half-to-even, no gradient
            ta["z_pres_probs"].append(torch.sigmoid(z_pres_log_odds))
            ta["z_pres"].append(z_pres)

            # ---- z_pres KL, masked by the PREVIOUS stopping_sum (:398-418)
            z_pres_kl = concrete_binary_kl_mc_sample(z_pres_pre_sigmoid, z_prior, z_temp, z_pres_log_odds, z_temp)
            running_loss = running_loss + torch.where(stopping_sum < thr, z_pres_kl, torch.zeros_like(running_loss))
            ta["z_pres_kls"].append(z_pres_kl)

            # ---- stop / count (:424-427)
            stopping_sum = stopping_sum + (1.0 - z_pres)
            live = stopping_sum < thr
            running_digits = running_digits + live.to(torch.int32)
            ta["stop_masks"].append(live)

            # ---- canvas (:429-439)
            running_recon = running_recon + torch.where(
                live.unsqueeze(1), z_pres.unsqueeze(1) * window_recon.reshape(-1, cs * cs),
                torch.zeros_like(running_recon))

            # ---- Gaussian KLs masked by the NEW sum (:441-496)
            scale_kl = 0.5 * torch.sum(
                scale_prior_lv - scale_log_variance - 1.0 + scale_variance / h["scale_prior_variance"] +
                torch.square(scale_mean - h["scale_prior_mean"]) / h["scale_prior_variance"], 1)
            running_loss = running_loss + torch.where(live, scale_kl, torch.zeros_like(running_loss))
            ta["scale_kls"].append(scale_kl)

            shift_kl = 0.5 * torch.sum(
                shift_prior_lv - shift_log_variance - 1.0 + shift_variance / h["shift_prior_variance"] +
                torch.square(shift_mean - h["shift_prior_mean"]) / h["shift_prior_variance"], 1)
            running_loss = running_loss + torch.where(live, shift_kl, torch.zeros_like(running_loss))
            ta["shift_kls"].append(shift_kl)

            vae_kl = 0.5 * torch.sum(
                vae_prior_lv - vae_log_variance - 1.0 + torch.exp(vae_log_variance) / h["vae_prior_variance"] +
                torch.square(vae_mean - h["vae_prior_mean"]) / h["vae_prior_variance"], 1)
            running_loss = running_loss + torch.where(live, vae_kl, torch.zeros_like(running_loss))
            ta["vae_kls"].append(vae_kl)

        if taps is not None:
            for k in ("thetas", "st_backward", "windows", "windows_in", "z_pres", "scales", "shifts"):
                taps[k] = ta[k]
                for tns in ta[k]:
                    if tns.requires_grad:
                        tns.retain_grad()
        # ---- post-loop (:569-611)
        out = {
            "rec_num_digits": running_digits,
            "rec_scales": torch.stack(ta["scales"]).permute(1, 0, 2),
            "rec_shifts": torch.stack(ta["shifts"]).permute(1, 0, 2),
            "rec_st_back": torch.stack(ta["st_backward"]).permute(1, 0, 2, 3),
            "rec_windows": torch.stack(ta["windows"]).permute(1, 0, 2),
            "rec_latents": torch.stack(ta["latents"]).permute(1, 0, 2),
            "z_pres_probs": torch.stack(ta["z_pres_probs"]).t(),
            "z_pres_kls": torch.stack(ta["z_pres_kls"]).t(),
            "scale_kls": torch.stack(ta["scale_kls"]).t(),
            "shift_kls": torch.stack(ta["shift_kls"]).t(),
            "vae_kls": torch.stack(ta["vae_kls"]).t(),
            # extras (not reference attributes) for stage-wise parity
            "z_pres": torch.stack(ta["z_pres"]).t(),
            "stop_masks": torch.stack(ta["stop_masks"]).t(),
            "stopping_sum": stopping_sum,
            "windows_in": torch.stack(ta["windows_in"]).permute(1, 0, 2),
            "thetas": torch.stack(ta["thetas"]).permute(1, 0, 2, 3),
            "canvas_raw": running_recon,
            "running_loss": running_loss,
            "executed_steps": executed_steps,
        }
        reconstruction = torch.clamp(running_recon, 0.0, 1.0)  # max(min(r,1),0), same tie rules (:582)
        reconstruction_loss = -torch.sum(
            x_img * torch.log(reconstruction + EPS) + (1.0 - x_img) * torch.log(1.0 - reconstruction + EPS), 1)
        loss_vec = running_loss + reconstruction_loss
        accuracy = (target_num_digits.to(torch.int32) == running_digits).to(torch.float32)
        out.update(reconstruction=reconstruction, reconstruction_loss=reconstruction_loss,
                   loss_per_item=loss_vec, loss=torch.mean(loss_vec), accuracy=torch.mean(accuracy))
        if keep is not None:
            keep.update(out)
        return out

    # -- gradients of the mean loss w.r.t. the 36 trainables (air_model.py:655)
    def loss_and_grads(self, input_images, target_num_digits, noise):
        leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in self.params.items())
        saved, self.params = self.params, leaves
        try:
            out = self.forward(input_images, target_num_digits, noise)
            grads = torch.autograd.grad(out["loss"], list(leaves.values()), allow_unused=True)
        finally:
            self.params = saved
        grads = OrderedDict((k, (g if g is not None else torch.zeros_like(v)).detach())
                            for (k, v), g in zip(leaves.items(), grads))
        return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}, grads

    @staticmethod
    def clip_by_global_norm(grads, clip_norm):
        """tf.clip_by_global_norm (air_model.py:673): norm = sqrt(2*sum(l2_loss(g)));
        scale = clip * min(1/norm, 1/clip)."""
        half = [torch.sum(g * g) / 2.0 for g in grads.values()]
        norm = torch.sqrt(torch.sum(torch.stack(half)) * 2.0)
        clip = torch.tensor(float(clip_norm), dtype=norm.dtype)
        scale = clip * torch.minimum(1.0 / norm, 1.0 / clip)
        return OrderedDict((k, g * scale) for k, g in grads.items()), norm

    def adam_apply(self, grads, lr=None, beta1=0.9, beta2=0.999, epsilon=1e-8):
        """tf.train.AdamOptimizer / ApplyAdam (air_model.py:654, 692): epsilon OUTSIDE the
        bias correction; beta powers are fp32 variables updated after the apply."""
        lr = self.hyper("learning_rate") if lr is None else lr
        lr = lr if torch.is_tensor(lr) else torch.tensor(float(lr), dtype=torch.float32)
        alpha = (lr * torch.sqrt(1.0 - self.beta2_power) / (1.0 - self.beta1_power)).to(self.dtype)
        for k, g in grads.items():
            m, v = self.adam_m[k], self.adam_v[k]
            m += (g - m) * (1.0 - beta1)
            v += (g * g - v) * (1.0 - beta2)
            self.params[k] = self.params[k] - (m * alpha) / (torch.sqrt(v) + epsilon)
        self.beta1_power = self.beta1_power * beta1
        self.beta2_power = self.beta2_power * beta2
        self.global_step += 1

    def train_step(self, input_images, target_num_digits, noise):
        """One ``sess.run(model.training)`` (training.py:212-224)."""
        out, grads = self.loss_and_grads(input_images, target_num_digits, noise)
        norm = None
        if self.h["gradient_clipping_norm"] is not None:
            grads, norm = self.clip_by_global_norm(grads, self.h["gradient_clipping_norm"])
        self.adam_apply(grads)
        out["grad_global_norm"] = norm
        return out, grads


# --------------------------------------------------------------------------------------
# Synthetic multi-digit canvases (stand-in for multi_mnist.py:82-183)
# --------------------------------------------------------------------------------------
def synthetic_canvases(B, canvas_size=50, max_digits=2, seed=0, digit_size=(14, 24)):
    """0..max_digits stroke-like blobs per 50x50 canvas, uniform placement with
    pixel-overlap rejection, background exactly 0.0, values in (0,1]; digit count uniform
    over {0..max_digits}.  Returns (images [B, cs*cs] fp32, counts [B] int32)."""
    import numpy as np
    rng = np.random.RandomState(seed)
    imgs = np.zeros((B, canvas_size, canvas_size), dtype=np.float32)
    counts = rng.randint(0, max_digits + 1, size=B).astype(np.int32)
    yy, xx = np.mgrid[0:32, 0:32].astype(np.float32)
    for b in range(B):
        for _ in range(int(counts[b])):
            for _attempt in range(50):
                hh = rng.randint(digit_size[0], digit_size[1] + 1)
                ww = rng.randint(digit_size[0] // 2 + 3, digit_size[1] + 1)
                # a "stroke": thick ellipse outline + a bar, thresholded and blurred-ish
                cy, cx = (hh - 1) / 2.0, (ww - 1) / 2.0
                ry, rx = max(hh / 2.0 - 1.5, 2.0), max(ww / 2.0 - 1.5, 2.0)
                r = np.sqrt(((yy[:hh, :ww] - cy) / ry) ** 2 + ((xx[:hh, :ww] - cx) / rx) ** 2)
                ring = np.clip(1.0 - np.abs(r - 0.8) * 3.0, 0.0, 1.0)
                bar = np.clip(1.0 - np.abs(xx[:hh, :ww] - cx - rng.uniform(-2, 2)) / 1.6, 0.0, 1.0) * rng.randint(0, 2)
                blob = np.maximum(ring, bar).astype(np.float32)
                blob[blob < 0.15] = 0.0
                top = rng.randint(0, canvas_size - hh + 1)
                left = rng.randint(0, canvas_size - ww + 1)
                region = imgs[b, top:top + hh, left:left + ww]
                if np.any((region > 0) & (blob > 0)):
                    continue  # overlap rejection (multi_mnist.py:141-160)
                imgs[b, top:top + hh, left:left + ww] = np.maximum(region, blob)
                break
    return torch.from_numpy(imgs.reshape(B, -1)), torch.from_numpy(counts)
