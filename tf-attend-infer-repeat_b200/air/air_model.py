"""AIRModel with the reference constructor (air/air_model.py:13-22), B200-native underneath.

The reference builds a TF graph (tf.while_loop over body(), air_model.py:278-508) and the
caller fetches result attributes with session.run.  Here the object is eager:

    model = AIRModel(images, targets, max_steps=3, ..., train=True)       # same kwargs
    model.train_step()            # == sess.run(model.training): forward, backward, clip, Adam
    model.run()                   # forward only (what fetching .loss / .rec_* evaluates)
    model.loss, model.accuracy, model.rec_num_digits, model.rec_scales, ...   # CUDA tensors

``input_images`` / ``target_num_digits`` play the role of the reference's input tensors:
persistent CUDA buffers that the caller refills (``feed``).  All arithmetic is hand-written
CUDA behind the C ABI (csrc/); the schedule below only orders kernel launches on the current
stream, so a whole step can be captured in a CUDA graph (``capture()``).

Differences from the reference, by design:
  * the loop always runs ``max_steps`` iterations (no device->host sync for the batch-global
    early exit, air_model.py:271-275); stopped items contribute exact +0.0, so loss, canvas,
    digit counts and gradients are identical.  Per-step outputs therefore have
    time dimension ``max_steps``; ``executed_steps`` gives the reference's dynamic trip count.
  * sampling noise can be injected (``noise=`` dict of [T,B,...] tensors) for reproducibility;
    otherwise it is drawn with torch's CUDA generator into persistent buffers.
  * the image part of the LSTM input projection (x @ K[:2500]) is step-invariant
    (air_model.py:535: rnn_input is the raw image every step) and is computed once.

``reference_rounding`` (default True): the write-back backward accumulates the gradient terms of canvas pixels OUTSIDE
the attention window the way the reference's fp32 autodiff does -- four ~1e10 corner products per lit, not yet
reconstructed pixel that cancel only to rounding error -- instead of dropping them as the exact zeros they are on paper.
The reference's training depends on those residues (with the exactly-cancelled gradient the model never leaves loss
~1900); False selects the analytic kernel (st_wb_bwd_axis), whose gradient is the fp64 one.

``skip_nonfinite_updates`` (default False = the reference): with True an optimisation step whose global gradient norm is
not finite changes nothing (``skipped_updates`` counts them).  The un-cancelled corner products overflow fp32 when a
sampled window collapses to ~1e-14 of the canvas (seen once in twelve 25k-iteration runs); the reference, and the default
here, then write NaN into every variable for good.
"""
from __future__ import annotations

import math

import torch

from .. import _cabi as C
from .. import dp, ops
from .params import ParamStore
from .vae import (VAEWeights, _flat2, alloc_vae_scratch, dense_dw, vae_backward_dx, vae_bias_items, vae_forward,
                  vae_weight_grads)

_VARIABLE_SCOPES = {}
_GEMM_WORKSPACES = {}  # device -> float32 scratch tensor (split-K partial sums)

_DEVICE_ANNEALABLE = ("z_pres_prior_log_odds", "learning_rate")
# float hyper-parameters that reach the kernels as launch arguments: annealed on the host, step by step (eager steps)
_HOST_ANNEALABLE = ("scale_prior_mean", "scale_prior_variance", "shift_prior_mean", "shift_prior_variance", "vae_prior_mean",
                    "vae_prior_variance", "vae_likelihood_std", "z_pres_temperature", "stopping_threshold",
                    "gradient_clipping_norm")


def annealed_value(schedule, step, eps=10e-10):
    """air_model.py:94-121 in fp32, as TensorFlow evaluates it: exponential_decay(init, step, iters, factor[, staircase]),
    then max(., min), min(., max), log(. + eps)."""
    import numpy as np
    f = np.float32
    p = f(step) / f(schedule["iters"])
    if schedule.get("staircase", False):
        p = np.floor(p)
    v = f(schedule["init"]) * np.power(f(schedule["factor"]), p, dtype=np.float32)
    if "min" in schedule:
        v = np.maximum(v, f(schedule["min"]))
    if "max" in schedule:
        v = np.minimum(v, f(schedule["max"]))
    if schedule.get("log", False):
        v = np.log(v + f(eps), dtype=np.float32)
    return float(v)


def reset_variable_scopes():
    """Forget all shared variable stores (the equivalent of tf.reset_default_graph())."""
    _VARIABLE_SCOPES.clear()


class AIRModel:

    def __init__(self, input_images, target_num_digits,
                 max_steps=3, max_digits=2, rnn_units=256, canvas_size=50, windows_size=28,
                 vae_latent_dimensions=50, vae_recognition_units=(512, 256), vae_generative_units=(256, 512),
                 scale_prior_mean=-1.0, scale_prior_variance=0.1, shift_prior_mean=0.0, shift_prior_variance=1.0,
                 vae_prior_mean=0.0, vae_prior_variance=1.0, vae_likelihood_std=0.3,
                 scale_hidden_units=64, shift_hidden_units=64, z_pres_hidden_units=64,
                 z_pres_prior_log_odds=-2.0, z_pres_temperature=1.0, stopping_threshold=0.99,
                 learning_rate=1e-3, gradient_clipping_norm=100.0, cnn=True, cnn_filters=8,
                 num_summary_images=60, train=False, reuse=False, scope="air",
                 annealing_schedules=None, *, gemm_mode="fp32", seed=0, process_group=None, reference_rounding=True,
                 skip_nonfinite_updates=False):
        if cnn and canvas_size != 50:
            raise ValueError("the reference's CNN front-end hard-codes 50x50 canvases (air_model.py:512, 533)")
        if cnn and cnn_filters not in (4, 8, 16):
            raise NotImplementedError("the conv kernels are built for cnn_filters in (4, 8, 16); the reference's default is 8")
        if not input_images.is_cuda:
            raise C.AirError("AIRModel needs CUDA tensors (no CPU fallback)")
        C.lib()  # fail loudly if the extension is missing

        self.input_images = C.f32(input_images)
        self.target_num_digits = target_num_digits.to(torch.int32).contiguous()
        self.batch_size = int(input_images.shape[0])
        for k, v in list(locals().items()):
            if k not in ("self", "input_images", "target_num_digits", "k", "v"):
                setattr(self, k, v)
        self.vae_recognition_units = tuple(vae_recognition_units)
        self.vae_generative_units = tuple(vae_generative_units)
        self.gemm = C.GEMM_MODES[gemm_mode]
        self.device = input_images.device
        self.num_summaries, self.img_summaries, self.var_summaries, self.grad_summaries = [], [], [], []
        assert self.input_images.shape[1] == canvas_size * canvas_size
        # LSTM input: the canvas itself, or the CNN front-end's 12x12xF feature map (air_model.py:533-535)
        self.rnn_input_dim = 12 * 12 * cnn_filters if cnn else canvas_size * canvas_size

        # ---- variables: shared by scope, like tf.variable_scope(scope, reuse=reuse) (air_model.py:68)
        key = (scope, self.device)
        hu = scale_hidden_units if scale_hidden_units == shift_hidden_units == z_pres_hidden_units else \
            (scale_hidden_units, shift_hidden_units, z_pres_hidden_units)   # unequal: zero-padded to the widest
        want = dict(in_dim=self.rnn_input_dim, win=windows_size * windows_size, R=rnn_units, HU_spec=hu,
                    L=vae_latent_dimensions, rec_units=self.vae_recognition_units, gen_units=self.vae_generative_units,
                    cnn_filters=cnn_filters if cnn else None)
        if reuse:
            if key not in _VARIABLE_SCOPES:
                raise ValueError(f"variable scope {scope!r} does not exist (reuse=True)")
            self.store = _VARIABLE_SCOPES[key]
            diff = {k: (self.store.dims[k], v) for k, v in want.items() if self.store.dims[k] != v}
            if diff:   # TF: "Trying to share variable ..., but specified shape ... and found shape ..."
                raise ValueError(f"variable scope {scope!r} holds variables of other shapes (existing, requested): {diff}")
        else:
            if key in _VARIABLE_SCOPES:   # TF: "Variable air/global_step already exists, disallowed. Did you mean to set reuse=True"
                raise ValueError(f"variable scope {scope!r} already exists on {self.device}: pass reuse=True to share its "
                                 "variables, another scope= for independent ones, or call reset_variable_scopes()")
            self.store = ParamStore(self.device, want["in_dim"], want["win"], rnn_units, hu, vae_latent_dimensions,
                                    self.vae_recognition_units, self.vae_generative_units, seed=seed,
                                    cnn_filters=want["cnn_filters"])
            _VARIABLE_SCOPES[key] = self.store
        if train:
            self.store.state[4] = float(learning_rate if not self._annealed("learning_rate") else 0.0)

        # ---- annealed hyper-parameters (air_model.py:76-82): evaluated on device from global_step
        self.annealing_schedules = annealing_schedules or {}
        for name in self.annealing_schedules:
            if name not in _DEVICE_ANNEALABLE + _HOST_ANNEALABLE:
                raise ValueError(f"{name!r} cannot be annealed: it is not a float hyper-parameter of the loop "
                                 f"(annealable: {_DEVICE_ANNEALABLE + _HOST_ANNEALABLE})")
        self._host_annealed = [n for n in self.annealing_schedules if n in _HOST_ANNEALABLE]
        # derived constants the reference also keeps on the model (air_model.py:72-74), as Python floats
        _log = lambda v: math.log(v) if v > 0 else float("-inf")
        self.scale_prior_log_variance = _log(scale_prior_variance)
        self.shift_prior_log_variance = _log(shift_prior_variance)
        self.vae_prior_log_variance = _log(vae_prior_variance)
        self._prior = torch.full((1,), float(z_pres_prior_log_odds), device=self.device, dtype=torch.float32)
        self.hyper = C.Hyper(scale_prior_mean, scale_prior_variance, shift_prior_mean, shift_prior_variance,
                             vae_prior_mean, vae_prior_variance, vae_likelihood_std, z_pres_temperature,
                             stopping_threshold, 1 if train else 0)

        # ---- data parallelism: one process per GPU, gradients all-reduced over NCCL.  process_group="local" builds a
        #      single-GPU model inside a distributed job (e.g. a global-batch reference on rank 0).
        self.pg = None if process_group == "local" else process_group
        self.world = 1 if process_group == "local" else dp.world_size(process_group)

        # torch-owned split-K scratch for the weight-gradient GEMMs of the tensor-core modes, one per device, handed
        # to the library with every call (the library keeps no pointer and allocates nothing)
        self._gemm_ws = None
        if self.gemm != C.GEMM_MODES["fp32"]:
            if self.device not in _GEMM_WORKSPACES:
                # (zeros: the last 64 KB hold the split-K tickets, which every GEMM leaves at zero again)
                _GEMM_WORKSPACES[self.device] = torch.zeros(16 << 20, device=self.device, dtype=torch.float32)
            self._gemm_ws = _GEMM_WORKSPACES[self.device]
            self.gemm = ops.GemmMode(self.gemm, self._gemm_ws)   # every GEMM of the model may split K through it
        self._alloc()
        self._graphs = None
        self.noise = None
        self.rec_num_digits = self.rec_scales = None
        self.loss = self.accuracy = None
        self.training = self.train_step if train else None

    # ------------------------------------------------------------------------------------------
    def _annealed(self, name):
        return bool(getattr(self, "annealing_schedules", None)) and name in self.annealing_schedules

    @property
    def global_step(self):
        return self.store.global_step

    @property
    def skipped_updates(self):
        """optimisation steps skipped because of a non-finite gradient norm (skip_nonfinite_updates=True)"""
        return int(self.store.state[5].item())

    def _alloc(self):
        B, T, dev = self.batch_size, self.max_steps, self.device
        R, HU, L = self.rnn_units, self.store.dims["HU"], self.vae_latent_dimensions
        win, cs2 = self.windows_size ** 2, self.canvas_size ** 2
        z = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        w = self.w = {}
        w["xk"] = z(B, 4 * R)
        w["h0"] = torch.zeros(B, R, device=dev)
        w["gates"], w["c"], w["h"] = z(T, B, 4 * R), z(T, B, R), z(T, B, R)
        w["hh"] = z(T, B, 5 * HU)
        w["fields"] = torch.zeros(T, C.NF, B, device=dev)
        w["theta"], w["theta_inv"] = z(T, B, 6), z(T, B, 6)
        w["win"] = z(T, B, win)
        w["enc"] = [z(T, B, u) for u in self.vae_recognition_units]
        Lp = (L + 3) // 4 * 4  # TMA operands need a leading dimension that is a multiple of 4 floats
        w["ml"], w["zs"] = z(T, B, 2 * L), torch.zeros(T, B, Lp, device=dev)[:, :, :L]
        w["dec"] = [z(T, B, u) for u in self.vae_generative_units]
        w["recon"] = z(T, B, win)
        w["gen"] = z(B, win)
        Bp = (B + 3) // 4 * 4   # (padded to 16 bytes: the three accumulators are zeroed together by one launch)
        self._acc_base = [z(Bp), z(Bp), torch.zeros(Bp, device=dev, dtype=torch.int32)]
        w["stop"], w["loss"], w["digits"] = (t[:B] for t in self._acc_base)
        w["rec_loss"], w["loss_item"] = z(B), z(B)
        self._zero_fwd = ops.zero_items(self._acc_base)
        w["canvas"], w["reconstruction"] = z(B, cs2), z(B, cs2)
        w["out2"] = torch.zeros(2, device=dev)
        # the four Gaussian noise tensors are views of one buffer: one normal_ launch per step (sizes padded to
        # 16 bytes so that every view stays aligned for the vectorised consumers)
        sizes = dict(scale=T * B, shift=T * B * 2, vae_latent=T * B * L)
        offs, tot = {}, 0
        for k, nelem in sizes.items():
            offs[k] = tot
            tot += (nelem + 3) & ~3
        w["noise_flat"] = z(tot)
        nv = lambda k, *shape: w["noise_flat"][offs[k]:offs[k] + sizes[k]].view(*shape)
        w["noise"] = dict(scale=nv("scale", T, B, 1), shift=nv("shift", T, B, 2), vae_latent=nv("vae_latent", T, B, L),
                          vae_like=None, concrete_u=z(T, B))
        w["rng_state"] = ops.rng_state(self.seed + 0x9E3779B97F4A7C15 * (1 + dp.rank(self.pg) if self.world > 1 else 1), dev)
        self._like_noise = w["rng_state"]
        if self.train:
            w["dcanvas"] = z(B, cs2)
            # the VAE / ST backward of all T steps runs before the (sequential) LSTM backward: per-step buffers
            w["dwin"] = z(T, B, win)
            w["dtheta"], w["dtheta_inv"], w["dz"] = z(T, B, 6), z(T, B, 6), z(T, B)
            w["dh"], w["dh_next"], w["dc"] = z(B, R), z(B, R), z(B, R)
            # pre-activation gradients are kept for all T steps: the weight-gradient GEMMs run once
            # per train step over the time-batched [T*B, .] buffers (one long-K GEMM per layer)
            w["dhh"] = z(T, B, 5 * HU)
            w["dgates"], w["dgates_sum"] = z(T, B, 4 * R), z(B, 4 * R)
            self._zero_bwd = ops.zero_items([w["dgates_sum"]])
            w["vae_d"] = alloc_vae_scratch(B, win, self.vae_recognition_units, L, self.vae_generative_units, dev,
                                           lead=(T,))
            # per-CTA partial sums of d(heads/out_w|b) for every step; summed once per train step
            self._heads_rows = int(C.lib().air_heads_bwd_workspace(B, HU)) // (7 * HU + 7)
            w["heads_ws"] = torch.zeros(T, self._heads_rows * (7 * HU + 7), device=dev)
            w["adam_ws"] = torch.zeros(int(C.lib().air_adam_workspace(self.store.n)), device=dev)
        if self.cnn:
            F = self.cnn_filters
            u8 = lambda *s: torch.empty(*s, device=dev, dtype=torch.uint8)
            # (input H, W, cin, pool) of the three layers; pooled outputs + one-byte argmax are all the backward needs
            self._cnn_layers = (("conv1", 50, 50, 1, True), ("conv2", 25, 25, F, True), ("conv3", 12, 12, F, False))
            w["cnn_out"] = [z(B, 25 * 25 * F), z(B, 12 * 12 * F), z(B, 12 * 12 * F)]
            w["cnn_arg"] = [u8(B, 25 * 25 * F), u8(B, 12 * 12 * F), None]
            if self.train:
                w["cnn_d"] = [z(B, 25 * 25 * F), z(B, 12 * 12 * F), z(B, 12 * 12 * F)]  # d(loss)/d(layer output)
                w["cnn_ws"] = torch.zeros(ops.conv5x5_bwd_workspace(B, F, F), device=dev)
        p, g = self.store.p, self.store.g
        in_dim = self.rnn_input_dim
        self.Kx, self.Kh = p["rnn/kernel"][:in_dim], p["rnn/kernel"][in_dim:]
        self.gKx, self.gKh = g["rnn/kernel"][:in_dim], g["rnn/kernel"][in_dim:]
        self.vw = VAEWeights(p, g, len(self.vae_recognition_units), len(self.vae_generative_units))
        if self.train:
            # every bias gradient of the step = column sums of a time-batched dY buffer: one launch pair for all
            pairs = vae_bias_items(self.vw, w["vae_d"]) + [(_flat2(w["dhh"]), g["heads/hidden_b"], False),
                                                           (w["dgates_sum"], g["rnn/bias"], False)]
            self._colsum_items = ops.colsum_items(pairs)
            self._colsum_keep = pairs  # the ctypes array holds raw pointers: keep the views alive
            w["colsum_ws"] = torch.zeros(ops.colsum_multi_workspace(self._colsum_items), device=dev)
            # gradient buckets of the data-parallel all-reduce, contiguous in the flat buffer and in production order
            # (params.py): kernel image rows | kernel h rows + cnn | head hidden weights | vae weights | arena
            self._buckets = self.store.bucket_bounds()
            self._pending = []

    # ------------------------------------------------------------------------------------------
    def feed(self, input_images, target_num_digits=None):
        """Refill the input buffers in place (the feed_dict of the reference).  uint8 canvases (CUDA) are expanded on the
        device to k * fl(1/255), the values the reference's MNIST-derived data set holds (air_expand_u8)."""
        if input_images.dtype == torch.uint8:
            ops.expand_u8(input_images.contiguous(), self.input_images)
        else:
            self.input_images.copy_(input_images, non_blocking=True)
        if target_num_digits is not None:
            self.target_num_digits.copy_(target_num_digits, non_blocking=True)

    def set_noise(self, noise):
        """Inject the five noise tensors ([T,B,1], [T,B,2], [T,B,L], [T,B,win], [T,B]) for the NEXT evaluation only:
        the run() / loss_and_grads() / train_step() that follows uses them instead of drawing, every later call
        draws fresh noise again -- like the reference, where each session.run samples anew."""
        if self.w["noise"]["vae_like"] is None:   # (only injected noise needs the likelihood-noise tensor)
            T, B = self.max_steps, self.batch_size
            self.w["noise"]["vae_like"] = torch.empty(T, B, self.windows_size ** 2, device=self.device)
        for k, buf in self.w["noise"].items():
            buf.copy_(noise[k].reshape(buf.shape))
        self._like_noise = _flat2(self.w["noise"]["vae_like"])
        self.noise = "injected"

    def _draw_noise(self):
        """Fresh noise for one evaluation: one launch fills the small tensors from the counter-based generator and
        advances the device step counter; the VAE likelihood noise ([T*B, window^2], 38 MB at B = 4096) is generated
        inside the gen_mean GEMM epilogue from the same counter and never exists in memory."""
        n = self.w["noise"]
        ops.noise_fill(self.w["rng_state"], n["scale"], n["shift"], n["vae_latent"], n["concrete_u"],
                       self.max_steps * self.batch_size, self.vae_latent_dimensions)
        self._like_noise = self.w["rng_state"]

    def _update_scalars(self):
        st = self.store.state
        if self._host_annealed:
            # any other float hyper-parameter (air_model.py:76-82 anneals ANY attribute): these are launch arguments,
            # so the schedule is evaluated on the host from global_step (one device read) before the step's launches
            step = self.store.global_step
            for name in self._host_annealed:
                setattr(self, name, annealed_value(self.annealing_schedules[name], step))
            self.hyper = C.Hyper(self.scale_prior_mean, self.scale_prior_variance, self.shift_prior_mean,
                                 self.shift_prior_variance, self.vae_prior_mean, self.vae_prior_variance,
                                 self.vae_likelihood_std, self.z_pres_temperature, self.stopping_threshold,
                                 1 if self.train else 0)
        if self._annealed("z_pres_prior_log_odds"):
            ops.anneal(st, self.annealing_schedules["z_pres_prior_log_odds"], self._prior)
        if self.train and self._annealed("learning_rate"):
            ops.anneal(st, self.annealing_schedules["learning_rate"], st[4:5])

    # ------------------------------------------------------------------------------------------
    # forward: air_model.py:278-508 (loop body) + :580-611 (loss / accuracy)
    # ------------------------------------------------------------------------------------------
    def _forward(self):
        self._update_scalars()   # annealed hyper-parameters of this step (device scalars / the host struct)
        w, hp, mode = self.w, self.hyper, self.gemm
        B, T = self.batch_size, self.max_steps
        x = self.input_images
        p = self.store.p
        cs, wsz = self.canvas_size, self.windows_size
        n = w["noise"]
        ops.zero_many(self._zero_fwd)  # stopping sums, running loss, digit counts (the canvas starts at zero: first write-back)
        # step-invariant image projection x @ K[:in_dim] (bias added per step, after the h part)
        ops.gemm(self._rnn_input(), self.Kx, w["xk"], mode=mode)
        # ---- (1) the recurrent chain: LSTM -> heads (pose, z_pres, stop) -> attention crop, step by step.  Nothing
        #      here depends on the VAE or the canvas (air_model.py:284-333, 368-427).
        for t in range(T):
            c_prev = w["c"][t - 1] if t > 0 else None
            # LSTM: gates = ([x,h] K) + b, accumulated in concat order (x part first).  The initial state is
            # zero (air_model.py:540-542), so at t = 0 the h rows contribute fma(0, k, acc) == acc exactly:
            # a K = 0 GEMM (accumulator init + bias only) gives the same bits without the MACs.
            if t > 0:
                ops.gemm(w["h"][t - 1], self.Kh, w["gates"][t], Cinit=w["xk"], bias=p["rnn/bias"], mode=mode)
            else:
                ops.gemm(w["h0"][:, :0], self.Kh[:0], w["gates"][t], Cinit=w["xk"], bias=p["rnn/bias"], mode=mode)
            ops.lstm_fwd(w["gates"][t], c_prev, w["c"][t], w["h"][t])
            # five hidden head layers as one GEMM + ReLU; outputs, sampling, KLs, theta, z_pres fused
            ops.gemm(w["h"][t], p["heads/hidden_w"], w["hh"][t], bias=p["heads/hidden_b"], epi=C.EPI_RELU, mode=mode)
            ops.heads_fwd(w["hh"][t], p["heads/out_w"], p["heads/out_b"], n["scale"][t], n["shift"][t],
                          n["concrete_u"][t], self._prior, hp, w["stop"], w["loss"], w["digits"], w["fields"][t],
                          w["theta"][t], w["theta_inv"][t])
        ops.st_forward_steps(x, w["theta"], w["win"], cs, cs, 1, wsz, wsz)  # the T attention crops: one launch
        # ---- (2) the VAE of ALL steps as one evaluation on T*B rows (air_model.py:335-349, 479-496): six GEMMs
        #      with M = T*B instead of 6 T with M = B.  Only the latent step touches per-step state (live mask,
        #      running loss), so it stays per step; the running loss therefore adds the VAE KLs after the pose /
        #      z_pres KLs of all steps -- same terms, different fp32 summation order than the reference's loop.
        def latent_steps():
            for t in range(T):
                ops.vae_latent_fwd(w["ml"][t], n["vae_latent"][t], hp, w["zs"][t], w["fields"][t], w["loss"])
        allbuf = dict(enc=[_flat2(e) for e in w["enc"]], ml=_flat2(w["ml"]), zs=_flat2(w["zs"]),
                      dec=[_flat2(d) for d in w["dec"]], recon=_flat2(w["recon"]))
        vae_forward(_flat2(w["win"]), self.vw, None, self._like_noise, self.vae_likelihood_std, hp, allbuf,
                    w["gen"], None, w["loss"], mode, latent_fn=latent_steps)
        # ---- (3) write-back + canvas accumulation in step order (air_model.py:351-366, 429-439)
        #      -- all T of them in one pass over the canvas (the running canvas stays in registers; starts from zero)
        f0 = w["fields"][0]
        ops.writeback_canvas_fwd_steps(w["recon"], w["theta_inv"], f0[C.F_Z], f0[C.F_STOP_NEW], C.NF * B,
                                       self.stopping_threshold, None, w["canvas"], wsz, wsz, cs, cs)
        dscale = 1.0 / (B * self.world)
        # the clipped reconstruction is only an output: written in inference, derived lazily in training
        ops.bce_loss(w["canvas"], x, None if self.train else w["reconstruction"], w["rec_loss"],
                     w["dcanvas"] if self.train else None, dscale)
        ops.finalize_loss(w["loss"], w["rec_loss"], w["digits"], self.target_num_digits, w["out2"], w["loss_item"])

    def _rnn_input(self):
        """air_model.py:510-535: the canvas, or conv-relu-pool, conv-relu-pool, conv-relu of it (NHWC flatten)."""
        if not self.cnn:
            return self.input_images
        w, p, F = self.w, self.store.p, self.cnn_filters
        src = self.input_images
        for i, (name, H, W, cin, pool) in enumerate(self._cnn_layers):
            ops.conv5x5_fwd(src, p[f"cnn/{name}/kernel"], p[f"cnn/{name}/bias"], w["cnn_out"][i], w["cnn_arg"][i], H, W, cin,
                            F, pool)
            src = w["cnn_out"][i]
        return src

    def _cnn_backward(self):
        """d(loss)/d(rnn_input) = dgates_sum K_x^T, then back through the three conv layers (no d(images))."""
        w, p, g, F = self.w, self.store.p, self.store.g, self.cnn_filters
        ops.gemm(w["dgates_sum"], self.Kx, w["cnn_d"][2], tB=True, mode=self.gemm)
        for i in (2, 1, 0):
            name, H, W, cin, pool = self._cnn_layers[i]
            src = w["cnn_out"][i - 1] if i > 0 else self.input_images
            ops.conv5x5_bwd(src, p[f"cnn/{name}/kernel"], w["cnn_out"][i], w["cnn_arg"][i], w["cnn_d"][i],
                            w["cnn_d"][i - 1] if i > 0 else None, g[f"cnn/{name}/kernel"], g[f"cnn/{name}/bias"], False,
                            w["cnn_ws"], H, W, cin, F, pool)

    def _vae_buf(self, t):
        w = self.w
        return dict(enc=[e[t] for e in w["enc"]], ml=w["ml"][t], zs=w["zs"][t], dec=[d[t] for d in w["dec"]],
                    recon=w["recon"][t])

    # ------------------------------------------------------------------------------------------
    # backward: what TF autodiff derives for air_model.py:655 (gradients of the mean loss)
    # ------------------------------------------------------------------------------------------
    def _backward(self):
        self._backward_loop()
        self._weight_grads()

    def _backward_loop(self):
        w, hp, mode = self.w, self.hyper, self.gemm
        B, T = self.batch_size, self.max_steps
        x = self.input_images
        p, g = self.store.p, self.store.g
        cs, wsz = self.canvas_size, self.windows_size
        n = w["noise"]
        dscale = 1.0 / (B * self.world)
        ops.zero_many(self._zero_bwd)  # the summed gate gradients the LSTM backward accumulates into
        vd = w["vae_d"]
        # ---- (1) write-back backward of every step (the canvas is a plain sum: all steps see the same dcanvas)
        #      -- one launch for all T steps
        f0 = w["fields"][0]
        ops.writeback_canvas_bwd_steps(w["recon"], w["theta_inv"], f0[C.F_Z], f0[C.F_STOP_NEW], C.NF * B,
                                       self.stopping_threshold, w["dcanvas"], vd["dgen"], w["dtheta_inv"], w["dz"],
                                       wsz, wsz, cs, cs, window_is_sigmoid=True,  # SigmoidGrad fused into the store
                                       axis_aligned_theta=True,  # heads_bwd reads dtheta_inv[0,2,4,5] only
                                       reference_rounding=self.reference_rounding)
        # ---- (2) VAE backward of all steps as one evaluation on T*B rows, then the crop backward per step
        def latent_bwd_steps():
            for t in range(T):
                ops.vae_latent_bwd(w["ml"][t], n["vae_latent"][t], vd["dzs"][t], w["fields"][t], hp, dscale, vd["dml"][t])
        allbuf = dict(enc=[_flat2(e) for e in w["enc"]], ml=_flat2(w["ml"]), zs=_flat2(w["zs"]),
                      dec=[_flat2(d) for d in w["dec"]], recon=_flat2(w["recon"]))
        alld = dict(denc=[_flat2(d) for d in vd["denc"]], dml=_flat2(vd["dml"]), dzs=_flat2(vd["dzs"]),
                    ddec=[_flat2(d) for d in vd["ddec"]], dgen=_flat2(vd["dgen"]))
        vae_backward_dx(_flat2(w["win"]), self.vw, None, hp, allbuf, alld, dscale, None, mode, dx_out=_flat2(w["dwin"]),
                        dgen_is_presigmoid=True, latent_bwd_fn=latent_bwd_steps)
        ops.st_backward_steps(x, w["theta"], w["dwin"], w["dtheta"], cs, cs, 1, wsz, wsz)
        # ---- (3) the recurrent chain backwards: heads -> LSTM, step by step
        for t in range(T - 1, -1, -1):
            last = t == T - 1
            f = w["fields"][t]
            ops.heads_bwd(w["hh"][t], p["heads/out_w"], n["scale"][t], n["shift"][t], f, w["dtheta"][t], w["dtheta_inv"][t],
                          w["dz"][t], self._prior, hp, dscale, w["dhh"][t], None, None, False, w["heads_ws"][t])
            # dh_t = dhh_t W_hid^T (+ the LSTM path from step t+1)
            ops.gemm(w["dhh"][t], p["heads/hidden_w"], w["dh"], Cinit=None if last else w["dh_next"], tB=True, mode=mode)
            ops.lstm_bwd(w["gates"][t], w["c"][t - 1] if t > 0 else None, w["c"][t], w["dh"],
                         None if last else w["dc"], w["dgates"][t], w["dc"], w["dgates_sum"])
            if t > 0:
                ops.gemm(w["dgates"][t], self.Kh, w["dh_next"], tB=True, mode=mode)
            if getattr(self, "_debug", None) is not None:  # per-step intermediate gradients for diagnostics
                self._debug[t] = {k: w[k][t].clone() for k in ("dwin", "dtheta", "dtheta_inv", "dz")}
                self._debug[t]["dh"] = w["dh"].clone()

    # ---- weight gradients, once per train step, over the time-batched buffers, in the order of the flat buffer; each
    #      finished bucket is handed to the all-reduce at once (data parallel) and travels while the next GEMMs run.
    def _weight_grads(self):
        w, g, mode, T = self.w, self.store.g, self.gemm, self.max_steps
        # the image rows of the LSTM kernel see the same input every step: one GEMM on the summed dgates (64 % of all
        # gradient bytes: first out)
        rnn_in = w["cnn_out"][2] if self.cnn else self.input_images
        ops.gemm(rnn_in, w["dgates_sum"], self.gKx, tA=True, mode=mode)
        self._reduce_bucket(0)
        if T > 1:   # K_h sees h_{t-1}: rows t = 1..T-1 (h_{-1} = 0 contributes nothing)
            ops.gemm(_flat2(w["h"][:T - 1]), _flat2(w["dgates"][1:]), self.gKh, tA=True, mode=mode)
        else:
            self.gKh.zero_()
        if self.cnn:
            self._cnn_backward()
        self._reduce_bucket(1)
        dense_dw(_flat2(w["h"]), _flat2(w["dhh"]), g["heads/hidden_w"], g["heads/hidden_b"], None, mode)
        self._reduce_bucket(2)
        allbuf = dict(enc=w["enc"], ml=w["ml"], zs=w["zs"], dec=w["dec"], recon=w["recon"])
        vae_weight_grads(_flat2(w["win"]), self.vw, allbuf, w["vae_d"], None, mode)
        self._reduce_bucket(3)
        # the arena: head outputs and every bias gradient (one column-sum launch pair), 17 KB -- the only message
        # that is not hidden behind a GEMM
        nw = 7 * self.store.dims["HU"]
        ops.reduce_rows(w["heads_ws"], T * self._heads_rows, nw + 7, nw, g["heads/out_w"])
        ops.reduce_rows(w["heads_ws"].view(-1)[nw:], T * self._heads_rows, nw + 7, 7, g["heads/out_b"])
        ops.colsum_multi(self._colsum_items, w["colsum_ws"])
        self._reduce_bucket(4)
        self._reduce_wait()

    def _apply_gradients(self):
        """air_model.py:673, 692: clip by global norm, Adam, global_step += 1."""
        st = self.store
        ops.adam_step(st.flat, st.grad, st.adam_m, st.adam_v, st.state, self.gradient_clipping_norm, 0.9, 0.999, 1e-8,
                      1.0, self.w["adam_ws"], skip_nonfinite=self.skip_nonfinite_updates)

    def _reduce_bucket(self, i):
        """SUM all-reduce of gradient bucket i on the communicator's stream, ordered after the kernels queued so far
        (the per-item loss weight already carries 1/world).  Capturable: inside capture() the collective becomes a
        node of the step's CUDA graph on a forked branch."""
        if self.world > 1:
            lo, hi = self._buckets[i], self._buckets[i + 1]
            if hi > lo:
                self._pending.append(dp.allreduce_async(self.store.grad[lo:hi], self.pg))

    def _reduce_wait(self):
        """The current stream (not the host) waits for every pending bucket."""
        for work in self._pending:
            work.wait()
        self._pending = []

    # ------------------------------------------------------------------------------------------
    # public API
    # ------------------------------------------------------------------------------------------
    def run(self, noise=None):
        """Forward pass only (what session.run of any result attribute evaluates)."""
        if noise is not None:
            self.set_noise(noise)
        if self.noise == "injected":
            self.noise = None        # one-shot: consumed by this evaluation (the buffers keep it for the backward)
        elif self._graphs is not None and not self.train:
            self._graphs.replay()    # captured inference step: fresh noise + forward in one graph replay
            return self
        else:
            self._draw_noise()
        self._forward()
        self._publish()
        return self

    infer = run

    def loss_and_grads(self, noise=None):
        """Forward + backward without the optimizer (gradients in store.named_grads())."""
        if not self.train:
            raise C.AirError("model was built with train=False")
        self.run(noise)
        self._backward()
        return self.loss, self.store.named_grads()

    def train_step(self, noise=None):
        """One optimisation step == sess.run(model.training) (training.py:212-224)."""
        if not self.train:
            raise C.AirError("model was built with train=False")
        if self._graphs is not None and noise is None:
            self._graphs.replay()
            return self
        self.loss_and_grads(noise)
        self._apply_gradients()
        return self

    def capture(self, warmup=3):
        """Capture one whole optimisation step -- noise, forward, backward, the bucketed NCCL all-reduces (data
        parallel: forked branches of the graph that overlap the remaining weight-gradient GEMMs), clip + Adam -- as
        ONE CUDA graph; subsequent train_step() calls are a single replay with no host work in between.  A model built
        with train=False captures its inference step (noise + forward); run() then replays it."""
        if self.noise == "injected":
            raise C.AirError("capture() would freeze the injected noise into the graph: run the pending evaluation "
                             "first (injection is one-shot) or call capture() before set_noise()")
        if self._host_annealed:
            raise C.AirError(f"capture() would freeze the host-annealed hyper-parameters {self._host_annealed} into the "
                             "graph (they are kernel launch arguments): train eagerly, or anneal only "
                             f"{_DEVICE_ANNEALABLE}, which are evaluated on the device")
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):   # (also creates the NCCL communicator before anything is captured)
                self._draw_noise(); self._forward()
                if self.train:
                    self._backward()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        # thread_local: other host threads (NCCL's watchdog, a data-feeding thread, a clock sampler) may touch the CUDA API
        # while this thread captures; only this thread's calls belong to the graph.
        # The garbage collector stays off while capturing: a collection that happens to run inside the capture can free
        # pinned host buffers or multi-stream tensors of models dropped earlier (AIRModel holds reference cycles), and the
        # allocators then record / query events on other streams -- calls that invalidate the capture ("operation not
        # permitted when stream is capturing"; seen in bench.py after the train / inference measurements).
        import gc
        gc.collect()
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._draw_noise(); self._forward()
                if self.train:
                    self._backward(); self._apply_gradients()
        finally:
            if gc_was_on:
                gc.enable()
        self._graphs = g
        self._publish()
        return self

    def _publish(self):
        """Result attributes of the reference (air_model.py:569-611), as views (no copies)."""
        w, L = self.w, self.vae_latent_dimensions
        F = w["fields"]  # [T, NF, B]
        col = lambda i: F[:, i, :].t()  # [B, T]
        self.rec_num_digits = w["digits"]
        self.rec_scales = col(C.F_S).unsqueeze(2)
        self.rec_st_back = w["theta_inv"].permute(1, 0, 2).reshape(self.batch_size, self.max_steps, 2, 3)
        self.rec_windows = w["recon"].permute(1, 0, 2)
        self.rec_latents = w["ml"][:, :, :L].permute(1, 0, 2)
        self.z_pres_probs, self.z_pres_kls = col(C.F_ZPROB), col(C.F_KL_Z)
        self.scale_kls, self.shift_kls, self.vae_kls = col(C.F_KL_SCALE), col(C.F_KL_SHIFT), col(C.F_KL_VAE)
        self.z_pres = col(C.F_Z)
        self.reconstruction_loss = w["rec_loss"]
        self.loss_per_item = w["loss_item"]
        self.loss, self.accuracy = w["out2"][0], w["out2"][1]

    def save(self, prefix, with_optimizer=True):
        """Write a TensorFlow-bundle checkpoint (``prefix.index`` + ``prefix.data-00000-of-00001``) with the
        reference's variable names -- what tf.train.Saver.save does in training.py:203-207."""
        from .. import checkpoint
        return checkpoint.save_model(self.store, prefix, scope=self.scope, with_optimizer=with_optimizer)

    def restore(self, prefix, verify_crc=True):
        """Load a checkpoint written by the reference or by save() (demo.py:33, embeddings.py:168)."""
        from .. import checkpoint
        return checkpoint.restore_model(self.store, prefix, scope=self.scope, verify_crc=verify_crc)

    @property
    def reconstruction(self):
        """clip(canvas, 0, 1) (air_model.py:582).  Inference writes it in the loss kernel; a training model
        derives it on access so the train step does not spend 41 MB of HBM writes on an unused output."""
        if self.train:
            return torch.clamp(self.w["canvas"], 0.0, 1.0)
        return self.w["reconstruction"]

    @property
    def rec_shifts(self):
        """[B, T, 2] (air_model.py:570); a fresh stack of the x / y field rows."""
        F = self.w["fields"]
        return torch.stack([F[:, C.F_X, :].t(), F[:, C.F_Y, :].t()], dim=2)

    @property
    def stop_masks(self):
        """[B, T] bool: stopping_sum < threshold after each step (the mask of air_model.py:427-496)."""
        return self.w["fields"][:, C.F_STOP_NEW, :].t() < self.stopping_threshold

    @property
    def executed_steps(self):
        """Trip count of the reference's while_loop (air_model.py:271-275); needs a host sync."""
        live_before = (self.w["fields"][:, C.F_STOP_PREV, :] < self.stopping_threshold).any(dim=1).cpu().tolist()
        n = 0
        for t in range(self.max_steps):
            if not live_before[t]:
                break
            n = t + 1
        return n
