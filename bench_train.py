"""Train / inference workloads of bench.py (BASELINE.json configs[2], [3], [4])."""
from __future__ import annotations

import os
import time

import torch

from bench import ClockSampler, barrier, max_over_ranks, time_launches

METRIC = "AIR train images/sec"


def _model(B, gemm_mode, seed, train=True, max_steps=3, cnn=False):
    import air_b200 as ab
    from importlib import import_module
    data = import_module("tf-attend-infer-repeat_b200.data")
    imgs, cnt = data.synthetic_canvases(B, seed=seed)
    # like the reference's MNIST-derived data set, the canvases hold the 256 values k * fl(1/255) (see air_expand_u8):
    # the same batch exists as uint8 (what the end-to-end path ships over PCIe) and as fp32 (device-resident runs)
    imgs = torch.round(imgs * 255.0).to(torch.uint8).float() * torch.tensor(1.0 / 255.0, dtype=torch.float32)
    hyper = dict(data.TRAINING_HYPER)
    hyper["max_steps"] = max_steps
    hyper["cnn"] = bool(cnn)   # air_model.py:510-535 front-end (the reference's own scripts all run cnn=False)
    ab.reset_variable_scopes()
    m = ab.AIRModel(imgs.cuda(), cnt.cuda(), train=train, annealing_schedules=data.TRAINING_ANNEALING,
                    gemm_mode=gemm_mode, seed=0, **hyper)
    return m, imgs, cnt, data


def _covered(m, imgs):
    """The well-conditioned parity fixture of tests/parity_util.covered_fixture, applied to a model in place: every canvas
    pixel samples inside the window (scale ~1, shift ~0, all steps live), reconstructions well inside (0, 1), digits away
    from the border -- so the canvas-residue chaos of DESIGN.md section 2 (caveat 1) does not mask GEMM-level errors."""
    v = m.store.named_views()
    v["scale/mean/output/biases"] += 12.0
    v["scale/mean/output/weights"] *= 0.2
    v["scale/log_variance/output/biases"] -= 8.0
    v["shift/mean/output/weights"] *= 0.0
    v["shift/mean/output/biases"] -= 3e-5
    v["shift/log_variance/output/weights"] *= 0.1
    v["shift/log_variance/output/biases"] -= 30.0
    v["z_pres/log_odds/output/biases"] += 5.0
    v["vae/gen_mean/weights"] *= 0.5
    v["vae/gen_mean/biases"] -= 3.5
    im = imgs.reshape(-1, 50, 50).clone()
    im[:, :7] = 0; im[:, -7:] = 0; im[:, :, :7] = 0; im[:, :, -7:] = 0
    return im.reshape(imgs.shape[0], -1).contiguous()


def parity_block(mode, B=4096, T=3, seed=7):
    """Measured errors of GEMM mode ``mode`` on one train step (forward + backward) at the benchmark's batch size:
    the same weights, canvases and injected noise through a model in ``mode`` and through one in the exact-FP32 mode
    (SIMT FFMA chains -- the mode tests/ pin to the reference graph and the oracle).  No oracle involved here."""
    import air_b200 as ab
    g = torch.Generator(device="cuda").manual_seed(seed)
    L, win = 50, 784
    noise = dict(scale=torch.randn(T, B, 1, device="cuda", generator=g), shift=torch.randn(T, B, 2, device="cuda", generator=g),
                 vae_latent=torch.randn(T, B, L, device="cuda", generator=g), vae_like=torch.randn(T, B, win, device="cuda", generator=g),
                 concrete_u=torch.rand(T, B, device="cuda", generator=g))
    res = {}
    for fixture in ("covered", "default_init"):
        out = {}
        for md in ("fp32", mode):
            m, imgs, cnt, _ = _model(B, md, seed=seed, train=True, max_steps=T)
            m.store.global_step = 2000
            if fixture == "covered":
                m.feed(_covered(m, imgs).cuda())
            m.loss_and_grads(noise)
            torch.cuda.synchronize()
            out[md] = dict(loss=m.loss.double().item(), digits=m.rec_num_digits.clone(), masks=m.stop_masks.clone(),
                           grads={k: v.double().clone() for k, v in m.store.named_grads().items()},
                           windows=m.rec_windows.double().clone(), kls=m.vae_kls.double().clone())
            del m
        a, b = out[mode], out["fp32"]
        rel = lambda x, y: float((x - y).norm() / y.norm().clamp_min(1e-30))
        gerr = {k: rel(a["grads"][k], b["grads"][k]) for k in b["grads"] if float(b["grads"][k].norm()) > 0}
        tot = rel(torch.cat([v.reshape(-1) for v in a["grads"].values()]), torch.cat([v.reshape(-1) for v in b["grads"].values()]))
        res[fixture] = {"digit_count_mismatches": int((a["digits"] != b["digits"]).sum()),
                        "stop_mask_mismatches": int((a["masks"] != b["masks"]).sum()),
                        "loss_rel_err": abs(a["loss"] - b["loss"]) / abs(b["loss"]),
                        "rec_windows_rel_err": rel(a["windows"], b["windows"]), "vae_kl_rel_err": rel(a["kls"], b["kls"]),
                        "grad_rel_err_worst_tensor": max(gerr.values()), "grad_rel_err_all_params": tot}
        torch.cuda.empty_cache()
    res["batch"] = B
    res["against"] = ("this library's exact-FP32 mode (k-sequential FFMA GEMMs), same weights / canvases / injected noise; that "
                      "mode is what tests/ compare with the reference graph and the oracle")
    res["note"] = ("'covered': every canvas pixel samples inside the window (well conditioned); 'default_init': uncovered lit "
                   "pixels make the canvas-derived loss / gradients rounding-chaotic in ANY fp32 implementation "
                   "(DESIGN.md section 2), digit counts and masks are the meaningful figures there")
    return res


def dp_check(m, rank, world, mode, T=3):
    """Evidence that the data-parallel step is the single-process step: (a) after the timed steps every rank holds
    bit-identical parameters (checksums over the raw bits, all-gathered); (b) on a global batch of 64 sharded over the
    ranks, the all-reduced gradient equals the gradient a single-GPU model computes on the whole batch (same weights,
    canvases, injected noise), up to fp32 summation order."""
    import torch.distributed as dist
    import air_b200 as ab
    bits = m.store.flat.view(torch.int32).to(torch.int64)
    cs = torch.stack([bits.sum(), (bits * torch.arange(1, bits.numel() + 1, device=bits.device)).sum(),
                      torch.tensor(int(m.global_step), device=bits.device)])
    allcs = [torch.empty_like(cs) for _ in range(world)]
    dist.all_gather(allcs, cs)
    identical = all(torch.equal(allcs[0], c) for c in allcs)
    Bg = 64
    g = torch.Generator(device="cuda").manual_seed(123)
    L, win = 50, 784
    noise = dict(scale=torch.randn(T, Bg, 1, device="cuda", generator=g), shift=torch.randn(T, Bg, 2, device="cuda", generator=g),
                 vae_latent=torch.randn(T, Bg, L, device="cuda", generator=g), vae_like=torch.randn(T, Bg, win, device="cuda", generator=g),
                 concrete_u=torch.rand(T, Bg, device="cuda", generator=g))
    from importlib import import_module
    data = import_module("tf-attend-infer-repeat_b200.data")
    imgs, cnt = data.synthetic_canvases(Bg, seed=99)
    hyper = dict(data.TRAINING_HYPER)

    def build(images, counts, scope, pg):
        mm = ab.AIRModel(images.cuda(), counts.cuda(), train=True, annealing_schedules=data.TRAINING_ANNEALING, gemm_mode=mode,
                         seed=0, scope=scope, process_group=pg, **hyper)
        mm.feed(_covered(mm, images).cuda())
        mm.store.global_step = 2000
        return mm
    per = Bg // world
    sl = slice(rank * per, (rank + 1) * per)
    m_dp = build(imgs[sl], cnt[sl], "dpcheck_shard", None)
    m_dp.loss_and_grads({k: v[:, sl].contiguous() for k, v in noise.items()})
    m_ref = build(imgs, cnt, "dpcheck_full", "local")
    m_ref.loss_and_grads(noise)
    err = float((m_dp.store.grad.double() - m_ref.store.grad.double()).norm() / m_ref.store.grad.double().norm())
    worst = torch.tensor([err], device="cuda", dtype=torch.float64)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    return {"replicas_bit_identical_after_timed_steps": bool(identical), "global_step": int(m.global_step),
            "reduced_grad_vs_single_gpu_rel_err": float(worst.item()), "fixture": f"global batch {Bg} = {per} rows per rank, covered poses, injected noise",
            "gemm_mode": mode}


def gemm_roofline(B, mode, peaks, steps=20):
    """The dominant kernel of the step: the [B,2500]x[2500,1024] LSTM input projection GEMM
    (and its twin dK = x^T dgates), timed alone with CUDA events."""
    import air_b200 as ab
    from air_b200 import ops
    x = torch.rand(B, 2500, device="cuda")
    W = torch.randn(2500, 1024, device="cuda") * 0.02
    out = torch.empty(B, 1024, device="cuda")
    ms = time_launches(lambda: ops.gemm(x, W, out, mode=ab._cabi.GEMM_MODES[mode]), steps, 5)
    flops = 2.0 * B * 2500 * 1024
    tf = flops / (ms * 1e-3) / 1e12
    peak = peaks["bf16_tflops"] * (0.5 if mode == "tf32" else 1.0)
    kern = "gemm_simt<128,128,16,8,8> (fp32 FFMA, exact mode)" if mode == "fp32" else "gemm_tf32 (tcgen05)"
    return {"bound": "tensor", "achieved": round(tf, 2), "peak": round(peak, 1), "unit": "TFLOP/s",
            "frac": round(tf / peak, 4),
            # traffic: NOT measured in this run -- dram__bytes_read + write of one launch at B = 4096 from the ncu --set full
            # captures under profiles/ (C stays in L2 within one launch), scaled with B; null for other batches' kernels
            "traffic": (round((51.28e6 + 1.94e6) * B / 4096) if mode == "tf32" else round((51.41e6 + 3.21e6) * B / 4096)
                        if mode == "tf32x3" else None),
            "traffic_source": "profiles/r2_gemm_pair_ncu_full.md (ncu capture of round 2, not this run)" if mode != "fp32" else None,
            "algorithmic_bytes": int(4 * (B * 2500 + 2500 * 1024 + B * 1024)),
            "tensor_pipe_active_ncu_r2_capture": 0.503 if mode == "tf32" else 0.442 if mode == "tf32x3" else None,
            "peak_source": peaks["source"], "kernel": kern,
            "shape": f"[{B},2500]x[2500,1024]", "ms": round(ms, 4),
            "note": "peak = measured cuBLAS bf16 burst" + (" / 2 (TF32 dense estimate; ncu of the CTA-pair kernel: tensor pipe active "
                                                           "50 % of elapsed cycles at 512 TFLOP/s under the profiler, "
                                                           "profiles/r2_gemm_pair_ncu_full.md)" if mode == "tf32" else
                                                           "; exact-fp32 SIMT mode cannot approach a tensor-core peak")}


def run(args, rank, world, peaks):
    import air_b200 as ab
    B = args.batch or 4096
    if getattr(args, "global_batch", None):     # configs[3] as written: a fixed global batch sharded over the ranks
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
        B = args.global_batch // world
    mode = args.gemm or os.environ.get("AIR_GEMM", "tf32")
    infer = args.workload == "infer"
    T = 5 if infer else 3
    if infer and not args.batch:
        B = 65536
    m, imgs, cnt, data = _model(B, mode, seed=rank, train=not infer, max_steps=T, cnn=getattr(args, "cnn", False))
    m.capture()   # one CUDA graph per step: noise + forward (+ backward + bucketed all-reduce + clip + Adam)
    step = m.run if infer else m.train_step
    n0 = ab.launch_count()
    for _ in range(args.warmup):
        step()
    barrier(world)
    launches_eager = ab.launch_count() - n0
    # (the sampler's keep-busy load must not contain a collective: ranks call it a different number of times)
    busy = step if world == 1 else (lambda: m.run())
    with ClockSampler(torch.cuda.current_device(), busy=busy) as cs:
        barrier(world)  # nvidia-smi start-up differs per rank: re-align before the timed region
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier(world)
        ms = max_over_ranks(e0.elapsed_time(e1), world) / args.steps
    value = B * world / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers: every step copies that step's batch from
    # pinned host memory to the device and reads that step's loss back to the host.  Pipelined like a real
    # input pipeline: batch k+1 is staged on a copy stream while step k computes, and the loss of step k is
    # consumed one step later (async D2H into pinned memory), so PCIe overlaps the kernels.
    n_host = 4
    imgs_u8 = torch.round(imgs * 255.0).to(torch.uint8)     # exact: the canvases are k * fl(1/255)
    assert torch.equal(imgs_u8.float() * torch.tensor(1.0 / 255.0, dtype=torch.float32), imgs)
    host_batches = [(imgs_u8.clone().pin_memory(), cnt.clone().pin_memory()) for _ in range(n_host)]
    stage = [(torch.empty_like(imgs_u8, device="cuda"), torch.empty_like(cnt, device="cuda")) for _ in range(2)]
    loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    staged_ev = [torch.cuda.Event() for _ in range(2)]
    consumed_ev = [torch.cuda.Event() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    main = torch.cuda.current_stream()
    result_of = (lambda: m.loss) if not infer else (lambda: m.rec_num_digits[:1].float())

    def prefetch(k):
        slot = k % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed_ev[slot])          # the step that last used this slot has read it
            stage[slot][0].copy_(host_batches[k % n_host][0], non_blocking=True)
            stage[slot][1].copy_(host_batches[k % n_host][1], non_blocking=True)
            staged_ev[slot].record(copy_stream)

    def run_e2e(n):
        last = None
        for ev in consumed_ev:
            ev.record(main)
        prefetch(0)
        for k in range(n):
            slot = k % 2
            if k + 1 < n:
                prefetch(k + 1)
            main.wait_event(staged_ev[slot])
            m.feed(stage[slot][0], stage[slot][1])             # uint8 -> fp32 expansion into the model's input buffers
            consumed_ev[slot].record(main)
            step()
            loss_host[slot].copy_(result_of().reshape(1), non_blocking=True)
            loss_ev[slot].record(main)
            if k > 0:                                           # consume the previous step's result on the host
                loss_ev[1 - slot].synchronize()
                last = float(loss_host[1 - slot].item())
        loss_ev[(n - 1) % 2].synchronize()
        return float(loss_host[(n - 1) % 2].item())

    run_e2e(4)
    barrier(world)
    n_e2e = max(10, args.steps)
    t0 = time.perf_counter()
    last = run_e2e(n_e2e)
    torch.cuda.synchronize()
    barrier(world)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, world) / n_e2e
    imgs_h, cnt_h = host_batches[0]

    # kernels per step: count one eager step (graph replays do not pass through the launch counter)
    probe, _, _, _ = (m, None, None, None)
    c0 = ab.launch_count()
    g, m._graphs = m._graphs, None
    step()
    m._graphs = g
    per_step = ab.launch_count() - c0
    torch.cuda.synchronize()

    flops_img = data.train_flops_per_image(T) if not infer else 2 * T * data.MAC_FWD_STEP
    line = {
        "metric": METRIC if not infer else "AIR inference images/sec", "value": round(value, 1), "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "tf32": "tf32", "tf32x3": "tf32x3 (fp32-grade)"}[mode], "data": "synthetic",
        "config": {"workload": ("AIRModel default (training.py:100-122) full train step: forward, backward, "
                                "global-norm clip, TF-Adam; T=3" if not infer else
                                "AIRModel default inference (train=False), T=5"),
                   "cnn_frontend": bool(getattr(args, "cnn", False)),
                   "batch_per_gpu": B, "global_batch": B * world, "gemm_mode": mode, "parallelism": f"dp{world}",
                   "noise": "drawn on device inside the timed step", "cuda_graph": True,
                   "l2": f"per-step working set ~{B * 61e3 / 1e6:.0f} MB of activations + 16 MB weights > 126 MB L2"},
        "step_tflops_canonical": round(value * flops_img / 1e12, 2),
        "e2e": {"value": round(B * world / (e2e_ms * 1e-3), 1), "unit": "images/s",
                "h2d_bytes_per_step": int(imgs_h.numel() * imgs_h.element_size() + cnt_h.numel() * 4), "d2h_bytes_per_step": 4,
                "ms_per_step": round(e2e_ms, 4), "last_result": last,
                "input_format": "uint8 canvases: the data set's values are k * fl(1/255) like the reference's MNIST-derived "
                                "canvases; AIRModel.feed expands them bit for bit on the device (air_expand_u8)",
                "pipeline": "H2D of batch k+1 on a copy stream overlaps step k; loss of step k read back asynchronously"},
        "gpu_launches": int(per_step * args.steps),
        "kernels_per_step": int(per_step),
        "loss": float(m.loss.item()) if not infer else None,
        "clocks": cs.summary(),
    }
    if rank == 0:
        line["roofline"] = gemm_roofline(B, mode, peaks)
        if world == 1:
            line["cpu_baseline"] = cpu_baseline(seconds=12.0)
    if world > 1 and not infer:
        line["dp_check"] = dp_check(m, rank, world, mode, T)
        line["allreduce"] = {"buckets_floats": [b - a for a, b in zip(m._buckets[:-1], m._buckets[1:])],
                             "note": "NCCL all-reduces captured inside the step's CUDA graph on forked branches, one per gradient "
                                     "bucket in production order; only the last (the 17 KB arena of biases) is not behind a GEMM"}
        if world in (2, 4) and not getattr(args, "global_batch", None) and not args.batch:
            # configs[3] as written: global batch 32768 sharded over this many GPUs (at 8 GPUs that is the line itself)
            del m
            torch.cuda.empty_cache()
            Bc = 32768 // world
            m4, _, _, _ = _model(Bc, mode, seed=rank, train=True, max_steps=T)
            m4.capture()
            for _ in range(3):
                m4.train_step()
            barrier(world)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                m4.train_step()
            e1.record()
            barrier(world)
            ms4 = max_over_ranks(e0.elapsed_time(e1), world) / 10
            line["c4_global_batch_32768"] = {"batch_per_gpu": Bc, "ms_per_step": round(ms4, 4), "value": round(32768 / (ms4 * 1e-3), 1),
                                             "unit": "images/s", "steps": 10}
            del m4
            m = None
        if m is not None:
            m._graphs = None   # (graphs with captured NCCL kernels must be gone before the process group is destroyed)
    if world == 1 and not infer:
        # parity figures of the mode that was timed, and the FP32-grade modes beside it on the same workload
        del m
        torch.cuda.empty_cache()
        if mode != "fp32":
            line["parity"] = parity_block(mode, B, T)
        for other, key, note in (("tf32x3", "exact_fp32_mode", "3xTF32 on tcgen05 (hi/lo split in-kernel, FP32 accumulate): the tensor-core "
                                  "mode that passes the whole parity suite at the exact-mode bars (tests/test_gpu_model.py, "
                                  "test_gpu_zz_reference_graph.py: bit-exact counts / masks, <= 1e-5, <= 1e-4)"),
                                 ("fp32", "simt_fp32_mode", "k-sequential FFMA GEMMs on CUDA cores (bit-reproducible vs oracle_gemm_seq_fma)")):
            if other == mode:
                continue
            m2, _, _, _ = _model(B, other, seed=rank, train=True, max_steps=T)
            m2.capture()
            ms2 = time_launches(m2.train_step, 20 if other != "fp32" else 5, 3)
            line[key] = {"value": round(B / (ms2 * 1e-3), 1), "unit": "images/s", "ms_per_step": round(ms2, 3), "gemm_mode": other,
                         "note": note}
            del m2
            torch.cuda.empty_cache()
            if other == "tf32x3":
                line[key]["parity"] = parity_block(other, B, T)
    return line


REF_SAMPLE_BATCH = 1024   # bounded sample of the batch-4096 step; the CPU port saturates its threads from ~1024 rows


def cpu_baseline(seconds=12.0, B=REF_SAMPLE_BATCH, steps=None, warmup=2, budget_s=200.0):
    """The reference restated on the CPU (oracle/air_oracle.py, torch, NOT TensorFlow): full train step (forward,
    backward, clip, Adam) on a bounded sample of the batch-4096 workload, all host cores.  ``steps=None`` runs for
    ``seconds``; otherwise exactly ``steps`` timed steps, cut down only if they would exceed ``budget_s``."""
    from oracle import air_oracle as O
    torch.set_num_threads(os.cpu_count())
    imgs, cnt = O.synthetic_canvases(B, seed=0)
    m = O.AIROracle(annealing_schedules=O.DEFAULT_ANNEALING, train=True, seed=0)
    noise = O.make_noise(0, 3, B)
    t0 = time.perf_counter()
    for _ in range(max(warmup, 1)):
        m.train_step(imgs, cnt, noise)
    per = (time.perf_counter() - t0) / max(warmup, 1)
    if steps is not None:
        steps = max(1, min(steps, int(budget_s / per)))
    t0 = time.perf_counter()
    n = 0
    while (steps is None and time.perf_counter() - t0 < seconds) or (steps is not None and n < steps):
        m.train_step(imgs, cnt, noise)
        n += 1
    dt = (time.perf_counter() - t0) / n
    return {"value": round(B / dt, 1), "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"oracle (torch CPU restatement of the TF-1.3 reference) full train step, batch {B} "
                      f"(a quarter of the batch-4096 step), {n} steps, {dt * 1e3:.1f} ms/step",
            "steps": n, "ms_per_step": round(dt * 1e3, 2)}


def run_reference(args, peaks):
    cb = cpu_baseline(steps=args.steps, warmup=args.warmup)
    return {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": cb["steps"], "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "AIRModel default full train step (reference restated on CPU; TF 1.3 cannot run here)",
                       "batch_per_gpu": REF_SAMPLE_BATCH,
                       "note": f"each step is a bounded sample (batch {REF_SAMPLE_BATCH}) of the batch-4096 workload; "
                               "throughput in images/s is what is compared"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
