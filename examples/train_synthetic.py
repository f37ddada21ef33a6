"""Train AIR on synthetic canvases through the product API and watch the digit-count accuracy (the quantity behind the
reference README's "98 % after ~25k iterations" claim, there on multi-MNIST) -- the GPU counterpart of
oracle/train_convergence.py, with the reference's training configuration (training.py:100-122, batch 64).

    python examples/train_synthetic.py --iters 25000 --gemm tf32x3 --seed 3 --log run.log     # 25k iterations: ~18 s on a B200

The model is built with reference_rounding=True (default): the write-back backward sums the un-cancelled fp32 corner
terms the reference's training depends on (--clean-gradient shows what happens without them: loss ~1900 for ever).
Whether a run settles on the right digit COUNT is a seed lottery in the reference's arithmetic (5 of 12 seeds reach
>= 95 %, the CPU oracle 2 of 5); logs of round 2's runs: profiles/r2_reference_rounding/gpu_training_*.log, next to the CPU
oracle's profiles/r1_oracle_convergence.log; the story: profiles/r2_reference_rounding.md.
"""
import argparse
import os
import sys
import time
from importlib import import_module

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import air_b200 as ab   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=25000)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--train-images", type=int, default=60000)
    ap.add_argument("--val-images", type=int, default=4096)
    ap.add_argument("--every", type=int, default=500)
    ap.add_argument("--gemm", default="fp32", choices=["fp32", "tf32", "tf32x3"])
    ap.add_argument("--log", default=None)
    ap.add_argument("--seed", type=int, default=0, help="initialisation, sampling noise and batch order")
    ap.add_argument("--data", default="device", choices=["device", "oracle"],
                    help="device: air_synth_canvases (generate_multi_image placement); oracle: the CPU generator "
                         "oracle/train_convergence.py trains on (test infrastructure, used here only to repeat that run)")
    ap.add_argument("--clean-gradient", action="store_true", help="reference_rounding=False (the fp64-like gradient)")
    ap.add_argument("--skip-nonfinite", action="store_true",
                    help="skip optimisation steps with a non-finite gradient norm (the reference would go NaN for good)")
    a = ap.parse_args()
    data = import_module("tf-attend-infer-repeat_b200.data")
    if a.data == "oracle":
        from oracle import air_oracle as O
        train, train_cnt = (t.cuda() for t in O.synthetic_canvases(a.train_images, seed=0))
        val, val_cnt = (t.cuda() for t in O.synthetic_canvases(a.val_images, seed=12345))
    else:
        train, train_cnt = data.device_canvases(a.train_images, seed=0)
        val, val_cnt = data.device_canvases(a.val_images, seed=12345)
    ab.reset_variable_scopes()
    m = ab.AIRModel(train[:a.batch].clone(), train_cnt[:a.batch].clone(), train=True,
                    annealing_schedules=data.TRAINING_ANNEALING, gemm_mode=a.gemm, seed=a.seed, reference_rounding=not a.clean_gradient,
                    skip_nonfinite_updates=a.skip_nonfinite,
                    **data.TRAINING_HYPER)
    ev = ab.AIRModel(val, val_cnt, train=False, reuse=True, annealing_schedules=data.TRAINING_ANNEALING,
                     gemm_mode=a.gemm, **data.TRAINING_HYPER)
    m.capture()
    log = open(a.log, "w") if a.log else None

    def emit(s):
        print(s, flush=True)
        if log:
            log.write(s + "\n")
            log.flush()

    emit(f"# CUDA training, batch {a.batch}, gemm {a.gemm}; columns: iteration  train_loss  train_acc  "
         f"val_acc(test mode)  val_acc_by_count(0/1/2)  seconds")
    g = torch.Generator(device="cuda").manual_seed(1 + a.seed)
    t0 = time.time()
    loss_sum = torch.zeros((), device="cuda")
    acc_sum = torch.zeros((), device="cuda")
    for it in range(a.iters + 1):
        if it % a.every == 0:
            ev.run()
            hit = (ev.rec_num_digits == val_cnt).float()
            by = [hit[val_cnt == k].mean().item() for k in range(3)]
            n = max(1, min(it, a.every))
            emit(f"{it:6d}  {loss_sum.item() / n:10.3f}  {acc_sum.item() / n:.4f}  {hit.mean().item():.4f}  "
                 f"{by[0]:.3f}/{by[1]:.3f}/{by[2]:.3f}  {time.time() - t0:7.1f}")
            loss_sum.zero_()
            acc_sum.zero_()
        if it == a.iters:
            break
        idx = torch.randint(0, a.train_images, (a.batch,), generator=g, device="cuda")
        m.feed(train[idx], train_cnt[idx])
        m.train_step()
        loss_sum += m.loss.reshape(())
        acc_sum += m.accuracy.reshape(())


if __name__ == "__main__":
    main()
