"""Does the restated algorithm LEARN?  (test infrastructure / evidence, CPU only)

Trains the CPU oracle (oracle/air_oracle.py -- the restatement that tests/test_reference_graph.py pins to the
reference's own graph) with the reference's training configuration (training.py:100-122: batch 64, lr 1e-4, clip 1.0,
z_pres prior annealed 1e4 * 0.1^(step/3000) in log space) on the synthetic 0-2 blob canvases of SURVEY 8d, fresh noise
every step, and logs the digit-count accuracy of the test-mode model (train=False, rounded z_pres) on held-out canvases
-- the quantity behind the reference README's "98 % in ~25k iterations" claim (there on real multi-MNIST, which is not
available offline).

    python oracle/train_convergence.py --iters 25000 --log profiles/r1_oracle_convergence.log
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import air_oracle as O   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=25000)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--train-images", type=int, default=20000)
    ap.add_argument("--val-images", type=int, default=1024)
    ap.add_argument("--every", type=int, default=500)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--log", default=None)
    ap.add_argument("--seed", type=int, default=0, help="initialisation, sampling noise and batch order (0 = the committed run)")
    ap.add_argument("--save", default=None, help="write the final parameters (float16 .npz, reference variable names)")
    a = ap.parse_args()
    torch.set_num_threads(a.threads)
    torch.manual_seed(0)
    train, train_cnt = O.synthetic_canvases(a.train_images, seed=0)
    val, val_cnt = O.synthetic_canvases(a.val_images, seed=12345)
    m = O.AIROracle(annealing_schedules=O.DEFAULT_ANNEALING, train=True, seed=a.seed)
    log = open(a.log, "w") if a.log else None

    def emit(s):
        print(s, flush=True)
        if log:
            log.write(s + "\n")
            log.flush()

    emit(f"# oracle training, batch {a.batch}, {a.train_images} synthetic train canvases, {a.val_images} held out, "
         f"{a.threads} threads; columns: iteration  train_loss  train_acc  val_acc(test mode)  val_acc_by_count(0/1/2)  "
         f"z_pres_prior_log_odds  seconds")
    g = torch.Generator().manual_seed(1 + a.seed)
    t0 = time.time()
    run_loss = run_acc = 0.0
    for it in range(a.iters + 1):
        if it % a.every == 0:
            ev = O.AIROracle(params=m.params, annealing_schedules=O.DEFAULT_ANNEALING, train=False)
            ev.global_step = m.global_step
            with torch.no_grad():
                out = ev.forward(val, val_cnt, O.make_noise(10_000 + it, 3, a.val_images))
            hit = (out["rec_num_digits"] == val_cnt).float()
            by = [hit[val_cnt == k].mean().item() for k in range(3)]
            n = max(1, min(it, a.every))
            emit(f"{it:6d}  {run_loss / n:10.3f}  {run_acc / n:.4f}  {hit.mean().item():.4f}  "
                 f"{by[0]:.3f}/{by[1]:.3f}/{by[2]:.3f}  {float(m.hyper('z_pres_prior_log_odds')):8.3f}  {time.time() - t0:7.0f}")
            run_loss = run_acc = 0.0
        if it == a.iters:
            if a.save:
                import numpy as np
                np.savez(a.save, global_step=np.int32(m.global_step),
                         **{k: v.detach().numpy().astype(np.float16) for k, v in m.params.items()})
            break
        idx = torch.randint(0, a.train_images, (a.batch,), generator=g)
        out, _ = m.train_step(train[idx], train_cnt[idx], O.make_noise(it + 1_000_003 * a.seed, 3, a.batch))
        run_loss += float(out["loss"])
        run_acc += float(out["accuracy"])


if __name__ == "__main__":
    main()
