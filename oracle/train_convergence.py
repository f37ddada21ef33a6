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


def _use_graph_order():
    """transformer.py:56-117 with the four gathers expressed as ONE gather of the concatenated index list: autograd then
    accumulates their gradients as one list in the order a | b | c | d -- the order of the reference graph's
    gradients/concat -> UnsortedSegmentSum (TensorFlow's CPU kernel is sequential).  Forward values are bit-identical."""
    def interpolate(im, x, y, out_size):
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
        vals = im.reshape(-1, channels)[idx]
        n = x.numel()
        Ia, Ib, Ic, Id = vals[:n], vals[n:2 * n], vals[2 * n:3 * n], vals[3 * n:]
        x0_f, x1_f, y0_f, y1_f = x0.to(dtype), x1.to(dtype), y0.to(dtype), y1.to(dtype)
        wa = ((x1_f - x) * (y1_f - y)).unsqueeze(1); wb = ((x1_f - x) * (y - y0_f)).unsqueeze(1)
        wc = ((x - x0_f) * (y1_f - y)).unsqueeze(1); wd = ((x - x0_f) * (y - y0_f)).unsqueeze(1)
        return ((wa * Ia + wb * Ib) + wc * Ic) + wd * Id
    O._interpolate = interpolate


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
    ap.add_argument("--order", default="torch", choices=["torch", "graph"],
                    help="accumulation order of the gather gradients in _interpolate: torch = four buffers (autograd's own), "
                         "graph = ONE list a|b|c|d as the reference graph's gradients/concat -> UnsortedSegmentSum does")
    ap.add_argument("--save", default=None, help="write the final parameters (float16 .npz, reference variable names)")
    a = ap.parse_args()
    torch.set_num_threads(a.threads)
    torch.manual_seed(0)
    if a.order == "graph":
        _use_graph_order()
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
