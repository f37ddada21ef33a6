"""What does the tensor-core (TF32) GEMM mode do to the model's results?  CPU study (test infrastructure): the oracle
with every dense contraction's operands cut to TF32 (10 explicit mantissa bits; truncation = the pessimistic reading of
the tensor core's operand conversion, round-to-nearest also shown), fp32 accumulation -- against the plain fp32 oracle,
with the TRAINED weights on held-out canvases and with default-init weights.

    python oracle/tf32_study.py > profiles/r1_tf32_study.md
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import air_oracle as O                      # noqa: E402
from tests.test_trained_weights import heldout, trained_params   # noqa: E402


def to_tf32(t, mode):
    """mode: 'truncate' / 'rn' = TF32 (13 low mantissa bits dropped); 'bf16 rn' = BF16 operands (16 dropped), the
    candidate throughput mode of DESIGN.md section 8."""
    drop = 16 if mode.startswith("bf16") else 13
    i = t.contiguous().view(torch.int32)
    if mode.endswith("rn"):                             # round to nearest even on the first kept bit
        i = i + ((1 << (drop - 1)) - 1) + ((i >> drop) & 1)
    return (i & ~((1 << drop) - 1)).view(torch.float32)


class tf32_matmul:
    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.orig = torch.matmul
        torch.matmul = lambda a, b: self.orig(to_tf32(a, self.mode), to_tf32(b, self.mode))

    def __exit__(self, *a):
        torch.matmul = self.orig


def rel(a, b):
    return float(torch.linalg.norm((a - b).double()) / torch.linalg.norm(b.double()))


def main():
    print("# TF32 operand precision in the dense contractions: effect on the model (CPU study with the oracle)\n")
    print("Operands of every GEMM (LSTM, heads, VAE) cut to TF32, fp32 accumulation; everything else fp32.\n")
    params, step = trained_params()
    imgs, cnt, noise = heldout(B=1024, seed=991)
    rows = []
    for name, p, train in (("trained weights, test mode", params, False), ("trained weights, train mode", params, True),
                           ("default init, train mode", O.init_params(seed=0), True)):
        def run():
            orc = O.AIROracle(params=p, annealing_schedules=O.DEFAULT_ANNEALING, train=train)
            orc.global_step = step
            with torch.no_grad():
                return orc.forward(imgs, cnt, noise)
        ref = run()
        for mode in ("truncate", "rn", "bf16 rn"):
            with tf32_matmul(mode):
                out = run()
            rows.append((name, mode,
                         (out["rec_num_digits"] == ref["rec_num_digits"]).float().mean().item(),
                         (ref["rec_num_digits"] == cnt).float().mean().item(),
                         (out["rec_num_digits"] == cnt).float().mean().item(),
                         rel(out["rec_scales"], ref["rec_scales"]), rel(out["rec_shifts"], ref["rec_shifts"]),
                         rel(out["rec_windows"], ref["rec_windows"]),
                         abs(float(out["loss"]) - float(ref["loss"])) / abs(float(ref["loss"]))))
    print("| case | operand conversion | same digit count as fp32 | accuracy fp32 | accuracy reduced | scales rel | shifts rel | windows rel | loss rel |")
    print("|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r[0]} | {r[1]} | {r[2]:.4f} | {r[3]:.4f} | {r[4]:.4f} | {r[5]:.1e} | {r[6]:.1e} | {r[7]:.1e} | {r[8]:.1e} |")
    print("\n1024 held-out synthetic canvases, same noise in both runs.  The loss column includes the canvas-residue "
          "sensitivity of DESIGN.md section 2 (a 1e-4 change of a pose moves uncovered pixels' terms by nats).")


if __name__ == "__main__":
    main()
