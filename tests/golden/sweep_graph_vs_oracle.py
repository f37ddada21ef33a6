"""Seed sweep: the oracle against the reference graph (oracle/tfgraph) beyond the committed fixtures.
    python tests/golden/sweep_graph_vs_oracle.py > profiles/r1_graph_vs_oracle_sweep.md      (needs /root/reference)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import air_oracle as O                      # noqa: E402
from oracle.tfgraph import air_graph as G, pb           # noqa: E402
from tests import parity_util as PU                     # noqa: E402

KEYS = ("rec_scales", "rec_shifts", "rec_st_back", "rec_windows", "z_pres_probs", "z_pres_kls", "scale_kls", "shift_kls",
        "vae_kls")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def main(seeds=range(10, 20)):
    nodes = pb.load_metagraph(G.META)
    print("# Oracle vs reference graph, seed sweep (B = 64, global_step = 2000)\n")
    print("| fixture | dtype | seed | digits / masks equal | loss rel | worst per-step rel | worst of 36 gradients (norm-wise) | global norm rel |")
    print("|---|---|---|---|---|---|---|---|")
    worst = {}
    for name, fx, dt in (("covered", PU.covered_fixture, np.float32), ("realistic", PU.realistic_fixture, np.float64),
                         ("default-init", PU.default_fixture, np.float64)):
        for seed in seeds:
            imgs, cnt, params, noise = fx(64, seed=seed)
            out = G.run_train_step(nodes, params, imgs, cnt, noise, global_step=2000, float_dtype=dt)
            tdt = torch.float32 if dt is np.float32 else torch.float64
            orc = O.AIROracle(params={k: v.to(tdt) for k, v in params.items()}, annealing_schedules=O.DEFAULT_ANNEALING,
                              dtype=tdt)
            orc.global_step = 2000
            o, g = orc.loss_and_grads(imgs.to(tdt), cnt, {k: v.to(tdt) for k, v in noise.items()})
            same = np.array_equal(out["rec_num_digits"], o["rec_num_digits"].numpy()) and \
                np.array_equal(out["stop_masks"], o["stop_masks"].numpy())
            lr = abs(float(out["loss"]) - float(o["loss"])) / abs(float(o["loss"]))
            ps = max(rel(out[k], o[k].numpy().reshape(out[k].shape)) for k in KEYS)
            gr = max(rel(out["raw_grads"][k], g[k].numpy()) for k in g if np.linalg.norm(g[k].numpy()) > 0)
            _, norm = O.AIROracle.clip_by_global_norm(g, 1.0)
            nr = abs(float(out["global_norm"]) - float(norm)) / float(norm)
            print(f"| {name} | {np.dtype(dt).name} | {seed} | {same} | {lr:.1e} | {ps:.1e} | {gr:.1e} | {nr:.1e} |")
            w = worst.setdefault((name, np.dtype(dt).name), [True, 0, 0, 0, 0])
            w[0] &= same
            w[1:] = [max(a, b) for a, b in zip(w[1:], (lr, ps, gr, nr))]
    print("\n| fixture | dtype | all discrete equal | max loss rel | max per-step rel | max gradient rel | max global-norm rel |")
    print("|---|---|---|---|---|---|---|")
    for (name, dt), w in worst.items():
        print(f"| {name} | {dt} | {w[0]} | {w[1]:.1e} | {w[2]:.1e} | {w[3]:.1e} | {w[4]:.1e} |")


if __name__ == "__main__":
    main()
