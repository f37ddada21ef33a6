"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN GRAPH (/root/reference/model/air-model.meta, the
MetaGraphDef training.py:141 saved: train model ``air``, its autodiff gradient graph, clip + ApplyAdam, and the
test model ``air_1``) with the numpy graph interpreter oracle/tfgraph.  Run in the build container, where
/root/reference exists:   python tests/golden/make_golden_ref_graph.py
The fixtures are tests/parity_util.py's (seeded, regenerated at test time), so only outputs are stored."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.tfgraph import air_graph as G, pb            # noqa: E402
from oracle.tfgraph.interp import Interpreter, run_in_big_stack   # noqa: E402
from tests import parity_util as PU                      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SMALL = ("rec_scales", "rec_shifts", "rec_st_back", "rec_latents", "z_pres_probs", "z_pres_kls", "scale_kls", "shift_kls",
         "vae_kls", "stop_masks", "rec_num_digits", "running_loss", "stopping_sum", "reconstruction_loss", "loss_per_item")


def st_inputs(B=64, seed=1):
    """SURVEY 8d microbench poses: s~U(0.3,0.9), x,y~U(-0.5,0.5) -> most canvas pixels fall outside the window."""
    from oracle import air_oracle as O
    rng = np.random.default_rng(seed)
    s = rng.uniform(0.3, 0.9, B).astype(np.float32)
    x = rng.uniform(-0.5, 0.5, B).astype(np.float32)
    y = rng.uniform(-0.5, 0.5, B).astype(np.float32)
    imgs = O.synthetic_canvases(B, seed=0)[0].numpy()
    rec = rng.uniform(0, 1, (B, 784)).astype(np.float32)
    return s, x, y, imgs, rec


def benign_upstream(B=64, seed=3):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((B, 50, 50)).astype(np.float32), rng.standard_normal((B, 28, 28)).astype(np.float32)


def graph_st(nodes, s, x, y, imgs, rec):
    """The two transformer() instances of the loop body (air_model.py:322-333, 351-366) evaluated in isolation."""
    def body():
        I = Interpreter(nodes, {}, {"pipeline/shuffle_batch:0": imgs, "air/rnn/while/scale/strided_slice:0": s,
                                    "air/rnn/while/shift/strided_slice:0": x,
                                    "air/rnn/while/shift/strided_slice_1:0": y,
                                    "air/rnn/while/vae/gen_sample/Sigmoid:0": rec})
        return (I.eval("air/rnn/while/st_forward/strided_slice", 0, 0),
                I.eval("air/rnn/while/st_backward/strided_slice", 0, 0),
                I.eval("air/rnn/while/st_forward/stack_2", 0, 0), I.eval("air/rnn/while/st_backward/stack_2", 0, 0))
    return run_in_big_stack(body)


def pack_train(out, n=1024):
    d = {k: out[k] for k in SMALL if k in out}
    d.update(loss=out["loss"], accuracy=out["accuracy"], executed_steps=np.int32(out["executed_steps"]),
             global_norm=out["global_norm"], rec_windows_sub=out["rec_windows"][:, :, ::8],
             reconstruction_sub=out["reconstruction"][:, ::8])
    for k, g in out["raw_grads"].items():
        d["g:" + k], d["gn:" + k] = G.digest(g, n)
    d["clip_scale"] = np.float64(np.linalg.norm(out["clipped_grads"]["rnn/bias"].astype(np.float64)) /
                                 np.linalg.norm(out["raw_grads"]["rnn/bias"].astype(np.float64)))
    for k, v in out["new_variables"].items():              # updated variables; Adam slots are implied by them
        if "/Adam" in k:
            continue
        if np.ndim(v):
            d["v:" + k], d["vn:" + k] = G.digest(v, n)
        else:
            d["v:" + k] = np.asarray(v)
    return d


def main():
    nodes = pb.load_metagraph(G.META)
    assert nodes[1] == "1.3.0"
    s, x, y, imgs, rec = st_inputs()
    win, back, th, thinv = graph_st(nodes, s, x, y, imgs, rec)
    np.savez_compressed(os.path.join(HERE, "ref_graph_st.npz"), crop=win, back=back, theta=th, theta_inv=thinv)

    imgs, cnt, params, noise = PU.covered_fixture(64, seed=3)
    out = G.run_train_step(nodes, params, imgs, cnt, noise, global_step=2000)
    np.savez_compressed(os.path.join(HERE, "ref_graph_train_covered.npz"), **pack_train(out))

    # gradients TF's autodiff graph carries into and out of the two Spatial Transformers, tapped inside the gradient
    # loop of an fp32 train step on realistic poses (step t = 1, first 16 items; the ST is per-item independent)
    imgs, cnt, params, noise = PU.realistic_fixture(64, seed=1)
    out = G.run_train_step(nodes, params, imgs, cnt, noise, global_step=2000, taps=G.ST_TAPS)
    np.savez_compressed(os.path.join(HERE, "ref_graph_st_grad.npz"),
                        **{k[4:]: v[1][:16] for k, v in out.items() if k.startswith("tap:")})

    # the same taps with a benign upstream gradient (N(0,1), fed at the tensors that enter the two ST gradient
    # subgraphs): well-conditioned enough to hold a CUDA kernel with its own summation order to a tolerance
    d_wb, d_crop = benign_upstream()
    out = G.run_train_step(nodes, params, imgs, cnt, noise, global_step=2000, taps=G.ST_TAPS,
                           extra_feeds={G.ST_TAPS["d_writeback"][0] + ":0": d_wb, G.ST_TAPS["d_crop"][0] + ":0": d_crop})
    np.savez_compressed(os.path.join(HERE, "ref_graph_st_grad_benign.npz"),
                        **{k[4:]: v[1][:16] for k, v in out.items() if k.startswith("tap:") and not k.startswith("tap:d_crop")
                           and not k.startswith("tap:d_writeback")})

    out = G.run_train_step(nodes, params, imgs, cnt, noise, global_step=2000, float_dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "ref_graph_train_realistic_fp64.npz"), **pack_train(out))

    for name, fx in (("default", PU.default_fixture(64, seed=1)), ("realistic", PU.realistic_fixture(64, seed=2))):
        imgs, cnt, params, noise = fx
        out = G.run_test_model(nodes, params, imgs, cnt, noise)
        d = {k: out[k] for k in SMALL if k in out}
        d.update(summaries=np.array(json.dumps(out["summaries"])), accuracy=out["accuracy"], executed_steps=np.int32(out["executed_steps"]),
                 rec_windows_sub=out["rec_windows"][:, :, ::8])
        np.savez_compressed(os.path.join(HERE, f"ref_graph_test_{name}.npz"), **d)
        if name == "realistic":
            # the tensor behind tf.summary.image("reconstruction") (air_model.py:211-267) with its inputs, first 12
            # images: 4 stored in full, all 12 as SHA-256 of the float32 bytes (the computation is bit-reproducible)
            im = out["reconstruction_image"][:12]
            np.savez_compressed(os.path.join(HERE, "ref_graph_vis.npz"), reconstruction=out["reconstruction"][:12],
                                rec_st_back=out["rec_st_back"][:12], rec_num_digits=out["rec_num_digits"][:12],
                                image_full=im[:4],
                                sha256=np.array([hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() for a in im]))
    # inference configuration (BASELINE configs[4]): the same test graph with cond()'s max_steps constant fed as 5
    imgs, cnt, params, noise = PU.realistic_fixture(64, seed=8, T=5)
    out = G.run_test_model(nodes, params, imgs, cnt, noise, max_steps=5)
    d = {k: out[k] for k in SMALL if k in out}
    d.update(accuracy=out["accuracy"], executed_steps=np.int32(out["executed_steps"]),
             rec_windows_sub=out["rec_windows"][:, :, ::8])
    np.savez_compressed(os.path.join(HERE, "ref_graph_test_realistic_T5.npz"), **d)
    for f in sorted(os.listdir(HERE)):
        if f.startswith("ref_graph"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
