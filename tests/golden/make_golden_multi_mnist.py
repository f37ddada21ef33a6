"""Golden vectors for the canvas generator, produced by THE REFERENCE'S OWN generate_multi_image
(/root/reference/multi_mnist.py:82-183, imported unmodified through oracle/tfgraph/tf_shim.py).

The reference composes MNIST digits; MNIST is not available offline, so the "digits" fed to it are the stroke-like
blobs of oracle/synth_oracle.py (in 28x28 frames, like MNIST images), and its np.random.randint draws -- the (x, y)
placement attempts -- are served from the same counter-based stream the device generator uses.  Everything else is
the reference's code: crop_non_empty, the 100-attempt loop, pixels_overlap, the canvas restart, the positions / boxes
bookkeeping.  Output: tests/golden/ref_multi_mnist_gen.npz.

    python tests/golden/make_golden_multi_mnist.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import synth_oracle as S  # noqa: E402

SEED, FIRST, N = 11, 5, 96


class ServedRandom:
    """np.random for the reference: randint pops the recorded draws and checks the requested range."""

    def __init__(self, xy):
        self.xy, self.i = list(xy), 0

    def randint(self, low, high):
        if self.i >= len(self.xy):   # (not an IndexError: generate_multi_image swallows those, :175-176, and would loop forever)
            raise RuntimeError("the reference asked for more placement draws than the generator used")
        v = self.xy[self.i]
        self.i += 1
        assert low == 0 and 0 <= v < high, (low, high, v)
        return v

    def permutation(self, a):
        return np.asarray(a)


class Np:
    def __init__(self, rnd):
        self.random = rnd

    def __getattr__(self, name):
        return getattr(np, name)


def load_reference():
    from oracle.tfgraph import tf_shim as T
    extra = {"tensorflow.examples": types.ModuleType("e"), "tensorflow.examples.tutorials": types.ModuleType("t"),
             "tensorflow.examples.tutorials.mnist": types.ModuleType("m")}
    extra["tensorflow.examples.tutorials.mnist"].input_data = None
    with T.installed():
        sys.modules.update(extra)
        try:
            return T.load_reference_module("/root/reference/multi_mnist.py")
        finally:
            for k in extra:
                sys.modules.pop(k, None)


def reference_canvas(ref, seed, img, cs=50, max_digits=2):
    """generate_multi_image of the reference on the blobs / draws the generator uses for image ``img``."""
    trace = {}
    S.one_canvas(seed, img, cs, max_digits, trace=trace)
    frames = np.zeros((len(trace["blobs"]) + 1, 28 * 28), np.float32)      # (+1: the id pool never runs dry)
    for i, b in enumerate(trace["blobs"]):
        f = np.zeros((28, 28), np.float32)
        f[2:2 + b.shape[0], 1:1 + b.shape[1]] = b                           # anywhere in the frame: it gets cropped
        frames[i] = f.ravel()
    rnd = ServedRandom(trace["xy"])
    ref.np = Np(rnd)
    ref.digit_ids, ref.next_digit_id, ref.used_digit_ids = np.arange(len(frames)), 0, set()
    ref.num_digits = trace["count"]                                         # the global generate_multi_image reads (:95)
    canvas, ids, pos, box = ref.generate_multi_image(frames, trace["count"], 28, cs, use_pixel_overlap=True)
    assert rnd.i == len(rnd.xy), "the reference consumed a different number of placement draws"
    return canvas, trace["count"], ids, pos, box


def main():
    ref = load_reference()
    canv, cnt, pos, box = [], [], [], []
    for b in range(N):
        c, k, ids, p, bx = reference_canvas(ref, SEED, FIRST + b)
        assert len(ids) == k and len(p) == 2 * k and len(bx) == 2 * k
        canv.append(c.ravel())
        cnt.append(k)
        pos.append(list(p) + [0] * (4 - len(p)))
        box.append(list(bx) + [0] * (4 - len(bx)))
    out = os.path.join(ROOT, "tests", "golden", "ref_multi_mnist_gen.npz")
    np.savez_compressed(out, seed=SEED, first_index=FIRST, canvases=np.asarray(canv, np.float32),
                        counts=np.asarray(cnt, np.int32), positions=np.asarray(pos, np.int32), boxes=np.asarray(box, np.int32))
    print("wrote", out, "counts:", np.bincount(cnt))


if __name__ == "__main__":
    main()
