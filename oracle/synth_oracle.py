"""TEST INFRASTRUCTURE ONLY (see oracle/air_oracle.py): numpy restatement, bit for bit, of the device-side
canvas generator air_synth_canvases (csrc/model_ops.cu: synth_canvases_k).

The generator stands in for the reference's data set (multi_mnist.py needs the MNIST download) and follows
generate_multi_image's placement semantics (multi_mnist.py:82-183 with use_pixel_overlap, gap = margin = 0):
every "digit" (here: a stroke-like blob) is cropped to its non-empty bounding box (crop_non_empty, :36-43), placed
at a uniform position -- x drawn before y (:143-144) -- the first one always fits, later ones are re-drawn up to
100 times (:141) until no pixel overlaps the canvas (pixels_overlap, :61-65), and a digit that cannot be placed
restarts the whole canvas with fresh digits (:95-171).  tests/test_reference_source.py runs the reference's own
generate_multi_image on the blobs and position draws of this file and requires identical canvases / positions / boxes.
"""
import numpy as np

M64 = (1 << 64) - 1
ATTEMPTS = 100   # multi_mnist.py:141
RESTARTS = 64    # the reference retries forever; see synth_canvases_k


def _rnd(seed, image, draw):
    z = (seed + (image + 1) * 0x9E3779B97F4A7C15 + draw * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def blob(seed, img, draw):
    """The hh x ww frame of one "digit" from draws draw .. draw+3 (values in (0, 1], background 0)."""
    f = np.float32
    hh = 14 + (_rnd(seed, img, draw + 0) >> 33) % 11
    ww = 10 + (_rnd(seed, img, draw + 1) >> 33) % 15
    bar_on = f((_rnd(seed, img, draw + 2) >> 33) & 1)
    u = f(_rnd(seed, img, draw + 3) >> 40) * f(2.0 ** -24)
    bar_off = f(-2.0) + f(4.0) * u
    cy, cx = f(hh - 1) / f(2), f(ww - 1) / f(2)
    ry, rx = max(f(hh) / f(2) - f(1.5), f(2)), max(f(ww) / f(2) - f(1.5), f(2))
    y, x = np.mgrid[0:hh, 0:ww].astype(np.float32)
    dy, dx = (y - cy) / ry, (x - cx) / rx
    r = np.sqrt(dy * dy + dx * dx, dtype=np.float32)
    ring = np.clip(f(1) - np.abs(r - f(0.8)) * f(3), f(0), f(1))
    bar = np.clip(f(1) - np.abs((x - cx) - bar_off) / f(1.6), f(0), f(1)) * bar_on
    v = np.maximum(ring, bar).astype(np.float32)
    v[v < f(0.15)] = 0
    return v


def crop(v):
    ys, xs = np.nonzero(v.sum(axis=1))[0], np.nonzero(v.sum(axis=0))[0]
    return v[ys[0]:ys[-1] + 1, xs[0]:xs[-1] + 1]


def one_canvas(seed, img, cs=50, max_digits=2, trace=None):
    """-> (canvas [cs, cs], count, positions [2*max_digits], boxes [2*max_digits]).  ``trace`` (a dict) receives the
    blobs and the (x, y) position draws in the order they were consumed, for the reference-side replay."""
    count = int((_rnd(seed, img, 0) >> 33) % (max_digits + 1))
    draw = 1
    if trace is not None:
        trace.update(blobs=[], xy=[], count=count)
    for _restart in range(RESTARTS):
        canvas = np.zeros((cs, cs), np.float32)
        positions, boxes = np.zeros(2 * max_digits, np.int32), np.zeros(2 * max_digits, np.int32)
        failed = False
        for k in range(count):
            full = blob(seed, img, draw)
            draw += 4
            v = crop(full)
            if trace is not None:
                trace["blobs"].append(full)
            h, w = v.shape
            found = False
            for _attempt in range(ATTEMPTS):
                x = int((_rnd(seed, img, draw + 0) >> 33) % (cs - w + 1))
                y = int((_rnd(seed, img, draw + 1) >> 33) % (cs - h + 1))
                draw += 2
                if trace is not None:
                    trace["xy"] += [x, y]
                found = k == 0 or not np.any((v > 0) & (canvas[y:y + h, x:x + w] > 0))
                if found:
                    break
            if not found:
                failed = True
                break
            canvas[y:y + h, x:x + w] += v
            positions[2 * k:2 * k + 2] = (x, y)
            boxes[2 * k:2 * k + 2] = (w, h)
        if not failed:
            return canvas, count, positions, boxes
    z = np.zeros(2 * max_digits, np.int32)
    return np.zeros((cs, cs), np.float32), 0, z, z.copy()


def synth_canvases(B, seed=0, first_index=0, canvas_size=50, max_digits=2, with_boxes=False):
    images = np.zeros((B, canvas_size, canvas_size), np.float32)
    counts = np.zeros(B, np.int32)
    positions, boxes = np.zeros((B, 2 * max_digits), np.int32), np.zeros((B, 2 * max_digits), np.int32)
    for b in range(B):
        images[b], counts[b], positions[b], boxes[b] = one_canvas(seed, first_index + b, canvas_size, max_digits)
    if with_boxes:
        return images.reshape(B, -1), counts, positions, boxes
    return images.reshape(B, -1), counts
