"""TEST INFRASTRUCTURE ONLY (see oracle/air_oracle.py): numpy restatement, bit for bit, of the device-side
canvas generator air_synth_canvases (csrc/model_ops.cu: synth_canvases_k).  The generator itself is a stand-in
for the reference's multi_mnist.py:82-183 (MNIST is not available offline): 0..max_digits stroke-like blobs per
canvas, uniform placement with pixel-overlap rejection (generate_multi_image, multi_mnist.py:141-160)."""
import numpy as np

M64 = (1 << 64) - 1


def _rnd(seed, image, draw):
    z = (seed + (image + 1) * 0x9E3779B97F4A7C15 + draw * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def synth_canvases(B, seed=0, first_index=0, canvas_size=50, max_digits=2):
    f = np.float32
    cs = canvas_size
    images = np.zeros((B, cs, cs), np.float32)
    counts = np.zeros(B, np.int32)
    for b in range(B):
        img = first_index + b
        count = (_rnd(seed, img, 0) >> 33) % (max_digits + 1)
        counts[b] = count
        draw = 1
        canvas = images[b]
        for _ in range(count):
            for _attempt in range(20):
                hh = 14 + (_rnd(seed, img, draw + 0) >> 33) % 11
                ww = 10 + (_rnd(seed, img, draw + 1) >> 33) % 15
                bar_on = f((_rnd(seed, img, draw + 2) >> 33) & 1)
                u = f(_rnd(seed, img, draw + 3) >> 40) * f(2.0 ** -24)
                bar_off = f(-2.0) + f(4.0) * u
                top = (_rnd(seed, img, draw + 4) >> 33) % (cs - hh + 1)
                left = (_rnd(seed, img, draw + 5) >> 33) % (cs - ww + 1)
                draw += 6
                cy, cx = f(hh - 1) / f(2), f(ww - 1) / f(2)
                ry, rx = max(f(hh) / f(2) - f(1.5), f(2)), max(f(ww) / f(2) - f(1.5), f(2))
                y, x = np.mgrid[0:hh, 0:ww].astype(np.float32)
                dy, dx = (y - cy) / ry, (x - cx) / rx
                r = np.sqrt(dy * dy + dx * dx, dtype=np.float32)
                ring = np.clip(f(1) - np.abs(r - f(0.8)) * f(3), f(0), f(1))
                bar = np.clip(f(1) - np.abs((x - cx) - bar_off) / f(1.6), f(0), f(1)) * bar_on
                v = np.maximum(ring, bar).astype(np.float32)
                v[v < f(0.15)] = 0
                region = canvas[top:top + hh, left:left + ww]
                if np.any((v > 0) & (region > 0)):
                    continue
                np.maximum(region, v, out=region)
                break
    return images.reshape(B, -1), counts
