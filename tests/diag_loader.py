"""Diagnostic: throughput of the TFRecord input path (tfrecords.TFRecordFile) on this host: index build, batch gather into
pinned buffers, and gather + H2D copy + uint8/fp32 feed.  python tests/diag_loader.py [n_records]"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import air_b200 as ab  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
imgs, cnt = ab.data.synthetic_canvases(2048, seed=0)
images = [im.reshape(50, 50).numpy() for im in imgs]
digits = cnt.numpy().tolist()
reps = (n + 2047) // 2048
pad = lambda d, k: [0] * k
with tempfile.TemporaryDirectory() as d:
    t0 = time.perf_counter()
    ab.tfrecords.write_to_records(os.path.join(d, "common"), (images * reps)[:n], [pad(x, 2) for x in (digits * reps)[:n]],
                                  [pad(x, 4) for x in (digits * reps)[:n]], [pad(x, 4) for x in (digits * reps)[:n]],
                                  [pad(x, 2) for x in (digits * reps)[:n]], (digits * reps)[:n])
    path = os.path.join(d, "common.tfrecords")
    print(f"wrote {n} records ({os.path.getsize(path) / 1e6:.0f} MB) in {time.perf_counter() - t0:.1f} s; host cores: {os.cpu_count()}")
    t0 = time.perf_counter()
    f = ab.tfrecords.TFRecordFile(path)
    print(f"index (framing + CRC-32C + Example parse, one pass): {time.perf_counter() - t0:.2f} s = {n / (time.perf_counter() - t0) / 1e6:.2f} M records/s")
    for B in (4096, 16384):
        it = f.batches(B, shuffle_buffer=10000, seed=0, epochs=20, pin_memory=torch.cuda.is_available())
        next(it)
        t0, k = time.perf_counter(), 0
        for im, dg in it:
            k += len(dg)
        dt = time.perf_counter() - t0
        print(f"batches of {B} into pinned buffers (shuffle queue 10000): {k / dt / 1e6:.2f} M images/s = {k * 10000 / dt / 1e9:.1f} GB/s")
    if torch.cuda.is_available():
        B = 4096
        dev = torch.empty(B, 2500, device="cuda")
        it = f.batches(B, shuffle_buffer=10000, seed=0, epochs=20)
        next(it)
        torch.cuda.synchronize()
        t0, k = time.perf_counter(), 0
        for im, dg in it:
            dev.copy_(im, non_blocking=True)
            k += len(dg)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"gather + H2D (fp32 records, batch {B}): {k / dt / 1e6:.2f} M images/s")
