"""Diagnostic: which earlier phase of bench.py makes a later AIRModel.capture() fail?  PHASE=st|train|infer|none"""
import copy
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import bench_train  # noqa: E402

phase = os.environ.get("PHASE", "none")
peaks = bench.measured_peaks()
args = type("A", (), dict(gpus=1, steps=5, warmup=3, impl="b200", workload="train", batch=None, global_batch=None, gemm=None,
                          cnn=False, no_convergence=True))()
for ph in phase.split("+"):
    if ph == "st":
        bench.st_summary(peaks)
    elif ph in ("train", "infer"):
        a = copy.copy(args)
        a.workload = ph
        if ph == "infer":
            torch.cuda.empty_cache()
            a.batch, a.steps = None, 20
        bench_train.run(a, 0, 1, peaks)
torch.cuda.synchronize()
print("phase", phase, "done; capturing", flush=True)
try:
    print(bench.training_convergence((3,), 200))
except Exception:
    traceback.print_exc()
