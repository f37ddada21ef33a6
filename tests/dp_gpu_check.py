"""Multi-GPU data-parallel check (run under torchrun, one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_gpu_check.py
Each rank runs the CUDA AIRModel on its row shard (injected noise shard); after the flat-gradient
all-reduce every rank must hold the single-GPU global-batch gradient (up to summation order), and
after identical Adam steps the replicas must stay bit-identical."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import air_b200 as ab  # noqa: E402
from oracle import air_oracle as O  # noqa: E402
from tests.parity_util import covered_fixture, relnorm  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
B = 32 * world
imgs, cnt, params, noise = covered_fixture(B, seed=0)
hyper = dict(O.DEFAULT_HYPER)


def build(images, counts, scope):
    m = ab.AIRModel(images.cuda(), counts.cuda(), train=True, annealing_schedules=O.DEFAULT_ANNEALING, scope=scope, **hyper)
    m.store.load_named({k: v.cuda() for k, v in params.items()})
    m.store.global_step = 2000
    return m


m = build(ab.dp.shard_rows(imgs, rank, world), ab.dp.shard_rows(cnt, rank, world), "dp")
assert m.world == world
m.loss_and_grads({k: v.cuda() for k, v in ab.dp.shard_noise(noise, rank, world).items()})
g_dp = m.store.grad.clone()
ok = True
if rank == 0:
    # single-GPU reference on the full batch (no process group -> world 1)
    ref = ab.AIRModel(imgs.cuda(), cnt.cuda(), train=True, annealing_schedules=O.DEFAULT_ANNEALING, scope="ref",
                      process_group="local", **hyper)
    ref.store.load_named({k: v.cuda() for k, v in params.items()})
    ref.store.global_step = 2000
    ref.set_noise({k: v.cuda() for k, v in noise.items()})
    ref._forward(); ref._publish(); ref._backward()
    e = relnorm(g_dp, ref.store.grad)
    print(f"[dp] world={world} reduced gradient vs single-GPU global batch: rel err {e:.3e}")
    ok &= e < 1e-5
# replicas stay identical after optimizer steps
for _ in range(3):
    m.train_step()
flat = m.store.flat.clone()
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
same = all(torch.equal(gathered[0], g) for g in gathered)
# the CUDA-graph path: ONE graph per step with the five bucketed NCCL all-reduces captured on forked branches
m.capture()
for _ in range(3):
    m.train_step()
torch.cuda.synchronize()
flat2 = m.store.flat.clone()
gathered2 = [torch.empty_like(flat2) for _ in range(world)]
dist.all_gather(gathered2, flat2)
same_graph = all(torch.equal(gathered2[0], g) for g in gathered2) and bool(torch.isfinite(flat2).all()) and not torch.equal(flat2, flat)
if rank == 0:
    print(f"[dp] replicas bit-identical after 3 steps: {same}; global_step={m.global_step}")
    print(f"[dp] replicas bit-identical after 3 more CUDA-graph steps with the overlapped all-reduce: {same_graph}")
    ok &= same and same_graph
    print("DP CHECK", "OK" if ok else "FAILED")
dist.barrier()
# the step's CUDA graph holds captured NCCL kernels; ncclCommDestroy waits for such graphs: drop them first, and do
# not let a blocking teardown hang the job
import gc      # noqa: E402
import threading  # noqa: E402
m._graphs = None
del m
gc.collect()
torch.cuda.synchronize()
t = threading.Thread(target=dist.destroy_process_group, daemon=True)
t.start()
t.join(20.0)
sys.stdout.flush()
os._exit(0 if ok else 1)
