"""World-size-2 data-parallel test on CPU (gloo): the host-side sharding + flat all-reduce used by
AIRModel reproduces the single-process global-batch gradient.  The per-rank compute is the oracle
(no GPU here); the product code under test is tf-attend-infer-repeat_b200/dp.py."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import air_oracle as O
from tests.parity_util import covered_fixture

B, T = 8, 3   # covered fixture: the default-init one has noise-dominated gradients (SURVEY hard part 2)


def _worker(rank, world, init_file, out_dir):
    import air_b200 as ab
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    torch.set_num_threads(2)
    imgs, cnt, params, noise = covered_fixture(B, seed=0, T=T)
    my_imgs = ab.dp.shard_rows(imgs, rank, world)
    my_cnt = ab.dp.shard_rows(cnt, rank, world)
    my_noise = ab.dp.shard_noise(noise, rank, world)
    m = O.AIROracle(params=params, annealing_schedules=O.DEFAULT_ANNEALING, train=True)
    out, grads = m.loss_and_grads(my_imgs, my_cnt, my_noise)      # gradient of the LOCAL mean
    flat = torch.cat([g.reshape(-1) for g in grads.values()]) / world   # per-item weight 1/(B_local*world)
    # the bucketed, asynchronous form AIRModel uses (first bucket reduced while the rest is still being produced)
    # must give exactly what one flat all-reduce gives
    bucketed = flat.clone()
    k = bucketed.numel() * 7 // 10
    wa = ab.dp.allreduce_async(bucketed[:k])
    wb = ab.dp.allreduce_async(bucketed[k:])
    wa.wait(); wb.wait()
    ab.dp.allreduce_flat(flat)
    assert torch.equal(bucketed, flat)
    scal = ab.dp.allreduce_scalars(torch.stack([out["loss"], out["accuracy"]]).clone())
    torch.save({"flat": flat, "scal": scal, "digits": out["rec_num_digits"]}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_global_batch():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "init")
        mp.spawn(_worker, args=(world, init_file, d), nprocs=world, join=True)
        res = [torch.load(os.path.join(d, f"r{r}.pt")) for r in range(world)]
    imgs, cnt, params, noise = covered_fixture(B, seed=0, T=T)
    m = O.AIROracle(params=params, annealing_schedules=O.DEFAULT_ANNEALING, train=True)
    out, grads = m.loss_and_grads(imgs, cnt, noise)
    want = torch.cat([g.reshape(-1) for g in grads.values()])
    assert torch.equal(res[0]["flat"], res[1]["flat"])                     # identical reduced gradients on every rank
    rel = (res[0]["flat"] - want).norm() / want.norm()
    assert rel < 1e-5, rel                                                 # equal up to summation order
    assert res[0]["scal"][0].item() == pytest.approx(out["loss"].item(), rel=1e-6)
    assert torch.equal(torch.cat([r["digits"] for r in res]), out["rec_num_digits"])


def test_shard_helpers():
    import air_b200 as ab
    t = torch.arange(24).reshape(6, 4)
    assert torch.equal(ab.dp.shard_rows(t, 1, 3), t[2:4])
    n = {"a": torch.arange(2 * 6 * 3).reshape(2, 6, 3)}
    assert torch.equal(ab.dp.shard_noise(n, 2, 3)["a"], n["a"][:, 4:6])
    with pytest.raises(ValueError):
        ab.dp.shard_rows(t, 0, 4)
    assert ab.dp.world_size() == 1 and ab.dp.rank() == 0
    g = torch.ones(5)
    assert ab.dp.allreduce_flat(g) is g
