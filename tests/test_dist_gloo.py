"""Host-side data-parallel logic on CPU: world size 2, gloo backend (no GPU).  Covers frame sharding, the
flat gradient bucket and its single all-reduce — the only exchange step of the path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from echoglad_b200 import dist as egdist
    r, _, w = egdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)  # identical replicas
    model = torch.nn.Sequential(torch.nn.Linear(16, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
    # two groups in backward order (last layer first): each is all-reduced from its post-accumulate hooks as soon as
    # its gradients exist (overlap path), the rest in all_reduce_mean()
    bucket = egdist.FlatGradBucket(model.parameters(), groups=[list(model[2].parameters()), list(model[0].parameters())])
    assert bucket.overlap
    frames = torch.randn(8, 16, generator=torch.Generator().manual_seed(1))  # the global batch, same on all ranks
    mine = egdist.shard_range(frames.shape[0], r, w)
    bucket.zero()
    loss = model(frames[mine.start:mine.stop]).square().mean()  # per-rank normaliser, as in the reference replicas
    loss.backward()
    bucket.all_reduce_mean()
    for p, v in zip(bucket.params, bucket.views):  # after the exchange every .grad IS its view of the flat buffer
        assert p.grad.data_ptr() == v.data_ptr()
    assert all(bucket._launched)
    # uneven loss normalisers: rank r holds V_r = r + 1 valid entries; scaled per-rank losses average to the global one
    v = torch.tensor(float(rank + 1))
    scale = egdist.global_normaliser_scale(v)
    assert abs(float(scale) - world * (rank + 1) / sum(range(1, world + 1))) < 1e-6
    torch.save({"flat": bucket.flat.clone(), "range": (mine.start, mine.stop)}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_equals_global_batch_gradient(tmp_path):
    from echoglad_b200 import dist as egdist
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    assert [o["range"] for o in outs] == [(0, 4), (4, 8)]
    assert torch.equal(outs[0]["flat"], outs[1]["flat"])  # every rank holds the same averaged gradient
    # equal shards => mean of per-rank mean-losses == global mean loss => same gradient as one rank on 8 frames
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
    bucket = egdist.FlatGradBucket(model.parameters(), groups=[list(model[2].parameters()), list(model[0].parameters())])
    assert not bucket.overlap  # single process: no hooks, one gather at the end
    frames = torch.randn(8, 16, generator=torch.Generator().manual_seed(1))
    bucket.zero()
    model(frames).square().mean().backward()
    bucket.all_reduce_mean()  # world size 1: gathers the fresh gradients into the flat buffer
    assert torch.allclose(bucket.flat, outs[0]["flat"], rtol=1e-5, atol=1e-7)


def test_shard_range_requires_equal_shards():
    from echoglad_b200 import dist as egdist
    assert list(egdist.shard_range(64, 3, 8)) == list(range(24, 32))
    with pytest.raises(ValueError):
        egdist.shard_range(10, 0, 4)
