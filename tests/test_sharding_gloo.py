"""CPU, world_size 2 over gloo: the multi-GPU host logic (batch sharding, flat gradient bucket,
all-reduce arithmetic, ragged slices, parameter broadcast, max-over-ranks timing)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kws_b200 import sharding


def test_shard_bounds_cover_exactly():
    for total in (0, 1, 7, 8192, 8193, 65536):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def test_bucket_views_follow_c_abi_order():
    from kws_b200 import engine
    p = {"W": torch.zeros(32, 128), "U": torch.zeros(128, 128), "bias_gate": torch.zeros(1, 128),
         "bias_update": torch.zeros(1, 128), "zeta": torch.zeros(1, 1), "nu": torch.zeros(1, 1)}
    assert engine.grad_bucket_numel(p) == 20738
    flat = torch.arange(20738, dtype=torch.float32)
    v = engine.bucket_views(flat, p)
    assert v["W"][0, 0] == 0 and v["U"][0, 0] == 4096 and v["bias_gate"][0, 0] == 20480 and v["nu"][0, 0] == 20737
    with pytest.raises(RuntimeError):
        engine.bucket_views(flat[:100], p)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, total_rows):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                      # same "global batch" and weights on every rank
        x = torch.randn(total_rows, 6)
        y = torch.randn(total_rows, 3)
        lin = torch.nn.Linear(6, 3)
        if rank != 0:
            with torch.no_grad():
                for p in lin.parameters():
                    p.add_(1.0)                   # diverge on purpose; broadcast must repair it
        sharding.broadcast_parameters(list(lin.parameters()), src=0)
        ref = torch.nn.Linear(6, 3)
        torch.manual_seed(0); torch.randn(total_rows, 6); torch.randn(total_rows, 3)
        ref = torch.nn.Linear(6, 3)               # same RNG position as `lin` on rank 0
        for a, b in zip(lin.parameters(), ref.parameters()):
            assert torch.equal(a, b)
        # data-parallel step on this rank's contiguous slice
        xs, ys = sharding.shard_batch(x, 0), sharding.shard_batch(y, 0)
        b, e = sharding.shard_bounds(total_rows, world, rank)
        assert xs.shape[0] == e - b and torch.equal(xs, x[b:e])
        bucket = sharding.GradBucket(lin.parameters())
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
        bucket.zero()
        ((lin(xs) - ys) ** 2).mean().backward()
        bucket.all_reduce_weighted(xs.shape[0], total_rows)
        ((ref(x) - y) ** 2).mean().backward()     # single-process gradient on the concatenated batch
        for p, q in zip(lin.parameters(), ref.parameters()):
            assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-6)
        if total_rows % world == 0:               # equal slices: plain mean of the per-rank grads
            bucket.zero()
            ((lin(xs) - ys) ** 2).mean().backward()
            bucket.all_reduce_mean()
            for p, q in zip(lin.parameters(), ref.parameters()):
                assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-6)
        # sharded outputs gather back to the global batch (ragged last slice included)
        full = sharding.gather_states(xs * 2.0, 0, total_rows)
        assert torch.equal(full, x * 2.0)
        assert sharding.max_over_ranks(float(rank + 1), torch.device("cpu")) == float(world)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total_rows", [16, 17])
def test_data_parallel_logic_world2_gloo(total_rows):
    mp.spawn(_worker, args=(2, _free_port(), total_rows), nprocs=2, join=True)


def _peer_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # No GPU here: fgrnn_peer_alloc fails on every rank.  Construction is a collective, so each rank must carry its error
        # through both object all-gathers and ALL ranks must raise the same RuntimeError -- nobody may be left waiting
        with pytest.raises(RuntimeError, match="peer memory is not available"):
            sharding.PeerReducer(1000, torch.device("cuda", rank), None)
        with pytest.raises(RuntimeError, match="needs a CUDA device"):
            sharding.PeerReducer(1000, torch.device("cpu"), None)
        t = torch.tensor([float(rank)])
        dist.all_reduce(t)                         # the group is still usable afterwards
        assert float(t) == float(sum(range(world)))
    finally:
        dist.destroy_process_group()


def test_peer_reducer_fails_together_without_peer_memory_world2_gloo():
    mp.spawn(_peer_worker, args=(2, _free_port()), nprocs=2, join=True)
