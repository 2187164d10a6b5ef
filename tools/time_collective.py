"""Developer tool (torchrun, one rank per GPU): latency of the gradient exchange + SGD update of the data-parallel step in
isolation -- ncclAllReduce(bucket) + sgd_flat_kernel against the fused peer-memory kernel (fgrnn_sgd_allreduce_peer) -- as
CUDA-graph replays of 20 back-to-back updates, device-timed, max over ranks.  argv: [numel]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from kws_b200 import sharding, train_step  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20738 + 13 * 128 + 13
    cap_group = dist.new_group(backend="nccl")
    params = torch.zeros(n, device=dev)
    grads = torch.randn(n, device=dev)
    reduced = torch.empty(n, device=dev)
    peer = sharding.PeerReducer(n, dev, cap_group)
    peer.bucket.copy_(grads)
    INNER, REPS = 20, 20

    def nccl_update():
        dist.all_reduce(grads, group=cap_group)
        train_step.sgd_flat(params, grads, 1e-3, 1.0 / world)

    def peer_update():
        peer.step(params, 1e-3, reduced=reduced)

    def sgd_only():
        train_step.sgd_flat(params, grads, 1e-3, 1.0)

    res = {}
    for name, fn in (("sgd_flat alone", sgd_only), ("ncclAllReduce + sgd_flat", nccl_update), ("fused peer kernel", peer_update)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier(device_ids=[local])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(INNER):
                fn()
        g.replay()
        torch.cuda.synchronize()
        dist.barrier(device_ids=[local])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(REPS):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = torch.tensor([e0.elapsed_time(e1) * 1e3 / (INNER * REPS)], device=dev)
        dist.all_reduce(us, op=dist.ReduceOp.MAX)
        res[name] = float(us)
        grads.normal_()                                  # keep the sums finite
        del g
    peer.check()
    if rank == 0:
        print("COLLECTIVE world %d, %d floats (%.1f KB): " % (world, n, n * 4 / 1e3) + "; ".join("%s %.2f us" % kv for kv in res.items()), flush=True)
    dist.barrier(device_ids=[local])
    os._exit(0)


if __name__ == "__main__":
    main()
