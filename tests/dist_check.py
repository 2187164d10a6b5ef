"""Multi-GPU correctness of the two sharded paths, run under torchrun with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py

1. batch-sharded inference: every rank's hidden states == the same rows of a single-GPU run, bit for bit;
2. data-parallel training: FastGRNN gradients in one flat bucket, summed with ONE NCCL all-reduce and divided by the
   world size == the gradients of a single-GPU run over the concatenated batch (rtol 1e-4 with the per-tensor atol floor
   of the north star), for every parameter;
3. the same step replayed as one CUDA graph with the all-reduce captured on its own communicator == the eager step;
4. the keyword spotter's step on the last state (kws_b200.train_step.LastStateTrainStep: fused head + loss kernel, BPTT from
   the last state's gradient, ONE all-reduce of the flat bucket, flat SGD with the 1 / world folded in): two data-parallel
   steps leave every rank with the parameters of two single-GPU steps over the concatenated batch.
5. the same step with the gradient exchange FUSED into the SGD kernel over NVLink peer memory (sharding.PeerReducer,
   fgrnn_sgd_allreduce_peer): parameters after eager steps and after CUDA-graph replays of the whole step == the single-GPU
   run within the tolerance, p.grad == the NCCL sum within rounding, and every rank holds the very same bits.
Prints "DIST_CHECK ok" on rank 0; any mismatch raises."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from kws_b200 import graphs, rnn, sharding, train_step  # noqa: E402


def grad_ratio(got, ref, rtol=1e-4):
    atol = rtol * float(ref.abs().max())
    return float(((got - ref).abs() / (atol + rtol * ref.abs())).max())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    rows, T, I, H = 96, 17, 32, 128
    Bg = rows * world
    g = torch.Generator().manual_seed(7)
    x_all = torch.randn(T, Bg, I, generator=g).to(dev)
    go_all = (torch.randn(T, Bg, H, generator=g) / Bg).to(dev)
    b, e = sharding.shard_bounds(Bg, world, rank)

    def make_layer():
        torch.manual_seed(3)
        return rnn.FastGRNN(I, H).to(dev)

    # ---- 1. inference: shards == slices of the full run, bitwise
    layer = make_layer()
    with torch.no_grad():
        full = layer(x_all)
        mine = layer(x_all[:, b:e].contiguous())
    assert torch.equal(mine, full[:, b:e]), "rank %d: sharded inference differs from the single-GPU run" % rank
    gathered = sharding.gather_states(mine, 1, Bg)
    assert torch.equal(gathered, full), "rank %d: all-gathered states differ" % rank

    # ---- 2. data-parallel gradients
    ref_layer = make_layer()
    ref_layer(x_all).backward(go_all)                                   # single GPU, concatenated batch
    ref = {k: v.grad.clone() for k, v in ref_layer.cell.named_parameters()}
    layer = make_layer()
    sharding.broadcast_parameters(list(layer.cell.parameters()))
    bucket = sharding.GradBucket(layer.cell.parameters())
    xs, gs = x_all[:, b:e].contiguous(), go_all[:, b:e].contiguous() * world     # local mean-loss scaling

    def step():
        bucket.zero()
        layer(xs).backward(gs)
        bucket.all_reduce_mean()
    step()
    torch.cuda.synchronize()
    worst = 0.0
    for k, p in layer.cell.named_parameters():
        r = grad_ratio(p.grad, ref[k])
        worst = max(worst, r)
        assert r <= 1.0, "rank %d: all-reduced gradient of %s off by %.3f x tolerance" % (rank, k, r)
    eager = bucket.flat.clone()

    # ---- 3. the step as one CUDA graph, all-reduce captured on a communicator of its own
    cap_group = dist.new_group(backend="nccl")

    def step_g():
        bucket.zero()
        layer(xs).backward(gs)
        bucket.all_reduce_mean(group=cap_group)
    for _ in range(3):
        step_g()
    torch.cuda.synchronize()
    cap = graphs.CapturedStep(step_g, warmup=1)
    bucket.flat.fill_(float("nan"))
    cap()
    torch.cuda.synchronize()
    assert torch.equal(bucket.flat, eager), "rank %d: captured data-parallel step differs from the eager one" % rank
    del cap                                   # a live graph that holds a captured collective blocks the communicator teardown
    torch.cuda.synchronize()

    # ---- 4. the fused last-state training step, data parallel vs one GPU on the concatenated batch
    def make_model():
        torch.manual_seed(5)
        return rnn.FastGRNN(I, H).to(dev), torch.nn.Linear(H, 13).to(dev)
    labels_all = torch.randint(0, 13, (Bg,), generator=g).to(dev)
    x2_all = torch.randn(T, Bg, I, generator=g).to(dev)
    l_ref, h_ref = make_model()
    ref_step = train_step.LastStateTrainStep(l_ref, h_ref, 0.05, data_parallel=False)
    l_dp, h_dp = make_model()
    sharding.broadcast_parameters(list(l_dp.cell.parameters()) + list(h_dp.parameters()))
    dp_step = train_step.LastStateTrainStep(l_dp, h_dp, 0.05, data_parallel=True, collective="nccl")
    assert dp_step.world == world and dp_step.collective == "nccl"
    for xa in (x_all, x2_all):
        loss_ref = ref_step(xa, labels_all)
        loss_dp = dp_step(xa[:, b:e].contiguous(), labels_all[b:e].contiguous())
        lsum = loss_dp.clone()
        dist.all_reduce(lsum)
        assert abs(float(lsum) / world - float(loss_ref)) <= 1e-5 * abs(float(loss_ref)) + 1e-6, "rank %d: mean of the rank losses differs" % rank
    torch.cuda.synchronize()
    worst_p = float(((dp_step.flat_params - ref_step.flat_params).abs() / (1e-6 + 1e-5 * ref_step.flat_params.abs())).max())
    assert worst_p <= 1.0, "rank %d: parameters after two data-parallel fused steps off by %.3f x (rtol 1e-5, atol 1e-6)" % (rank, worst_p)

    # ---- 5. the gradient exchange fused into the SGD kernel over NVLink peer memory
    l_ref2, h_ref2 = make_model()
    ref2 = train_step.LastStateTrainStep(l_ref2, h_ref2, 0.05, data_parallel=False)
    l_pp, h_pp = make_model()
    sharding.broadcast_parameters(list(l_pp.cell.parameters()) + list(h_pp.parameters()))
    pp_step = train_step.LastStateTrainStep(l_pp, h_pp, 0.05, data_parallel=True, collective="peer")
    assert pp_step.collective == "peer" and pp_step.peer is not None
    l_nc, h_nc = make_model()
    sharding.broadcast_parameters(list(l_nc.cell.parameters()) + list(h_nc.parameters()))
    nc_step = train_step.LastStateTrainStep(l_nc, h_nc, 0.05, data_parallel=True, collective="nccl")
    xs2, ls2 = x2_all[:, b:e].contiguous(), labels_all[b:e].contiguous()
    for xa in (x_all, x2_all):                         # two eager steps
        ref2(xa, labels_all)
        pp_step(xa[:, b:e].contiguous(), labels_all[b:e].contiguous())
        nc_step(xa[:, b:e].contiguous(), labels_all[b:e].contiguous())
    torch.cuda.synchronize()
    pp_step.peer.check()
    worst_g = float(((pp_step.flat_grads - nc_step.flat_grads).abs() / (1e-7 + 1e-5 * nc_step.flat_grads.abs())).max())
    assert worst_g <= 1.0, "rank %d: peer-reduced gradient sum differs from the NCCL sum by %.3f x (rtol 1e-5)" % (rank, worst_g)
    cap5 = graphs.CapturedStep(lambda: pp_step(xs2, ls2), warmup=1)      # the whole step incl. the peer exchange as ONE graph
    torch.cuda.synchronize()
    for _ in range(3):
        cap5()
    torch.cuda.synchronize()
    pp_step.peer.check()
    for _ in range(1 + 3):                             # the warm-up call of CapturedStep and three replays (capture itself runs nothing)
        ref2(x2_all, labels_all)
    torch.cuda.synchronize()
    worst_pp = float(((pp_step.flat_params - ref2.flat_params).abs() / (1e-6 + 1e-5 * ref2.flat_params.abs())).max())
    assert worst_pp <= 1.0, "rank %d: parameters after six peer-fused steps off by %.3f x (rtol 1e-5, atol 1e-6)" % (rank, worst_pp)
    allp = [torch.empty_like(pp_step.flat_params) for _ in range(world)]
    dist.all_gather(allp, pp_step.flat_params)
    assert all(torch.equal(allp[0], q) for q in allp), "rank %d: the replicas' parameters are not bit-identical" % rank
    del cap5
    dist.barrier(device_ids=[local])
    if rank == 0:
        print("DIST_CHECK ok: world %d, sharded inference bitwise, all-reduced gradients at %.3f of tolerance, captured step bitwise, "
              "fused last-state step parameters at %.3f of (1e-5, 1e-6); peer-memory all-reduce + SGD kernel: gradient sum at %.3f of "
              "rtol 1e-5 against NCCL, parameters after 6 steps (3 through one CUDA graph) at %.3f, replicas bit-identical"
              % (world, worst, worst_p, worst_g, worst_pp), flush=True)
    sys.stdout.flush()
    os._exit(0)                               # skip the process-group teardown: nothing left to do, and it must never hang a test


if __name__ == "__main__":
    main()
