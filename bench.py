#!/usr/bin/env python
"""bench.py -- FastGRNN sequences/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5]

Default workload (N=1): BASELINE config 2 -- FastGRNN KWS inference, input 32, hidden 128, T=99,
batch 8192 per GPU, full-rank W/U, fp32 state.  A "step" is one forward pass of the hot path over
one batch of synthetic MFCC tensors.  N>1 (launched by torch.distributed.run, one rank per GPU):
every rank runs its own 8192-row batch (weak scaling; no data-path collective for inference;
workload c3 adds the NCCL gradient all-reduce).

`value`     = whole-job sequences/s with inputs resident in HBM (device-timed, max over ranks).
`e2e`       = the same metric through kws_b200.streaming.HostPipeline with pinned HOST buffers,
              H2D of the inputs and D2H of all hidden states inside the timed region.
`roofline`  = dominant kernel (the persistent forward recurrence) against the measured HBM peak.
`cpu_baseline` / `--impl reference` = the CPU restatement of the reference rnn.py FastGRNN
              (oracle/, bit-identical to the reference in the build container) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: per-GPU batch, T, I, H, wRank, uRank, x dtype, mode
    "c2": dict(B=8192, T=99, I=32, H=128, wR=None, uR=None, x="f32", mode="infer",
               desc="FastGRNN KWS inference: input 32, hidden 128, T=99, batch 8192/GPU, full-rank, fp32 state"),
    "c3": dict(B=2048, T=99, I=32, H=128, wR=None, uR=None, x="f32", mode="train",
               desc="FastGRNN KWS training fwd+BPTT data-parallel: 2048 rows/GPU, hidden 128, T=99, NCCL grad allreduce"),
    "c4": dict(B=32768, T=99, I=32, H=256, wR=16, uR=32, x="f32", mode="infer",
               desc="Low-rank FastGRNN (wRank=16,uRank=32), hidden 256, T=99, batch 32768/GPU"),
    "c5": dict(B=8192, T=1000, I=32, H=128, wR=None, uR=None, x="bf16", mode="infer",
               desc="Long-sequence FastGRNN: T=1000, hidden 128, 8192 rows/GPU, bf16 inputs, fp32 state"),
}


def algorithmic_bytes_per_seq(w):
    """forward = x read once + every hidden state written once."""
    xb = 2 if w["x"] == "bf16" else 4
    fwd = xb * w["T"] * w["I"] + 4 * w["T"] * w["H"]
    if w["mode"] == "infer":
        return fwd
    # + backward: grad_h, h (for h_{t-1}) and x read once (z,c recomputed in the accounting)
    return fwd + 4 * w["T"] * w["H"] * 2 + xb * w["T"] * w["I"]


def flops_per_seq(w):
    I, H, T = w["I"], w["H"], w["T"]
    wf = 2 * I * H if w["wR"] is None else 2 * (I * w["wR"] + w["wR"] * H)
    uf = 2 * H * H if w["uR"] is None else 2 * (H * w["uR"] + w["uR"] * H)
    return T * (wf + uf)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region: an NVML polling thread (a sample
    every ~0.5 ms, so even a 3 ms region is covered); `nvidia-smi -lms` is the fallback when NVML is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASON_BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        import threading
        self.proc = None
        self.samples = []            # (sm_mhz, power_w, reasons bitmask)
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        except Exception:                         # noqa: BLE001 -- no NVML binding: fall back to nvidia-smi
            self._nvml = None
            try:
                self.proc = subprocess.Popen(
                    ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            except OSError:
                self.proc = None

    def _poll(self):
        n = self._nvml
        while not self._stop.is_set():
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
                try:
                    power = n.nvmlDeviceGetPowerUsage(self._h) / 1000.0
                except Exception:                 # noqa: BLE001
                    power = 0.0
                try:
                    reasons = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:                 # noqa: BLE001 -- older bindings
                    reasons = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                self.samples.append((mhz, power, reasons))
            except Exception:                     # noqa: BLE001
                pass
            time.sleep(0.0005)

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
            mask = 0
            for _, _, r in self.samples:
                mask |= r
            reasons = sorted(name for bit, name in self.REASON_BITS.items() if mask & bit)
            return {"sm_mhz": statistics.median(m for m, _, _ in self.samples), "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "power_w_max": max(p for _, p, _ in self.samples),
                    "samples": len(self.samples), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm), "source": "nvidia-smi"}


def make_params(w, device, layout="IH"):
    """Reference init (rnn.py:246-261) under torch.manual_seed(0), generated on the CPU."""
    from oracle.fastgrnn_oracle import init_params
    torch.manual_seed(0)
    p = init_params(w["I"], w["H"], w["wR"], w["uR"])
    return p, {k: v.to(device).contiguous() for k, v in p.tensors().items()}


class CpuReference:
    """The reference's CPU implementation of the path (restated in oracle/, bit-identical to rnn.py FastGRNN +
    BaseRNN) on a bounded sample of the workload.  Three ways of using the host cores are tried once each and the
    fastest is timed: torch intra-op threads over one 1024-row call (the reference as written), one thread, and one
    1024-row call per core on a thread pool (torch releases the GIL inside its kernels) -- the batch-sharded analogue
    of what the GPU arm does."""

    ROWS = 1024

    def __init__(self, w):
        from oracle import fastgrnn_oracle as O
        self.O, self.w = O, w
        self.cores = os.cpu_count() or 1
        torch.manual_seed(0)
        self.p = O.init_params(w["I"], w["H"], w["wR"], w["uR"])
        x = torch.randn(self.ROWS, w["T"], w["I"])
        if w["x"] == "bf16":
            x = x.bfloat16().float()   # the reference path is fp32-only
        self.x = x
        self.pool = None
        self.mode = None

    def _call(self, mode):
        """one step; returns the number of sequences it processed"""
        with torch.no_grad():
            if mode == "pool":
                list(self.pool.map(lambda _: self.O.unroll(self.x, self.p, None, True), range(self.cores)))
                return self.ROWS * self.cores
            self.O.unroll(self.x, self.p, None, True)
            return self.ROWS

    def _enter(self, mode):
        torch.set_num_threads(self.cores if mode == "intra" else 1)
        if mode == "pool" and self.pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self.pool = ThreadPoolExecutor(self.cores)

    def choose(self):
        best = None
        modes = ["intra", "single"] + (["pool"] if self.cores > 1 else [])
        for mode in modes:
            self._enter(mode)
            self._call(mode)                                  # warm-up
            t0 = time.perf_counter()
            n = self._call(mode)
            v = n / (time.perf_counter() - t0)
            if best is None or v > best[0]:
                best = (v, mode)
        self.mode = best[1]
        self._enter(self.mode)
        return self.mode

    def step(self):
        t0 = time.perf_counter()
        n = self._call(self.mode)
        return n, time.perf_counter() - t0

    def describe(self):
        how = {"intra": "one %d-row call with %d torch intra-op threads" % (self.ROWS, self.cores),
               "single": "one %d-row call on one thread" % self.ROWS,
               "pool": "%d concurrent %d-row calls, one per core (thread pool, 1 intra-op thread each)" % (self.cores, self.ROWS)}[self.mode]
        return ("%s per step of the %d-sequence workload, T=%d; fastest of {intra-op threads, single thread, per-core "
                "pool}; torch CPU restatement of rnn.py FastGRNN (bit-identical to the reference in the build container)"
                % (how, self.w["B"], self.w["T"]))

    def cores_used(self):
        return 1 if self.mode == "single" else self.cores


def cpu_reference_seq_per_s(w, seconds=10.0):
    ref = CpuReference(w)
    ref.choose()
    n_tot, t_tot, it = 0, 0.0, 0
    while True:
        n, dt = ref.step()
        n_tot += n; t_tot += dt; it += 1
        if (t_tot >= seconds and it >= 2) or it >= 1000:
            break
    return n_tot / t_tot, ref


def run_reference(args, w, rank, world):
    """--impl reference: the CPU reference arm. Rank 0 alone runs; other ranks exit 0."""
    if rank != 0:
        return
    ref = CpuReference(w)
    ref.choose()
    n_tot, t_tot = 0, 0.0
    for i in range(args.warmup + args.steps):
        n, dt = ref.step()
        if i >= args.warmup:
            n_tot += n; t_tot += dt
    value = n_tot / t_tot
    line = {
        "impl": "reference", "metric": "FastGRNN sequences/sec (fwd infer)", "value": value, "unit": "sequences/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "sample": ref.describe()},
        "cpu_baseline": {"value": value, "unit": "sequences/s", "cores": ref.cores_used(), "kind": "port",
                         "sample": ref.describe(), "host_cores": ref.cores},
        "e2e": {"value": value, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--layout", default="IH", choices=["IH", "HI"])
    ap.add_argument("--path", default="auto", choices=["auto", "generic", "smem", "tcgen05"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="training workloads: time the eager step, not the CUDA graph")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = dict(WORKLOADS[args.workload])

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    import torch.distributed as dist
    from kws_b200 import _lib, engine, sharding, streaming

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the FastGRNN path has no CPU fallback")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    if world != args.gpus and rank == 0:
        print("warning: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world), file=sys.stderr)

    force = {"auto": -1, "generic": 0, "smem": 1, "tcgen05": 2}[args.path]
    p_cpu, params = make_params(w, device)
    if args.layout == "HI":
        from oracle.fastgrnn_oracle import to_cuda_layout
        params = {k: v.to(device).contiguous() for k, v in to_cuda_layout(p_cpu).items()}
    B, T, I, H = w["B"], w["T"], w["I"], w["H"]
    xdt = torch.bfloat16 if w["x"] == "bf16" else torch.float32
    torch.manual_seed(1000 + rank)
    x_host = torch.randn(B, T, I).to(xdt).pin_memory()
    x = x_host.to(device)
    out = torch.empty(B, T, H, dtype=torch.float32, device=device)
    plan = engine.forward_plan(x, params, None, layout=args.layout, batch_first=True, force_path=force)

    train = w["mode"] == "train"
    if train:
        from kws_b200 import rnn as krnn
        torch.manual_seed(0)
        layer = krnn.FastGRNN(I, H, wRank=w["wR"], uRank=w["uR"], batch_first=False).to(device)
        head = torch.nn.Linear(H, 13).to(device)
        sharding.broadcast_parameters(list(layer.parameters()) + list(head.parameters())) if world > 1 else None
        plist = list(layer.cell.parameters()) + list(head.parameters())
        bucket = sharding.GradBucket(plist)
        opt = torch.optim.SGD(plist, lr=1e-3)
        x_tm = x.transpose(0, 1).contiguous()
        labels = torch.randint(0, 13, (B,), device=device)

    dbg = (lambda m: print("[bench rank %d] %s" % (rank, m), file=sys.stderr, flush=True)) if os.environ.get("FGRNN_BENCH_DEBUG") else (lambda m: None)

    def step_compute():                           # forward, loss head, BPTT: gradients land in the flat bucket
        bucket.zero()
        hs = layer(x_tm)
        logp = torch.nn.functional.log_softmax(head(hs[-1]), dim=1)      # model.py:228-230
        loss = torch.nn.functional.nll_loss(logp, labels)
        loss.backward()

    def step():
        if not train:
            engine.forward(x, params, None, layout=args.layout, batch_first=True, out=out, force_path=force)
            return
        step_compute()
        if world > 1:
            bucket.all_reduce_mean()
        opt.step()

    graphed = False
    run_step = step
    if train and not args.no_graph:
        try:                                      # the step as CUDA graphs (kws_b200/graphs.py)
            from kws_b200 import graphs
            if world == 1:
                cap = graphs.CapturedStep(step, warmup=args.warmup)
                run_step, graphed = cap, True
            else:
                # data parallel: the NCCL all-reduce stays an ordinary launch between two captured graphs (a collective
                # captured next to eager collectives on the same communicator hung on this pool's NCCL 2.28 / torch 2.11)
                for _ in range(args.warmup):
                    step()
                cap = graphs.CapturedStep(step_compute, warmup=1)
                cap_opt = graphs.CapturedStep(opt.step, warmup=1)
                n_lib = cap.launches + cap_opt.launches

                def run_step():
                    cap()
                    bucket.all_reduce_mean()
                    cap_opt()
                cap.launches = n_lib
                graphed = True
        except Exception as e:                    # noqa: BLE001 -- report and time the eager step instead
            print("bench.py: CUDA-graph capture failed (%s); timing the eager step" % (e,), file=sys.stderr)
            run_step = step

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(device)

    dbg("graphed=%s" % graphed)
    for _ in range(args.warmup):
        run_step()
    dbg("warmup done")
    barrier()
    dbg("barrier done")
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _lib.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        run_step()
        evs[i + 1].record()
    dbg("timed loop enqueued")
    barrier()
    dbg("timed loop done")
    launches = cap.launches * args.steps if graphed else _lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    total_ms = evs[0].elapsed_time(evs[-1])
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    rank_ms = [total_ms / args.steps]
    if world > 1:                                 # per-rank step times (diagnostic; the headline uses the max)
        tl = [torch.zeros(1, device=device) for _ in range(world)]
        dist.all_gather(tl, torch.tensor([total_ms / args.steps], device=device))
        rank_ms = [float(t) for t in tl]
    total_ms = sharding.max_over_ranks(total_ms, device)
    ms_per_step = total_ms / args.steps
    value = world * B * args.steps / (total_ms * 1e-3)

    # roofline of the dominant kernel (the forward recurrence; one launch per inference step)
    peak, peak_src = measured_peaks()
    kernel_ms = statistics.mean(step_ms) if not train else None
    roof = None
    if kernel_ms:
        abytes = algorithmic_bytes_per_seq(w) * B
        achieved = abytes / (kernel_ms * 1e-3) / 1e9
        prof = os.path.join(ROOT, "profiles", "traffic_%s_%s.json" % (args.workload, plan))
        traffic = None
        if os.path.isfile(prof):
            with open(prof) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "fgrnn forward recurrence (%s path)" % plan,
                "algorithmic_bytes_per_launch": abytes, "kernel_ms": kernel_ms, "peak_source": peak_src,
                "fp32_tflops": flops_per_seq(w) * B / (kernel_ms * 1e-3) / 1e12}

    # e2e through the host-buffer API
    e2e = None
    if not args.no_e2e and not train:
        out_host = torch.empty(B, T, H, dtype=torch.float32).pin_memory()
        pipe = streaming.HostPipeline(params, layout=args.layout, T=T, I=I, H=H, chunk_rows=1024, x_dtype=xdt, device=device)
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            pipe.run(x_host, out_host)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pipe.run(x_host, out_host)
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_s = sharding.max_over_ranks(e2e_s, device)
        e2e = {"value": world * B * e2e_steps / e2e_s, "unit": "sequences/s",
               "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
               "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps,
               "api": "kws_b200.streaming.HostPipeline.run(pinned x_host -> pinned out_host)"}
    elif train:
        # training e2e: inputs + labels from pinned host memory, loss read back each step
        xh = x_tm.cpu().pin_memory(); lh = labels.cpu().pin_memory()
        e2e_steps = max(3, min(args.steps, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            x_tm.copy_(xh, non_blocking=True); labels.copy_(lh, non_blocking=True)
            step()
            torch.cuda.synchronize(device)
        barrier()
        e2e_s = sharding.max_over_ranks(time.perf_counter() - t0, device)
        e2e = {"value": world * B * e2e_steps / e2e_s, "unit": "sequences/s",
               "h2d_bytes_per_step": xh.numel() * 4 + lh.numel() * 8, "d2h_bytes_per_step": 4,
               "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ref = cpu_reference_seq_per_s(w, seconds=args.cpu_seconds)
        cpu = {"value": v, "unit": "sequences/s", "cores": ref.cores_used(), "kind": "port", "sample": ref.describe(),
               "host_cores": ref.cores}

    if rank == 0:
        line = {
            "metric": "FastGRNN sequences/sec (%s)" % ("fwd+bwd train" if train else "fwd infer"),
            "value": value, "unit": "sequences/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "rank_ms_per_step": rank_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "name": args.workload, "per_gpu_batch": B, "global_batch": B * world,
                       "T": T, "input": I, "hidden": H, "wRank": w["wR"], "uRank": w["uR"], "x_dtype": w["x"],
                       "layout": "(B,T,F) contiguous" if not train else "(T,B,F) contiguous",
                       "weight_layout": args.layout, "kernel_path": plan, "cuda_graph": graphed,
                       "l2": "inputs+outputs per step = %.0f MB > 126 MB L2, no flush needed"
                             % ((algorithmic_bytes_per_seq(w) * B) / 1e6),
                       "parallelism": "batch-sharded x%d%s" % (world, ", NCCL grad all-reduce" if train else ", no collective")},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
