#!/usr/bin/env python
"""bench.py -- FastGRNN sequences/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5|m1|m2] [--no-extra]

Default workload (N=1): BASELINE config 2 -- FastGRNN KWS inference, input 32, hidden 128, T=99,
batch 8192 per GPU, full-rank W/U, fp32 state.  A "step" is one forward pass of the hot path over
one batch of synthetic MFCC tensors.  N>1 (launched by torch.distributed.run, one rank per GPU):
every rank runs its own 8192-row batch (weak scaling; no data-path collective for inference;
workload c3 adds the gradient all-reduce: our fused peer-memory all-reduce + SGD kernel, or NCCL with --collective nccl).

`value`     = whole-job sequences/s with inputs resident in HBM (device-timed, max over ranks).
`e2e`       = the same metric through kws_b200.streaming.HostPipeline with pinned HOST buffers,
              H2D of the inputs and D2H of all hidden states inside the timed region.
`roofline`  = dominant kernel (the persistent forward recurrence) against the measured HBM peak.
`cpu_baseline` / `--impl reference` = the CPU restatement of the reference rnn.py FastGRNN
              (oracle/, bit-identical to the reference in the build container) on the host cores.
`extra`     = short runs of the other BASELINE configurations at the same N (c3: data-parallel training step with the
              gradient all-reduce -- `collective` says which one ran --, c4: low rank H=256, c5: T=1000 bf16, m1/m2: the two layers of the reference's
              default model), each with its own value / ms_per_step / roofline, so that the scaling record covers the
              path that has a collective.
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
    # the reference's default model (trainingConfig.py:9-30): 64 delta-MFCC features -> 256 -> 128
    "m1": dict(B=8192, T=99, I=64, H=256, wR=None, uR=None, x="f32", mode="infer",
               desc="Default-model layer 1: input 64, hidden 256, T=99, batch 8192/GPU, full-rank, fp32 state"),
    "m2": dict(B=8192, T=99, I=256, H=128, wR=None, uR=None, x="f32", mode="infer",
               desc="Default-model layer 2: input 256, hidden 128, T=99, batch 8192/GPU, full-rank, fp32 state"),
}
EXTRA_STEPS = {"c2": 20, "c3": 30, "c4": 6, "c5": 6, "m1": 10, "m2": 10}


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
            except Exception as e:                # noqa: BLE001 -- keep polling, remember why a poll failed
                self.error = "%s: %s" % (type(e).__name__, e)
            time.sleep(0.0001)

    def begin(self, timeout=0.5):
        """Call right before the timed region: waits until the polling thread is up (first samples taken), then drops
        what it has seen so far, so that every sample kept lies inside the region."""
        if self._thread is not None:
            t0 = time.perf_counter()
            while len(self.samples) < 2 and time.perf_counter() - t0 < timeout:
                time.sleep(0.001)
            self._skip = len(self.samples)

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
            self.samples = self.samples[getattr(self, "_skip", 0):]
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"], "error": getattr(self, "error", None)}
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


def init_params_reference(I, H, wR, uR):
    """The reference's initialisation (rnn.py:246-261): matrices 0.1*randn in the order W|W1,W2 then U|U1,U2, both
    biases ones, zeta = 1, nu = -4; FastGRNNCell layout ([I,H] / [H,H] ...).  Drawn on the CPU under the caller's seed."""
    p = {}
    if wR is None:
        p["W"] = 0.1 * torch.randn(I, H)
    else:
        p["W1"] = 0.1 * torch.randn(I, wR); p["W2"] = 0.1 * torch.randn(wR, H)
    if uR is None:
        p["U"] = 0.1 * torch.randn(H, H)
    else:
        p["U1"] = 0.1 * torch.randn(H, uR); p["U2"] = 0.1 * torch.randn(uR, H)
    p["bias_gate"] = torch.ones(1, H); p["bias_update"] = torch.ones(1, H)
    p["zeta"] = torch.ones(1, 1); p["nu"] = -4.0 * torch.ones(1, 1)
    return p


def make_params(w, device, layout="IH"):
    torch.manual_seed(0)
    p = init_params_reference(w["I"], w["H"], w["wR"], w["uR"])
    if layout == "HI":                                 # FastGRNNCUDA stores every matrix transposed (rnn.py:782-805)
        p = {k: (v.t().contiguous() if v.dim() == 2 and k[0] in "WU" else v) for k, v in p.items()}
    return {k: v.to(device).contiguous() for k, v in p.items()}


def config_of(w, name, world):
    """Identical in both arms (--impl ours / reference): what is measured, not how."""
    train = w["mode"] == "train"
    return {"workload": w["desc"], "name": name, "per_gpu_batch": w["B"], "global_batch": w["B"] * world,
            "T": w["T"], "input": w["I"], "hidden": w["H"], "wRank": w["wR"], "uRank": w["uR"], "x_dtype": w["x"],
            "layout": "(T,B,F) contiguous" if train else "(B,T,F) contiguous",
            "l2": "inputs+outputs per step = %.0f MB > 126 MB L2, no flush needed" % ((algorithmic_bytes_per_seq(w) * w["B"]) / 1e6),
            "parallelism": "batch-sharded x%d%s" % (world, ", gradient all-reduce" if train else ", no collective")}


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
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        # time the CPU path needs for ONE step of the workload (per_gpu_batch sequences), from the measured rate
        "ms_per_step": 1e3 * w["B"] / value, "sample_ms_per_step": 1e3 * t_tot / max(args.steps, 1),
        "sample_sequences_per_step": n_tot // max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(w, args.workload, world),
        "cpu_baseline": {"value": value, "unit": "sequences/s", "cores": ref.cores_used(), "kind": "port",
                         "sample": ref.describe(), "host_cores": ref.cores},
        "e2e": {"value": value, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as local to the GPU, BEFORE any pinned host buffer is allocated, so that
    the staging memory of the host-buffer API is first-touched on the GPU's NUMA node (round 1: eight ranks on node 0 made
    e2e scale 0.21 at 8 GPUs)."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i * 64 + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info = {"bound": True, "cpus": "%d-%d (%d)" % (allowed[0], allowed[-1], len(allowed))}
        else:
            info["why"] = "NVML affinity mask does not intersect the allowed CPUs"
    except Exception as e:                       # noqa: BLE001 -- affinity is an optimisation, never a failure
        info["why"] = "%s: %s" % (type(e).__name__, e)
    return info


class Bench:
    """One process = one rank = one GPU.  run(name, steps) measures one workload and returns its record."""

    def __init__(self, args, rank, world, local_rank):
        import torch.distributed as dist
        from kws_b200 import _lib, engine, sharding, streaming
        self.dist, self._lib, self.engine, self.sharding, self.streaming = dist, _lib, engine, sharding, streaming
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        self.device = torch.device("cuda", local_rank)
        self.force = {"auto": -1, "generic": 0, "smem": 1, "tcgen05": 2}[args.path]
        self.peak, self.peak_src = measured_peaks()
        self.cap_group = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local_rank])
        torch.cuda.synchronize(self.device)

    def run(self, name, steps, warmup, *, full):
        """full=True: the headline record (e2e, cpu baseline, clock record); False: a short `extra` record."""
        args, dev, world, rank = self.args, self.device, self.world, self.rank
        engine, sharding, _lib = self.engine, self.sharding, self._lib
        w = dict(WORKLOADS[name])
        B, T, I, H = w["B"], w["T"], w["I"], w["H"]
        train = w["mode"] == "train"
        params = make_params(w, dev, args.layout)
        xdt = torch.bfloat16 if w["x"] == "bf16" else torch.float32
        torch.manual_seed(1000 + rank)
        x_host = torch.randn(B, T, I).to(xdt).pin_memory()
        x = x_host.to(dev)
        out = torch.empty(B, T, H, dtype=torch.float32, device=dev)
        plan = engine.forward_plan(x, params, None, layout=args.layout, batch_first=True, force_path=self.force)
        graphed, cap = False, None
        self.collective = None

        if train:
            from kws_b200 import graphs, rnn as krnn, train_step
            torch.manual_seed(0)
            layer = krnn.FastGRNN(I, H, wRank=w["wR"], uRank=w["uR"], batch_first=False).to(dev)
            head = torch.nn.Linear(H, 13).to(dev)
            plist = list(layer.cell.parameters()) + list(head.parameters())
            if world > 1:
                sharding.broadcast_parameters(plist)
            x_tm = x.transpose(0, 1).contiguous()
            labels = torch.randint(0, 13, (B,), device=dev)
            fused = args.train_api == "fused"
            if world > 1 and self.cap_group is None and not args.no_graph:
                # the captured collective gets a communicator of its own: the eager barrier / timing collectives of the harness
                # stay on the default one (a collective captured next to eager collectives on the SAME communicator hung in round 1)
                self.cap_group = self.dist.new_group(backend="nccl")
            if fused:
                # kws_b200.train_step: recurrence, fused head+loss kernel, BPTT from the last state's gradient, then (one GPU) flat
                # SGD or (data parallel) the fused all-reduce + SGD kernel over NVLink peer memory; --collective nccl = ncclAllReduce + SGD
                stepper = train_step.LastStateTrainStep(layer, head, 1e-3, group=self.cap_group, data_parallel=world > 1,
                                                        collective=args.collective)
                self.collective = stepper.collective

                def step_compute():
                    stepper.compute(x_tm, labels)

                def step():
                    step_compute()
                    stepper.update()
                opt_step = stepper.update
            else:
                # the module API through autograd, as the unchanged trainClassifier.py drives it
                bucket = sharding.GradBucket(plist)
                opt = torch.optim.SGD(plist, lr=1e-3)
                opt_step = opt.step

                def step_compute():                   # forward, loss head, BPTT: gradients land in the flat bucket
                    bucket.zero()
                    hs = layer(x_tm)
                    logp = torch.nn.functional.log_softmax(head(hs[-1]), dim=1)      # model.py:228-230
                    loss = torch.nn.functional.nll_loss(logp, labels)
                    loss.backward()

                def step():
                    step_compute()
                    if world > 1:
                        bucket.all_reduce_mean(group=self.cap_group)
                    opt.step()

            run_step = step
            if not args.no_graph:
                try:
                    for _ in range(max(warmup, 3)):
                        step()                        # also initialises the communicator outside of any capture
                    torch.cuda.synchronize(dev)
                    mode = os.environ.get("FGRNN_BENCH_DP_GRAPH", "one" if world > 1 else "one")
                    if world == 1 or mode == "one":
                        cap = graphs.CapturedStep(step, warmup=1)          # whole step incl. ncclAllReduce = ONE graph
                        run_step, graphed = cap, True
                    else:
                        cap = graphs.CapturedStep(step_compute, warmup=1)
                        cap_opt = graphs.CapturedStep(opt_step, warmup=1)
                        cap.launches += cap_opt.launches

                        def run_step():
                            cap()
                            if fused and stepper.peer is None:
                                self.dist.all_reduce(stepper.flat_grads, group=self.cap_group)
                            elif not fused:
                                bucket.all_reduce_mean(group=self.cap_group)
                            cap_opt()
                        graphed = True
                except Exception as e:                # noqa: BLE001 -- report and time the eager step instead
                    print("bench.py: CUDA-graph capture failed (%s); timing the eager step" % (e,), file=sys.stderr)
                    run_step, graphed, cap = step, False, None
        else:
            def run_step():
                engine.forward(x, params, None, layout=args.layout, batch_first=True, out=out, force_path=self.force)

        for _ in range(warmup):
            run_step()
        self.barrier()
        sampler = ClockSampler(self.local_rank) if rank == 0 else None
        old_switch = sys.getswitchinterval()
        sys.setswitchinterval(1e-4)                   # let the NVML polling thread run inside millisecond-long regions
        if sampler:
            sampler.begin()
        self.barrier()                                # every rank: the sampler's start-up must not skew rank 0
        launches0 = _lib.launch_count()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record()
        for i in range(steps):
            run_step()
            evs[i + 1].record()
        self.barrier()
        sys.setswitchinterval(old_switch)
        launches = cap.launches * steps if graphed else _lib.launch_count() - launches0
        if sampler and len(sampler.samples) - getattr(sampler, "_skip", 0) < 3:
            # one NVML poll takes longer than this timed region (a few ms): repeat the same steps, untimed, for 0.3 s right
            # after it and sample the clocks there -- said so in the record
            t_end = time.perf_counter() + 0.3
            while time.perf_counter() < t_end:
                for _ in range(steps):
                    run_step()
                torch.cuda.synchronize(dev)
            sampler.note = "timed region shorter than one NVML poll: sampled during an immediate 0.3 s repetition of the same steps"
        clocks = sampler.stop() if sampler else None
        if clocks is not None and getattr(sampler, "note", None):
            clocks["note"] = sampler.note
        self.barrier()
        total_ms = evs[0].elapsed_time(evs[-1])
        step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        rank_ms = [total_ms / steps]
        if world > 1:                                 # per-rank step times (diagnostic; the headline uses the max)
            tl = [torch.zeros(1, device=dev) for _ in range(world)]
            self.dist.all_gather(tl, torch.tensor([total_ms / steps], device=dev))
            rank_ms = [float(t) for t in tl]
        total_ms = sharding.max_over_ranks(total_ms, dev)
        ms_per_step = total_ms / steps
        value = world * B * steps / (total_ms * 1e-3)

        # roofline: algorithmic bytes of one step over the measured duration of the step's kernels
        kernel_ms = statistics.mean(step_ms)
        abytes = algorithmic_bytes_per_seq(w) * B
        achieved = abytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "traffic_%s_%s.json" % (name, plan))
        if os.path.isfile(prof):
            with open(prof) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        kernel = (("forward recurrence + fused head/loss kernel + BPTT from the last state's gradient (reverse recurrence, "
                   "contraction, reduce) + flat SGD, one CUDA graph" if args.train_api == "fused" else
                   "module API through autograd: forward recurrence + loss head + BPTT + SGD, one CUDA graph")
                  if train else "fgrnn forward (%s path)" % plan)
        roof = {"bound": "hbm", "achieved": achieved, "peak": self.peak, "unit": "GB/s", "frac": achieved / self.peak,
                "traffic": traffic, "kernel": kernel, "algorithmic_bytes_per_launch": abytes, "kernel_ms": kernel_ms,
                "peak_source": self.peak_src, "fp32_tflops": flops_per_seq(w) * B / (kernel_ms * 1e-3) / 1e12 if not train else None}

        rec = {"metric": "FastGRNN sequences/sec (%s)" % ("fwd+bwd train" if train else "fwd infer"),
               "value": value, "unit": "sequences/s", "ms_per_step": ms_per_step, "rank_ms_per_step": rank_ms, "steps": steps,
               "kernel_path": plan, "cuda_graph": graphed, "roofline": roof, "gpu_launches": int(launches), "clocks": clocks,
               "config": dict(config_of(w, name, world), weight_layout=args.layout, kernel_path=plan, cuda_graph=graphed)}
        if train:
            # which collective ran the gradient exchange: "peer" = our fused all-reduce + SGD kernel over NVLink peer memory
            rec["collective"] = self.collective if (args.train_api == "fused" and world > 1) else ("nccl" if world > 1 else "none")
            if args.train_api == "fused" and stepper.peer is not None:
                stepper.peer.check()                  # raises if a peer failed to arrive in any step
        if not full:
            del x, out, x_host
            torch.cuda.empty_cache()
            return rec

        # ---- e2e through the host-buffer API: inputs from pinned host memory, results back in pinned host memory
        e2e = None
        if not args.no_e2e and not train:
            def time_pipe(last_only):
                oh = (torch.empty(B, H, dtype=torch.float32) if last_only else torch.empty(B, T, H, dtype=torch.float32)).pin_memory()
                pipe = self.streaming.HostPipeline(params, layout=args.layout, T=T, I=I, H=H, chunk_rows=args.chunk_rows,
                                                   x_dtype=xdt, device=dev, last_state_only=last_only)
                n = max(3, min(steps, 10))
                for _ in range(2):
                    pipe.run(x_host, oh)
                self.barrier()
                t0 = time.perf_counter()
                for _ in range(n):
                    pipe.run(x_host, oh)
                self.barrier()
                dt = sharding.max_over_ranks(time.perf_counter() - t0, dev)
                return {"value": world * B * n / dt, "unit": "sequences/s", "h2d_bytes_per_step": pipe.h2d_bytes,
                        "d2h_bytes_per_step": pipe.d2h_bytes, "ms_per_step": 1e3 * dt / n, "steps": n,
                        "host_gbs": (pipe.h2d_bytes + pipe.d2h_bytes) * world * n / dt / 1e9}
            e2e = time_pipe(False)
            # what the host side can take: a plain pinned D2H copy of the same result, all ranks at once
            oh = torch.empty(B, T, H, dtype=torch.float32).pin_memory()
            oh.copy_(out, non_blocking=True); self.barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                oh.copy_(out, non_blocking=True)
            self.barrier()
            dt = sharding.max_over_ranks(time.perf_counter() - t0, dev)
            e2e["pinned_d2h_copy_gbs_per_gpu"] = 3 * oh.numel() * 4 / dt / 1e9
            e2e["d2h_gbs_per_gpu"] = e2e["d2h_bytes_per_step"] / (e2e["ms_per_step"] * 1e-3) / 1e9
            del oh
            e2e["api"] = "kws_b200.streaming.HostPipeline.run(pinned x_host -> pinned out_host), all T hidden states returned (the operator contract)"
            e2e["last_state_only"] = time_pipe(True)
            e2e["last_state_only"]["api"] = ("HostPipeline(last_state_only=True): h_T only, what model.py:228 consumes "
                                             "(hidden2keyword(model_output[-1]))")
        elif train:
            xh = x_tm.cpu().pin_memory(); lh = labels.cpu().pin_memory()
            n = max(3, min(steps, 10))
            self.barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                x_tm.copy_(xh, non_blocking=True); labels.copy_(lh, non_blocking=True)
                run_step()
                torch.cuda.synchronize(dev)
            self.barrier()
            dt = sharding.max_over_ranks(time.perf_counter() - t0, dev)
            e2e = {"value": world * B * n / dt, "unit": "sequences/s", "h2d_bytes_per_step": xh.numel() * 4 + lh.numel() * 8,
                   "d2h_bytes_per_step": 4, "ms_per_step": 1e3 * dt / n, "steps": n}
        rec["e2e"] = e2e
        return rec


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
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other configurations")
    ap.add_argument("--no-graph", action="store_true", help="training workloads: time the eager step, not the CUDA graph")
    ap.add_argument("--train-api", default="fused", choices=["fused", "module"],
                    help="training workloads: kws_b200.train_step (default) or the module API through autograd")
    ap.add_argument("--collective", default="auto", choices=["auto", "peer", "nccl"],
                    help="data-parallel training: fused all-reduce + SGD kernel over NVLink peer memory, or ncclAllReduce + SGD")
    ap.add_argument("--chunk-rows", type=int, default=1024, help="rows per pipeline chunk of the host-buffer API")
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

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the FastGRNN path has no CPU fallback")
    numa = bind_to_gpu_numa_node(local_rank)
    import torch.distributed as dist
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    if world != args.gpus and rank == 0:
        print("warning: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world), file=sys.stderr)

    b = Bench(args, rank, world, local_rank)
    rec = b.run(args.workload, args.steps, args.warmup, full=True)
    extra = {}
    if not args.no_extra:
        for name in ("c3", "c4", "c5", "m1", "m2", "c2"):
            if name == args.workload:
                continue
            try:
                r = b.run(name, EXTRA_STEPS[name], 3, full=False)
                extra[name] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "rank_ms_per_step", "steps", "kernel_path",
                                                 "cuda_graph", "gpu_launches", "clocks")}
                extra[name]["roofline"] = {k: r["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "kernel", "kernel_ms")}
                extra[name]["workload"] = r["config"]["workload"]
                if "collective" in r:
                    extra[name]["collective"] = r["collective"]
            except Exception as e:                   # noqa: BLE001 -- an extra must never take the headline down
                extra[name] = {"error": "%s: %s" % (type(e).__name__, e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ref = cpu_reference_seq_per_s(w, seconds=args.cpu_seconds)
        cpu = {"value": v, "unit": "sequences/s", "cores": ref.cores_used(), "kind": "port", "sample": ref.describe(),
               "host_cores": ref.cores}

    if rank == 0:
        line = {
            "metric": rec["metric"], "value": rec["value"], "unit": "sequences/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "rank_ms_per_step": rec["rank_ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(w, args.workload, world),
            "impl_detail": {"weight_layout": args.layout, "kernel_path": rec["kernel_path"], "cuda_graph": rec["cuda_graph"],
                            "numa": numa, "e2e_chunk_rows": args.chunk_rows, "collective": rec.get("collective")},
            "roofline": rec["roofline"], "cpu_baseline": cpu, "e2e": rec.get("e2e"), "gpu_launches": rec["gpu_launches"],
            "clocks": rec["clocks"], "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
