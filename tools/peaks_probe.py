"""Developer tool: the derived peaks BASELINE.md section 3 asks for, measured on this box -- cuBLAS TF32 and fp32 (FFMA)
GEMM throughput (8192^3, best of 10 and a 2 s sustained loop), next to the driver's MEASURED_PEAKS.json."""
import json, os, sys, time
import torch
dev = torch.device("cuda:0")
N = 8192
a = torch.randn(N, N, device=dev); b = torch.randn(N, N, device=dev)
def measure(allow_tf32):
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    torch.backends.cudnn.allow_tf32 = allow_tf32
    for _ in range(3): a @ b
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = max(best, 2 * N ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0; t0 = time.time(); e0.record()
    while time.time() - t0 < 2.0:
        for _ in range(5): a @ b
        n += 5; torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return best, 2 * N ** 3 * n / (e0.elapsed_time(e1) * 1e-3) / 1e12
tf32 = measure(True); fp32 = measure(False)
out = {"tf32_tflops": tf32[0], "tf32_tflops_sustained": tf32[1], "fp32_ffma_tflops": fp32[0], "fp32_ffma_tflops_sustained": fp32[1],
       "how": "torch.matmul fp32 8192^3 on cuBLAS with allow_tf32 on / off: best of 10 and a 2 s loop", "gpu": torch.cuda.get_device_name(0)}
print(json.dumps(out))
