"""Developer tool: how exact is torch's CPU fp32 matmul on this host?  (The CPU oracle of the parity tests is built on it.)"""
import os, torch, platform
print("cpu:", platform.processor(), "| threads", torch.get_num_threads(), "| matmul precision", torch.get_float32_matmul_precision())
print("mkldnn enabled", torch.backends.mkldnn.enabled, "| env", {k: v for k, v in os.environ.items() if "DNN" in k or "MKL" in k or "OMP" in k})
try:
    print(open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0])
    flags = open("/proc/cpuinfo").read().split("flags")[1].split("\n")[0]
    print("amx:", "amx_tile" in flags, "avx512_bf16:", "avx512_bf16" in flags)
except Exception as e:
    print(e)
g = torch.Generator().manual_seed(0)
for M in (20, 64, 70, 2048):
    for K in (32, 64, 128, 256):
        for N in (128, 256):
            a = torch.randn(M, K, generator=g); b = 0.1 * torch.randn(K, N, generator=g)
            ref = a.double() @ b.double()
            e1 = float(((a @ b).double() - ref).abs().max() / ref.abs().max())
            with torch.backends.mkldnn.flags(enabled=False):
                e2 = float(((a @ b).double() - ref).abs().max() / ref.abs().max())
            flag = "  <-- reduced precision" if e1 > 5e-6 else ""
            print("M=%4d K=%3d N=%3d: rel err %.2e (mkldnn off: %.2e)%s" % (M, K, N, e1, e2, flag))
