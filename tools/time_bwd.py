import torch, sys, os
sys.path.insert(0, os.getcwd())
from kws_b200 import engine
from oracle import fastgrnn_oracle as O
dev = torch.device("cuda:0")
for B in (2048, 8192):
    torch.manual_seed(0)
    p = O.init_params(32, 128)
    params = {k: v.to(dev).contiguous() for k, v in p.tensors().items()}
    x = torch.randn(99, B, 32, device=dev); go = torch.randn(99, B, 128, device=dev) / B
    out, z_s, c_s, _ = engine.forward(x, params, None, layout="IH", batch_first=False, save_for_backward=True)
    for _ in range(3): g = engine.backward(go, x, out, z_s, c_s, params, None, layout="IH", batch_first=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): g = engine.backward(go, x, out, z_s, c_s, params, None, layout="IH", batch_first=False)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B}: backward (rec + contract + reduce) {e0.elapsed_time(e1)/20*1e3:.1f} us")
