import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from monohair_b200.gabor import calOrientationGabor
torch.manual_seed(0)
for (H, W) in ((96, 160), (130, 257), (1080, 1920)):
    x = torch.rand((1, 1, H, W), device="cuda") * 0.2 - 0.1
    ref = calOrientationGabor(tensor_cores=False)
    tc = calOrientationGabor(tensor_cores=True)
    t2, o2, c2 = ref(x)
    t1, o1, c1 = tc(x)
    torch.cuda.synchronize()
    same = (o1 == o2).float().mean().item()
    print(H, W, "orient identical", same, "max|dconf|", (c1 - c2).abs().max().item(), "conf mean", c2.mean().item(), flush=True)
    for m, name in ((ref, "fp32"), (tc, "tc")):
        for _ in range(2): m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): m(x)
        e1.record(); torch.cuda.synchronize()
        print("  ", name, e0.elapsed_time(e1) / 5, "ms", flush=True)
