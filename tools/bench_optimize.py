"""Time optimize_kernel (PMVO.forward) alone on a slice of the BASELINE workload and print a bit-level checksum, so
kernel variants (MH_LIB=<variant .so>) can be compared for speed AND for bit-identity.
    python tools/bench_optimize.py [n_points] [variant.so ...]     (the committed library is always timed first)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from monohair_b200 import pmvo as P  # noqa: E402
from monohair_b200.camera import cameras_from_scene  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 150000
    dev = torch.device("cuda:0")
    cfg = bench.WORKLOADS["full"]
    sc, cand_np, scalp = bench.make_workload(cfg, dev)
    pm = P.PMVO.from_u8(cameras_from_scene(sc), sc.depth, sc.ori_gray, sc.conf_u8, sc.mask_u8, device=dev,
                        image_size=[cfg["H"], cfg["W"]], patch_size=cfg["patch"], visible_threshold=cfg["visible_thr"],
                        conf_threshold=cfg["conf_thr"])
    cand = torch.from_numpy(cand_np).to(dev).float().contiguous()
    _, pts, _ = pm.filter_points(cand[:cand.size(0) // 30 * 30])
    pts = pts[:n].float().contiguous()
    cks = lambda t: int(t.contiguous().view(torch.int32).to(torch.int64).sum().item())
    from monohair_b200 import _lib
    for lib in [None] + sys.argv[2:]:
        if lib is not None:                                   # tuning only: swap the loaded library under the same host code
            _lib._lib, _lib.LIB_PATH = None, os.path.abspath(lib)
            _lib.lib()
        for _ in range(2):
            _, ori, loss, hc = pm.forward(pts)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); _, ori, loss, hc = pm.forward(pts); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"optimize[{os.path.basename(lib or _lib.LIB_PATH)}]: n={pts.size(0)} median {np.median(ts):.2f} ms "
              f"({np.median(ts) * 1e3 / pts.size(0):.4f} us/pt)  checksum ori {cks(ori)} loss {cks(loss)} hc {int(hc.sum().item())}",
              flush=True)


if __name__ == "__main__":
    main()
