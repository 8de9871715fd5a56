"""Micro-benchmark of mh_knn (k=100) on shell-like points, for ncu."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monohair_b200 import synthetic as syn  # noqa: E402
from monohair_b200 import pmvo as P  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    cand = syn.candidate_points(num_per_grid=4, seed=0)
    rng = np.random.default_rng(0)
    pts = torch.from_numpy(cand[rng.random(cand.shape[0]) < 0.47]).to(dev).float().contiguous()
    for _ in range(2):
        idx = P.knn(pts, pts, 100, dev)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); idx = P.knn(pts, pts, 100, dev); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"knn: n={pts.size(0)} k=100 median {np.median(ts):.2f} ms -> {pts.size(0) / np.median(ts) / 1e3:.1f} M queries/s")


if __name__ == "__main__":
    main()
