"""Golden vectors for calOrientationGabor from the UNMODIFIED reference (GaborFilter.py) run on CPU."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_import  # noqa: E402


def texture(H, W, seed):
    """sum of oriented line textures + noise, DoG-like zero-mean float image in [-0.2, 0.2]."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    img = np.zeros((H, W))
    for _ in range(40):
        th, wl, ph = rng.uniform(0, np.pi), rng.uniform(3, 6), rng.uniform(0, 2 * np.pi)
        cx, cy, s = rng.uniform(0, W), rng.uniform(0, H), rng.uniform(8, 25)
        env = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
        img += env * np.cos(2 * np.pi * (xx * np.cos(th) + yy * np.sin(th)) / wl + ph)
    img += rng.normal(0, 0.05, (H, W))
    return (img / np.abs(img).max() * 0.2).astype(np.float32)


def main():
    G = ref_import.import_reference_gabor()
    torch.manual_seed(0)
    img = texture(96, 128, 0)
    m = G.calOrientationGabor()
    t = torch.from_numpy(img)[None, None]
    with torch.no_grad():
        two, best, conf = m(t, torch.ones_like(t), iter=1, threshold=0.0)
        bank = torch.stack([m.gabor_fn(17, 1, 1, torch.ones(1) * (np.pi * i / 180), 1.8, 2.4, 4)[0, 0] for i in range(180)])
    np.savez_compressed(os.path.join(HERE, "gabor_small.npz"), image=img, two=two[0].numpy(), orient=best[0, 0].numpy(),
                        conf=conf[0, 0].numpy(), bank=bank.numpy(),
                        versions=np.array([f"torch={torch.__version__}"]))
    print("gabor golden:", two.shape, "conf mean", float(conf.mean()), "distinct orientations", len(np.unique(best.numpy())))


if __name__ == "__main__":
    main()
