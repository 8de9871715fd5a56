"""CPU: Gabor oracle against goldens from the unmodified reference, and the skimage restatement's invariants."""
import math

import numpy as np
import torch

from golden_util import load
from oracle import gabor_oracle as G


def test_bank_bit_exact_vs_reference():
    g = load("gabor_small")
    assert np.array_equal(G.gabor_bank().numpy(), g["bank"])


def test_orientation_and_confidence_vs_reference():
    g = load("gabor_small")
    two, orient, conf, _ = G.gabor_orientation(g["image"])
    assert np.array_equal(orient.numpy(), g["orient"])
    assert np.array_equal(conf.numpy(), g["conf"])
    assert np.array_equal(two.numpy(), g["two"])


def test_skimage_restatement_shapes_and_symmetry():
    ks = G.generate_gabor_filters()
    assert len(ks) == 180
    assert ks[0].shape == (17, 13) or ks[0].shape == (13, 17)          # variable support
    assert all(k.shape[0] % 2 == 1 and k.shape[1] % 2 == 1 and max(k.shape) <= 17 for k in ks)
    k = ks[37]
    assert np.allclose(k, k[::-1, ::-1])                                # real Gabor with offset 0 is even
    img = np.random.default_rng(0).integers(0, 255, (40, 48)).astype(np.uint8)
    d = G.difference_of_gaussians(img, 0.4, 10)
    assert d.dtype == np.float64 and abs(d.mean()) < 0.05
