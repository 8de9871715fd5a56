"""GPU parity of the Gabor kernels: float32 bank against goldens from the unmodified reference (CPU), float64
calc_orientation_maps path against the numpy/scipy oracle.

Tolerances (float): |responses| 1e-4 relative to the image's max response (289-term fp32 sums in a different order
than oneDNN's), orientation index exact wherever the top-2 response margin exceeds 1e-4*max (BASELINE.md §3 /
SURVEY.md §8d config 1), confidence 2e-3 abs on those pixels.  The float64 path is compared at 1e-12.
"""
import math

import numpy as np
import pytest
import torch

from golden_util import load

pytestmark = pytest.mark.gpu


def test_bank_matches_reference():
    from monohair_b200.gabor import calOrientationGabor
    g = load("gabor_small")
    assert np.array_equal(calOrientationGabor().bank("cuda:0").cpu().numpy(), g["bank"])


@pytest.mark.parametrize("tensor_cores", [False, True], ids=["fp32_cuda_cores", "tcgen05_tf32x3"])
def test_gabor_orientation_vs_reference_golden(tensor_cores):
    """both kernels against the unmodified reference (CPU, fp32): the CUDA-core one and the tensor-core one (tcgen05,
    3-term tf32 split, fused epilogue); same margin gate, same tolerances."""
    from monohair_b200.gabor import calOrientationGabor
    from oracle import gabor_oracle as G
    g = load("gabor_small")
    img = torch.from_numpy(g["image"])[None, None].cuda()
    two, orient, conf = calOrientationGabor(tensor_cores=tensor_cores)(img, None, iter=1, threshold=0.0)
    orient, conf, two = orient[0, 0].cpu().numpy(), conf[0, 0].cpu().numpy(), two[0].cpu().numpy()
    _, _, _, res = G.gabor_orientation(g["image"])
    res = res.numpy()
    top2 = np.sort(res, axis=0)[-2:]
    margin = (top2[1] - top2[0]) / res.max()
    ok = margin > 1e-4
    print(f"\ngabor: {ok.mean() * 100:.2f}% pixels with top-2 margin > 1e-4; orientation identical on "
          f"{np.mean(orient == g['orient']) * 100:.3f}% of all pixels")
    assert ok.mean() > 0.5
    assert np.array_equal(orient[ok], g["orient"][ok])
    assert np.mean(orient == g["orient"]) >= 0.995
    assert np.abs(conf - g["conf"])[ok].max() <= 2e-3
    assert np.abs(two - g["two"])[:, ok].max() <= 1e-6
    # distinct orientation values are exactly the reference's float32 k*pi/180
    assert set(np.unique(orient)) <= set(np.unique(g["orient"])) | {np.float32(0)}


@pytest.mark.parametrize("tensor_cores", [False, True], ids=["fp32_cuda_cores", "tcgen05_tf32x3"])
def test_full_frame_properties(tensor_cores):
    """1080p frame: a pure sinusoidal grating must come back with its own orientation (up to the bank's 1 degree),
    confidence in [0,1], and the result must not depend on the tile decomposition (shifted crop equality)."""
    from monohair_b200.gabor import calOrientationGabor
    H, W = 1080, 1920
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    th = math.radians(30.0)
    img = (0.1 * torch.cos(2 * math.pi * (yy * math.cos(th) + xx * math.sin(th)) / 4.0)).float()
    m = calOrientationGabor(tensor_cores=tensor_cores)
    two, orient, conf = m(img[None, None].cuda())
    o = orient[0, 0, 100:-100, 100:-100]
    assert float(conf.min()) >= 0 and float(conf.max()) <= 1
    # the bank's own argmax is periodic-pattern sensitive: ~1 % of pixels land a few degrees off (same on the oracle)
    assert torch.all((o - th).abs() < math.radians(10.0)) and ((o - th).abs() < math.radians(0.01)).float().mean() > 0.95
    crop = img[37:37 + 400, 53:53 + 600].contiguous()
    _, o2, _ = m(crop[None, None].cuda())
    assert torch.equal(o2[0, 0, 20:-20, 20:-20], orient[0, 0, 57:417, 73:633])


def test_calc_orientation_maps_f64_vs_oracle():
    from monohair_b200 import gabor as MG
    from oracle import gabor_oracle as G
    rng = np.random.default_rng(1)
    H = W = 96
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.zeros((H, W, 3))
    for _ in range(30):
        th, wl = rng.uniform(0, np.pi), rng.uniform(3, 6)
        img += (np.cos(2 * np.pi * (xx * np.cos(th) + yy * np.sin(th)) / wl) * 20)[..., None]
    img = np.clip(img + 128 + rng.normal(0, 4, img.shape), 0, 255).astype(np.uint8)
    ks_o = G.generate_gabor_filters()
    ks = MG.generate_gabor_filters(1.8, 2.4, 0.23, 180)
    assert all(np.array_equal(a, b) for a, b in zip(ks, ks_o))
    F_o = G.calc_orients(img.astype(np.float64), ks_o)
    F = MG.calc_orients(img.astype(np.float64), ks).cpu().numpy()
    assert np.abs(F - F_o).max() <= 1e-12 * max(1.0, F_o.max())
    print(f"\ncalc_orients float64: bit-identical {np.mean(F == F_o) * 100:.2f}%")
    om_o = F_o.argmax(0)
    om = F.argmax(0)
    assert np.mean(om == om_o) > 0.9999
    V_o = G.calc_confidences(F_o, om_o / 180 * math.pi)
    V = MG.calc_confidences(torch.from_numpy(F).cuda(), om / 180 * math.pi).cpu().numpy()
    assert np.allclose(V, V_o, rtol=1e-10, atol=1e-12)
    d_o = G.difference_of_gaussians(G.rgb2gray(img.astype(np.float64)), 0.4, 10)
    d = MG.difference_of_gaussians(G.rgb2gray(img.astype(np.float64)), 0.4, 10).cpu().numpy()
    assert np.abs(d - d_o).max() <= 1e-12


def _line_texture(H, W, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    img = np.zeros((H, W))
    for _ in range(40):
        th, lam, ph = rng.uniform(0, np.pi), rng.uniform(3, 6), rng.uniform(0, 2 * np.pi)
        img += np.cos(2 * np.pi * (xx * np.cos(th) + yy * np.sin(th)) / lam + ph)
    img = (img - img.min()) / (img.max() - img.min()) * 255
    return np.clip(img + rng.normal(0, 4, img.shape), 0, 255).astype(np.uint8)


def test_calculate_orientation_files_and_batch_generate(tmp_path):
    """GaborFilter.py:164-237: the three files per image with the reference's encodings: best_ori through
    cv2.imwrite of a float array (round half to even, saturate: 0.5->0, 1.5->2, 2.5->2), conf through torchvision's
    save_image (floor(x*255+0.5), three equal channels), Ori as BGR of {1, (sin+1)/2, (cos+1)/2}*255."""
    import cv2
    from PIL import Image
    from monohair_b200 import gabor as GB
    root = tmp_path
    (root / "capture_images").mkdir()
    (root / "hair_mask").mkdir()
    for i in range(2):
        Image.fromarray(np.stack([_line_texture(96, 128, 20 + i)] * 3, -1)).save(root / "capture_images" / f"{i:03d}.png")
        Image.fromarray(np.full((96, 128), 255, np.uint8)).save(root / "hair_mask" / f"{i:03d}.png")
    GB.batch_generate(str(root), "capture_images")
    for i in range(2):
        name = f"{i:03d}.png"
        image = np.array(Image.open(root / "capture_images" / name).convert('L'))
        gray = GB.difference_of_gaussians(image, 0.4, 10).type(torch.float)[None, None]
        ori, best, conf = GB.calOrientationGabor()(gray, None, 1, threshold=0.0)
        deg = best[0, 0].cpu().numpy().astype(np.float64) / math.pi * 180
        f_best = cv2.imread(str(root / "best_ori" / name), cv2.IMREAD_UNCHANGED)
        assert f_best.dtype == np.uint8 and f_best.shape == (96, 128)
        assert np.array_equal(f_best, np.clip(np.rint(deg), 0, 255).astype(np.uint8))      # np.rint: half to even, like cv2
        f_conf = np.array(Image.open(root / "conf" / name))
        c8 = np.floor(conf[0, 0].cpu().numpy().astype(np.float32) * np.float32(255) + np.float32(0.5)).clip(0, 255).astype(np.uint8)
        assert f_conf.shape == (96, 128, 3) and np.array_equal(f_conf[..., 0], c8) and np.array_equal(f_conf[..., 1], c8)
        f_ori = cv2.imread(str(root / "Ori" / name), cv2.IMREAD_UNCHANGED)                # BGR as stored
        o = (ori[0].cpu().numpy().transpose(1, 2, 0).astype(np.float64) + 1) / 2
        exp = np.concatenate([np.ones((96, 128, 1)), o], axis=2)[..., ::-1] * 255
        assert np.array_equal(f_ori, np.clip(np.rint(exp), 0, 255).astype(np.uint8))
        # what PMVO reads back (Load_Ori_And_Conf) decodes to the orientation the bank found, to the 1-degree quantisation
        back = (180 - f_best.astype(np.float64)) / 180 * math.pi
        assert np.abs(np.sin(back) - np.sin(math.pi - np.deg2rad(np.rint(deg)))).max() < 1e-12
    # the encode rules themselves, on exact half-way values
    probe = np.array([[0.5, 1.5, 2.5, 254.5, 255.5, 300.0, -3.0]], np.float64)
    cv2.imwrite(str(root / "probe.png"), probe)
    assert cv2.imread(str(root / "probe.png"), cv2.IMREAD_UNCHANGED).tolist() == [[0, 2, 2, 254, 255, 255, 0]]


def test_calc_orientation_maps_main_writes_reference_files(tmp_path):
    """calc_orientation_maps.py:51-92 end to end on one 64x80 frame: file names, dtypes, and contents equal to the numpy /
    scipy oracle pushed through the same encodings."""
    import types
    import cv2
    from PIL import Image
    from monohair_b200 import gabor as GB
    from oracle import gabor_oracle as G
    rgb = np.stack([_line_texture(64, 80, 30 + c) for c in range(3)], -1)
    for d in ("img", "mask"):
        (tmp_path / d).mkdir()
    Image.fromarray(rgb).save(tmp_path / "img" / "f0.png")
    Image.fromarray(np.full((64, 80), 255, np.uint8)).save(tmp_path / "mask" / "f0.png")
    args = types.SimpleNamespace(img_path=str(tmp_path / "img"), mask_path=str(tmp_path / "mask"), orient_dir=str(tmp_path / "orient"),
                                 conf_dir=str(tmp_path / "conf"), sigma_x=1.8, sigma_y=2.4, freq=0.23, num_filters=180)
    GB.main(args)
    F = G.calc_orients(rgb, G.generate_gabor_filters(1.8, 2.4, 0.23, 180))
    om = F.argmax(0)
    idx = cv2.imread(str(tmp_path / "orient" / "f0.png"), cv2.IMREAD_UNCHANGED)
    assert idx.dtype == np.uint8 and np.array_equal(idx, om.astype('uint8'))
    conf = np.load(tmp_path / "conf" / "f0.npy")
    assert conf.dtype == np.float16
    ref_conf = (1 / G.calc_confidences(F, om / 180 * math.pi) ** 2)
    assert np.allclose(conf.astype(np.float64), ref_conf.astype(np.float16).astype(np.float64), rtol=2e-3)
    for f in ("f0_ori.png", "f01.png", "f0_conf.png"):
        assert (tmp_path / "orient" / f).exists()
    rad = om / 180 * math.pi
    cm = (np.stack([np.cos(rad) * 0.5 + 0.5, np.sin(rad) * 0.5 + 0.5, np.zeros_like(rad)], 2).astype(np.float32) * 255)
    assert np.array_equal(cv2.imread(str(tmp_path / "orient" / "f0_ori.png"), cv2.IMREAD_UNCHANGED), np.clip(np.rint(cm), 0, 255).astype(np.uint8))


def test_tensor_core_bank_agrees_with_fp32_bank_on_ragged_sizes():
    """tcgen05 kernel vs the CUDA-core kernel on sizes that are not multiples of its 128 x 2 tile (edges, odd heights):
    orientation identical wherever the fp32 kernel's top-2 margin is above 1e-5 of the largest response."""
    from monohair_b200.gabor import calOrientationGabor
    from oracle import gabor_oracle as G
    for (H, W), seed in (((33, 129), 1), ((2, 128), 2), ((65, 40), 3), ((131, 300), 4)):
        img = _line_texture(H, W, seed).astype(np.float32) / 255.0 - 0.5
        x = torch.from_numpy(img)[None, None].cuda()
        _, o_ref, c_ref = calOrientationGabor(tensor_cores=False)(x)
        _, o_tc, c_tc = calOrientationGabor(tensor_cores=True)(x)
        _, _, _, res = G.gabor_orientation(img)
        res = res.numpy()
        top2 = np.sort(res, axis=0)[-2:]
        ok = torch.from_numpy((top2[1] - top2[0]) / res.max() > 1e-5).cuda()
        assert ok.float().mean() > 0.5
        assert torch.equal(o_tc[0, 0][ok], o_ref[0, 0][ok]), (H, W)
        assert float((c_tc - c_ref)[0, 0][ok].abs().max()) <= 2e-3
