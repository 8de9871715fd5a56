"""TEST INFRASTRUCTURE — CPU oracle for the Gabor orientation bank.  NOT part of the product.

(1) calOrientationGabor (GaborFilter.py:16-145) restated with torch CPU ops (one batched conv2d instead of the
    180-launch loop + quadratic torch.cat).  Pinned against the unmodified reference run on CPU
    (tests/golden/gabor_small.npz, tests/golden/make_golden_gabor.py).
(2) calc_orientation_maps.py (generate_gabor_filters :18-24, calc_orients :27-32, calc_confidences :35-49) needs
    scikit-image ==0.23.2 (requirements.txt:36), which is absent from /root/reference and from this image.
    gabor_kernel / difference_of_gaussians are restated from skimage 0.23's published algorithm:
        gabor_kernel(f, theta, sigma_x, sigma_y, n_stds=3): x0 = ceil(max(|n sx cos|, |n sy sin|, 1)),
        y0 = ceil(max(|n sy cos|, |n sx sin|, 1)); y,x = meshgrid(-y0..y0, -x0..x0, 'ij'); rotx = x cos + y sin,
        roty = -x sin + y cos; g = exp(-.5 (rotx^2/sx^2 + roty^2/sy^2)) / (2 pi sx sy) * exp(i 2 pi f rotx)
        difference_of_gaussians(img, lo, hi) = gaussian(img, lo) - gaussian(img, hi), img_as_float first,
        gaussian = scipy.ndimage.gaussian_filter(mode='nearest', truncate=4.0)
    PARITY UNPINNED for (2): no reference test or golden vector fixes these outputs and skimage itself could not be
    run here; the restatement is validated only against the formulas above and scipy.ndimage.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from scipy import ndimage as ndi


# ----------------------------------------------------------------------------------------------- (1)
def gabor_fn(theta, kernel_size=17, sigma_x=1.8, sigma_y=2.4, Lambda=4.0, psi=0.0):
    """GaborFilter.py:115-145 for one theta (float32 tensor [1]) -> [17,17]."""
    half = kernel_size // 2
    y0 = torch.arange(-half, half + 1).float() - 0.5
    x0 = torch.arange(-half, half + 1).float() - 0.5
    y = y0.view(1, -1).repeat(kernel_size, 1)
    x = x0.view(-1, 1).repeat(1, kernel_size)
    sx, sy = torch.ones(1) * sigma_x, torch.ones(1) * sigma_y
    lam, ps = torch.ones(1) * Lambda, torch.ones(1) * psi
    x_t = x * torch.cos(theta.view(-1, 1)) + y * torch.sin(theta.view(-1, 1))
    y_t = -x * torch.sin(theta.view(-1, 1)) + y * torch.cos(theta.view(-1, 1))
    return torch.exp(-.5 * (x_t ** 2 / sx.view(-1, 1) ** 2 + y_t ** 2 / sy.view(-1, 1) ** 2)) \
        * torch.cos(2 * math.pi * x_t / lam.view(-1, 1) + ps.view(-1, 1))


def gabor_bank(n=180, kernel_size=17):
    """the 180 kernels of calOrientationGabor.filter (GaborFilter.py:31-34): theta_i = float32(pi*i/180)."""
    return torch.stack([gabor_fn(torch.ones(1) * (math.pi * i / n), kernel_size) for i in range(n)])


def gabor_orientation(image, n=180, lo=0.0, hi=0.2):
    """calOrientationGabor.forward(iter=1, threshold=0) (GaborFilter.py:29-113).  image [H,W] float32 ->
    (two_channel [2,H,W], orient [H,W], conf [H,W], |responses| [n,H,W])."""
    bank = gabor_bank(n)
    img = torch.as_tensor(image, dtype=torch.float32)[None, None]
    res = torch.cat([F.conv2d(img, bank[i][None, None], padding=8) for i in range(n)], 1)
    res = torch.abs(res)
    max_resp = torch.max(res, dim=1, keepdim=True)[0]
    am = torch.argmax(res, dim=1, keepdim=True).float()
    best = am * math.pi / n
    H, W = img.shape[2:]
    orient = torch.cat([torch.ones(1, 1, H, W) * math.pi * i / n for i in range(n)], 1)
    diff = torch.minimum(torch.abs(best - orient),
                         torch.minimum(torch.abs(best - orient - math.pi), torch.abs(best - orient + math.pi)))
    rd = res - max_resp
    var = torch.sum(diff * rd * rd, dim=1, keepdim=True) ** (1 / 2)
    zero = torch.zeros_like(var)
    orient_data = torch.where(var > zero, best, zero)
    var_data = torch.where(var > zero, var, zero)
    var_data = var_data / torch.max(var_data)
    conf = ((var_data - lo) / (hi - lo)).clamp(0, 1)
    two = torch.cat([torch.sin(orient_data), torch.cos(orient_data)], 1)
    return two[0], orient_data[0, 0], conf[0, 0], res[0]


# ----------------------------------------------------------------------------------------------- (2)
def sk_gabor_kernel(frequency, theta, sigma_x, sigma_y, n_stds=3):
    ct, st = np.cos(theta), np.sin(theta)
    x0 = math.ceil(max(abs(n_stds * sigma_x * ct), abs(n_stds * sigma_y * st), 1))
    y0 = math.ceil(max(abs(n_stds * sigma_y * ct), abs(n_stds * sigma_x * st), 1))
    y, x = np.meshgrid(np.arange(-y0, y0 + 1), np.arange(-x0, x0 + 1), indexing='ij', sparse=True)
    rotx = x * ct + y * st
    roty = -x * st + y * ct
    g = np.empty(np.broadcast(rotx, roty).shape, dtype=np.complex128)
    np.exp(-0.5 * (rotx ** 2 / sigma_x ** 2 + roty ** 2 / sigma_y ** 2) + 1j * (2 * np.pi * frequency * rotx), out=g)
    g *= 1 / (2 * np.pi * sigma_x * sigma_y)
    return g


def generate_gabor_filters(sigma_x=1.8, sigma_y=2.4, freq=0.23, num_filters=180):
    """calc_orientation_maps.py:18-24."""
    thetas = np.linspace(0, math.pi * (num_filters - 1) / num_filters, num_filters)
    return [np.real(sk_gabor_kernel(freq, theta=math.pi - t, sigma_x=sigma_x, sigma_y=sigma_y)) for t in thetas]


def rgb2gray(rgb):
    """calc_orientation_maps.py:12-15."""
    return 0.2989 * rgb[:, :, 0] + 0.5870 * rgb[:, :, 1] + 0.1140 * rgb[:, :, 2]


def difference_of_gaussians(img, lo, hi):
    img = np.asarray(img)
    if img.dtype == np.uint8:
        img = img / 255.0
    img = img.astype(np.float64)
    return ndi.gaussian_filter(img, lo, mode='nearest', truncate=4.0) - ndi.gaussian_filter(img, hi, mode='nearest', truncate=4.0)


def calc_orients(img, kernels):
    """calc_orientation_maps.py:27-32 -> F_orients [n,H,W] float64."""
    filt = difference_of_gaussians(rgb2gray(img), 0.4, 10)
    return np.abs(np.stack([ndi.convolve(filt, k, mode='wrap') for k in kernels]))


def calc_confidences(F_orients, orientation_map, num_filters=180):
    """calc_orientation_maps.py:35-49 -> V_F."""
    bins = np.linspace(0, math.pi * (num_filters - 1) / num_filters, num_filters)[:, None, None]
    om = orientation_map[None]
    d = np.minimum(np.abs(om - bins), np.minimum(np.abs(om - bins - math.pi), np.abs(om - bins + math.pi)))
    return (d ** 2 * (F_orients / F_orients.sum(axis=0, keepdims=True))).sum(0)
