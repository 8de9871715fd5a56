"""B200-native Gabor orientation maps: host-side mirror of preprocess_capture_data/GaborFilter.py
(calOrientationGabor, calculate_orientation, batch_generate) and calc_orientation_maps.py
(generate_gabor_filters, calc_orients, calc_confidences), driving csrc/gabor.cu through the C ABI.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch
import torch.nn as nn

from ._lib import check, lib, ptr, stream_ptr


def bank_for_tensor_cores(bank):
    """[n<=192,17,17] float32 bank -> the operand mh_gabor_orientation_tc streams: per kernel row r the matrix
    B_r[filter][tap] (192 x 24, zero padded) split into tf32 parts hi = rna(x), lo = rna(x - hi) (round to nearest, ties
    away: what cvt.rna.tf32.f32 does to the image operand on the device) and laid out as the tensor core reads it from
    shared memory: [k-group of 4 taps][group of 8 filters][8 filters][4 taps]."""
    b = np.asarray(bank, dtype=np.float32)
    n, ks, _ = b.shape
    assert ks == 17 and n <= 192

    def rna(x):
        u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
        return ((u + np.uint32(0x1000)) & np.uint32(0xffffe000)).view(np.float32)
    out = np.zeros((17, 2, 6, 24, 8, 4), dtype=np.float32)
    full = np.zeros((17, 192, 24), dtype=np.float32)
    full[:, :n, :17] = b.transpose(1, 0, 2)                     # [r][filter][tap]
    hi = rna(full)
    lo = rna(full - hi)
    for part, m in enumerate((hi, lo)):
        out[:, part] = m.reshape(17, 24, 8, 6, 4).transpose(0, 3, 1, 2, 4)     # [r][ng][nr][kg][ke] -> [r][kg][ng][nr][ke]
    return out


class calOrientationGabor(nn.Module):
    """GaborFilter.py:16-145.  forward() keeps the reference signature; only iter=1 (the only value the pipeline
    uses, GaborFilter.py:237) is implemented on the device.  Two kernels: the tensor-core one (tcgen05, 3-term tf32 split,
    fused epilogue; default) and the fp32 CUDA-core one that reproduces the reference's arg-max on every pixel of the
    goldens (tensor_cores=False, or MH_GABOR_FP32=1)."""

    def __init__(self, channel_in=1, channel_out=1, stride=1, tensor_cores=None):
        super().__init__()
        self.tensor_cores = (os.environ.get("MH_GABOR_FP32", "0") != "1") if tensor_cores is None else bool(tensor_cores)
        self._bank_tc = None
        self.channel_in = channel_in
        self.channel_out = channel_out
        self.numKernels = 180
        self.clamp_confidence_low = 0.0
        self.clamp_confidence_high = 0.2
        self._bank = None

    def gabor_fn(self, kernel_size, channel_in, channel_out, theta, sigma_x, sigma_y, Lambda, phase=0.):
        """GaborFilter.py:115-145 (built on the host with the reference's float32 torch ops; 180 x 289 values)."""
        theta = torch.as_tensor(theta, dtype=torch.float).cpu()
        sigma_x = torch.ones(channel_out) * sigma_x
        sigma_y = torch.ones(channel_out) * sigma_y
        Lambda = torch.ones(channel_out) * Lambda
        psi = torch.ones(channel_out) * phase
        xmax = ymax = kernel_size // 2
        ksize = 2 * xmax + 1
        y_0 = torch.arange(-ymax, ymax + 1).float() - 0.5
        y = y_0.view(1, -1).repeat(channel_out, channel_in, ksize, 1).float()
        x_0 = torch.arange(-xmax, xmax + 1).float() - 0.5
        x = x_0.view(-1, 1).repeat(channel_out, channel_in, 1, ksize).float()
        x_theta = x * torch.cos(theta.view(-1, 1, 1, 1)) + y * torch.sin(theta.view(-1, 1, 1, 1))
        y_theta = -x * torch.sin(theta.view(-1, 1, 1, 1)) + y * torch.cos(theta.view(-1, 1, 1, 1))
        return torch.exp(-.5 * (x_theta ** 2 / sigma_x.view(-1, 1, 1, 1) ** 2 + y_theta ** 2 / sigma_y.view(-1, 1, 1, 1) ** 2)) \
            * torch.cos(2 * math.pi * x_theta / Lambda.view(-1, 1, 1, 1) + psi.view(-1, 1, 1, 1))

    def bank(self, device, sigma_x=1.8, sigma_y=2.4, Lambda=4, kernel_size=17):
        if self._bank is None or self._bank.device != torch.device(device):
            ks = [self.gabor_fn(kernel_size, 1, 1, torch.ones(1) * (math.pi * i / self.numKernels), sigma_x, sigma_y, Lambda)[0, 0]
                  for i in range(self.numKernels)]
            self._bank = torch.stack(ks).contiguous().to(device)
        return self._bank

    def forward(self, image, label=None, iter=1, threshold=0.0):
        """image [1,1,H,W] float32 on a CUDA device -> (orientTwoChannel [1,2,H,W], orient [1,1,H,W], conf [1,1,H,W])."""
        if iter != 1:
            raise NotImplementedError("only iter=1 is implemented (the pipeline's value, GaborFilter.py:237)")
        assert image.dim() == 4 and image.size(0) == 1 and image.size(1) == 1
        if not image.is_cuda:
            image = image.cuda()
        dev = image.device
        H, W = int(image.size(2)), int(image.size(3))
        img = image.type(torch.float).contiguous()
        bank = self.bank(dev)
        orient = torch.empty((1, 1, H, W), dtype=torch.float32, device=dev)
        conf = torch.empty((1, 1, H, W), dtype=torch.float32, device=dev)
        two = torch.empty((1, 2, H, W), dtype=torch.float32, device=dev)
        if self.tensor_cores:
            if self._bank_tc is None or self._bank_tc.device != dev:
                self._bank_tc = torch.from_numpy(bank_for_tensor_cores(bank.cpu().numpy())).to(dev).contiguous()
                assert self._bank_tc.numel() * 4 == lib().mh_gabor_tc_bank_bytes()
            wsb = 4 * H * W + 256
            ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                check(lib().mh_gabor_orientation_tc(stream_ptr(dev), ptr(img), H, W, ptr(self._bank_tc), self.numKernels,
                                                    float(self.clamp_confidence_low), float(self.clamp_confidence_high),
                                                    ptr(orient), ptr(conf), ptr(two), ptr(ws), wsb), "mh_gabor_orientation_tc")
            conf[conf < threshold] = 0
            return two, orient, conf
        wsb = lib().mh_gabor_workspace_bytes(H, W, self.numKernels)
        ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib().mh_gabor_orientation(stream_ptr(dev), ptr(img), H, W, ptr(bank), self.numKernels, 17,
                                             float(self.clamp_confidence_low), float(self.clamp_confidence_high),
                                             ptr(orient), ptr(conf), ptr(two), ptr(ws), wsb), "mh_gabor_orientation")
        conf[conf < threshold] = 0
        return two, orient, conf


def _gauss_kernel1d(sigma, truncate=4.0):
    """scipy.ndimage._gaussian_kernel1d (order 0)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return (phi / phi.sum()), radius


def difference_of_gaussians(image, low_sigma, high_sigma, device="cuda:0"):
    """skimage.filters.difference_of_gaussians restated (see oracle/gabor_oracle.py) on the device, float64."""
    img = np.asarray(image)
    if img.dtype == np.uint8:
        img = img / 255.0
    d = torch.from_numpy(np.ascontiguousarray(img, dtype=np.float64)).to(device)
    H, W = d.shape
    k_lo, r_lo = _gauss_kernel1d(low_sigma)
    k_hi, r_hi = _gauss_kernel1d(high_sigma)
    k_lo_d = torch.from_numpy(k_lo).to(device)
    k_hi_d = torch.from_numpy(k_hi).to(device)
    out = torch.empty_like(d)
    scratch = torch.empty((2, H, W), dtype=torch.float64, device=device)
    with torch.cuda.device(d.device):
        check(lib().mh_dog_f64(stream_ptr(d.device), ptr(d), H, W, ptr(k_lo_d), r_lo, ptr(k_hi_d), r_hi, ptr(out), ptr(scratch)),
              "mh_dog_f64")
    return out


def calculate_orientation(image_dir, label_dir, save_root, filename=None, iter=1, threshold=0.0, device="cuda:0"):
    """GaborFilter.py:164-224: writes best_ori/<file>, conf/<file>, Ori/<file>."""
    import cv2
    from PIL import Image
    for sub in ("Ori", "conf", "best_ori"):
        os.makedirs(os.path.join(save_root, sub), exist_ok=True)
    image = np.array(Image.open(image_dir).convert('L'))
    dog = difference_of_gaussians(image, 0.4, 10, device=device)
    gray = dog.type(torch.float)[None, None]
    ori, best_ori, confidence = calOrientationGabor()(gray, None, iter, threshold=threshold)
    cv2.imwrite(os.path.join(save_root, 'best_ori', filename), best_ori[0].cpu().numpy().transpose(1, 2, 0) / math.pi * 180,
                [int(cv2.IMWRITE_JPEG_QUALITY), 100])
    # torchvision.utils.save_image(confidence, path) (GaborFilter.py:208): make_grid repeats the single channel three
    # times, x*255 + 0.5, clamp, uint8, then PIL writes the file in the format of the extension with PIL's defaults
    # (JPEG quality 75 for .jpg/.JPG captures, lossless for .png)
    c8 = confidence[0, 0].mul(255).add_(0.5).clamp_(0, 255).to(torch.uint8).cpu().numpy()
    Image.fromarray(np.repeat(c8[..., None], 3, axis=2)).save(os.path.join(save_root, 'conf', filename))
    o = ori[0].cpu().numpy().transpose(1, 2, 0)
    o = (o + 1) / 2
    H, W = o.shape[:2]
    cv2.imwrite(os.path.join(save_root, 'Ori', filename), np.concatenate([np.ones((H, W, 1)), o], axis=2)[..., ::-1] * 255,
                [int(cv2.IMWRITE_JPEG_QUALITY), 100])


def batch_generate(root, image_folder, device="cuda:0"):
    """GaborFilter.py:231-237."""
    for file in os.listdir(os.path.join(root, image_folder)):
        calculate_orientation(os.path.join(root, image_folder, file), os.path.join(root, 'hair_mask', file), save_root=root,
                              filename=file, iter=1, threshold=0.0, device=device)


# ---------------------------------------------------------------------- calc_orientation_maps.py
def gabor_kernel(frequency, theta, sigma_x, sigma_y, n_stds=3):
    """skimage.filters.gabor_kernel (scikit-image 0.23) restated; see oracle/gabor_oracle.py for the citation."""
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


def generate_gabor_filters(sigma_x, sigma_y, freq, num_filters):
    """calc_orientation_maps.py:18-24."""
    thetas = np.linspace(0, math.pi * (num_filters - 1) / num_filters, num_filters)
    return [np.real(gabor_kernel(freq, theta=math.pi - t, sigma_x=sigma_x, sigma_y=sigma_y)) for t in thetas]


def rgb2gray(rgb):
    r, g, b = rgb[:, :, 0], rgb[:, :, 1], rgb[:, :, 2]
    return 0.2989 * r + 0.5870 * g + 0.1140 * b


def calc_orients(img, kernels, device="cuda:0"):
    """calc_orientation_maps.py:27-32 on the device (float64, periodic true convolution) -> F_orients [n,H,W] tensor."""
    gray_img = rgb2gray(np.asarray(img))
    filtered = difference_of_gaussians(gray_img, 0.4, 10, device=device)
    K = max(max(k.shape) for k in kernels)
    n = len(kernels)
    bank = np.zeros((n, K, K))
    for i, k in enumerate(kernels):                       # centred zero padding to K x K (all supports are odd)
        oy, ox = (K - k.shape[0]) // 2, (K - k.shape[1]) // 2
        bank[i, oy:oy + k.shape[0], ox:ox + k.shape[1]] = k
    bank_d = torch.from_numpy(bank).to(device)
    H, W = filtered.shape
    out = torch.empty((n, H, W), dtype=torch.float64, device=device)
    with torch.cuda.device(filtered.device):
        check(lib().mh_filterbank_wrap_f64(stream_ptr(filtered.device), ptr(filtered), H, W, ptr(bank_d), n, K, ptr(out)),
              "mh_filterbank_wrap_f64")
    return out


def main(args, device="cuda:0"):
    """calc_orientation_maps.py:51-92: the first image of args.img_path -> orientation / confidence files with the
    reference's names, dtypes and cv2 encodings (float arrays handed to cv2.imwrite are rounded half-to-even and
    saturated to 8 bit by OpenCV; the RGB->BGR conversion of :78 only touches `<basename>1.png`)."""
    import cv2
    from PIL import Image
    os.makedirs(args.orient_dir, exist_ok=True)
    os.makedirs(args.conf_dir, exist_ok=True)
    kernels = generate_gabor_filters(args.sigma_x, args.sigma_y, args.freq, args.num_filters)
    for img_name in sorted(os.listdir(args.img_path))[:1]:
        basename = img_name.split('.')[0]
        img = np.array(Image.open(os.path.join(args.img_path, img_name)))
        mask = np.array(Image.open(os.path.join(args.mask_path, img_name)))
        mask = mask / np.max(mask)                                       # computed and unused, as in the reference
        F_orients = calc_orients(img, kernels, device=device)
        orientation_map = F_orients.argmax(0).cpu().numpy()
        orientation_map_rad = orientation_map / args.num_filters * math.pi
        indices_cm2 = np.stack([np.cos(orientation_map_rad) * 0.5 + 0.5, np.sin(orientation_map_rad) * 0.5 + 0.5,
                                np.zeros_like(orientation_map_rad)], axis=2)
        indices_cm2 = indices_cm2.astype(np.float32) * 255
        cv2.imwrite(f'{args.orient_dir}/{basename}_ori.png', indices_cm2)
        indices_cm2 = cv2.cvtColor(indices_cm2, cv2.COLOR_RGB2BGR)
        confidence_map = calc_confidences(F_orients, orientation_map_rad, args).cpu().numpy()
        cv2.imwrite(f'{args.orient_dir}/{basename}1.png', indices_cm2)
        confidence_map = 1 / confidence_map ** 2
        cv2.imwrite(f'{args.orient_dir}/{basename}_conf.png', confidence_map * 255 / 10)
        cv2.imwrite(f'{args.orient_dir}/{basename}.png', orientation_map.astype('uint8'))
        np.save(f'{args.conf_dir}/{basename}.npy', confidence_map.astype('float16'))


def calc_confidences(F_orients, orientation_map, args=None, num_filters=180):
    """calc_orientation_maps.py:35-49 (torch ops on the device)."""
    nf = args.num_filters if args is not None else num_filters
    F_orients = torch.as_tensor(F_orients)
    om = torch.as_tensor(orientation_map, device=F_orients.device, dtype=torch.float64)[None]
    bins = torch.from_numpy(np.linspace(0, math.pi * (nf - 1) / nf, nf)).to(F_orients.device)[:, None, None]
    d = torch.minimum(torch.abs(om - bins), torch.minimum(torch.abs(om - bins - math.pi), torch.abs(om - bins + math.pi)))
    Fn = F_orients / F_orients.sum(dim=0, keepdim=True)
    return (d ** 2 * Fn).sum(0)
