"""Device side of the reference's training-image transform (codes/get_dataloader.py:30-36, identical for the COCO and the WikiArt
dataset classes):

    ToPILImage -> Resize((512, 512)) -> RandomCrop((256, 256)) -> ToTensor -> Normalize(ImageNet mean / std)

`GpuTrainTransform` takes the decoded RGB uint8 image (what `cv2.imread` + `cvtColor` hand to the transform, :64-70) and produces the
normalised fp32 [3, 256, 256] tensor ON THE GPU with one kernel (csrc/norm_misc.cu: resize_crop_normalize_kernel): the host only
uploads the raw pixels -- 3 bytes per input pixel instead of running Pillow's resample and shipping 12 bytes per output pixel --
and only the cropped window of the resized image is computed.  Pillow's antialiased bilinear resample is restated exactly (its
fixed-point coefficients are built here, on the host, the way libImaging/Resample.c builds them), and the crop offsets are drawn
like torchvision's RandomCrop.get_params, so a run seeded like the reference crops the same windows: the result is bit-identical
to the reference's pipeline (tests/test_gpu_kernels.py).  JPEG decoding stays where the reference has it (cv2, host).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import ops

PRECISION_BITS = 32 - 8 - 2  # Pillow: libImaging/Resample.c


def pil_resize_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR filter (support 1, antialiased when shrinking):
    per output coordinate the first input coordinate, the number of taps and the taps as 22-bit fixed-point integers."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin_a = np.zeros(out_size, np.int32)
    cnt_a = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = []
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            a = -a if a < 0.0 else a
            v = 1.0 - a if a < 1.0 else 0.0
            w.append(v)
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        xmin_a[xx], cnt_a[xx] = xmin, xmax
    return xmin_a, cnt_a, kk


class GpuTrainTransform:
    """Drop-in for the `transform` argument of the reference's dataset classes when the images are to be prepared on the GPU:
    `transform(img_rgb_uint8_hwc)` -> normalised fp32 CUDA tensor [3, crop, crop].  `img` may be a numpy array (as cv2 returns it)
    or a uint8 tensor (host or device)."""

    def __init__(self, device, size=(512, 512), crop=(256, 256), mean=ops.IMAGENET_MEAN, std=ops.IMAGENET_STD):
        self.device = torch.device(device)
        self.size, self.crop, self.mean, self.std = tuple(size), tuple(crop), mean, std
        self._coeffs: Dict[Tuple[int, int], tuple] = {}

    def coeffs(self, in_size: int, out_size: int):
        key = (in_size, out_size)
        c = self._coeffs.get(key)
        if c is None:
            c = tuple(torch.from_numpy(a).to(self.device) for a in pil_resize_coeffs(in_size, out_size))
            if len(self._coeffs) > 256:  # datasets hold a handful of distinct image sizes; bound the cache anyway
                self._coeffs.clear()
            self._coeffs[key] = c
        return c

    def crop_params(self) -> Tuple[int, int]:
        """torchvision.transforms.RandomCrop.get_params on the resized image: same generator, same order of draws."""
        h, w = self.size
        th, tw = self.crop
        if h == th and w == tw:
            return 0, 0
        i = int(torch.randint(0, h - th + 1, size=(1,)).item())
        j = int(torch.randint(0, w - tw + 1, size=(1,)).item())
        return i, j

    def __call__(self, img, out: Optional[torch.Tensor] = None, top_left: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        if isinstance(img, np.ndarray):
            img = torch.from_numpy(np.ascontiguousarray(img))
        if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3:
            raise ValueError("GpuTrainTransform: expected an RGB uint8 image [H, W, 3]")
        img = img.to(self.device, non_blocking=True).contiguous()
        H, W = int(img.shape[0]), int(img.shape[1])
        top, left = top_left if top_left is not None else self.crop_params()
        if out is None:
            out = torch.empty(3, self.crop[0], self.crop[1], dtype=torch.float32, device=self.device)
        ops.resize_crop_normalize(img, self.coeffs(W, self.size[1]), self.coeffs(H, self.size[0]), top, left, out, self.mean, self.std)
        return out
