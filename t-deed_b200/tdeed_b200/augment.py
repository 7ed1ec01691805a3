"""Per-clip training augmentation of the reference (model/model.py:77-84,154-157) on the fused kernels of train_aug.cu.

The reference applies, to every clip of the batch independently (`Impl.augment` loops over `x[i]`):
    RandomApply([ColorJitter(hue=0.2)], 0.25) -> RandomApply([ColorJitter(saturation=(0.7, 1.2))], 0.25)
    -> RandomApply([ColorJitter(brightness=(0.7, 1.2))], 0.25) -> RandomApply([ColorJitter(contrast=(0.7, 1.2))], 0.25)
    -> RandomApply([GaussianBlur(5)], 0.25) -> RandomHorizontalFlip()
on float frames in [0, 1].  `ClipAugment.sample()` draws the same decisions / factors (torch's CPU generator, like torchvision);
`apply()` runs them as at most three kernel passes per clip (crop and the x/255 scaling ride on the first one).
"""
import ctypes

import torch

from . import _lib as L


class ClipAugment:
    P_APPLY = 0.25
    HUE = (-0.2, 0.2)
    SATURATION = (0.7, 1.2)
    BRIGHTNESS = (0.7, 1.2)
    CONTRAST = (0.7, 1.2)
    SIGMA = (0.1, 2.0)
    P_FLIP = 0.5

    @classmethod
    def sample(cls):
        """One clip's parameters: dict(hue, sat, bri, con, sigma: float | None, flip: bool)."""
        def maybe(lo_hi):
            if float(torch.rand(1)) < cls.P_APPLY:
                return float(torch.empty(1).uniform_(lo_hi[0], lo_hi[1]))
            return None
        return dict(hue=maybe(cls.HUE), sat=maybe(cls.SATURATION), bri=maybe(cls.BRIGHTNESS), con=maybe(cls.CONTRAST),
                    sigma=maybe(cls.SIGMA), flip=bool(float(torch.rand(1)) < cls.P_FLIP))

    @staticmethod
    def gaussian_kernel1d(sigma, ksize=5):
        """torchvision.transforms._functional_tensor._get_gaussian_kernel1d (fp32)."""
        half = (ksize - 1) * 0.5
        x = torch.linspace(-half, half, steps=ksize, dtype=torch.float32)
        pdf = torch.exp(-0.5 * (x / sigma).pow(2))
        return (pdf / pdf.sum()).tolist()

    @classmethod
    def apply(cls, frames, crop, params, out=None):
        """frames (B,T,3,H,W) uint8 | float valued 0..255 on the device; crop (cy, cx, h, w); params: list of B dicts (sample()).
        -> float32 (B,T,3,h,w) in [0, 1] (the state the reference's pipeline leaves the clip in before Normalize)."""
        lib = L.load()
        b, t, _, in_h, in_w = frames.shape
        cy, cx, h, w = crop
        dev = frames.device
        if out is None:
            out = torch.empty((b, t, 3, h, w), dtype=torch.float32, device=dev)
        tmp = None
        mean = None
        st = L.stream()
        for i, p in enumerate(params):
            src = frames[i]
            dst = out[i]
            second = p['con'] is not None or p['sigma'] is not None or p['flip']
            if second and tmp is None:
                tmp = torch.empty((t, 3, h, w), dtype=torch.float32, device=dev)
            first_out = tmp if second else dst
            L.check(lib.tdeed_aug_color(L.ptr(src), L.dtype_code(src.dtype), 1.0 / 255.0, t, in_h, in_w, cy, cx, h, w,
                                        int(p['hue'] is not None), p['hue'] or 0.0, int(p['sat'] is not None), p['sat'] or 0.0,
                                        int(p['bri'] is not None), p['bri'] or 0.0, L.ptr(first_out), st), 'aug_color')
            if not second:
                continue
            if p['con'] is not None:
                if mean is None:
                    mean = torch.empty(t, dtype=torch.float32, device=dev)
                L.check(lib.tdeed_aug_gray_mean(L.ptr(tmp), t, h * w, L.ptr(mean), st), 'aug_gray_mean')
            k = (ctypes.c_float * 5)(*(cls.gaussian_kernel1d(p['sigma']) if p['sigma'] is not None else [0.0] * 5))
            L.check(lib.tdeed_aug_contrast_blur_flip(L.ptr(tmp), t, h, w, int(p['con'] is not None), p['con'] or 0.0, L.ptr(mean),
                                                     int(p['sigma'] is not None), k, int(p['flip']), L.ptr(dst), st),
                    'aug_contrast_blur_flip')
        return out
