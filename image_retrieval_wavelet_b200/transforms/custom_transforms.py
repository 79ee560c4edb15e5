"""Drop-in mirror of the reference's SWT transform plugin, running on the B200 kernels.

Mirrors ``/root/reference/main/transforms/custom_transforms.py:126-188`` (``BaseWaveletTransform``, ``SWTTransform``,
``RawStackTransform``): same class names, constructor arguments, ``__call__`` contract (PIL RGB image in,
``torch.float32 [3, 4, H', W']`` out with bands LL, LH, HL, HH and values scaled by 1/255), ``fix_size`` and
``__repr__`` — so ``Getter.get_transform`` (main/getter.py:25-35) resolves them by name from the same YAML.

Added on top (SURVEY.md §7 hard part 6): a batched device entry ``forward(x)`` for ``[B, C, H, W]`` uint8/float32
CUDA tensors, meant to run right after ``batch["image"].cuda()`` (main/engine/base_update.py:65,
main/engine/evaluate.py:92-93) so that the 4x-inflated float32 sub-bands never cross PCIe.

No CPU fallback: without the CUDA library / a device these classes raise.
"""
import ctypes

import numpy as np
import torch

from .. import _cabi
from .wavelets import filter_bank

__all__ = ["BaseWaveletTransform", "SWTTransform", "RawStackTransform", "DWTTransform", "swt2", "dwt2", "resize_u8"]


def _filters(wavelet):
    lo, hi = filter_bank(wavelet)
    f = len(lo)
    return (ctypes.c_float * f)(*lo), (ctypes.c_float * f)(*hi), f


def swt2(x, wavelet="haar", level=1, out=None):
    """Batched stationary wavelet transform on the current CUDA device.

    ``x``: ``[..., H, W]`` uint8 (scaled by 1/255 like the reference) or float32 CUDA tensor, H and W divisible by
    ``2**level``.  Returns float32 ``[..., 4, H, W]`` = (cA, cH, cV, cD) of level ``level`` — ``pywt.swt2(...)[0]``.
    """
    _cabi.require_cuda()
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise TypeError("swt2 expects a CUDA tensor (use SWTTransform.__call__ for PIL images)")
    if x.dtype not in (torch.uint8, torch.float32):
        raise TypeError(f"swt2 expects uint8 or float32, got {x.dtype}")
    if x.dim() < 2:
        raise ValueError("swt2 expects [..., H, W]")
    level = int(level)
    h, w = int(x.shape[-2]), int(x.shape[-1])
    if level < 1:
        raise ValueError("level must be >= 1")
    if h % (1 << level) or w % (1 << level):
        # pywt.swt2 raises ValueError for sizes that are not a multiple of 2**level
        raise ValueError(f"swt2: H={h} and W={w} must be divisible by 2**level={1 << level}")
    lead = tuple(x.shape[:-2])
    planes = int(np.prod(lead)) if lead else 1
    xc = x.contiguous()
    if planes == 0:
        return torch.empty(lead + (4, h, w), dtype=torch.float32, device=x.device)
    if xc.data_ptr() % 16:
        xc = xc.clone()
    if out is None:
        out = torch.empty(lead + (4, h, w), dtype=torch.float32, device=x.device)
    elif out.shape != lead + (4, h, w) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != x.device:
        raise ValueError("out must be a contiguous float32 tensor of shape [..., 4, H, W] on x's device")
    lo, hi, f = _filters(wavelet)
    with torch.cuda.device(x.device), _cabi.nvtx_range("b200/swt2"):
        rc = _cabi.load().b200_swt2_fwd(_cabi.ptr(xc), int(x.dtype == torch.uint8), _cabi.ptr(out), planes, 1, h, w, lo, hi, f,
                                        level, _cabi.stream_ptr())
    _cabi.check(rc, "b200_swt2_fwd")
    return out


def resize_u8(x, size, resample="bicubic"):
    """PIL's ``Image.resize`` on the device, bit for bit: ``[..., H, W]`` uint8 CUDA -> ``[..., size[0], size[1]]`` uint8.

    ``resample``: ``"bicubic"`` (``fix_size``, custom_transforms.py:132-139) or ``"bilinear"`` (torchvision ``Resize`` on a PIL
    image), both with Pillow's antialiasing window and 22-bit fixed-point weights (``b200_resize_u8``)."""
    _cabi.require_cuda()
    if not isinstance(x, torch.Tensor) or not x.is_cuda or x.dtype != torch.uint8 or x.dim() < 2:
        raise TypeError("resize_u8 expects a [..., H, W] uint8 CUDA tensor")
    filt = {"bicubic": _cabi.RESIZE_BICUBIC, "bilinear": _cabi.RESIZE_BILINEAR}.get(resample)
    if filt is None:
        raise ValueError("resample must be 'bicubic' or 'bilinear'")
    h, w = int(x.shape[-2]), int(x.shape[-1])
    ho, wo = int(size[0]), int(size[1])
    if ho < 1 or wo < 1:
        raise ValueError("size must be positive")
    lead = tuple(x.shape[:-2])
    planes = int(np.prod(lead)) if lead else 1
    out = torch.empty(lead + (ho, wo), dtype=torch.uint8, device=x.device)
    if planes == 0 or h == 0 or w == 0:
        return out
    xc = x.contiguous()
    lib = _cabi.load()
    ws = torch.empty(max(int(lib.b200_resize_workspace_bytes(planes, h, w, ho, wo, filt)), 1), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.b200_resize_u8(_cabi.ptr(xc), _cabi.ptr(out), planes, h, w, ho, wo, filt, _cabi.ptr(ws), ws.numel(),
                                _cabi.stream_ptr())
    _cabi.check(rc, "b200_resize_u8")
    return out


class BaseWaveletTransform(object):
    """Shared pipeline: resize to fit the level, transform each RGB channel, stack into ``[3, S, H, W]``."""

    def __init__(self, level=1, wavelet="haar"):
        self.level = level
        self.wavelet = wavelet

    def fix_size(self, image):
        """custom_transforms.py:132-139 — PIL bicubic resize up to a multiple of ``2**level`` (host side)."""
        from PIL import Image

        w, h = image.size
        factor = 2 ** self.level
        new_w = int(np.ceil(w / factor) * factor)
        new_h = int(np.ceil(h / factor) * factor)
        if new_w != w or new_h != h:
            image = image.resize((new_w, new_h), resample=Image.BICUBIC)
        return image

    def fix_size_cuda(self, x):
        """``fix_size`` for a batch already on the device: ``[..., H, W]`` uint8 CUDA, resized with PIL's bicubic arithmetic
        (bit-identical to resizing every image on the host) when H or W is not a multiple of ``2**level``."""
        h, w = int(x.shape[-2]), int(x.shape[-1])
        factor = 2 ** self.level
        new_h, new_w = int(np.ceil(h / factor) * factor), int(np.ceil(w / factor) * factor)
        if (new_h, new_w) == (h, w):
            return x
        return resize_u8(x, (new_h, new_w), "bicubic")

    def _image_array(self, img):
        img = self.fix_size(img)
        arr = np.array(img)
        if arr.ndim != 3 or arr.shape[2] < 3:
            # the reference indexes img_np[:, :, c] for c in range(3)
            raise IndexError("expected an RGB image (H, W, 3)")
        return np.ascontiguousarray(arr[:, :, :3])

    def forward(self, x):
        raise NotImplementedError("Cette méthode doit être définie dans la sous-classe.")

    def _apply_wavelet(self, channel_pixels):
        """One channel, ``[H, W]`` float32 in [0, 1] -> numpy ``[S, H, W]`` (kept for API parity; runs on the GPU)."""
        _cabi.require_cuda()
        x = torch.as_tensor(np.ascontiguousarray(channel_pixels, dtype=np.float32)).cuda()
        return self.forward(x[None, None])[0, 0].cpu().numpy()

    def __call__(self, img):
        _cabi.require_cuda()
        arr = self._image_array(img)
        if arr.dtype != np.uint8:
            x = torch.as_tensor(np.ascontiguousarray(arr.astype(np.float32).transpose(2, 0, 1)) / np.float32(255.0)).cuda()
            return self.forward(x[None])[0].cpu()
        return self._call_u8_hwc(arr)

    def _call_u8_hwc(self, arr):
        x = torch.from_numpy(arr).cuda().permute(2, 0, 1).contiguous()
        return self.forward(x[None])[0].cpu()


class SWTTransform(BaseWaveletTransform):
    """Stationary wavelet transform (size preserved: H, W)."""

    def forward(self, x):
        """``[B, C, H, W]`` uint8/float32 CUDA -> float32 ``[B, C, 4, H', W']``.  uint8 batches whose size is not a multiple
        of ``2**level`` get ``fix_size`` on the device first (518 -> 520 at levels 2-3), exactly what ``__call__`` does to
        a PIL image; float32 batches must already have a valid size (PIL only resizes 8-bit images this way)."""
        if isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.uint8:
            x = self.fix_size_cuda(x)
        return swt2(x, self.wavelet, self.level)

    def _call_u8_hwc(self, arr):
        # one C call: H2D of the HWC bytes, de-interleave, SWT, D2H of the sub-bands
        h, w, c = arr.shape
        if h % (1 << self.level) or w % (1 << self.level):
            raise ValueError("image size must be divisible by 2**level after fix_size")
        lo, hi, f = _filters(self.wavelet)
        out = torch.empty((c, 4, h, w), dtype=torch.float32)
        rc = _cabi.load().b200_swt2_fwd_host(ctypes.c_void_p(arr.ctypes.data), 1, 1, _cabi.ptr(out), 1, c, h, w, lo, hi, f,
                                             int(self.level))
        _cabi.check(rc, "b200_swt2_fwd_host")
        return out

    def __repr__(self):
        return f"SWTTransform(shape='C,S,H,W', wavelet={self.wavelet}, level={self.level})"


class RawStackTransform(BaseWaveletTransform):
    """Parameter-matched control: the same ``[C, copies, H, W]`` layout, every 'sub-band' a copy of the channel."""

    def __init__(self, level=1, wavelet="haar", copies=4):
        super().__init__(level=level, wavelet=wavelet)
        self.copies = copies

    def forward(self, x):
        _cabi.require_cuda()
        if not x.is_cuda or x.dtype not in (torch.uint8, torch.float32) or x.dim() != 4:
            raise TypeError("RawStackTransform.forward expects a [B, C, H, W] uint8/float32 CUDA tensor")
        if x.dtype == torch.uint8:
            x = self.fix_size_cuda(x)                 # __call__ resizes every image first (custom_transforms.py:146)
        b, c, h, w = (int(v) for v in x.shape)
        xc = x.contiguous()
        out = torch.empty((b, c, int(self.copies), h, w), dtype=torch.float32, device=x.device)
        if out.numel() == 0:
            return out
        with torch.cuda.device(x.device):
            rc = _cabi.load().b200_raw_stack(_cabi.ptr(xc), int(x.dtype == torch.uint8), _cabi.ptr(out), b, c, h, w,
                                             int(self.copies), _cabi.stream_ptr())
        _cabi.check(rc, "b200_raw_stack")
        return out

    def __repr__(self):
        return f"RawStackTransform(shape='C,{self.copies},H,W', copies={self.copies})"


def dwt2(x, wavelet="haar", level=1):
    """Batched decimated wavelet transform on the current CUDA device: ``pywt.wavedec2(x, wavelet, level=level)`` in the
    default ``'symmetric'`` mode, coarsest level only.

    ``x``: ``[..., H, W]`` uint8 (scaled by 1/255) or float32 CUDA tensor.  Returns float32 ``[..., 4, H_L, W_L]`` =
    (cA, cH, cV, cD) with ``H_l = (H_{l-1} + F - 1) // 2`` (``b200_dwt2_fwd``)."""
    _cabi.require_cuda()
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise TypeError("dwt2 expects a CUDA tensor (use DWTTransform.__call__ for PIL images)")
    if x.dtype not in (torch.uint8, torch.float32):
        raise TypeError(f"dwt2 expects uint8 or float32, got {x.dtype}")
    if x.dim() < 2:
        raise ValueError("dwt2 expects [..., H, W]")
    level = int(level)
    if level < 1:
        raise ValueError("level must be >= 1")
    lo, hi, f = _filters(wavelet)
    h, w = int(x.shape[-2]), int(x.shape[-1])
    ho, wo = h, w
    for _ in range(level):
        ho, wo = (ho + f - 1) // 2, (wo + f - 1) // 2
    lead = tuple(x.shape[:-2])
    planes = int(np.prod(lead)) if lead else 1
    out = torch.empty(lead + (4, ho, wo), dtype=torch.float32, device=x.device)
    if planes == 0 or h == 0 or w == 0:
        return out
    xc = x.contiguous()
    lib = _cabi.load()
    ws = torch.empty(max(int(lib.b200_dwt2_workspace_bytes(planes, h, w, f, level)), 1), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.b200_dwt2_fwd(_cabi.ptr(xc), int(x.dtype == torch.uint8), _cabi.ptr(out), planes, h, w, lo, hi, f, level,
                               _cabi.ptr(ws), ws.numel(), _cabi.stream_ptr())
    _cabi.check(rc, "b200_dwt2_fwd")
    return out


class DWTTransform(BaseWaveletTransform):
    """Discrete multi-level wavelet transform (size divided by 2^level): ``pywt.wavedec2`` in its default symmetric mode,
    coarsest (cA, cH, cV, cD) only (custom_transforms.py:191-205; config/transform/cifar_dwt.yaml)."""

    def __init__(self, level=1, wavelet="haar"):
        super().__init__(level=level, wavelet=wavelet)

    def forward(self, x):
        """``[B, C, H, W]`` uint8/float32 CUDA -> float32 ``[B, C, 4, H_L, W_L]`` (uint8 batches get ``fix_size`` first)."""
        if isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.uint8:
            x = self.fix_size_cuda(x)
        return dwt2(x, self.wavelet, self.level)

    def __repr__(self):
        factor = 2 ** self.level
        return f"DWTTransform(shape='C,S,H/{factor},W/{factor}', wavelet={self.wavelet}, level={self.level})"
