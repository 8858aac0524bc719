"""GAN contrast adaptation of synthetic rasters (BASELINE config #5), host side over the C ABI.

Mirrors the reference's inference path for `docker/trained_models/GAN/config.yml`:

    test.py:58-90  ->  GanSegModel.inference (models/gan_seg_model.py:65-79)  ->  resnetGenerator9 (models/networks.py:502)
    transforms     ->  ScaleIntensityd, Rotate90d/Flipd (background), AddRandomBackgroundNoised (data/data_transforms.py:498-516)
    writer         ->  utils/visualizer.py:330-338  (`G_<name>.png`, uint8(pred * 255))

`ResnetGenerator9` loads the reference's checkpoint unchanged (same state-dict keys) and runs on the GPU through
`octa_gan_forward_dev` (csrc/octa_gan.cu: tcgen05 3x3 convolutions, bf16 activations, fp32 accumulation).  There is no CPU
path: without the CUDA library / a device the calls raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import numpy as np

from . import _lib, graph_io

STEM, DOWN, UP, HEAD = "model.1", ("model.4", "model.8"), ("model.22", "model.26"), "model.30"
BLOCKS = tuple("model.%d" % i for i in range(12, 21))


class OctaGanWeights(ctypes.Structure):
    _fields_ = [("stem_w", ctypes.c_void_p), ("conv_w", ctypes.c_void_p * 22), ("head_w", ctypes.c_void_p),
                ("head_b", ctypes.c_float), ("reserved", ctypes.c_int)]


def _bind():
    L = _lib.lib()
    L.octa_gan_create.argtypes = [ctypes.POINTER(OctaGanWeights), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  ctypes.POINTER(ctypes.c_void_p)]
    L.octa_gan_forward_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p]
    L.octa_gan_destroy.argtypes = [ctypes.c_void_p]
    L.octa_gan_destroy.restype = None
    L.octa_gan_speckle_dev.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    L.octa_gan_input_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.octa_test_gan_conv3_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return L


def conv_keys() -> list:
    """The 22 3x3 convolutions in the order of OctaGanWeights.conv_w (include/octa_b200.h)."""
    keys = list(DOWN)
    for b in BLOCKS:
        keys += [b + ".conv_block.1", b + ".conv_block.5"]
    return keys + list(UP)


def random_init_state_dict(seed: int = 0) -> dict:
    """Random-init weights with the reference's state-dict keys and shapes (benchmarks / smoke tests without a checkpoint)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    shapes = {STEM: (64, 1, 7, 7), DOWN[0]: (128, 64, 3, 3), DOWN[1]: (256, 128, 3, 3), UP[0]: (128, 256, 3, 3),
              UP[1]: (64, 128, 3, 3), HEAD: (1, 64, 7, 7)}
    for b in BLOCKS:
        shapes[b + ".conv_block.1"] = shapes[b + ".conv_block.5"] = (256, 256, 3, 3)
    sd = {}
    for k, s in shapes.items():
        sd[k + ".weight"] = torch.randn(*s, generator=g) * (2.0 / (s[1] * s[2] * s[3])) ** 0.5
        sd[k + ".bias"] = torch.zeros(s[0])
    return sd


def conv_flops_per_image(H: int, W: int) -> int:
    """Algorithmic flops of the 22 3x3 convolutions of one image: 2 * 9 * sum(pixels * Cin * Cout)."""
    return 2 * 9 * (H * W * 64 * 128 + (H // 2) * (W // 2) * 128 * 256 + 18 * (H // 4) * (W // 4) * 256 * 256
                    + (H // 2) * (W // 2) * 256 * 128 + H * W * 128 * 64)


class ResnetGenerator9:
    """`resnetGenerator9()` of models/networks.py:502-503 for inference: x float32 [N,1,H,W] in [0,1] -> [N,1,H,W]."""

    def __init__(self, state_dict: dict, image_size: Sequence[int] = (304, 304), max_images: int = 32, device=None):
        import torch

        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.H, self.W = int(image_size[0]), int(image_size[1])
        self.max_images = int(max_images)
        sd = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in state_dict.items() if hasattr(v, "detach")}
        need = [STEM + ".weight", HEAD + ".weight", HEAD + ".bias"] + [k + ".weight" for k in conv_keys()]
        missing = [k for k in need if k not in sd]
        if missing:
            raise KeyError("state dict lacks %s" % missing)
        exp = {STEM + ".weight": (64, 1, 7, 7), HEAD + ".weight": (1, 64, 7, 7)}
        for k, shp in exp.items():
            if tuple(sd[k].shape) != shp:
                raise ValueError("%s has shape %s, expected %s" % (k, tuple(sd[k].shape), shp))
        w = OctaGanWeights()
        keep = []
        w.stem_w = sd[STEM + ".weight"].data_ptr()
        w.head_w = sd[HEAD + ".weight"].data_ptr()
        w.head_b = float(sd[HEAD + ".bias"][0])
        for i, k in enumerate(conv_keys()):
            t = sd[k + ".weight"]
            if t.shape[2:] != (3, 3):
                raise ValueError("%s.weight is not a 3x3 kernel" % k)
            keep.append(t)
            w.conv_w[i] = t.data_ptr()
        L = _bind()
        self._L = L
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.octa_gan_create(ctypes.byref(w), self.max_images, self.H, self.W, ctypes.byref(self._h)))
        del keep

    @classmethod
    def from_checkpoint(cls, path: str, **kw):
        """Loads `<epoch>_G_model.pth` as written by the reference ({'epoch', 'model': state_dict}; utils/visualizer.py)."""
        import torch

        ck = torch.load(path, map_location="cpu", weights_only=False)
        return cls(ck["model"] if isinstance(ck, dict) and "model" in ck else ck, **kw)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.octa_gan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, x, out=None, out_u8=None, stream=None):
        """x: CUDA float32 [N,1,H,W] or [N,H,W].  Returns float32 like x; `out_u8` (CUDA uint8 [N,H,W]) also receives
        uint8(pred * 255), the pixels of the reference's PNG."""
        torch = self.torch
        if not (x.is_cuda and x.dtype == torch.float32):
            raise ValueError("x must be a CUDA float32 tensor")
        x = x.contiguous()
        n = x.shape[0]
        if tuple(x.shape[-2:]) != (self.H, self.W) or x.numel() != n * self.H * self.W:
            raise ValueError("x must be [N,1,%d,%d]" % (self.H, self.W))
        if out is None:
            out = torch.empty_like(x)
        if stream is None:
            stream = torch.cuda.current_stream(x.device)
        with torch.cuda.device(x.device):
            for i in range(0, n, self.max_images):
                m = min(self.max_images, n - i)
                _lib.check(self._L.octa_gan_forward_dev(self._h, x[i:i + m].data_ptr(), m, out[i:i + m].data_ptr(),
                                                         out_u8[i:i + m].data_ptr() if out_u8 is not None else None,
                                                         ctypes.c_void_p(stream.cuda_stream)))
        return out

    __call__ = forward


def speckle_device(seeds: Sequence[int], H: int, W: int, device=None, stream=None):
    """float64 CUDA tensor [N,H,W]: row i = `np.random.seed(seeds[i]); np.random.uniform(0, 1, (H, W))`, generated on the GPU."""
    import torch

    L = _bind()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    s = torch.from_numpy(np.array([int(v) & 0xFFFFFFFF for v in seeds], dtype=np.uint32).view(np.int32)).to(device)
    out = torch.empty((len(seeds), H, W), dtype=torch.float64, device=device)
    if stream is None:
        stream = torch.cuda.current_stream(device)
    with torch.cuda.device(device):
        _lib.check(L.octa_gan_speckle_dev(s.data_ptr(), len(seeds), H, W, out.data_ptr(), ctypes.c_void_p(stream.cuda_stream)))
    return out


def prepare_input(raster_u8, background_u8=None, speckle=None, out=None, stream=None):
    """ScaleIntensityd + background Rotate90d/Flipd + AddRandomBackgroundNoised + CastToTyped on the GPU.
    raster_u8 / background_u8: CUDA uint8 [N,H,W]; speckle: CUDA float64 [N,H,W].  Returns CUDA float32 [N,1,H,W]."""
    import torch

    L = _bind()
    if not (raster_u8.is_cuda and raster_u8.dtype == torch.uint8 and raster_u8.dim() == 3):
        raise ValueError("raster_u8 must be a CUDA uint8 tensor [N,H,W]")
    raster_u8 = raster_u8.contiguous()
    n, H, W = raster_u8.shape
    if (background_u8 is None) != (speckle is None):
        raise ValueError("background_u8 and speckle go together")
    if background_u8 is not None:
        background_u8 = background_u8.contiguous()
        speckle = speckle.contiguous()
        if tuple(background_u8.shape) != (n, H, W) or tuple(speckle.shape) != (n, H, W) or speckle.dtype != torch.float64:
            raise ValueError("background_u8 uint8 [N,H,W] and speckle float64 [N,H,W] expected")
    if out is None:
        out = torch.empty((n, 1, H, W), dtype=torch.float32, device=raster_u8.device)
    mm = torch.empty((n, 4), dtype=torch.int32, device=raster_u8.device)
    if stream is None:
        stream = torch.cuda.current_stream(raster_u8.device)
    with torch.cuda.device(raster_u8.device):
        _lib.check(L.octa_gan_input_dev(raster_u8.data_ptr(), background_u8.data_ptr() if background_u8 is not None else None,
                                        speckle.data_ptr() if speckle is not None else None, n, H, W, mm.data_ptr(),
                                        out.data_ptr(), ctypes.c_void_p(stream.cuda_stream)))
    return out


def conv3_test(x: np.ndarray, w: np.ndarray, reflect: bool) -> np.ndarray:
    """Test hook: one 3x3 convolution through the tcgen05 kernel (host arrays in torch layouts)."""
    L = _bind()
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    n, cin, H, W = x.shape
    cout = w.shape[0]
    y = np.empty((n, cout, H, W), dtype=np.float32)
    _lib.check(L.octa_test_gan_conv3_host(x.ctypes.data, w.ctypes.data, n, H, W, cin, cout, int(bool(reflect)), y.ctypes.data))
    return y


def contrast_adapt(generator: ResnetGenerator9, raster_u8, background_u8=None, speckle_seeds: Optional[Sequence[int]] = None):
    """Config #5 tail: uint8 rasters [N,H,W] (CUDA) -> uint8 GAN images [N,H,W] (CUDA), the pixels of `G_<name>.png`."""
    import torch

    n, H, W = raster_u8.shape
    sp = None
    if background_u8 is not None:
        if speckle_seeds is None:
            raise ValueError("speckle_seeds are required with a background")
        sp = speckle_device(speckle_seeds, H, W, device=raster_u8.device)
    x = prepare_input(raster_u8, background_u8, sp)
    out8 = torch.empty((n, H, W), dtype=torch.uint8, device=raster_u8.device)
    generator.forward(x, out_u8=out8)
    return out8


def save_images(out_dir: str, names: Sequence[str], images_u8: np.ndarray, prefix: str = "G_"):
    """utils/visualizer.py:330-338 + test.py:85: `<save_dir>/<inference>_<csv name without extension>.png`."""
    from PIL import Image

    os.makedirs(out_dir, exist_ok=True)
    for name, img in zip(names, images_u8):
        stem = ".".join(os.path.basename(name).split(".")[:-1]) or os.path.basename(name)
        graph_io.save_png(os.path.join(out_dir, prefix + stem + ".png"), np.asarray(img, dtype=np.uint8))
