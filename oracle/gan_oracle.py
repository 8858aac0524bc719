"""TEST INFRASTRUCTURE -- plain PyTorch fp32 restatement of the GAN contrast-adaptation inference path.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product
(octa_autosegmentation_b200/gan.py + csrc/octa_gan.cu) never does.

What is restated (file:line relative to the reference root):
  * `resnetGenerator9()`            models/networks.py:502-503 = ResnetGenerator(1, 1, ngf=64, instance norm, 9 blocks)
  * `ResnetGenerator.__init__`      models/networks.py:355-420 (reflect-pad 7x7 stem, two stride-1 3x3 convs each followed
                                    by an anti-aliased Downsample, 9 ResnetBlocks, two Upsample + 3x3 conv stages,
                                    reflect-pad 7x7 head, Sigmoid)
  * `ResnetBlock`                   models/networks.py:291-348 (reflect pad 1, conv, IN, ReLU, reflect pad 1, conv, IN, skip)
  * `Downsample` / `Upsample`       models/networks.py:244-289 (blur-pool [1,2,1]^2/16 stride 2 after reflect pad 1;
                                    replicate pad 1 + transposed conv with [1,3,3,1]^2/64*4, cropped = x2 bilinear)
  * `get_norm_layer('instance')`    models/networks.py:224-242 (InstanceNorm2d, affine=False, eps 1e-5, biased variance)
  * `AddRandomBackgroundNoised`     data/data_transforms.py:498-516, `ScaleIntensityd(minv=0, maxv=1)` (MONAI) and the
                                    uint8 PNG writer utils/visualizer.py:330-338

The state-dict key names are the reference's (`model.<i>.weight`, `model.<i>.conv_block.<j>.weight`), so the shipped
checkpoint docker/trained_models/GAN/checkpoints/150_G_model.pth loads into `generator_forward` unchanged.
Pinned by tests/test_oracle_gan.py: the reference's own classes (parsed out of models/networks.py in the build container,
unmodified) and this restatement give the same tensor with the shipped checkpoint.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# layer indices of nn.Sequential `model` in ResnetGenerator (networks.py:372-417) that carry parameters
STEM, DOWN1, DOWN2, UP1, UP2, HEAD = 1, 4, 8, 22, 26, 30
BLOCKS = list(range(12, 21))


def random_state_dict(seed: int, scale: float = 1.0) -> dict:
    """Seeded synthetic weights with the reference's shapes and key names (torch CPU generator: identical on every host)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(key, cout, cin, k):
        fan_in = cin * k * k
        sd[key + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * (scale * (2.0 / fan_in) ** 0.5)
        sd[key + ".bias"] = torch.randn(cout, generator=g) * 0.1

    conv("model.%d" % STEM, 64, 1, 7)
    conv("model.%d" % DOWN1, 128, 64, 3)
    conv("model.%d" % DOWN2, 256, 128, 3)
    for b in BLOCKS:
        conv("model.%d.conv_block.1" % b, 256, 256, 3)
        conv("model.%d.conv_block.5" % b, 256, 256, 3)
    conv("model.%d" % UP1, 128, 256, 3)
    conv("model.%d" % UP2, 64, 128, 3)
    conv("model.%d" % HEAD, 1, 64, 7)
    return sd


def _inorm(x):
    return F.instance_norm(x, eps=1e-5)


def _down(x):   # networks.py:266-289
    a = torch.tensor([1.0, 2.0, 1.0])
    f = a[:, None] * a[None, :]
    f = (f / f.sum())[None, None].repeat(x.shape[1], 1, 1, 1).to(x)
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), f, stride=2, groups=x.shape[1])


def _up(x):     # networks.py:244-264
    a = torch.tensor([1.0, 3.0, 3.0, 1.0])
    f = a[:, None] * a[None, :]
    f = (f / f.sum() * 4.0)[None, None].repeat(x.shape[1], 1, 1, 1).to(x)
    y = F.conv_transpose2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), f, stride=2, padding=2, groups=x.shape[1])
    return y[:, :, 1:, 1:][:, :, :-1, :-1]


def generator_forward(sd: dict, x: torch.Tensor) -> torch.Tensor:
    """x: float32 [N,1,H,W] in [0,1] -> float32 [N,1,H,W] in (0,1)."""
    w = lambda k: sd[k + ".weight"].to(x)
    b = lambda k: sd[k + ".bias"].to(x)
    k = "model.%d" % STEM
    h = F.relu(_inorm(F.conv2d(F.pad(x, (3, 3, 3, 3), mode="reflect"), w(k), b(k))))
    for idx in (DOWN1, DOWN2):
        k = "model.%d" % idx
        h = _down(F.relu(_inorm(F.conv2d(h, w(k), b(k), padding=1))))
    for blk in BLOCKS:
        k1, k2 = "model.%d.conv_block.1" % blk, "model.%d.conv_block.5" % blk
        t = F.relu(_inorm(F.conv2d(F.pad(h, (1, 1, 1, 1), mode="reflect"), w(k1), b(k1))))
        t = _inorm(F.conv2d(F.pad(t, (1, 1, 1, 1), mode="reflect"), w(k2), b(k2)))
        h = h + t
    for idx in (UP1, UP2):
        k = "model.%d" % idx
        h = F.relu(_inorm(F.conv2d(_up(h), w(k), b(k), padding=1)))
    k = "model.%d" % HEAD
    return torch.sigmoid(F.conv2d(F.pad(h, (3, 3, 3, 3), mode="reflect"), w(k), b(k)))


def scale_intensity(img: np.ndarray) -> np.ndarray:
    """MONAI ScaleIntensity(minv=0, maxv=1) on one image: (x - min) / (max - min); a constant image becomes zeros."""
    x = img.astype(np.float32)
    mn, mx = x.min(), x.max()
    if mx - mn == 0:
        return x * np.float32(0.0)
    return ((x - mn) / (mx - mn)).astype(np.float32)


def speckle(seed: int, shape) -> np.ndarray:
    """The draw of data_transforms.py:510 with the legacy global stream seeded right before it."""
    rs = np.random.RandomState(seed)
    return rs.uniform(0, 1, shape)


def load_image_like_monai(path) -> np.ndarray:
    """LoadImaged(image_only=True) on a PNG (config.yml:49-55): monai.data.PILReader with its default reverse_indexing=True
    returns the pixel array with the two spatial axes swapped (third-party behaviour, restated from MONAI's documentation:
    "reverse_indexing: whether to use a reversed spatial indexing convention for the returned data array ... default True").
    The config's Rotate90d(k=1) + Flipd(0) in prepare_input undo exactly that."""
    from PIL import Image
    return np.ascontiguousarray(np.asarray(Image.open(path).convert("L"), dtype=np.uint8).T)


def prepare_input(raster_u8: np.ndarray, background_u8: np.ndarray | None, speckle64: np.ndarray) -> np.ndarray:
    """ScaleIntensityd on both, background Rotate90d(k=1) + Flipd(axis 0) as in docker/trained_models/GAN/config.yml:60-86,
    then img = maximum(img, noise * speckle) (float32 x float64 -> float64) and CastToTyped(float32)."""
    img = scale_intensity(raster_u8)
    if background_u8 is None:
        raise ValueError("the oracle needs an explicit background (torch.rand_like is not reproducible across devices)")
    bg = scale_intensity(background_u8)
    bg = np.flip(np.rot90(bg, 1, axes=(0, 1)), axis=0)
    out = np.maximum(img.astype(np.float64), bg.astype(np.float64) * speckle64)
    return out.astype(np.float32)


def to_png_u8(pred: np.ndarray) -> np.ndarray:
    """utils/visualizer.py:338: (pred * 255).astype(uint8) -- float32 multiply, truncation."""
    return (pred.astype(np.float32) * 255).astype(np.uint8)
