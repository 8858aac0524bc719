"""CPU: the GAN oracle (oracle/gan_oracle.py) is pinned to the reference.

In the build container the reference's own classes are parsed out of models/networks.py (unmodified source, only the
definitions the generator needs -- the module itself imports MONAI, which is not installed) and run with the shipped
checkpoint; the restatement must give the same tensor.  On hosts without /root/reference the test is skipped (the GPU
parity tests then rely on the oracle as pinned here).  Also CPU-side checks of the host wrappers' argument handling.
"""
import ast
import functools
import os

import numpy as np
import pytest

REF = "/root/reference"
CKPT = os.path.join(REF, "docker/trained_models/GAN/checkpoints/150_G_model.pth")


def _reference_generator():
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    tree = ast.parse(open(os.path.join(REF, "models/networks.py")).read())
    want = {"get_filter", "get_pad_layer", "get_norm_layer", "Identity", "Upsample", "Downsample", "ResnetBlock",
            "ResnetGenerator", "resnetGenerator9"}
    ns = {"torch": torch, "nn": nn, "np": np, "functools": functools, "F": F}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in want:
            exec(compile(ast.Module([node], []), "networks.py", "exec"), ns)
    return ns["resnetGenerator9"]()


@pytest.mark.skipif(not os.path.exists(CKPT), reason="reference checkout not present on this host")
def test_oracle_equals_reference_generator_with_shipped_checkpoint():
    import torch
    from oracle import gan_oracle as go

    G = _reference_generator()
    ck = torch.load(CKPT, map_location="cpu", weights_only=False)
    assert not any(G.load_state_dict(ck["model"]))           # no missing / unexpected keys
    G.eval()
    x = torch.rand(2, 1, 64, 64, generator=torch.Generator().manual_seed(0))
    xr = torch.rand(1, 1, 48, 80, generator=torch.Generator().manual_seed(1))      # non-square: pads / resampling per axis
    with torch.no_grad():
        assert torch.equal(G(x), go.generator_forward(ck["model"], x))
        assert torch.equal(G(xr), go.generator_forward(ck["model"], xr))
        sd = go.random_state_dict(1)
        G.load_state_dict(sd, strict=False)                   # (the blur filters are buffers, not in the synthetic dict)
        assert torch.equal(G(x), go.generator_forward(sd, x))


def test_oracle_input_transform_semantics():
    from oracle import gan_oracle as go

    img = np.array([[0, 10], [20, 40]], dtype=np.uint8)
    bg = np.array([[0, 100], [200, 50]], dtype=np.uint8)
    sp = np.full((2, 2), 0.5)
    x = go.prepare_input(img, bg, sp)
    # background is transposed (Rotate90d(k=1) then Flipd(0)), scaled to [0,1], multiplied by the speckle
    exp = np.maximum(img / 40.0, (bg.T / 200.0) * 0.5).astype(np.float32)
    assert np.allclose(x, exp, atol=1e-7)
    assert np.array_equal(go.scale_intensity(np.full((3, 3), 9, np.uint8)), np.zeros((3, 3), np.float32))
    np.random.seed(5)
    assert np.array_equal(go.speckle(5, (4, 4)), np.random.uniform(0, 1, (4, 4)))


def test_generator_wrapper_rejects_bad_state_dict():
    from octa_autosegmentation_b200 import gan
    from oracle import gan_oracle as go

    sd = go.random_state_dict(0)
    assert len(gan.conv_keys()) == 22 and gan.conv_keys()[0] == "model.4" and gan.conv_keys()[-1] == "model.26"
    bad = dict(sd)
    del bad["model.15.conv_block.5.weight"]
    with pytest.raises(KeyError):
        gan.ResnetGenerator9(bad, device="cpu")


def test_background_keeps_its_own_orientation(tmp_path):
    """LoadImaged (PILReader, reverse_indexing=True) delivers the background transposed and the config's Rotate90d(k=1) +
    Flipd(0) transposes it back (docker/trained_models/GAN/config.yml:49-86): the noise the generator sees has the PNG's own
    orientation.  The product's loader (octa_autosegmentation_b200/test.py:load_background) must hand the device the same array."""
    from PIL import Image
    from octa_autosegmentation_b200 import test as gan_cli
    rs = np.random.RandomState(3)
    A = rs.randint(0, 255, (304, 304)).astype(np.uint8)
    A[:40, :] = 255                                   # a bright band along the TOP rows: not symmetric under transposition
    p = tmp_path / "bg.png"
    Image.fromarray(A).save(p)
    from oracle import gan_oracle as go
    loaded = go.load_image_like_monai(p)
    assert np.array_equal(loaded, A.T) and np.array_equal(gan_cli.load_background(str(p), 304, 304), loaded)
    x = go.prepare_input(np.zeros((304, 304), np.uint8), loaded, np.ones((304, 304)))
    assert np.array_equal(x, go.scale_intensity(A))   # effective background == the PNG as stored
