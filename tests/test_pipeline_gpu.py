"""GPU: the batched pipeline (growth -> edge rows -> voxelize -> 2-D rasters -> CSV) and its software-pipelined variant
must give identical results; a sample's result depends on its seed only."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def small_config():
    from octa_autosegmentation_b200.config import default_config
    cfg = default_config()
    for m, i in zip(cfg["Greenhouse"]["modes"], (10, 10)):
        m["I"], m["N"] = i, 400
    return cfg


def digest(out):
    vol = out["volume_host"] if "volume_host" in out else out["volume"].cpu().numpy()
    return (tuple(hashlib.sha256(c).hexdigest() for c in out["csv"]), hashlib.sha256(vol.tobytes()).hexdigest(),
            hashlib.sha256(out["label_host"].tobytes()).hexdigest(), hashlib.sha256(out["image_host"].tobytes()).hexdigest())


def test_pipelined_batches_equal_single_steps():
    from octa_autosegmentation_b200.pipeline import Pipeline
    dims = (304, 304, 16)
    batches = [[11, 12, 13], [21, 22, 23], [31, 32, 33], [41, 42, 43]]
    pipe = Pipeline(small_config(), volume_dims=dims, label_res=(304, 304), image_res=(152, 152))
    want = [digest(pipe.run(b)) for b in batches]
    got = [digest(o) for o in pipe.run_pipelined(batches, d2h_volume=True)]      # (host mode recycles the device buffers)
    assert got == want
    # the oracle pins the first sample of the first batch (growth + voxelizer)
    from oracle import growth_oracle as go
    from oracle import vox_oracle
    oa, ov, _ = go.run(small_config(), 11)
    e7 = np.concatenate([oa, ov])
    first = pipe.run(batches[0])
    assert np.array_equal(first["volume"][0].cpu().numpy(), vox_oracle.voxelize_edges(e7, list(dims)))
