"""CPU: host-side logic of the package (config surface, CLI overrides, sharding, multi-process gather)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import yaml

from conftest import ROOT
from octa_autosegmentation_b200 import config as cfgmod
from octa_autosegmentation_b200 import growth, pipeline


def test_default_config_shape_and_overrides():
    c = cfgmod.default_config()
    assert set(c) == {"Greenhouse", "output", "Forest"} and len(c["Greenhouse"]["modes"]) == 2
    cfgmod.apply_cli_overrides_from_unknown_args(c, ["--output.directory", "/tmp/x", "--Greenhouse.param_scale=6",
                                                     "--output.save_stats", "--threads", "3", "--Forest.N_trees", "16"])
    assert c["output"]["directory"] == "/tmp/x" and c["Greenhouse"]["param_scale"] == 6
    assert c["output"]["save_stats"] is True and c["Forest"]["N_trees"] == 16 and "threads" not in c


def test_read_config_roundtrip(tmp_path):
    p = tmp_path / "c.yml"
    p.write_text(yaml.dump(cfgmod.default_config()))
    assert cfgmod.read_config(str(p)) == cfgmod.default_config()


def test_make_config_mirrors_reference_flags():
    c = cfgmod.default_config()
    g = growth.make_config(c)
    assert g.n_modes == 2 and g.modes[0].first_mode == 1 and g.modes[1].first_mode == 0
    assert g.modes[0].reinit == 0 and g.modes[1].reinit == 1
    assert [g.walls[i] for i in range(g.n_walls)] == [0, 1, 2, 3]
    c["Forest"]["type"] = "grid"
    with pytest.raises(NotImplementedError):
        growth.make_config(c)


def test_round_robin_sharding_covers_every_sample_once():
    for world in (1, 2, 4, 8):
        allseeds = sorted(s for r in range(world) for s in pipeline.shard_seeds(100, 37, r, world))
        assert allseeds == list(range(100, 137))
    assert pipeline.shard_seeds(0, 10, 1, 4) == [1, 5, 9]


def test_two_rank_gloo_gather_of_edge_tables():
    """N > 1 host path: ranks own disjoint samples and rank 0 gathers the packed edge tables (gloo, CPU)."""
    script = os.path.join(ROOT, "tests", "_gloo_gather_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", script],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GATHER_OK" in r.stdout


def test_gan_host_helpers():
    """Host-side pieces of the GAN path that need no device: weight order of the C ABI, random-init shapes, flop count,
    resolution parsing of the reference's Test.data_augmentation stanza, buffer-set count of the pipelined API."""
    from octa_autosegmentation_b200 import gan, test as gan_cli
    from octa_autosegmentation_b200.pipeline import Pipeline
    from oracle import gan_oracle as go

    keys = gan.conv_keys()
    assert keys[:3] == ["model.4", "model.8", "model.12.conv_block.1"] and keys[-3:] == ["model.20.conv_block.5", "model.22", "model.26"]
    a, b = gan.random_init_state_dict(3), go.random_state_dict(3)
    assert set(a) == set(b) and all(tuple(a[k].shape) == tuple(b[k].shape) for k in a)
    assert abs(gan.conv_flops_per_image(304, 304) / 1e9 - 177.2) < 0.1
    cfg = {"Test": {"data_augmentation": [{"name": "LoadImaged"}, {"name": "LoadGraphAndFilterByRandomRadiusd", "image_resolutions": [[608, 304]],
                                                                   "min_radius": [0.002]}]}}
    assert gan_cli._resolution(cfg) == (608, 304, 0.002)
    assert gan_cli._resolution({"Test": {}}) == (304, 304, 0.0)
    assert Pipeline.buffer_sets(8) == 10 and Pipeline.buffer_sets(0) == 3 and Pipeline.buffer_sets(8, True) == 21


def test_png_writer_stores_the_pixels_pil_stores(tmp_path):
    """graph_io.png_bytes (zlib level 1, filter 0) against PIL: same pixels for 8-bit gray and 1-bit images, incl. odd widths."""
    import io
    import numpy as np
    from PIL import Image
    from octa_autosegmentation_b200 import graph_io
    rs = np.random.RandomState(0)
    for shape in [(1216, 1216), (304, 304), (7, 13), (1, 1)]:
        gray = (rs.rand(*shape) * 255 * (rs.rand(*shape) > 0.5)).astype(np.uint8)
        back = np.array(Image.open(io.BytesIO(graph_io.png_bytes(gray))))
        assert back.dtype == np.uint8 and np.array_equal(back, gray)
        bits = np.array(Image.fromarray(gray).convert("1"))
        im = Image.open(io.BytesIO(graph_io.png_bytes(bits)))
        assert im.mode == "1" and np.array_equal(np.array(im), bits)
    p = tmp_path / "x.png"
    graph_io.save_png(str(p), gray)
    assert np.array_equal(np.array(Image.open(p)), gray)


def test_bench_configs_module_and_buffer_sets():
    import bench_configs
    from octa_autosegmentation_b200.pipeline import Pipeline
    assert all(callable(getattr(bench_configs, "run_config%d" % k)) for k in (3, 4, 5))
    assert Pipeline.buffer_sets(7, True, 1) == 9 and Pipeline.buffer_sets(7, False) == 9 and Pipeline.buffer_sets(7, True) == 20


def test_nifti1_writer_fields_and_roundtrip(tmp_path):
    """graph_io.save_nifti stands in for nib.save(nib.Nifti1Image(vol, np.eye(4)), ...) (generate_vessel_graph.py:75-77,
    visualize_vessel_graphs.py:85-87).  No nibabel here, so the check is the NIfTI-1 standard itself: field offsets and values of
    the 348-byte header, Fortran voxel order, gzip container."""
    import gzip
    import struct
    from octa_autosegmentation_b200 import graph_io
    rng = np.random.default_rng(3)
    vol = rng.integers(0, 65535, size=(7, 5, 3), dtype=np.uint16)
    p = str(tmp_path / "v.nii.gz")
    graph_io.save_nifti(p, vol)
    raw = gzip.open(p, "rb").read()
    assert len(raw) == 352 + vol.size * 2
    assert struct.unpack_from("<i", raw, 0)[0] == 348                                  # sizeof_hdr
    assert struct.unpack_from("<8h", raw, 40) == (3, 7, 5, 3, 1, 1, 1, 1)              # dim
    assert struct.unpack_from("<hh", raw, 70) == (512, 16)                             # datatype uint16, bitpix
    assert struct.unpack_from("<8f", raw, 76) == (1.0,) * 8                            # pixdim, qfac = 1
    assert struct.unpack_from("<f", raw, 108)[0] == 352.0                              # vox_offset
    assert all(np.isnan(struct.unpack_from("<2f", raw, 112)))                          # scl_slope / scl_inter: no scaling
    assert struct.unpack_from("<2h", raw, 252) == (0, 2)                               # qform unknown, sform aligned
    assert struct.unpack_from("<12f", raw, 280) == (1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0)  # srow_x/y/z = identity
    assert raw[344:348] == b"n+1\0" and raw[348:352] == b"\0\0\0\0"
    # Fortran order: the first axis runs fastest
    assert np.array_equal(np.frombuffer(raw, dtype="<u2", offset=352, count=7), vol[:, 0, 0])
    assert np.array_equal(graph_io.load_nifti(p), vol)
    u8 = rng.integers(0, 255, size=(4, 6, 2), dtype=np.uint8)
    graph_io.save_nifti(str(tmp_path / "u.nii"), u8)
    raw = open(tmp_path / "u.nii", "rb").read()
    assert struct.unpack_from("<hh", raw, 70) == (2, 8) and np.array_equal(graph_io.load_nifti(str(tmp_path / "u.nii")), u8)
    with pytest.raises(ValueError):                # nibabel refuses bool arrays as well
        graph_io.nifti1_bytes(np.zeros((2, 2, 2), dtype=bool))
