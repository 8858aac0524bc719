"""GPU: the reference's `main` (generate_vessel_graph.py:24-86) written against the object API of
octa_autosegmentation_b200/vessel_graph_generation/ and run on the real engine: CSV bytes, per-step lists and sink lists equal
the goldens the unmodified reference wrote; the rasterizers take the edge dicts of that script as they are."""
import numpy as np
import pytest

from test_object_api import check_against_goldens, reference_main_csv, small_config

pytestmark = pytest.mark.gpu


def test_reference_main_on_the_object_api_gpu():
    from octa_autosegmentation_b200.vessel_graph_generation import tree2img
    from oracle import agg_oracle, vox_oracle
    data, greenhouse, art, ven, art_edges, ven_edges = reference_main_csv(small_config(), 0)
    check_against_goldens(data, greenhouse, art, ven)
    # generate_vessel_graph.py:43,69-86 on those edge dicts
    volume_dimension = [int(d) for d in greenhouse.simspace.shape * 152]
    radius_list = []
    art_mat, _ = tree2img.voxelize_forest(art_edges, volume_dimension, radius_list)
    ven_mat, _ = tree2img.voxelize_forest(ven_edges, volume_dimension, radius_list)
    e_art = np.array([[*e["node1"], *e["node2"], e["radius"]] for e in art_edges])
    e_ven = np.array([[*e["node1"], *e["node2"], e["radius"]] for e in ven_edges])
    assert np.array_equal(art_mat, vox_oracle.voxelize_edges(e_art, volume_dimension))
    assert np.array_equal(ven_mat, vox_oracle.voxelize_edges(e_ven, volume_dimension))
    assert len(radius_list) == len(art_edges) + len(ven_edges)
    image_res = [*volume_dimension]
    del image_res[2]
    img, _ = tree2img.rasterize_forest(art_edges, image_res, MIP_axis=2, radius_list=[])
    assert np.array_equal(img, agg_oracle.raster_edges(e_art, image_res))
