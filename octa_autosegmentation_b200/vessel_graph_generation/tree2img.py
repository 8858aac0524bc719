"""vessel_graph_generation.tree2img of the reference: the same function names over the GPU rasterizers (../tree2img.py)."""
from ..tree2img import rasterize_forest, voxelize_forest  # noqa: F401
from ..stats_plots import plot_vessel_radii  # noqa: F401


def save_2d_img(img, out_dir: str, name: str):
    """tree2img.py:282-292."""
    import numpy as np
    from .. import graph_io
    graph_io.save_png(f"{out_dir}/{name}.png", np.asarray(img).astype(np.uint8))
