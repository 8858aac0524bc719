"""vessel_graph_generation.utilities of the reference as far as scripts import it (utilities.py:17-38)."""
from ..config import read_config  # noqa: F401
from ..generate_vessel_graph import prepare_output_dir  # noqa: F401
