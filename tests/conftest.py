import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_graph_rows(name):
    """Rows exactly as every reference consumer reads them (csv.DictReader, string cells)."""
    import csv
    import gzip
    import io
    p = os.path.join(GOLDEN, name)
    if name.endswith(".gz"):
        text = gzip.open(p, "rt", newline="").read()
    else:
        text = open(p, newline="").read()
    return list(csv.DictReader(io.StringIO(text, newline="")))


def rows_to_edges7(rows):
    import numpy as np
    out = np.empty((len(rows), 7))
    for i, r in enumerate(rows):
        out[i, 0:3] = [float(c) for c in r["node1"][1:-1].split(" ") if c]
        out[i, 3:6] = [float(c) for c in r["node2"][1:-1].split(" ") if c]
        out[i, 6] = float(r["radius"])
    return out
