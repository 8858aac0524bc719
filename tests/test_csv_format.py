"""CPU: the product's CSV writer/reader (csrc/octa_csv.cu, host code) against numpy's str(ndarray),
Python's repr(float) and csv.writer, and against the reference's own files."""
import csv
import gzip
import io
import os

import numpy as np

from conftest import GOLDEN, rows_to_edges7
from octa_autosegmentation_b200 import graph_io


def python_csv(e7):
    buf = io.StringIO(newline="")
    w = csv.writer(buf)
    w.writerow(["node1", "node2", "radius"])
    for row in e7:
        w.writerow([row[0:3], row[3:6], float(row[6])])
    return buf.getvalue().encode()


def test_matches_numpy_and_repr_on_random_rows():
    rng = np.random.RandomState(0)
    rows = []
    for _ in range(4000):
        scale = 10.0 ** rng.randint(-9, 3, 7)
        v = rng.uniform(-1, 1, 7) * scale
        if rng.rand() < 0.3:
            v[rng.randint(0, 6)] = 0.0
        if rng.rand() < 0.2:
            v[:6] = np.round(v[:6], rng.randint(0, 9))
        v[6] = abs(v[6])
        rows.append(v)
    # values the growth path produces: stump roots on the walls, leaf radius, tiny z
    rows += [[0.0, 0.87404801, 0.0031498, 1 - 1e-6, 0.5, 0.25, 0.0025 / 3], [1.0, 2.0, 3.0, 5e-1, 2.5e-1, 1e-5, 1.0],
             [-2.03458325e-03, 8.09653209e-01, 7.84710126e-04, 1e-4, 0.1, 0.0999999999, 1e-5],
             [123456789.0, 1.0, 0.5, 1e16, 1e15, 1.5e-7, 1e16], [0.001953125, 0.5, 0.25, 0.1, 0.2, 0.30000000000000004, 1e22]]
    e7 = np.array(rows, dtype=np.float64)
    assert graph_io.csv_bytes(e7) == python_csv(e7)


def test_reproduces_reference_files_byte_for_byte():
    for name in ("graph_small_s0.csv", "graph_small_s1.csv", "graph_docker_s0.csv.gz"):
        p = os.path.join(GOLDEN, name)
        raw = gzip.open(p, "rb").read() if name.endswith(".gz") else open(p, "rb").read()
        e7 = graph_io.parse_csv_bytes(raw)
        # the parser sees what every reference consumer sees
        rows = list(csv.DictReader(io.StringIO(raw.decode(), newline="")))
        assert np.array_equal(e7, rows_to_edges7(rows))
        # 8-digit cells re-format to themselves; radii are full precision
        assert graph_io.csv_bytes(e7) == raw


def test_empty_and_errors():
    assert graph_io.csv_bytes(np.zeros((0, 7))) == b"node1,node2,radius\r\n"
    assert graph_io.parse_csv_bytes(b"node1,node2,radius\r\n").shape == (0, 7)
    import pytest
    from octa_autosegmentation_b200._lib import OctaError
    with pytest.raises(OctaError):
        graph_io.csv_bytes(np.array([[np.nan, 0, 0, 0, 0, 0, 1.0]]))
    with pytest.raises(OctaError):
        graph_io.parse_csv_bytes(b"node1,node2,radius\r\n[0.1 0.2,[0 0 0],1\r\n")
