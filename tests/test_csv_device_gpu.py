"""GPU: the device CSV writer (csrc/octa_csv_dev.cu) against the host writer, which tests/test_csv_format.py pins to numpy /
CPython / the reference's files: same bytes for reference graphs, for adversarial cells (exponential notation, negative and zero
coordinates, short decimals), and host fallback for what the device formatters decline."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_csv(graphs):
    import torch
    from octa_autosegmentation_b200 import graph_io
    offs = np.cumsum([0] + [len(g) for g in graphs]).astype(np.int64)
    e7 = np.concatenate([np.asarray(g, dtype=np.float64).reshape(-1, 7) for g in graphs]) if offs[-1] else np.zeros((0, 7))
    n = len(graphs)
    ws_b, cap_b = graph_io.csv_device_sizes(n, int(offs[-1]))
    dev = torch.from_numpy(np.ascontiguousarray(e7)).cuda() if offs[-1] else torch.zeros((1, 7), dtype=torch.float64, device="cuda")
    text = torch.zeros(cap_b, dtype=torch.uint8, device="cuda")
    toff = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    fb = torch.zeros(n, dtype=torch.int32, device="cuda")
    ws = torch.zeros(ws_b, dtype=torch.uint8, device="cuda")
    graph_io.csv_batch_device(dev, offs, text, toff, fb, ws)
    torch.cuda.synchronize()
    t, o, f = text.cpu().numpy(), toff.cpu().numpy(), fb.cpu().numpy()
    return [None if f[i] else t[o[i]:o[i + 1]].tobytes() for i in range(n)]


def test_reference_graphs_get_the_bytes_of_the_host_writer():
    from conftest import load_graph_rows, rows_to_edges7
    from octa_autosegmentation_b200 import graph_io
    graphs = [rows_to_edges7(load_graph_rows(n)) for n in ("graph_small_s0.csv", "graph_small_s1.csv", "graph_docker_s0.csv.gz", "graph_nerve_s0.csv")]
    graphs.append(np.zeros((0, 7)))                       # an empty graph is a header
    got = _device_csv(graphs)
    for g, text in zip(graphs, got):
        assert text is not None and text == graph_io.csv_bytes(g)


def test_adversarial_cells_and_fallbacks():
    from octa_autosegmentation_b200 import graph_io
    rs = np.random.RandomState(3)
    n = 20000
    a = rs.uniform(0, 1, (n, 7))
    a[:, 6] = 10 ** rs.uniform(-4, 0, n)
    a[: n // 4, :6] *= 10 ** rs.uniform(-7, 1, (n // 4, 1))          # exponential notation, mixed magnitudes
    a[n // 4: n // 2, :6] = np.round(a[n // 4: n // 2, :6], 3)       # short decimals
    a[rs.randint(0, n, 500), rs.randint(0, 6, 500)] = 0.0
    a[rs.randint(0, n, 500), rs.randint(0, 6, 500)] *= -1.0
    a[:, 6] = np.where((a[:, 6] >= 1e-4) & (a[:, 6] < 1), a[:, 6], 0.0123)
    graphs = [a[i:i + 1000] for i in range(0, n, 1000)]
    bad = a[:50].copy(); bad[7, 6] = 0.5                              # a power of two: the device declines, the host formats
    big = a[:50].copy(); big[3, 0] = 3.0e9                            # outside the range of the 9-digit exponential path
    graphs += [bad, big]
    got = _device_csv(graphs)
    assert got[-2] is None and got[-1] is None
    for g, text in zip(graphs[:-2], got[:-2]):
        assert text is not None and text == graph_io.csv_bytes(g)


def test_pipeline_csv_is_the_same_with_both_writers(monkeypatch):
    from octa_autosegmentation_b200.config import default_config
    from octa_autosegmentation_b200.pipeline import Pipeline
    cfg = default_config()
    for m, i in zip(cfg["Greenhouse"]["modes"], (12, 12)):
        m["I"], m["N"] = i, 400
    seeds = [[5, 6, 7], [8, 9, 10]]
    outs = {}
    for mode in ("device", "host"):
        monkeypatch.setenv("OCTA_CSV", mode)
        pipe = Pipeline(cfg, volume_dims=[152, 152, 8], label_res=(304, 304), image_res=(152, 152))
        assert pipe.device_csv == (mode == "device")
        outs[mode] = [[bytes(c) for c in o["csv"]] for o in pipe.run_pipelined(seeds, in_flight=2)]
    assert outs["device"] == outs["host"] and all(len(c) > 1000 for b in outs["device"] for c in b)
