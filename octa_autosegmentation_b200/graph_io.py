"""Graph CSV surface (byte-compatible with the reference's files), over the C ABI.

write: generate_vessel_graph.py:59-66;  read: the `[x y z]` string cells every consumer parses
(tree2img.py:73-76, visualize_vessel_graphs.py:72-75, data_transforms.py:377-381)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


def _bind():
    L = _lib.lib()
    L.octa_format_csv.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_size_t,
                                  ctypes.POINTER(ctypes.c_size_t)]
    L.octa_parse_csv.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int64]
    L.octa_parse_csv.restype = ctypes.c_int64
    return L


def csv_bytes(edges7: np.ndarray) -> bytes:
    """Header + one row per edge, exactly as the reference writes them."""
    L = _bind()
    e = np.ascontiguousarray(edges7, dtype=np.float64).reshape(-1, 7)
    cap = 64 + 110 * len(e)
    buf = ctypes.create_string_buffer(cap)
    n = ctypes.c_size_t(0)
    rc = L.octa_format_csv(e.ctypes.data, len(e), buf, cap, ctypes.byref(n))
    if rc == _lib.OCTA_E_NOMEM:
        cap = n.value
        buf = ctypes.create_string_buffer(cap)
        rc = L.octa_format_csv(e.ctypes.data, len(e), buf, cap, ctypes.byref(n))
    _lib.check(rc)
    return buf.raw[:n.value]


def write_csv(path: str, edges7: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(csv_bytes(edges7))


def parse_csv_bytes(data: bytes) -> np.ndarray:
    L = _bind()
    n = L.octa_parse_csv(data, len(data), None, 0)
    if n < 0:
        _lib.check(int(n))
    out = np.empty((n, 7), dtype=np.float64)
    n2 = L.octa_parse_csv(data, len(data), out.ctypes.data, n)
    if n2 < 0:
        _lib.check(int(n2))
    return out


def read_csv(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        return parse_csv_bytes(f.read())


def png_bytes(image: np.ndarray, level: int = 1) -> bytes:
    """PNG file of a 2-D uint8 (gray, 8 bit) or bool (PIL mode "1", 1 bit) array: the same pixels `PIL.Image.save` stores
    (utils of generate_vessel_graph.py:85 / visualize_vessel_graphs.py:99-101), written with zlib directly.  PIL's encoder takes
    73 ms for a 1216^2 label (adaptive filter search + zlib level 6, under the GIL); filter 0 + zlib level 1 takes 9 ms,
    releases the GIL and gives a file ~10 % larger -- the host cost per sample otherwise exceeds the GPU's by 50x."""
    import struct
    import zlib

    a = np.asarray(image)
    if a.ndim != 2 or a.dtype not in (np.uint8, np.bool_):
        raise ValueError("png_bytes: 2-D uint8 or bool array expected")
    h, w = a.shape
    depth = 1 if a.dtype == np.bool_ else 8
    rows = np.packbits(a, axis=1) if depth == 1 else a
    raw = np.empty((h, rows.shape[1] + 1), dtype=np.uint8)
    raw[:, 0] = 0                                  # filter type 0 for every row
    raw[:, 1:] = rows

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data))

    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, 0, 0, 0, 0))
            + chunk(b"IDAT", zlib.compress(raw.tobytes(), level)) + chunk(b"IEND", b""))


def save_png(path: str, image: np.ndarray) -> None:
    """`Image.fromarray(image).save(path)` for uint8 / bool images (OCTA_PNG=pil: through PIL's own encoder)."""
    import os
    if os.environ.get("OCTA_PNG") == "pil":
        from PIL import Image
        Image.fromarray(image).save(path)
        return
    with open(path, "wb") as f:
        f.write(png_bytes(image))


def csv_batch_device(edges_dev, offsets, text_dev, text_offsets_dev, fallback_dev, workspace, stream=None):
    """Device writer (octa_format_csv_batch_dev): the CSV files of a batch, back to back, into `text_dev` (uint8 CUDA tensor);
    `text_offsets_dev` int64 [n+1], `fallback_dev` int32 [n] (!= 0: format that graph with csv_bytes).  Enqueued on `stream`."""
    import torch
    L = _lib.lib()
    offs = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offs) - 1
    L.octa_format_csv_batch_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    st = torch.cuda.current_stream() if stream is None else stream
    _lib.check(L.octa_format_csv_batch_dev(edges_dev.data_ptr(), offs.ctypes.data, n, text_dev.data_ptr(), text_dev.numel(),
                                           text_offsets_dev.data_ptr(), fallback_dev.data_ptr(), workspace.data_ptr(), workspace.numel(),
                                           ctypes.c_void_p(st.cuda_stream)))


def csv_device_sizes(n_graphs: int, n_edges: int):
    """(workspace bytes, text capacity bytes) of csv_batch_device."""
    L = _lib.lib()
    L.octa_format_csv_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int64]
    L.octa_format_csv_workspace_bytes.restype = ctypes.c_size_t
    L.octa_format_csv_text_cap.argtypes = [ctypes.c_int, ctypes.c_int64]
    L.octa_format_csv_text_cap.restype = ctypes.c_size_t
    return int(L.octa_format_csv_workspace_bytes(n_graphs, n_edges)), int(L.octa_format_csv_text_cap(n_graphs, n_edges))
