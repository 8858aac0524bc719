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
    cap = data.count(b"\n") + 1               # a row ends with a newline: one pass instead of count + fill
    out = np.empty((cap, 7), dtype=np.float64)
    n = L.octa_parse_csv(data, len(data), out.ctypes.data, cap)
    if n < 0:
        _lib.check(int(n))
    return out[:n] if n <= cap else parse_csv_bytes_two_pass(data)


def parse_csv_bytes_two_pass(data: bytes) -> np.ndarray:
    """Count, then fill (the sizing protocol of octa_parse_csv for callers that cannot bound the row count)."""
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


_NIFTI_CODES = {"uint8": (2, 8), "int16": (4, 16), "int32": (8, 32), "float32": (16, 32), "float64": (64, 64), "int8": (256, 8),
                "uint16": (512, 16), "uint32": (768, 32), "int64": (1024, 64), "uint64": (1280, 64)}


def nifti1_bytes(vol: np.ndarray) -> bytes:
    """The single-file NIfTI-1 image `nib.Nifti1Image(vol, np.eye(4))` stands for (generate_vessel_graph.py:75-77,
    visualize_vessel_graphs.py:85-87), uncompressed: the 348-byte header of the NIfTI-1 standard as nibabel fills it for an
    array and an identity affine (dim = [ndim, *shape, 1...], pixdim = 1, sform 'aligned' = identity rows, qform 'unknown' with
    qfac 1, scl_slope / scl_inter NaN = "no scaling", vox_offset 352, magic "n+1"), four zero bytes (no extensions), then the
    voxels in Fortran order.  Written without nibabel (not part of this stack); parity with nibabel's bytes is NOT pinned -- no
    nibabel here to compare with -- the fields follow the standard and tests/test_host_logic.py reads them back."""
    import struct

    a = np.asarray(vol)
    if a.dtype.name not in _NIFTI_CODES:
        raise ValueError('data dtype "%s" not supported' % a.dtype.name)      # nibabel: HeaderDataError with this text
    if not 1 <= a.ndim <= 7:
        raise ValueError("NIfTI-1 holds 1 to 7 dimensions")
    code, bitpix = _NIFTI_CODES[a.dtype.name]
    dim = [a.ndim] + list(a.shape) + [1] * (7 - a.ndim)
    if max(dim) > 32767:
        raise ValueError("NIfTI-1 dimensions are 16-bit")
    nan = float("nan")
    h = struct.pack("<i10s18sihcB", 348, b"", b"", 0, 0, b"\0", 0)                       # sizeof_hdr ... dim_info
    h += struct.pack("<8h", *dim)
    h += struct.pack("<3f4h", 0.0, 0.0, 0.0, 0, code, bitpix, 0)                           # intent_p1-3, intent_code, datatype, bitpix, slice_start
    h += struct.pack("<8f", 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0)                        # pixdim (pixdim[0] = qfac)
    h += struct.pack("<3fhBB", 352.0, nan, nan, 0, 0, 0)                                   # vox_offset, scl_slope, scl_inter, slice_end, slice_code, xyzt_units
    h += struct.pack("<4f2i", 0.0, 0.0, 0.0, 0.0, 0, 0)                                    # cal_max, cal_min, slice_duration, toffset, glmax, glmin
    h += struct.pack("<80s24s2h", b"", b"", 0, 2)                                          # descrip, aux_file, qform_code 0, sform_code 2 (aligned)
    h += struct.pack("<6f", 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)                                  # quatern_b/c/d, qoffset_x/y/z
    h += struct.pack("<12f", 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0)   # srow_x / y / z
    h += struct.pack("<16s4s", b"", b"n+1\0")
    assert len(h) == 348
    return h + b"\0\0\0\0" + a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes(order="F")


def save_nifti(path: str, vol: np.ndarray) -> None:
    """`nib.save(nib.Nifti1Image(vol, np.eye(4)), path)` for `.nii.gz` / `.nii` paths (gzip level 1, nibabel's default)."""
    import gzip
    data = nifti1_bytes(vol)
    if path.endswith(".gz"):
        with gzip.GzipFile(path, "wb", compresslevel=1) as f:
            f.write(data)
    else:
        with open(path, "wb") as f:
            f.write(data)


def load_nifti(path: str) -> np.ndarray:
    """Reads back what save_nifti (or any little-endian single-file NIfTI-1 writer) stored; used by the tests."""
    import gzip
    import struct
    raw = gzip.open(path, "rb").read() if path.endswith(".gz") else open(path, "rb").read()
    if struct.unpack_from("<i", raw, 0)[0] != 348 or raw[344:348] != b"n+1\0":
        raise ValueError("not a little-endian single-file NIfTI-1 image")
    dim = struct.unpack_from("<8h", raw, 40)
    code = struct.unpack_from("<h", raw, 70)[0]
    name = next(k for k, v in _NIFTI_CODES.items() if v[0] == code)
    off = int(struct.unpack_from("<f", raw, 108)[0])
    shape = dim[1:1 + dim[0]]
    return np.frombuffer(raw, dtype=np.dtype(name).newbyteorder("<"), count=int(np.prod(shape)), offset=off).reshape(shape, order="F")


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
