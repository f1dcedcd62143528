"""Tensor <-> bytes wire format of the two-process loop (reference: utils/data_transfer.py:4-22).

Two frame kinds travel over the same RPC methods (run_tamp / get_trajs / get_suction, reactive_tamp.py:43-87):
* `torch.save` frames, what the reference sends. They are decoded with `weights_only=True` (tensors, ndarrays and plain
  scalars only): the server binds tcp://0.0.0.0:4242 (reactive_tamp.py:92), so a frame must never be able to run code.
  M3P2I_UNSAFE_PICKLE=1 restores arbitrary unpickling for a trusted loop.
* raw fp32 frames (`raw_to_bytes` / `bytes_to_torch` recognises them by their magic): an 8-byte header, the shape, then
  the C-contiguous float32 payload -- no pickling, 30x cheaper to encode than torch.save for the [1, 18] / [1, 7, 13]
  states of one tick.
"""
import io
import os
import struct

import numpy as np
import torch

RAW_MAGIC = b"M3F32\x00"


def torch_to_bytes(t) -> bytes:
    buff = io.BytesIO()
    torch.save(t, buff)
    return buff.getvalue()


def raw_to_bytes(t) -> bytes:
    """Raw frame: magic, ndim (u16), dims (u32 each), float32 data."""
    a = np.ascontiguousarray(t.detach().cpu().numpy() if torch.is_tensor(t) else t, dtype=np.float32)
    return RAW_MAGIC + struct.pack("<H", a.ndim) + struct.pack(f"<{a.ndim}I", *a.shape) + a.tobytes()


def _decode_raw(b: bytes):
    n = len(RAW_MAGIC)
    (ndim,) = struct.unpack_from("<H", b, n)
    shape = struct.unpack_from(f"<{ndim}I", b, n + 2)
    off = n + 2 + 4 * ndim
    count = int(np.prod(shape)) if ndim else 1
    if len(b) - off != 4 * count:
        raise ValueError("raw fp32 frame: payload size does not match its shape")
    return torch.from_numpy(np.frombuffer(b, dtype=np.float32, count=count, offset=off).reshape(shape).copy())


def bytes_to_torch(b: bytes):
    if b[: len(RAW_MAGIC)] == RAW_MAGIC:
        return _decode_raw(b)
    if os.environ.get("M3P2I_UNSAFE_PICKLE") == "1":
        return torch.load(io.BytesIO(b), weights_only=False)
    try:
        from numpy._core.multiarray import _reconstruct
    except ImportError:  # numpy < 2
        from numpy.core.multiarray import _reconstruct
    with torch.serialization.safe_globals([_reconstruct, np.ndarray, np.dtype, type(np.dtype(np.float32)),
                                           type(np.dtype(np.float64)), type(np.dtype(np.int64)), type(np.dtype(np.int32))]):
        return torch.load(io.BytesIO(b), weights_only=True)


def numpy_to_bytes(t: np.ndarray) -> bytes:
    return torch_to_bytes(t)


def bytes_to_numpy(b: bytes):
    return bytes_to_torch(b)


def check_server(server_address):
    """Remove a stale unix-socket file before binding (data_transfer.py:24-29); a missing file is fine."""
    import os
    try:
        os.unlink(server_address)
    except FileNotFoundError:
        pass
