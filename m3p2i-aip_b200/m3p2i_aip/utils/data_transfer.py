"""Tensor <-> bytes wire format of the two-process loop (reference: utils/data_transfer.py:4-22).

Two frame kinds travel over the same RPC methods (run_tamp / get_trajs / get_suction, reactive_tamp.py:43-87):
* `torch.save` frames, what the reference sends. They are decoded with `weights_only=True` (tensors, ndarrays and plain
  scalars only): the server binds tcp://0.0.0.0:4242 (reactive_tamp.py:92), so a frame must never be able to run code.
  M3P2I_UNSAFE_PICKLE=1 restores arbitrary unpickling for a trusted loop.
* raw fp32 frames (`raw_to_bytes` / `bytes_to_torch` recognises them by their magic): an 8-byte header, the shape, then
  the C-contiguous float32 payload -- no pickling, 30x cheaper to encode than torch.save for the [1, 18] / [1, 7, 13]
  states of one tick.
* shared-memory frames (`ShmFrames.put` returns a ~50-byte token, `bytes_to_torch` recognises it by its magic and reads
  the float32 payload from the named segment): for the two processes of the loop on one host (scripts/sim.py and
  scripts/reactive_tamp.py) only the token travels through the RPC layer.
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


SHM_MAGIC = b"M3SHM\x00"
_SHM_HEADER = 8 + 4 + 4 * 8   # per slot: sequence number (u64), ndim (u32), up to 8 dims (u32)
_shm_attached = {}


def _shm_open(name, create=False, size=0):
    from multiprocessing import resource_tracker, shared_memory
    seg = shared_memory.SharedMemory(name=name, create=create, size=size)
    if not create:
        # a reader must not unlink the writer's segment when it exits (Python < 3.13 registers every attach)
        try:
            resource_tracker.unregister(seg._name, "shared_memory")
        except Exception:
            pass
    return seg


class ShmFrames:
    """Writer side of the shared-memory transport: a ring of `slots` raw fp32 frames (each up to `max_floats` floats) in
    one named segment. `put(t)` copies the tensor into the next slot and returns the token to send instead of the frame;
    any process on the host decodes it with `bytes_to_torch`. A token whose slot has been overwritten since (more than
    `slots` frames later) is refused, not silently mixed up."""

    def __init__(self, name=None, slots=8, max_floats=1 << 16):
        self.slots, self.max_floats = int(slots), int(max_floats)
        self.stride = _SHM_HEADER + 4 * self.max_floats
        self.seg = _shm_open(name, create=True, size=self.slots * self.stride)
        self.name = self.seg.name
        self.seq = 0

    def put(self, t) -> bytes:
        a = np.ascontiguousarray(t.detach().cpu().numpy() if torch.is_tensor(t) else t, dtype=np.float32)
        if a.size > self.max_floats or a.ndim > 8:
            raise ValueError(f"frame of {a.size} floats / {a.ndim} dims does not fit a slot ({self.max_floats} floats, 8 dims)")
        self.seq += 1
        slot = self.seq % self.slots
        off = slot * self.stride
        buf = self.seg.buf
        struct.pack_into("<Q", buf, off, 0)                     # invalid while the payload is being written
        struct.pack_into(f"<I{a.ndim}I", buf, off + 8, a.ndim, *a.shape)
        np.frombuffer(buf, dtype=np.float32, count=a.size, offset=off + _SHM_HEADER)[:] = a.ravel()
        struct.pack_into("<Q", buf, off, self.seq)
        nm = self.name.encode()
        return SHM_MAGIC + struct.pack("<H", len(nm)) + nm + struct.pack("<IIQ", slot, self.stride, self.seq)

    def close(self):
        try:
            self.seg.close()
            self.seg.unlink()
        except FileNotFoundError:
            pass


def _decode_shm(b: bytes):
    n = len(SHM_MAGIC)
    (ln,) = struct.unpack_from("<H", b, n)
    name = b[n + 2: n + 2 + ln].decode()
    slot, stride, seq = struct.unpack_from("<IIQ", b, n + 2 + ln)
    seg = _shm_attached.get(name)
    if seg is None:
        seg = _shm_attached[name] = _shm_open(name)
    off = slot * stride
    if off + _SHM_HEADER > seg.size:
        raise ValueError("shared-memory frame: slot outside the segment")
    (have,) = struct.unpack_from("<Q", seg.buf, off)
    (ndim,) = struct.unpack_from("<I", seg.buf, off + 8)
    if have != seq or ndim > 8:
        raise ValueError("shared-memory frame: the slot has been overwritten since this token was issued")
    shape = struct.unpack_from(f"<{ndim}I", seg.buf, off + 12)
    count = int(np.prod(shape)) if ndim else 1
    if off + _SHM_HEADER + 4 * count > seg.size:
        raise ValueError("shared-memory frame: payload outside the segment")
    out = np.frombuffer(seg.buf, dtype=np.float32, count=count, offset=off + _SHM_HEADER).reshape(shape).copy()
    (again,) = struct.unpack_from("<Q", seg.buf, off)
    if again != seq:
        raise ValueError("shared-memory frame: the slot was overwritten while it was being read")
    return torch.from_numpy(out)


def bytes_to_torch(b: bytes):
    if b[: len(RAW_MAGIC)] == RAW_MAGIC:
        return _decode_raw(b)
    if b[: len(SHM_MAGIC)] == SHM_MAGIC:
        return _decode_shm(b)
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
