"""Tensor <-> bytes wire format of the two-process loop (reference: utils/data_transfer.py:4-22)."""
import io

import numpy as np
import torch


def torch_to_bytes(t) -> bytes:
    buff = io.BytesIO()
    torch.save(t, buff)
    return buff.getvalue()


def bytes_to_torch(b: bytes):
    return torch.load(io.BytesIO(b), weights_only=False)


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
