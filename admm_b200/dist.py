"""Host-side plumbing for row-sharded runs: one process per GPU (torchrun), torch.distributed for
the rendezvous, the library's own NCCL communicator for the data path.

The row split is the reference's (/root/reference/src/PADMMLasso.h:163-178): block r holds rows
[r * floor(n / N), (r + 1) * floor(n / N)); the last block also takes the n mod N remainder.
"""
import ctypes as C

import numpy as np

from . import _capi as K


def row_block(n, nblocks, r):
    """(first_row, n_rows) of block r among nblocks, the reference's split."""
    if not (0 <= r < nblocks):
        raise ValueError("block index out of range")
    chunk = n // nblocks
    if chunk < 1:
        raise ValueError("more blocks than observations")
    return r * chunk, (chunk if r < nblocks - 1 else chunk + n % nblocks)


def broadcast_bytes(payload, src=0):
    """Broadcast a bytes object of known length from rank `src` with torch.distributed (any backend)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    buf = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone().to(dev)
    dist.broadcast(buf, src)
    return bytes(buf.cpu().numpy().tobytes())


def init_comm():
    """Create the library's communicator over the ranks of the initialised torch.distributed group."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    L = K.lib()
    raw = C.create_string_buffer(K.COMM_ID_BYTES)
    if rank == 0:
        K.check(L.b200admm_comm_id(raw))
    ident = broadcast_bytes(raw.raw, 0)
    K.check(L.b200admm_comm_init(ident, rank, world))
    return rank, world


def destroy_comm():
    K.lib().b200admm_comm_destroy()
