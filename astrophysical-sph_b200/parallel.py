"""Host-side helpers of the multi-GPU path (one process per GPU).

Every rank holds the full particle state; after the key sort each rank owns one contiguous chunk of the sorted
order (a Morton/octant-key range) as TARGETS of search / density / force / tree walk, and NCCL all-gathers
(K-th distances, rho, cross-rank reverse pairs, force outputs, g, PHI) rebuild the replicated state; nothing is
reduced across ranks.  The arithmetic here must stay identical to set_partition() in csrc/sph_api.cu.
"""
from __future__ import annotations

TILE = 128            # targets per rank are a multiple of the 128-target tiles (csrc/sph_api.cu: set_partition)
MAX_RANKS = 16


def chunk_size(N: int, nranks: int) -> int:
    """Targets per rank: ceil(N / nranks) rounded up to a multiple of 128."""
    if not 1 <= nranks <= MAX_RANKS:
        raise ValueError("unsupported rank count (1..16)")
    per = (N + nranks - 1) // nranks
    return (per + TILE - 1) // TILE * TILE


def padded_size(N: int, nranks: int = 1) -> int:
    """Stride of the sorted-space component arrays: nranks * chunk."""
    return chunk_size(N, nranks) * nranks


def target_range(N: int, nranks: int, rank: int) -> tuple[int, int]:
    """[t0, t1) of sorted slots owned by `rank`."""
    chunk = chunk_size(N, nranks)
    t0 = min(rank * chunk, N)
    return t0, max(t0, min((rank + 1) * chunk, N))


def upload_slice(N: int, nranks: int, rank: int) -> tuple[int, int]:
    """Rows [r0, r1) of every state column that `rank` moves over PCIe in sph_upload (csrc/sph_api.cu)."""
    per = (N + nranks - 1) // nranks
    r0 = min(rank * per, N)
    return r0, min(r0 + per, N)


def share_unique_id(dist, make_id, src: int = 0) -> bytes:
    """Rank `src` creates the 128-byte NCCL unique id (SphB200.comm_unique_id) and every rank receives it through
    torch.distributed (any backend)."""
    box = [make_id() if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    assert isinstance(box[0], (bytes, bytearray)) and len(box[0]) == 128
    return bytes(box[0])


def init_handle_comm(handle, dist):
    """Join `handle` to a communicator spanning dist's world."""
    from .libsph import SphB200

    uid = share_unique_id(dist, SphB200.comm_unique_id)
    handle.comm_init(dist.get_world_size(), dist.get_rank(), uid)
