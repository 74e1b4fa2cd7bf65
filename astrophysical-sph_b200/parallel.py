"""Host-side helpers of the multi-GPU path (one process per GPU).

Every rank holds the full particle state; after the key sort each rank owns one contiguous chunk of the sorted
order (a Morton/octant-key range) as TARGETS of search / density / force / tree walk, and NCCL all-gathers
(h, rho, g, PHI, row reductions) plus one all-reduce (hydro reactions that land on other ranks' particles)
rebuild the replicated state.  The arithmetic here must stay identical to eval_internal() in csrc/sph_api.cu.
"""
from __future__ import annotations

PAD_QUANTUM = 1680   # divisible by 1..8, 10, 12, 14, 15, 16 (csrc/sph_api.cu: sph_create)


def padded_size(N: int) -> int:
    return (N + PAD_QUANTUM - 1) // PAD_QUANTUM * PAD_QUANTUM


def target_range(N: int, nranks: int, rank: int) -> tuple[int, int]:
    """[t0, t1) of sorted slots owned by `rank`."""
    NS = padded_size(N)
    if NS % nranks:
        raise ValueError("unsupported rank count (use 1-8, 10, 12, 14, 15 or 16)")
    chunk = NS // nranks
    t0 = min(rank * chunk, N)
    t1 = min(t0 + chunk, N) if rank * chunk < N else N
    return t0, max(t0, min((rank + 1) * chunk, N))


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
