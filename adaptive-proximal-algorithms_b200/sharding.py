"""Row partition of the data matrix over the ranks.

Rows of A (and b / labels) are split into P contiguous blocks; x, the gradient
and the stepsize state are replicated.  The only exchange per iteration is one
sum all-reduce of the A'r partials and the value sums (n + 2 doubles).
"""
from __future__ import annotations


def shard_rows(m: int, nranks: int, rank: int) -> tuple[int, int]:
    """(row0, rows) of ``rank``: contiguous, balanced to within one row, and
    every block a multiple of 8 rows except possibly the last (work units of the
    GEMV kernels are blocks of >= 8 rows)."""
    if not (0 <= rank < nranks):
        raise ValueError("rank out of range")
    if nranks > m:
        raise ValueError("more ranks than rows")
    blocks = (m + 7) // 8
    b0 = (blocks * rank) // nranks
    b1 = (blocks * (rank + 1)) // nranks
    row0, row1 = min(b0 * 8, m), min(b1 * 8, m)
    if rank == nranks - 1:
        row1 = m
    return row0, row1 - row0


def all_shards(m: int, nranks: int) -> list[tuple[int, int]]:
    return [shard_rows(m, nranks, r) for r in range(nranks)]


def attach_communicator(dev, dist=None):
    """Create the library's NCCL communicator for this process from an initialised
    ``torch.distributed`` process group (used only to broadcast the 128-byte id)."""
    if dist is None:
        import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [dev.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    dev.comm_init(world, rank, box[0])
    return world, rank


def attach_p2p(dev, n_max, dist=None) -> bool:
    """Enable the all-reduce inside the kernels over NVLink peer memory (include/adaprox.h: adaprox_p2p_*): every rank
    exports its exchange block, the 64-byte CUDA IPC handles are all-gathered through ``torch.distributed`` and every rank
    maps its peers' blocks.  Call after ``attach_communicator``; at most 8 ranks on one node.

    Collective and failure-safe: every rank takes part in both gathers whatever happens locally, and the function
    returns True only if EVERY rank attached; otherwise it returns False on every rank and sets ``ADAPROX_NO_P2P=1`` so
    that the library keeps using ``ncclAllReduce`` (row-sharded primal-dual solves then raise, they need the blocks)."""
    import os
    if dist is None:
        import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = None
    try:
        mine = dev.p2p_export(n_max)
    except Exception as e:                                          # noqa: BLE001 - reported to every rank below
        mine = None
        err = repr(e)
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    ok = all(h is not None for h in handles)
    if ok:
        try:
            dev.p2p_attach(world, rank, b"".join(handles))
        except Exception as e:                                      # noqa: BLE001
            ok = False
            err = repr(e)
    oks = [None] * world
    dist.all_gather_object(oks, ok)
    if not all(oks):
        os.environ["ADAPROX_NO_P2P"] = "1"
        if rank == 0:
            import sys
            print("[adaprox] peer exchange blocks not available on every rank; falling back to ncclAllReduce", file=sys.stderr)
        return False
    return True
