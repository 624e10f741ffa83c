"""Host-side helpers of the two ways the registration path shards across GPUs
(SURVEY.md §8e). One process per GPU; torch.distributed is only the plumbing.

1. batches of independent scan/submap pairs -> `shard_pairs` (no collective);
2. one very large registration split over the GPUs -> `init_comm`: every rank sets
   the same full clouds; inside the library (apd_comm_init) each rank computes its
   slice of the covariances (all-gathered) and linearizes its slice of the source
   (28 / 1 doubles all-reduced over NCCL). `shard_range` is the partition arithmetic
   for callers that split their own data (pairs, frames).
"""
import ctypes


def shard_range(n, rank, world):
    """contiguous split of n items: ranks [0, n % world) get one extra"""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_pairs(n_pairs, rank, world):
    """pair i -> rank i % world (round-robin keeps per-rank work even when pair sizes vary)"""
    return list(range(rank, n_pairs, world))


def broadcast_unique_id(lib, rank, dist, device=None):
    """rank 0 creates an ncclUniqueId through the library, everybody receives it"""
    import torch

    buf = (ctypes.c_char * 128)()
    if rank == 0:
        rc = lib.apd_comm_unique_id(buf)
        if rc != 0:
            raise RuntimeError(f"apd_comm_unique_id failed ({rc})")
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


def init_comm(reg, lib, rank, world, n_source_total, dist, device=None, fused=True):
    """attach an NCCL communicator to the registration handle `reg`; fused: also exchange the peer-memory mailboxes
    (cudaIpc handles, all-gathered here) so that the H/b/err all-reduce runs inside the reduction kernels"""
    import torch

    uid = broadcast_unique_id(lib, rank, dist, device)
    reg.comm_init(uid, rank, world, n_source_total)
    if fused and world > 1:
        mine = torch.tensor(list(reg.comm_peer_handle()), dtype=torch.uint8, device=device)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        reg.comm_peer_attach([bytes(t.cpu().tolist()) for t in every])
