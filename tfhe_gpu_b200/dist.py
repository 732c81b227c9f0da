"""Multi-GPU plumbing: one process per GPU, keys replicated once, batch sharded, no collective in the loop.

The reference replicates its keys with a host loop of cudaMemcpy to every device and round-robins ciphertexts from a
single host thread (bootstrapping.cu:1007-1069, :1617).  Here rank 0 owns the key arrays, `broadcast_keys` sends the
raw uint64 key arrays to every rank with one torch.distributed broadcast each (NCCL over NVLink/NVSwitch on GPUs, gloo
in the CPU tests), every rank re-encodes them on its own GPU (tfhe_b200_setup with key_space = DEVICE), and
`shard_range` gives each rank a contiguous slice of the batch.
"""
import numpy as np


def shard_range(batch, world, rank):
    """Contiguous, balanced split (the first batch % world ranks get one extra ciphertext)."""
    base, rem = divmod(batch, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def broadcast_keys(params_dict, bk, ksk, device, src=0):
    """Replicate (params, bk, ksk) from rank `src` to all ranks.  `bk`/`ksk` are uint64 numpy arrays on the source
    rank (ignored elsewhere); returns (params_dict, bk_tensor, ksk_tensor) with int64 tensors on `device`."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    if rank == src:
        bk_t = torch.from_numpy(np.ascontiguousarray(bk).view(np.int64)).to(device)
        ksk_t = torch.from_numpy(np.ascontiguousarray(ksk).view(np.int64)).to(device)
    if world == 1:
        torch.cuda.synchronize(device) if torch.device(device).type == "cuda" else None
        return params_dict, bk_t, ksk_t
    meta = [params_dict, int(bk_t.numel()), int(ksk_t.numel())] if rank == src else [None, 0, 0]
    dist.broadcast_object_list(meta, src=src)
    if rank != src:
        bk_t = torch.empty(meta[1], dtype=torch.int64, device=device)
        ksk_t = torch.empty(meta[2], dtype=torch.int64, device=device)
    dist.broadcast(bk_t, src=src)
    dist.broadcast(ksk_t, src=src)
    # NCCL broadcasts are asynchronous to the host and ordered only on torch's stream; the engine re-encodes the keys on
    # its own streams, so the data must have landed before the tensors are handed to tfhe_b200_setup
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)
    return meta[0], bk_t, ksk_t
