"""Multi-GPU plumbing for the two ways the path shards. One process per GPU, torch.distributed for the
collective; the math stays in the CUDA library.

  * extraction / stereo: frames are independent -> contiguous chunks of the sequence per rank, no data-path collective;
  * brute-force kNN (BASELINE.json configs[4]): train rows are split in contiguous index ranges, queries are replicated,
    every rank emits one packed 64-bit partial per query (best << 48 | second << 32 | global index), the partials are
    all-gathered rank-major over NCCL (8 B x Q per rank) and folded by the merge kernel. The fold equals the reference's
    single ascending scan (src/ORBmatcher.cc:477-507) because its result is order independent: idx = lowest-index argmin,
    second = minimum over the rest.
"""
import numpy as np


def shard_range(n, world, rank):
    """Contiguous [begin, end) of n items owned by `rank`; sizes differ by at most one, lower ranks get the extra."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def pack_partial(best, second, idx):
    """numpy mirror of the device packing, for host-side plumbing tests: idx < 0 means none."""
    best = np.asarray(best, np.uint64); second = np.asarray(second, np.uint64)
    idx = np.asarray(idx, np.int64).astype(np.uint32).astype(np.uint64)
    return ((best << np.uint64(48)) | (second << np.uint64(32)) | idx).view(np.int64)


def knn2_sharded(d_query, d_train_shard, index_base, th_low=50, nnratio=0.6, group=None):
    """Train-sharded best/second scan on the calling rank's GPU. d_query: (Q, 32) uint8 CUDA tensor, identical on every
    rank; d_train_shard: this rank's contiguous slice of the train set starting at global row `index_base`.
    Returns (idx, best, second, match) CUDA tensors, identical on every rank."""
    import torch
    import torch.distributed as dist
    from . import api
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    nq = d_query.shape[0]
    part = api.knn2_partial_device(d_query, d_train_shard, int(index_base))
    if world == 1:
        gathered = part.view(1, nq)
    else:
        gathered = torch.empty((world, nq), dtype=torch.int64, device=d_query.device)
        dist.all_gather_into_tensor(gathered.view(-1), part, group=group)
    return api.knn2_merge_device(gathered, world, nq, th_low, nnratio)


def extract_sharded(extractor, frames_host, group=None):
    """Frame-sharded extraction: every rank runs ExtractBatch on its contiguous share of `frames_host` (F, H, W) and
    returns (begin, end, keypoints, descriptors) for that share; nothing crosses GPUs."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    b, e = shard_range(len(frames_host), world, rank)
    if e == b:
        return b, e, [], []
    k, d = extractor.ExtractBatch(frames_host[b:e])
    return b, e, k, d
