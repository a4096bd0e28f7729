"""Multi-GPU plumbing of the generator path.

Inference shards by utterance: every op of the path is per-sample (InstanceNorm is per (b, c)), so rank r
simply takes utterances ``i % world == r``; weights are replicated and there is NO data-path collective.
The only collectives are bookkeeping: a barrier around timed regions and a MAX all-reduce of the step time.
These helpers are backend-agnostic (``nccl`` on the GPU box, ``gloo`` in the CPU tests).
"""

import torch
import torch.distributed as dist


def shard_utterances(n_utts, rank, world):
    """Indices of the utterances rank ``rank`` converts (round-robin, BASELINE config 5)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_utts, world))


def max_over_ranks(value, device="cpu"):
    """MAX of a python float over all ranks (identity without a process group)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def aggregate_throughput(units_per_rank, world, ms_max):
    """Whole-job units/s when every rank processed ``units_per_rank`` units in ``ms_max`` (max over ranks)."""
    return world * units_per_rank / (ms_max * 1e-3)


def gather_in_order(local_outputs, local_indices, n_utts):
    """Reassemble per-utterance results (python objects / CPU tensors) from all ranks in utterance order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = [None] * n_utts
        for i, o in zip(local_indices, local_outputs):
            out[i] = o
        return out
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, list(zip(local_indices, local_outputs)))
    out = [None] * n_utts
    for part in gathered:
        for i, o in part:
            out[i] = o
    return out
