"""Multi-GPU plumbing: games are independent, so the path shards with NO data-path
collective -- rank r of R owns game ids {first + r, first + r + R, ...}, its own resident game slots and
its own output files.  The single collective is one broadcast of the packed float32 weight blob from
rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests) at start and after every weight reload."""
import numpy as np
import torch
import torch.distributed as dist

from .agent import model as M


def rank_game_ids(rank, world_size, first_game_id, slots, games_per_slot):
    """ids the engine of `rank` hands to its slots (rz_engine_cfg.first_game_id / game_id_stride):
    slot s, k-th game -> first + rank + (s + k * slots) * world_size."""
    return [first_game_id + rank + (s + k * slots) * world_size for k in range(games_per_slot) for s in range(slots)]


def broadcast_blob(model_config, blob_or_none, device):
    """rank 0 passes the float32 blob (numpy), the others None; returns a float32 tensor on `device`
    holding the same blob on every rank."""
    n = M.blob_size(model_config)
    t = torch.empty(n, dtype=torch.float32, device=device)
    if dist.get_rank() == 0:
        t.copy_(torch.from_numpy(np.ascontiguousarray(blob_or_none, dtype=np.float32)))
    dist.broadcast(t, src=0)
    return t


def control_device(cuda_index):
    """device of the small control tensors: the GPU under NCCL, the host under gloo (CPU tests)"""
    return f"cuda:{cuda_index}" if dist.get_backend() == "nccl" else "cpu"


def sum_over_ranks(values, device):
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()
