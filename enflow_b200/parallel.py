"""Data parallelism for the flow: the stand-in for the reference's DDP wrapper (`enflow/main.py:159`).

Molecules never interact (`enflow/data/base.py:129-142`, `enflow/flow/loss.py:13`), so the batch is sharded across
ranks and the model is replicated.  The only collective on the training path is ONE all-reduce (mean) over the
flat gradient buffer, issued from the flow's backward node right after the last weight-gradient kernel.
"""
import torch
import torch.distributed as dist


def init_data_parallel(flow, group=None, src=0):
    """Broadcast rank `src`'s parameters (what DDP does at construction) and enable the gradient all-reduce."""
    group = dist.group.WORLD if group is None else group
    dist.broadcast(flow.flat_params, src, group=group)
    flow._dp_group = group
    return flow


def allreduce_mean_(flat, group):
    """In-place mean over ranks of a flat buffer: ONE collective (NCCL's AVG over NVLink/NVSwitch on GPUs, which is also
    what a captured CUDA graph replays; SUM followed by a scale under gloo, which has no AVG, in the CPU tests)."""
    if flat.is_cuda and dist.get_backend(group) == 'nccl':
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / dist.get_world_size(group))
    return flat


def env_ranks(world_size=None, world_rank=None, local_rank=None):
    """(world, rank, local) from explicit SLURM-style arguments (`__main__.py:6` of the reference) or torchrun's env."""
    import os
    if world_size is None:
        world_size = os.environ.get('WORLD_SIZE', os.environ.get('SLURM_NTASKS'))
        world_rank = os.environ.get('RANK', os.environ.get('SLURM_PROCID'))
        local_rank = os.environ.get('LOCAL_RANK', os.environ.get('SLURM_LOCALID'))
    if world_size is None or int(world_size) <= 1:
        return 1, 0, 0
    return int(world_size), int(world_rank or 0), int(local_rank or 0)
