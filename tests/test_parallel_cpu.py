"""World-size-2 gloo tests of the N>1 host logic (no GPU): parameter broadcast, flat-gradient all-reduce,
dataset sharding as `enflow/main.py:142-143` does it."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from enflow_b200.data.synthetic import SYNTHETICDataset
    from enflow_b200.flow.dynamics import LFIntegrator
    from enflow_b200.nn.argmax import ArgMax
    from enflow_b200.nn.egcl import EGCL
    from enflow_b200.parallel import allreduce_mean_, init_data_parallel
    from torch.utils.data.distributed import DistributedSampler
    torch.manual_seed(100 + rank)                      # different init per rank, like DDP before its broadcast
    m = LFIntegrator([EGCL(4, 4, 128) for _ in range(2)], ArgMax(4, 128), dt=0.01)
    init_data_parallel(m)
    gathered = [torch.empty_like(m.flat_params) for _ in range(world)]
    dist.all_gather(gathered, m.flat_params)
    assert all(torch.equal(g, gathered[0]) for g in gathered), 'parameters must match rank 0 after the broadcast'
    assert torch.equal(next(m.parameters()).data.reshape(-1), m.flat_params[:next(m.parameters()).numel()])
    m.flat_grads.fill_(float(rank + 1))                # rank r contributes r+1 everywhere
    allreduce_mean_(m.flat_grads, m._dp_group)
    assert torch.allclose(m.flat_grads, torch.full_like(m.flat_grads, (1 + world) / 2))
    ds = SYNTHETICDataset(config='c2', num_mols=10, n_atoms=5)
    idx = list(DistributedSampler(ds, num_replicas=world, rank=rank, shuffle=True, seed=0))
    allidx = [None] * world
    dist.all_gather_object(allidx, idx)
    assert sorted(sum(allidx, [])) == list(range(10)), 'shards must cover the dataset exactly once'
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, f'ok{rank}'), 'w').write('ok')


def test_world_size_2_gloo(tmp_path):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(2))


def test_env_ranks():
    from enflow_b200.parallel import env_ranks
    assert env_ranks('4', '3', '1') == (4, 3, 1)
    assert env_ranks(None, None, None) == (1, 0, 0) or 'WORLD_SIZE' in os.environ
