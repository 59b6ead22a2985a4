"""Batch sharding over the GPUs of one box.

Samples are independent (the reference maps over seeds with Beam,
scripts/run_evaluation.py:147-150,215-217), so each rank integrates a contiguous
block of samples with no communication; the only collective is the final gather
of the saved snapshots (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import os

import numpy as np


def shard_bounds(total, rank, world_size):
  """Contiguous block [start, stop) of `total` samples owned by `rank`; the first
  total % world_size ranks take one extra sample."""
  if not 0 <= rank < world_size:
    raise ValueError('rank {} outside world of {}'.format(rank, world_size))
  base, extra = divmod(total, world_size)
  start = rank * base + min(rank, extra)
  return start, start + base + (1 if rank < extra else 0)


def init_from_env(backend=None):
  """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
  import torch
  import torch.distributed as dist
  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if world > 1 and not dist.is_initialized():
    if backend is None:
      backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
      torch.cuda.set_device(local)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group(backend=backend, rank=rank, world_size=world)
  elif torch.cuda.is_available():
    torch.cuda.set_device(local)
  return rank, world, local


def gather_snapshots(local, total_samples, sample_axis=1, group=None):
  """All ranks receive the snapshots of all samples, concatenated along
  `sample_axis` in rank order.  `local` holds this rank's shard_bounds() block.
  Uneven shards are padded to the largest block for the collective."""
  import torch
  import torch.distributed as dist
  if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
    return local
  world = dist.get_world_size(group)
  counts = [b - a for a, b in (shard_bounds(total_samples, r, world) for r in range(world))]
  largest = max(counts)
  x = local.movedim(sample_axis, 0).contiguous()
  if x.shape[0] != counts[dist.get_rank(group)]:
    raise ValueError('rank holds {} samples, expected {}'.format(x.shape[0], counts[dist.get_rank(group)]))
  if x.shape[0] < largest:
    pad = torch.zeros((largest - x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    x = torch.cat([x, pad], dim=0)
  out = torch.empty((world * largest,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
  dist.all_gather_into_tensor(out, x, group=group)
  if min(counts) == largest:                    # even shards: the gathered buffer is the result, no second copy
    return out.movedim(0, sample_axis)
  pieces = [out[r * largest:r * largest + counts[r]] for r in range(world)]
  return torch.cat(pieces, dim=0).movedim(0, sample_axis)


def max_over_ranks(value, device=None, group=None):
  """Max of a Python float over ranks (timing: the slowest rank defines the step)."""
  import torch
  import torch.distributed as dist
  if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
    return float(value)
  if device is None:
    device = 'cuda' if dist.get_backend(group) == 'nccl' else 'cpu'
  t = torch.tensor([float(value)], dtype=torch.float64, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
  return float(t.item())
