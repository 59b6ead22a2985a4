"""Batch sharding + the final snapshot gather on 2 CPU processes (gloo)."""
import os
import socket

import numpy as np
import pytest

from ddd1d_b200 import distributed


def test_shard_bounds_cover_the_batch():
  for total in (0, 1, 7, 8, 4096, 4097, 32768):
    for world in (1, 2, 3, 4, 8):
      blocks = [distributed.shard_bounds(total, r, world) for r in range(world)]
      assert blocks[0][0] == 0 and blocks[-1][1] == total
      for (a0, b0), (a1, b1) in zip(blocks[:-1], blocks[1:]):
        assert b0 == a1 and b0 >= a0
      sizes = [b - a for a, b in blocks]
      assert max(sizes) - min(sizes) <= 1
  with pytest.raises(ValueError):
    distributed.shard_bounds(10, 2, 2)


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _worker(rank, world, port, total, queue):
  import torch
  import torch.distributed as dist
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
  r, w, _ = distributed.init_from_env(backend='gloo')
  assert (r, w) == (rank, world)
  start, stop = distributed.shard_bounds(total, rank, world)
  # snapshots [time, sample, x] whose value encodes (time, global sample, x)
  t, n = 3, 5
  local = torch.stack([torch.stack([torch.arange(n, dtype=torch.float32) + 100.0 * s + 10000.0 * k
                                    for s in range(start, stop)]) if stop > start
                       else torch.zeros((0, n)) for k in range(t)])
  full = distributed.gather_snapshots(local, total, sample_axis=1)
  slowest = distributed.max_over_ranks(1.0 + rank)
  queue.put((rank, full.numpy(), slowest))
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.parametrize('total', (7, 8))
def test_gather_snapshots_two_ranks(total):
  import torch.multiprocessing as mp
  ctx = mp.get_context('spawn')
  queue = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, total, queue)) for r in range(2)]
  for p in procs:
    p.start()
  results = [queue.get(timeout=120) for _ in procs]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  want = np.stack([np.stack([np.arange(5, dtype=np.float32) + 100.0 * s + 10000.0 * k
                             for s in range(total)]) for k in range(3)])
  for rank, full, slowest in results:
    np.testing.assert_array_equal(full, want)
    assert slowest == 2.0


def test_single_process_is_identity():
  import torch
  x = torch.arange(24.).reshape(2, 3, 4)
  assert distributed.gather_snapshots(x, 3) is x
  assert distributed.max_over_ranks(3.5) == 3.5


def _eval_worker(rank, world, port, total, queue):
  """run_integrate_batch(distributed=True) plumbing with a stand-in for the CUDA integration of one block."""
  import torch.distributed as dist
  from ddd1d_b200 import evaluation
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
  distributed.init_from_env(backend='gloo')
  times = np.array([0.0, 0.5, 1.0])
  y0 = np.arange(total * 4, dtype=np.float64).reshape(total, 4)

  def local_runner(block, seed0):          # "integrates" sample s: y(t) = y0 * (1 + t) + seed
    seeds = seed0 + np.arange(block.shape[0])
    y = block[:, None, :] * (1.0 + times)[None, :, None] + seeds[:, None, None]
    return {'y': y, 'num_evals': 100 + seeds, 'x': np.arange(4) * 0.25}

  class HP(object):                        # only reached when a rank holds no sample
    pass
  res = evaluation._run_sharded(None, HP(), y0, times, 2.0, 'RK23', None, 7, None, local_runner=local_runner)
  queue.put((rank, res))
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.parametrize('total', (5, 8))
def test_sharded_evaluation_two_ranks(total):
  """Every rank integrates its block of seeds and receives all samples: equal to the unsharded result."""
  import torch.multiprocessing as mp
  ctx = mp.get_context('spawn')
  queue = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, total, queue)) for r in range(2)]
  for p in procs:
    p.start()
  results = [queue.get(timeout=120) for _ in procs]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  times = np.array([0.0, 0.5, 1.0])
  y0 = np.arange(total * 4, dtype=np.float64).reshape(total, 4)
  seeds = 7 + np.arange(total)
  want = y0[:, None, :] * (1.0 + times)[None, :, None] + seeds[:, None, None]
  for rank, res in results:
    np.testing.assert_array_equal(res['y'], want)
    np.testing.assert_array_equal(res['num_evals'], 100 + seeds)
    np.testing.assert_array_equal(res['sample'], seeds)
    np.testing.assert_array_equal(res['time'], 2.0 + times)
