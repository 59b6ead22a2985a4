"""Multi-GPU product path (needs >= 2 GPUs: `gpurun --gpus 2`): evaluation.run_integrate_batch(distributed=True)
-- every rank integrates a contiguous block of seeds, one NCCL all-gather at the end -- must equal the
single-GPU call BIT FOR BIT (rows are independent: scripts/run_evaluation.py:147-150,212-221)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gpus():
  import torch
  return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _case():
  import ddd1d_b200.workloads as wl
  from tests import gpu_helpers as G
  n, total = 64, 301                                   # uneven shards, packed rows on the tensor engine
  hp = G.product_hparams('burgers', 'plain', n)
  weights = wl.synthetic_weights('burgers')
  y0 = G.smooth_rows(total, n, seed=21)
  times = np.linspace(0.0, 0.04, 5)
  return hp, weights, y0, times


def _worker(rank, world, port, queue):
  import torch
  import torch.distributed as dist
  from ddd1d_b200 import distributed, evaluation
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                    LOCAL_RANK=str(rank))
  distributed.init_from_env(backend='nccl')
  hp, weights, y0, times = _case()
  fixed = evaluation.run_integrate_batch(weights, hp, y0, times, fixed_dt=1e-3, first_seed=5, distributed=True)
  adaptive = evaluation.run_integrate_batch(weights, hp, y0, times, first_seed=5, distributed=True)
  queue.put((rank, fixed['y'], adaptive['y'], adaptive['num_evals']))
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.skipif(_gpus() < 2, reason='needs two GPUs')
def test_sharded_evaluation_equals_single_gpu():
  import torch.multiprocessing as mp
  from ddd1d_b200 import evaluation
  hp, weights, y0, times = _case()
  want_fixed = evaluation.run_integrate_batch(weights, hp, y0, times, fixed_dt=1e-3, first_seed=5)
  want_adaptive = evaluation.run_integrate_batch(weights, hp, y0, times, first_seed=5)
  world = 2
  ctx = mp.get_context('spawn')
  queue = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, queue)) for r in range(world)]
  for p in procs:
    p.start()
  results = [queue.get(timeout=600) for _ in procs]
  for p in procs:
    p.join(timeout=120)
    assert p.exitcode == 0
  for rank, fixed, adaptive, nfev in results:
    np.testing.assert_array_equal(fixed, want_fixed['y'])             # bit for bit
    np.testing.assert_array_equal(adaptive, want_adaptive['y'])
    np.testing.assert_array_equal(nfev, want_adaptive['num_evals'])
