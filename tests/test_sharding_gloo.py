"""CPU, world_size 2 over gloo: the N > 1 path shards the batch with no data-path collective; the optional final
gather reassembles exactly what a single rank would have produced.  The per-sample compute is stood in for by the
oracle (tests may use it), because the CUDA path needs a GPU -- what is under test is the sharding/gather logic."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import asm_oracle as ao


def _worker(rank, world, port, batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from style_transfer_based_holographic_imaging_b200 import parallel
    rng = np.random.default_rng(42)                      # same data on every rank
    O = (rng.standard_normal((batch, 1, 32, 32)) + 1j * rng.standard_normal((batch, 1, 32, 32))).astype(np.complex64)
    d = ((0.2 + 0.8 * rng.random((batch, 1, 1, 1))) * 1e-3).astype(np.float32)
    lo, hi = parallel.shard_bounds(batch, rank, world)
    assert parallel.shard(torch.from_numpy(O)).shape[0] == hi - lo
    local = np.abs(ao.asm(O[lo:hi], 532e-9, d[lo:hi], 1.5e-6)) ** 2 if hi > lo else np.zeros((0, 1, 32, 32))
    full = parallel.gather_batch(torch.from_numpy(local.astype(np.float32)), batch)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), full.numpy())
        np.save(os.path.join(out_dir, "single.npy"), (np.abs(ao.asm(O, 532e-9, d, 1.5e-6)) ** 2).astype(np.float32))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [6, 5])
def test_two_rank_shard_and_gather(tmp_path, batch):
    port = 29650 + batch
    mp.spawn(_worker, args=(2, port, batch, str(tmp_path)), nprocs=2, join=True)
    g = np.load(tmp_path / "gathered.npy")
    s = np.load(tmp_path / "single.npy")
    assert g.shape == s.shape
    assert np.array_equal(g, s)        # sharding must not change any sample (bitwise)
