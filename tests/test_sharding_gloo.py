"""CPU suite, part 3: the N > 1 path (read sharding + the final sum-reduce of the accumulators) on the gloo
backend, world_size 2.  Each rank maps its shard with the ORACLE (test infrastructure; the GPU kernels have their
own parity tests) and the reduced accumulators must equal the single-process result on all reads."""
import os
import socket
import sys

import numpy as np
import pytest

from gnumap_b200 import _abi, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_slices_partition_the_reads():
    for n, world, sl in ((0, 2, 2048), (1, 2, 2048), (5000, 2, 2048), (10_000, 4, 1000), (4096, 8, 2048), (7, 3, 2)):
        seen = np.zeros(n, dtype=np.int32)
        for r in range(world):
            for lo, hi in sharding.shard_slices(n, r, world, sl):
                assert hi - lo <= sl and lo % sl == 0
                seen[lo:hi] += 1
            assert np.array_equal(np.sort(sharding.shard_indices(n, r, world, sl)), sharding.shard_indices(n, r, world, sl))
        assert np.all(seen == 1)
    with pytest.raises(ValueError):
        sharding.shard_slices(10, 2, 2)


def _worker(rank, world, port, mode, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from gnumap_b200 import index
    from oracle import oracle as O
    from tests import common
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    contigs, batch, _ = common.world_plain(seed=21, length=60_000, n_reads=700, read_len=64)
    ix = index.build_index(contigs)
    p = common.set_mode(O.default_params(), mode)
    mine = sharding.shard_batch(batch, rank, world, slice_reads=128)
    res = O.process_batch(O.OracleIndex(ix), p, mine)
    tensors = [torch.from_numpy(res["amount"])]
    if res["planes"] is not None:
        tensors.append(torch.from_numpy(res["planes"].reshape(-1)))
    sharding.all_reduce_accumulators(tensors)
    np.save(os.path.join(out_dir, f"status_{rank}.npy"), res["results"]["status"])
    np.save(os.path.join(out_dir, f"pos_{rank}.npy"), res["results"]["best_first_pos"])
    if rank == 0:
        np.save(os.path.join(out_dir, "amount.npy"), tensors[0].numpy())
        if len(tensors) > 1:
            np.save(os.path.join(out_dir, "planes.npy"), tensors[1].numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", [_abi.MODE_NORMAL, _abi.MODE_SNP])
def test_sharded_run_equals_single_process(tmp_path, mode):
    import torch.multiprocessing as mp
    from gnumap_b200 import index
    from oracle import oracle as O
    from tests import common
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, mode, str(tmp_path)), nprocs=world, join=True)
    contigs, batch, _ = common.world_plain(seed=21, length=60_000, n_reads=700, read_len=64)
    ix = index.build_index(contigs)
    p = common.set_mode(O.default_params(), mode)
    want = O.process_batch(O.OracleIndex(ix), p, batch)
    amount = np.load(tmp_path / "amount.npy")
    assert np.allclose(amount, want["amount"], rtol=1e-5, atol=1e-6)
    assert want["amount"].sum() > 0
    if mode != _abi.MODE_NORMAL:
        assert np.allclose(np.load(tmp_path / "planes.npy"), want["planes"].reshape(-1), rtol=1e-5, atol=1e-6)
    for r in range(world):
        idx = sharding.shard_indices(batch.n_reads, r, world, 128)
        assert np.array_equal(np.load(tmp_path / f"status_{r}.npy"), want["results"]["status"][idx])
        assert np.array_equal(np.load(tmp_path / f"pos_{r}.npy"), want["results"]["best_first_pos"][idx])
