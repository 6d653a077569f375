"""-m gpu: several contexts of one process whose accumulators are terms of one sum (gmx_comm): the reads are dealt to the
contexts in 2048-read slices round robin (the reference's MPI "burn" scheme, reference inc/SeqManager.h:329-345), every
context maps its share, and gmx_finish on the root must deliver the accumulators of the whole batch (oracle)."""
import numpy as np
import pytest

from gnumap_b200 import _abi, index, sharding
from tests import common

pytestmark = pytest.mark.gpu


def _run(api, O, mode, devices, backend, world="plain"):
    contigs, batch, _ = getattr(common, "world_" + world)()
    ix = index.build_index(contigs)
    pg = common.set_mode(api.default_params(), mode); po = common.set_mode(O.default_params(), mode)
    want = O.process_batch(O.OracleIndex(ix), po, batch)
    ms = [api.Mapper(ix, pg, device=d) for d in devices]
    comm = api.Comm(ms, backend)
    n = len(ms)
    status = np.zeros(batch.n_reads, dtype=np.int32)
    for rank, m in enumerate(ms):
        for lo, hi in sharding.shard_slices(batch.n_reads, rank, n, slice_reads=256):
            r = m.process_batch(batch.slice(lo, hi))["results"]
            status[lo:hi] = r["status"]
    assert np.array_equal(status, want["results"]["status"])
    # per-context partial sums differ from the total unless everything went to one context
    amount, planes = ms[0].finish()                                   # reduces into the root first
    common.accum_close(amount, want["amount"], want["hits"], batch.offsets, pg.gen_size, ix.l_pac)
    if mode != _abi.MODE_NORMAL:
        for b in range(5):
            common.accum_close(planes[b], want["planes"][b], want["hits"], batch.offsets, pg.gen_size, ix.l_pac, what=f"plane {b}")
    st = comm.stats()
    assert st["bytes"] == 4 * (len(amount) + (planes.size if planes is not None else 0))
    # the other contexts were zeroed by the reduce: a second finish adds nothing
    again, _ = ms[0].finish()
    assert np.array_equal(again, amount)
    for m in ms[1:]:
        with pytest.raises(api.GmxError):
            m.finish()                                                # only the root holds the sum
    # all-reduce: every context ends with the total
    for m in ms:
        m.reset_accumulators()
    for rank, m in enumerate(ms):
        for lo, hi in sharding.shard_slices(batch.n_reads, rank, n, slice_reads=256):
            m.process_batch(batch.slice(lo, hi), fetch=False)
    comm.reduce(all=True)
    comm.close()
    tot = [m.finish()[0] for m in ms]
    for t in tot[1:]:
        assert np.array_equal(t, tot[0])
    assert np.allclose(tot[0], want["amount"], rtol=1e-5, atol=1e-6)
    for m in ms:
        m.close()
    return st


@pytest.mark.parametrize("mode", [_abi.MODE_NORMAL, _abi.MODE_SNP])
def test_contexts_sharing_one_gpu(mode):
    from gnumap_b200 import api
    from oracle import oracle as O
    _run(api, O, mode, [0, 0, 0], api.COMM_PEER)


@pytest.mark.parametrize("backend", ["peer", "nccl"])
@pytest.mark.parametrize("mode", [_abi.MODE_NORMAL, _abi.MODE_BS])
def test_one_context_per_gpu(mode, backend):
    import torch
    from gnumap_b200 import api
    from oracle import oracle as O
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs two or more GPUs")
    st = _run(api, O, mode, list(range(n)), api.COMM_PEER if backend == "peer" else api.COMM_NCCL)
    assert st["backend"] == backend


def test_comm_refuses_mismatched_contexts():
    from gnumap_b200 import api
    contigs, _, _ = common.world_plain(length=60_000, n_reads=4)
    ix = index.build_index(contigs)
    a = api.Mapper(ix)
    b = api.Mapper(ix, common.set_mode(api.default_params(), _abi.MODE_SNP))
    with pytest.raises(api.GmxError) as e:
        api.Comm([a, b])
    assert e.value.code == _abi.GMX_ERR_INVALID
    with pytest.raises(api.GmxError):
        api.Comm([a, a])
    with pytest.raises(api.GmxError):
        api.Comm([a], api.COMM_NCCL)
    c = api.Comm([a])
    amount, _ = a.finish()
    assert not amount.any()
    c.close(); a.close(); b.close()
