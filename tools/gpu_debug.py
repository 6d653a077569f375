import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gnumap_b200 import api, index, _abi
from oracle import oracle as O
from tests import common

world = sys.argv[1] if len(sys.argv) > 1 else "repeats"
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
contigs, batch, _ = getattr(common, "world_" + world)()
ix = index.build_index(contigs)
m = api.Mapper(ix, common.set_mode(api.default_params(), mode))
got = m.process_batch(batch)
amount, planes = m.finish()
want = O.process_batch(O.OracleIndex(ix), common.set_mode(O.default_params(), mode), batch)
d = np.abs(amount - want["amount"])
rel = d / np.maximum(np.abs(want["amount"]), 1e-30)
bad = np.nonzero(~np.isclose(amount, want["amount"], rtol=1e-5, atol=1e-6))[0]
print("n bins", len(amount), "bad", len(bad), "max abs", d.max(), "max rel (where want>1e-3)", rel[want["amount"] > 1e-3].max())
for b in bad[:20]:
    print(b, amount[b], want["amount"][b], d[b], rel[b])
print("sum", amount.sum(dtype=np.float64), want["amount"].sum(dtype=np.float64))
print(m.stage_stats())
