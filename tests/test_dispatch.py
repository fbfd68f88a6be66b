"""Host logic of the multi-GPU dispatcher with two processes over gloo (no GPU needed): shards are
parsed by the oracle standing in for the device, then indices are rebased and compared with a
single-process parse of the whole batch."""
import os
import subprocess
import sys

import numpy as np

import oracle_lib as O
from nutdb_b200 import dispatch, workload as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_is_contiguous_and_balanced():
    text, offs = W.generate(2, 256 << 10)
    for parts in (1, 2, 3, 8):
        r = dispatch.split_statements(offs, parts)
        assert r[0][0] == 0 and r[-1][1] == len(offs) - 1
        assert all(r[i][1] == r[i + 1][0] for i in range(parts - 1))
        sizes = [int(offs[hi] - offs[lo]) for lo, hi in r]
        assert max(sizes) - min(sizes) < 400   # within a couple of statements
    # degenerate: more parts than statements
    r = dispatch.split_statements(np.array([0, 5, 9], np.uint64), 4)
    assert r[0][0] == 0 and r[-1][1] == 2 and sum(hi - lo for lo, hi in r) == 2


WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch.distributed as dist
import oracle_lib as O
from nutdb_b200 import dispatch, workload as W
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
text, offs = W.generate(3, 192 << 10)
class B: pass
def parse(t, o):
    r = O.parse_batch(t, o, 1)
    b = B(); b.stmt, b.err = r.stmt, r.err
    b.n_stmt, b.n_tok, b.n_node, b.n_err = len(r.stmt), len(r.tok_type), len(r.node), len(r.err)
    b.node = r.node
    return b
b, stmt, err, allt = dispatch.parse_sharded(parse, text, offs, rank, world, dist)
np.savez(os.path.join(sys.argv[2], f"rank{rank}.npz"), stmt=stmt, err=err, node=b.node, allt=allt)
dist.barrier(); dist.destroy_process_group()
"""


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29531", str(script), ROOT, str(tmp_path)],
                          env=env, timeout=300)
    text, offs = W.generate(3, 192 << 10)
    whole = O.parse_batch(text, offs)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    stmt = np.concatenate([p["stmt"] for p in parts])
    err = np.concatenate([p["err"] for p in parts])
    node = np.concatenate([p["node"] for p in parts])
    assert np.array_equal(stmt, whole.stmt)
    assert np.array_equal(err, whole.err)
    assert np.array_equal(node, whole.node)
    assert int(parts[0]["allt"][:, 0].sum()) == len(offs) - 1
