"""Extended differential fuzz on the GPU: mutated statements from every seed pool, tokens + nodes + errors vs the oracle."""
import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import fuzz, parity as P
from nutdb_b200 import gpu, workload as W
ctx = gpu.Context(0)
pools = {}
pools['corpus'] = W.corpus_statements() + fuzz.EXTRA_SEEDS
pools['simple'] = fuzz.SIMPLE_SEEDS
for cfg in (2, 3, 4):
    text, offs = W.generate(cfg, 96 << 10, seed=1000 + cfg)
    pools['cfg%d' % cfg] = [bytes(text[int(offs[i]):int(offs[i + 1])]) for i in range(len(offs) - 1)][:400]
pools['hex'] = [b"select 0x1F, 0X2a, 0x, 0x0 from t where a = 0xdeadBEEF limit 0x10", b"select 0x1G, 0xzz, 00x1, 0x1.5, .0x1, 1.0x2, 0x1_2, x0x1, 0x1x",
                b"insert into t values (0x1, '0x2', 0xFF), (0xa,0xb,0xc)", b"select 'a\\u{41}', 'b\\u{110000}', \"c\\u{+7a}\" from `t` where x = 'it''s' -- c\n and y"]
import test_emul_parity as T
pools['wide'] = T.WIDE + T.WIDE_AUTOMATON + [d[:400] for d in T.DEEP] + [s.encode() for s in T.fold_statements(n_random=200)[::5]]
pools['escaped'] = [s.encode() for s in fuzz.escaped_literal_statements(600, seed=77)]
pools['predicates'] = T.PREDICATES + T.PREDICATES_AUTOMATON + T.JOINS + T.JOINS_AUTOMATON + T.CASES + T.CASES_AUTOMATON + T.QUALIFIED + T.QUALIFIED_AUTOMATON + fuzz.SIMPLE_SEEDS[:40]
t0 = time.time(); total = 0; nbad = 0
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 240.0
sd = int(sys.argv[2]) if len(sys.argv) > 2 else 5000  # first seed: pass another base to cover new ground
while time.time() - t0 < budget:
    for name, pool in pools.items():
        st = fuzz.fuzz_statements(pool, 20000, seed=sd, max_mut=4)
        text, offs = P.make_batch(st)
        got = ctx.parse_batch(text, offs)
        bad = P.compare_with_oracle(got, text, offs)
        total += len(st)
        if bad:
            nbad += 1
            print('MISMATCH', name, sd, bad[:3], flush=True)
        sd += 1
print('fuzzed', total, 'statements in', round(time.time() - t0), 's; mismatching batches:', nbad)
