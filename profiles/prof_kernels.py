"""Per-kernel timing of one workload through the library (events around every launch): the command the ncu --set full\ncaptures in profiles/ were taken under.  usage: prof_kernels.py <config 1-4> <bytes>"""
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from nutdb_b200 import gpu, workload as W
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nbytes = int(sys.argv[2]) if len(sys.argv) > 2 else (256 << 20)
ctx = gpu.Context(0)
text, offs = W.corpus(nbytes) if cfg == 1 else W.generate(cfg, nbytes)
ctx.set_profiling(True)
for i in range(3):
    got = ctx.parse_batch(text, offs, flags=gpu.F_NO_HOST_COPY)
print("cfg", cfg, "bytes", int(offs[-1]), "stmts", len(offs) - 1, "tok", got.n_tok, "node", got.n_node, "punts", ctx.exact_lexed_statements(), "slow", ctx.slow_statements())
print(ctx.timing())
agg = {}
for k, ms in ctx.kernel_timing():
    agg[k] = agg.get(k, 0) + ms
print({k: round(v, 3) for k, v in agg.items()})
if ctx.exact_lexed_statements() and len(sys.argv) > 3:
    got = ctx.parse_batch(text, offs)
    tb = got.stmt["tok_begin"].astype(np.int64)
    d = np.nonzero(np.diff(tb) < 0)[0]
    cut = int(tb[d[0] + 1]) if len(d) else 0
    idx = np.nonzero(tb >= tb.max() - 5000)[0]
    mx = np.sort(tb)[::-1]
    # statements lexed into the extra region have the largest tok_begin values
    order = np.argsort(tb)[::-1][:ctx.exact_lexed_statements()]
    for i in order[:12]:
        print("punted stmt", i, "off%8192", int(offs[i]) % 8192, int(offs[i + 1]) % 8192, bytes(text[int(offs[i]):int(offs[i + 1])]))
