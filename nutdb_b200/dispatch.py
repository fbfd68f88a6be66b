"""Multi-GPU batch dispatcher: statements are independent (reference `Parser::parse` keeps no state
across calls, src/parser/mod.rs:21-37), so a batch shards by contiguous statement ranges balanced by
bytes -- one process per GPU, no collective on the data path.  Only the per-shard totals (statements,
tokens, nodes, errors) are exchanged so that shard-local indices can be rebased into batch-global ones.
"""
import numpy as np


def split_statements(offs, nparts):
    """-> [(lo, hi)] statement ranges, contiguous, covering everything, balanced by bytes."""
    offs = np.asarray(offs, np.uint64)
    n = len(offs) - 1
    total = int(offs[-1] - offs[0])
    cuts = [0]
    for p in range(1, nparts):
        target = np.uint64(int(offs[0]) + total * p // nparts)   # same dtype as offs: no O(n) conversion per cut
        s = int(np.searchsorted(offs, target, side="left"))
        cuts.append(min(max(s, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(nparts)]


def shard(text, offs, rank, world):
    """This rank's statements: (text view, offsets rebased to the view, first statement index)."""
    lo, hi = split_statements(offs, world)[rank]
    o = np.asarray(offs[lo:hi + 1], np.uint64)
    b, e = int(o[0]), int(o[-1])
    return text[b:e + 64] if e + 64 <= len(text) else np.concatenate([text[b:e], np.zeros(64, np.uint8)]), o - o[0], lo


def totals(batch):
    return np.array([batch.n_stmt, batch.n_tok, batch.n_node, batch.n_err], np.int64)


def rebase(stmt, err, bases):
    """Turn shard-local indices into batch-global ones. bases = exclusive prefix [stmt, tok, node, err]."""
    # the records hold 32-bit indices: a batch-global prefix beyond 2^32 (about 16 GiB of SQL in tokens) cannot be
    # expressed in them -- such batches keep chunk-local records plus the prefix (what nutdb_gpu_mctx_* delivers)
    top = [int(bases[1]) + (int(stmt["tok_begin"].max()) if len(stmt) else 0),
           int(bases[2]) + (int(stmt["node_begin"].max()) if len(stmt) else 0),
           int(bases[0]) + (int(err["stmt"].max()) if len(err) else 0)]
    if max(top) >= 1 << 32:
        raise OverflowError("batch-global indices exceed 32 bits: keep the records shard-local and carry the prefix")
    stmt = stmt.copy()
    stmt["tok_begin"] += np.uint32(bases[1])
    stmt["node_begin"] += np.uint32(bases[2])
    err = err.copy()
    err["stmt"] += np.uint32(bases[0])
    return stmt, err


def exchange_totals(local_totals, dist=None):
    """all-gather of the 4 per-shard totals (the only cross-GPU values). -> (world x 4) int64, exclusive prefix."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        allt = local_totals[None, :]
    else:
        import torch
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.from_numpy(local_totals.copy()).to(dev)
        out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        allt = np.stack([o.cpu().numpy() for o in out])
    prefix = np.zeros_like(allt)
    prefix[1:] = np.cumsum(allt, axis=0)[:-1]
    return allt, prefix


def parse_sharded(parse_fn, text, offs, rank, world, dist=None):
    """Parses this rank's shard with parse_fn(text, offs) -> batch and returns
    (batch, global statement records, global error records, all totals).  Outputs stay sharded."""
    t, o, first = shard(text, offs, rank, world)
    b = parse_fn(t, o)
    allt, prefix = exchange_totals(totals(b), dist)
    stmt, err = rebase(b.stmt, b.err, prefix[rank])
    return b, stmt, err, allt
