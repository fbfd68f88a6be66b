"""Shared comparison of a batch result (GPU or host-emulated device code) against the oracle."""
import re

import numpy as np

import oracle_lib as O

TT_ESCAPED_SQ, TT_ESCAPED_DQ = 5, 6   # include/nutdb_gpu.h NUTDB_TT_Escaped{SQ,DQ}StringLiteral (token.rs ordinals)
KW_SAMPLE = 20000                     # escaped literals whose side byte is recomputed from the text per comparison
SIDE_BYTE = True                      # (off for the emulated round-1 warp lexer, which no kernel launches any more)


def make_batch(stmts):
    raws = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in stmts]
    offs = np.zeros(len(raws) + 1, np.uint64)
    offs[1:] = np.cumsum([len(r) for r in raws])
    return np.frombuffer(b"".join(raws) + b"\0" * 16, np.uint8), offs


def describe(stmts, i):
    s = stmts[i]
    return repr(s if len(s) < 300 else s[:300])


def compare_with_oracle(got, text, offs, stmts=None, check_tokens=True):
    """Bit-exact comparison; returns list of mismatch descriptions (empty = parity)."""
    want = O.parse_batch(text, offs)
    bad = []
    n = len(offs) - 1
    gs, ws = got.stmt, want.stmt

    def sql(i):
        return bytes(text[int(offs[i]):int(offs[i + 1])])

    st_bad = np.nonzero(gs["status"] != ws["status"])[0]
    for i in st_bad[:5]:
        bad.append(f"stmt {i}: status {gs['status'][i]} != oracle {ws['status'][i]}: {sql(i)!r}")
    nc_bad = np.nonzero(gs["node_count"] != ws["node_count"])[0]
    for i in nc_bad[:5]:
        bad.append(f"stmt {i}: node_count {gs['node_count'][i]} != oracle {ws['node_count'][i]}: {sql(i)!r}")
    tu_bad = np.nonzero(gs["tok_used"] != ws["tok_used"])[0]
    for i in tu_bad[:5]:
        bad.append(f"stmt {i}: tok_used {gs['tok_used'][i]} != oracle {ws['tok_used'][i]}: {sql(i)!r}")
    if bad:
        return bad
    # nodes: same counts per statement => compare the concatenation when layouts agree
    if not np.array_equal(gs["node_begin"], ws["node_begin"]):
        bad.append("node_begin arrays differ although node counts agree")
        return bad
    if len(got.node) != len(want.node) or not np.array_equal(got.node, want.node):
        m = min(len(got.node), len(want.node))
        d = np.nonzero(got.node[:m] != want.node[:m])[0]
        k = int(d[0]) if len(d) else m
        i = int(np.searchsorted(ws["node_begin"], k, side="right") - 1)
        bad.append(f"stmt {i}: node {k - int(ws['node_begin'][i])} differs: got {got.node[k] if k < len(got.node) else None}"
                   f" oracle {want.node[k] if k < len(want.node) else None}: {sql(i)!r}")
        return bad
    # errors
    ge = np.sort(got.err, order="stmt")
    we = np.sort(want.err, order="stmt")
    if len(ge) != len(we) or not np.array_equal(ge, we):
        m = min(len(ge), len(we))
        d = np.nonzero(ge[:m] != we[:m])[0]
        k = int(d[0]) if len(d) else m
        i = int(we["stmt"][k]) if k < len(we) else -1
        bad.append(f"error record {k} differs: got {ge[k] if k < len(ge) else None} oracle {we[k] if k < len(we) else None}"
                   f": {sql(i)!r}")
        return bad
    if check_tokens:
        # tokens the reference pulled == the first tok_used tokens of each statement (one gather per array, no
        # per-statement loop: the 1 GiB-class workloads are compared with tokens too)
        u = ws["tok_used"].astype(np.int64)
        u[(offs[1:] == offs[:-1])] = 0
        tot = int(u.sum())
        if tot:
            first = np.cumsum(u) - u                       # position of each statement's first pulled token in the gather
            rel = np.arange(tot, dtype=np.int64) - np.repeat(first, u)
            gi = np.repeat(gs["tok_begin"].astype(np.int64), u) + rel
            wi = np.repeat(ws["tok_begin"].astype(np.int64), u) + rel
            if gi.max(initial=-1) >= len(got.tok_type):
                bad.append("pulled tokens run past the token arrays")
                return bad
            word = want.tok_type[wi] == O.TT_KEYWORD_OR_IDENTIFIER
            diff = ((got.tok_type[gi] != want.tok_type[wi]) | (got.tok_start[gi] != want.tok_start[wi]) |
                    (got.tok_end[gi] != want.tok_end[wi]) | (word & (got.tok_kw[gi] != want.tok_kw[wi])))
            for k in np.nonzero(diff)[0][:4]:
                i = int(np.searchsorted(first, k, side="right") - 1)
                j = int(k - first[i])
                bad.append(f"stmt {i}: pulled token {j} differs: got ({got.tok_type[gi[k]]}, {got.tok_start[gi[k]]}, "
                           f"{got.tok_end[gi[k]]}, kw {got.tok_kw[gi[k]]}) oracle ({want.tok_type[wi[k]]}, "
                           f"{want.tok_start[wi[k]]}, {want.tok_end[wi[k]]}, kw {want.tok_kw[wi[k]]}): {sql(i)!r}")
            # side byte of escaped literals (no oracle counterpart: the reference has no such field): 1 iff the literal
            # holds no backslash-u escape, recomputed here by walking backslash pairs (literal.rs:59-63)
            es = np.nonzero((want.tok_type[wi] == TT_ESCAPED_SQ) | (want.tok_type[wi] == TT_ESCAPED_DQ))[0]
            if SIDE_BYTE and not bad and len(es):
                if len(es) > KW_SAMPLE:
                    es = es[np.linspace(0, len(es) - 1, KW_SAMPLE).astype(np.int64)]
                st_of = np.searchsorted(first, es, side="right") - 1
                for k, i in zip(es.tolist(), st_of.tolist()):
                    lit = bytes(text[int(offs[i]) + int(want.tok_start[wi[k]]):int(offs[i]) + int(want.tok_end[wi[k]])])
                    clean = all(m.group(1) != b"u" for m in re.finditer(rb"\\(.)", lit, re.S))
                    if int(got.tok_kw[gi[k]]) != int(clean):
                        bad.append(f"stmt {i}: escaped literal {lit!r}: side byte {got.tok_kw[gi[k]]}, expected {int(clean)}")
                        if len(bad) >= 4:
                            break
    return bad
