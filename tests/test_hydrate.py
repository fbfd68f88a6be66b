"""Host-side re-hydration (nutdb_fmt_debug / nutdb_fmt_error in libnutdb_gpu.so): the flat arrays
carry everything needed to rebuild the reference's `Statement` / `ParseError`.  Here the flat
arrays come from the oracle, so this runs without a GPU; test_gpu_parity.py proves the GPU
produces the same arrays.  Also checks that the library exports every symbol of include/nutdb_gpu.h."""
import ctypes as C
import os
import re

import numpy as np

import fuzz
import oracle_lib as O
import parity as P
from nutdb_b200 import gpu, workload as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def batch_struct(b, keep):
    raw = gpu.NutdbBatch()
    raw.n_stmt, raw.n_tok, raw.n_node, raw.n_err = len(b.stmt), len(b.tok_type), len(b.node), len(b.err)
    for name in ("stmt", "node", "err"):
        arr = np.ascontiguousarray(getattr(b, name))
        keep.append(arr)
        setattr(raw, name, arr.ctypes.data if len(arr) else None)
    return raw


def fmt(fn, raw, i, sql):
    need = fn(C.byref(raw), i, sql, len(sql), None, 0)
    buf = C.create_string_buffer(max(need, 1))
    fn(C.byref(raw), i, sql, len(sql), buf, len(buf))
    return buf.value.decode("utf-8")


def check(stmts):
    L = gpu.lib()
    text, offs = P.make_batch(stmts)
    b = O.parse_batch(text, offs)
    keep = []
    raw = batch_struct(b, keep)
    n_ok = 0
    for i, s in enumerate(stmts):
        s = s.encode() if isinstance(s, str) else s
        want = O.parse(s)
        if want.ok:
            assert fmt(L.nutdb_fmt_debug, raw, i, s) == want.debug, s
            n_ok += 1
        else:
            assert fmt(L.nutdb_fmt_error, raw, i, s) == want.error, s
    return n_ok


def test_corpus_debug_strings():
    assert check(W.corpus_statements()) == len(W.corpus_statements())


def test_known_answer_app_d_1():
    # SURVEY.md App. D.1, derived by hand from the derive(Debug) layout of the reference AST
    L = gpu.lib()
    s = b"SELECT * FROM table WHERE 1 = 1"
    text, offs = P.make_batch([s])
    keep = []
    raw = batch_struct(O.parse_batch(text, offs), keep)
    assert fmt(L.nutdb_fmt_debug, raw, 0, s) == (
        "Select(SelectStmt { query: Single(QueryBody { with: None, distinct: None, columns: [QueryExpr { inner: "
        "Identifier(Identifier { name: Wildcard, qualifier: None }), alias: None }], from: Some(FromClause { source: "
        "QuerySource { inner: Table(\"table\"), alias: None } }), joins: [], where: Some(WhereClause { condition: "
        "Literal(Boolean(true)) }), group_by: None, having: None, order_by: None, limit: None }) })")


def test_extra_seeds_and_errors():
    errs = ["", "SELECT 1d", "SELECT a FROM t ORDER BY a ASC", "SELECT $0", "SELECT 'abc", "select café",
            "select `a\nb`", "select 1 /* x", "select @1", "select $", "select a ! b", "select 0q", "select 1.5.2",
            "CREATE VIEW v AS SELECT 1", "CREATE TABLE t (a Int8) COMMENT 'x' COMMENT 'y'",
            "INSERT INTO t VALUES (1, 2), (3)", "select '\\u{110000}'", "select 99999999999999999999999999999999999999999",
            "select a from t limit 0x", "FROB", "select (1", "select x not y", "select a from t join u"]
    check(fuzz.EXTRA_SEEDS + [e.encode() for e in errs])


def test_mutation_fuzz_strings():
    check(fuzz.fuzz_statements(W.corpus_statements() + fuzz.EXTRA_SEEDS, 1500, seed=21, max_mut=3))


def test_synthetic_strings():
    for cfg in (2, 3, 4):
        text, offs = W.generate(cfg, 96 << 10)
        check([bytes(text[int(offs[i]):int(offs[i + 1])]) for i in range(len(offs) - 1)])


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nutdb_gpu.h")).read()
    names = set(re.findall(r"\b(nutdb_(?:gpu|fmt)_\w+)\s*\(", hdr))
    assert len(names) >= 14
    L = gpu.lib()
    for n in names:
        assert hasattr(L, n), n


# ---- the 32-bit wire form of a node (NUTDB_PN_* in nutdb_gpu.h): a reference encoder in numpy / Python -------------
NODE_EXT_DT = np.dtype([("index", "<u4"), ("hdr", "<u4"), ("a", "<u4"), ("b", "<u4")])


def encode_wire(stmt, node, gap_max=1021, len_max=510, size_max=0x7FFFE):
    """NutdbNode records (oracle) -> (wire words, side table), statement by statement, by the rules of the header."""
    words = np.zeros(len(node), np.uint32)
    ext = []
    for s in stmt:
        if s["status"] != 0:
            continue
        b0, cnt = int(s["node_begin"]), int(s["node_count"])
        pos = 0
        for j in range(cnt):
            nd = node[b0 + j]
            kind, sub, aux, a, b = int(nd["kind"]), int(nd["sub"]), int(nd["aux"]), int(nd["a"]), int(nd["b"])
            w = kind | ((sub & 31) << 7) | ((aux & 1) << 12)
            if kind >= 32:
                size = j - a
                esc = sub > 31 or aux > 1 or size > size_max
                w |= (0x7FFFF if esc else size) << 13
                if esc:
                    ext.append((b0 + j, kind | sub << 8 | aux << 16, a, 0))
            elif a == 0 and b == 0:
                w |= (1023 << 13) | (511 << 23)
            else:
                esc = sub > 31 or aux > 1 or a < pos or a - pos > gap_max or b - a > len_max
                if esc:
                    w |= (1022 << 13) | (511 << 23)
                    ext.append((b0 + j, kind | sub << 8 | aux << 16, a, b - a))
                else:
                    w |= ((a - pos) << 13) | ((b - a) << 23)
                pos = b
            words[b0 + j] = w
    return words, np.array(ext, NODE_EXT_DT) if ext else np.zeros(0, NODE_EXT_DT)


def expand_wire(stmt, words, ext):
    keep = [np.ascontiguousarray(stmt), np.ascontiguousarray(words), np.ascontiguousarray(ext)]
    raw = gpu.NutdbBatch()
    raw.n_stmt, raw.n_node = len(stmt), len(words)
    raw.stmt, raw.pnode = keep[0].ctypes.data, keep[1].ctypes.data if len(words) else None
    raw.n_ext, raw.ext = len(ext), keep[2].ctypes.data if len(ext) else None
    out = np.zeros(len(words), gpu.NODE_DT)
    rc = gpu.lib().nutdb_batch_expand_nodes(C.byref(raw), out.ctypes.data)
    return rc, out


def test_wire_nodes_expand_to_the_oracle_records():
    """nutdb_batch_expand_nodes (what a host does with the 4-byte nodes that cross PCIe) reproduces the oracle's
    NutdbNode records from their wire form -- with the real field limits and with tiny ones that push most nodes
    through the side table."""
    long_lit = "select '" + "x" * 700 + "', a /* " + "c" * 3000 + " */ , 'y', " + ", ".join("f(%d)" % i for i in range(50))
    stmts = W.corpus_statements() + fuzz.EXTRA_SEEDS + [long_lit, "select 1d", "", "select a.b, t.* from t"] + \
        fuzz.fuzz_statements(W.corpus_statements(), 1500, seed=12, max_mut=3)
    text, offs = P.make_batch(stmts)
    b = O.parse_batch(text, offs)
    for limits in ((1021, 510, 0x7FFFE), (3, 2, 4), (0, 0, 0)):
        words, ext = encode_wire(b.stmt, b.node, *limits)
        rc, out = expand_wire(b.stmt, words, ext)
        assert rc == 0
        assert np.array_equal(out, b.node), limits
    assert len(encode_wire(b.stmt, b.node)[1]) >= 2          # the long literal and the token behind the long comment
    words, ext = encode_wire(b.stmt, b.node, 3, 2, 4)
    rc, _ = expand_wire(b.stmt, words, ext[:-1])             # an escape without its entry is refused, not guessed
    assert rc != 0
