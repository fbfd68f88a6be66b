"""ctypes binding of oracle/liboracle.so -- the CPU restatement of the reference parser.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module.  The product (nutdb_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None


class NutdbNode(C.Structure):
    _fields_ = [("kind", C.c_uint8), ("sub", C.c_uint8), ("aux", C.c_uint16), ("parent", C.c_uint32),
                ("a", C.c_uint32), ("b", C.c_uint32)]


class NutdbStmt(C.Structure):
    _fields_ = [("status", C.c_uint32), ("tok_begin", C.c_uint32), ("tok_count", C.c_uint32),
                ("node_begin", C.c_uint32), ("node_count", C.c_uint32), ("tok_used", C.c_uint32)]


class NutdbError(C.Structure):
    _fields_ = [("stmt", C.c_uint32), ("cls", C.c_uint16), ("code", C.c_uint16), ("line", C.c_uint32),
                ("col", C.c_uint32), ("pos", C.c_uint32), ("a", C.c_uint32), ("b", C.c_uint32), ("c", C.c_uint32)]


NODE_DT = np.dtype([("kind", "u1"), ("sub", "u1"), ("aux", "<u2"), ("parent", "<u4"), ("a", "<u4"), ("b", "<u4")])
STMT_DT = np.dtype([("status", "<u4"), ("tok_begin", "<u4"), ("tok_count", "<u4"), ("node_begin", "<u4"),
                    ("node_count", "<u4"), ("tok_used", "<u4")])
ERR_DT = np.dtype([("stmt", "<u4"), ("cls", "<u2"), ("code", "<u2"), ("line", "<u4"), ("col", "<u4"),
                   ("pos", "<u4"), ("a", "<u4"), ("b", "<u4"), ("c", "<u4")])


TT_KEYWORD_OR_IDENTIFIER = 0   # include/nutdb_gpu.h NUTDB_TT_KeywordOrIdentifier (token.rs:6)


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    src_newer = not os.path.exists(path) or any(
        os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(path)
        for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".hpp")))
    if src_newer:
        build()
    L = C.CDLL(path)
    L.ora_tokenize.restype = C.c_int64
    L.ora_tokenize.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                               C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint32),
                               C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_char_p, C.c_size_t]
    L.ora_get_pos.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.ora_unescape.restype = C.c_int
    L.ora_unescape.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.ora_keyword_id.argtypes = [C.c_char_p, C.c_size_t]
    L.ora_keyword_text.restype = C.c_char_p
    L.ora_parse.restype = C.c_void_p
    L.ora_parse.argtypes = [C.c_char_p, C.c_size_t]
    L.ora_free.argtypes = [C.c_void_p]
    L.ora_status.argtypes = [C.c_void_p]
    L.ora_debug.restype = C.c_char_p
    L.ora_debug.argtypes = [C.c_void_p]
    L.ora_error_display.restype = C.c_char_p
    L.ora_error_display.argtypes = [C.c_void_p]
    L.ora_error_record.argtypes = [C.c_void_p, C.POINTER(NutdbError)]
    L.ora_n_nodes.restype = C.c_size_t
    L.ora_n_nodes.argtypes = [C.c_void_p]
    L.ora_nodes.restype = C.c_void_p
    L.ora_nodes.argtypes = [C.c_void_p]
    L.ora_m_alg.restype = C.c_size_t
    L.ora_m_alg.argtypes = [C.c_void_p]
    L.ora_n_pulled.restype = C.c_size_t
    L.ora_n_pulled.argtypes = [C.c_void_p]
    L.ora_pulled.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ora_parse_batch.restype = C.c_void_p
    L.ora_parse_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
    L.ora_batch_free.argtypes = [C.c_void_p]
    L.ora_batch_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint64)] * 5
    for f in ("stmt", "node", "err", "tok_type", "tok_kw", "tok_start", "tok_end"):
        getattr(L, "ora_batch_" + f).restype = C.c_void_p
        getattr(L, "ora_batch_" + f).argtypes = [C.c_void_p]
    L.ora_bench.restype = C.c_double
    L.ora_bench.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64),
                            C.POINTER(C.c_uint64)]
    _LIB = L
    return L


def _b(sql):
    return sql.encode("utf-8") if isinstance(sql, str) else bytes(sql)


def tokenize(sql):
    """Reference tokenizer over the whole input.  -> (tokens [(type,start,end)], error | None)."""
    L = lib()
    raw = _b(sql)
    cap = len(raw) + 2
    ty = np.zeros(cap, np.uint8)
    st = np.zeros(cap, np.uint32)
    en = np.zeros(cap, np.uint32)
    et, es = C.c_int(0), C.c_int(0)
    ep, el, ec = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    ctx = C.create_string_buffer(256)
    n = L.ora_tokenize(raw, len(raw), ty.ctypes.data, st.ctypes.data, en.ctypes.data, cap, C.byref(et), C.byref(es),
                       C.byref(ep), C.byref(el), C.byref(ec), ctx, 256)
    toks = [(int(ty[i]), int(st[i]), int(en[i])) for i in range(n)]
    err = None
    if es.value:
        err = dict(type=et.value, site=es.value, pos=ep.value, line=el.value, col=ec.value,
                   ctx=ctx.value.decode("utf-8", "replace"))
    return toks, err


def tokenize_arrays(sql):
    """tokenize() without the per-token Python objects: (types u8, starts u32, ends u32, error | None)."""
    L = lib()
    raw = _b(sql)
    cap = len(raw) + 2
    ty = np.zeros(cap, np.uint8)
    st = np.zeros(cap, np.uint32)
    en = np.zeros(cap, np.uint32)
    et, es = C.c_int(0), C.c_int(0)
    ep, el, ec = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    ctx = C.create_string_buffer(256)
    n = L.ora_tokenize(raw, len(raw), ty.ctypes.data, st.ctypes.data, en.ctypes.data, cap, C.byref(et), C.byref(es),
                       C.byref(ep), C.byref(el), C.byref(ec), ctx, 256)
    err = dict(type=et.value, site=es.value, pos=ep.value) if es.value else None
    return ty[:n], st[:n], en[:n], err


TT_SEMICOLON, TT_WHITESPACE, TT_EOF = 13, 38, 39   # include/nutdb_gpu.h (token.rs ordinals)


def split_statements(buf):
    """Expected output of the raw-buffer statement splitter, derived from the REFERENCE TOKENIZER (not from the
    product's context automaton): one sequential tokenizer walk over the whole buffer (tokenizer/mod.rs:66-112); every
    SemiColon token ends a statement (the statement includes it); what follows the last one is a final statement
    unless it is only Whitespace tokens.  Returns (offsets u64, valid_upto): the tokenizer stops at its first error, so
    only the offsets up to `valid_upto` are defined by it (valid_upto = len(buf) + 1 when the buffer lexes cleanly)."""
    raw = _b(buf)
    ty, st, en, err = tokenize_arrays(raw)
    semi = en[ty == TT_SEMICOLON].astype(np.uint64)
    offs = np.concatenate([np.zeros(1, np.uint64), semi])
    if err is not None:   # (the error may be reported at EOF for a token that began earlier: pinned up to the last good token)
        return offs, int(en[-1]) if len(en) else 0
    last = int(semi[-1]) if len(semi) else 0
    tail = (en > last) & (ty != TT_WHITESPACE) & (ty != TT_EOF)
    if tail.any() or (last < len(raw) and not len(ty)):
        offs = np.concatenate([offs, np.array([len(raw)], np.uint64)])
    return offs, len(raw) + 1


def get_pos(sql, cursor):
    L = lib()
    raw = _b(sql)
    l, c = C.c_uint32(0), C.c_uint32(0)
    L.ora_get_pos(raw, len(raw), cursor, C.byref(l), C.byref(c))
    return l.value, c.value


def unescape(raw, quote):
    L = lib()
    r = _b(raw)
    out = C.create_string_buffer(4 * len(r) + 16)
    n = C.c_size_t(0)
    rc = L.ora_unescape(r, len(r), ord(quote), out, len(out), C.byref(n))
    return rc, out.raw[:n.value].decode("utf-8")


class ParseResult:
    def __init__(self, status, debug, error, rec, nodes, pulled, m_alg):
        self.status, self.debug, self.error, self.rec = status, debug, error, rec
        self.nodes, self.pulled, self.m_alg = nodes, pulled, m_alg

    @property
    def ok(self):
        return self.status == 0


def parse(sql):
    """Reference Parser::parse on one statement."""
    L = lib()
    raw = _b(sql)
    h = L.ora_parse(raw, len(raw))
    try:
        status = L.ora_status(h)
        debug = L.ora_debug(h).decode("utf-8")
        error = L.ora_error_display(h).decode("utf-8")
        rec = NutdbError()
        L.ora_error_record(h, C.byref(rec))
        nn = L.ora_n_nodes(h)
        nodes = np.zeros(nn, NODE_DT)
        if nn:
            C.memmove(nodes.ctypes.data, L.ora_nodes(h), nn * NODE_DT.itemsize)
        npul = L.ora_n_pulled(h)
        ty = np.zeros(npul, np.uint8)
        st = np.zeros(npul, np.uint32)
        en = np.zeros(npul, np.uint32)
        if npul:
            L.ora_pulled(h, ty.ctypes.data, st.ctypes.data, en.ctypes.data)
        pulled = [(int(ty[i]), int(st[i]), int(en[i])) for i in range(npul)]
        recd = {k: getattr(rec, k) for k, _ in NutdbError._fields_}
        return ParseResult(status, debug, error, recd, nodes, pulled, L.ora_m_alg(h))
    finally:
        L.ora_free(h)


class Batch:
    pass


def parse_batch(text, offs, nthreads=0):
    """Parse a batch; returns numpy arrays laid out like NutdbBatch."""
    L = lib()
    text = np.ascontiguousarray(np.frombuffer(text, np.uint8) if not isinstance(text, np.ndarray) else text)
    offs = np.ascontiguousarray(offs, np.uint64)
    n = len(offs) - 1
    if nthreads <= 0:
        nthreads = os.cpu_count() or 1
    h = L.ora_parse_batch(text.ctypes.data, offs.ctypes.data, n, nthreads)
    try:
        cnt = [C.c_uint64(0) for _ in range(5)]
        L.ora_batch_counts(h, *[C.byref(c) for c in cnt])
        n_node, n_err, n_tok, t_alg, m_alg = [c.value for c in cnt]

        def arr(ptr, count, dt):
            a = np.zeros(count, dt)
            if count:
                C.memmove(a.ctypes.data, ptr, count * a.itemsize)
            return a

        b = Batch()
        b.stmt = arr(L.ora_batch_stmt(h), n, STMT_DT)
        b.node = arr(L.ora_batch_node(h), n_node, NODE_DT)
        b.err = arr(L.ora_batch_err(h), n_err, ERR_DT)
        b.tok_type = arr(L.ora_batch_tok_type(h), n_tok, np.uint8)
        b.tok_kw = arr(L.ora_batch_tok_kw(h), n_tok, np.uint8)
        b.tok_start = arr(L.ora_batch_tok_start(h), n_tok, np.uint32)
        b.tok_end = arr(L.ora_batch_tok_end(h), n_tok, np.uint32)
        b.t_alg, b.m_alg = t_alg, m_alg
        return b
    finally:
        L.ora_batch_free(h)


def bench(text, offs, nthreads, reps=3):
    """CPU baseline: seconds (best of reps) to parse+drop the whole batch on nthreads threads."""
    L = lib()
    text = np.ascontiguousarray(text)
    offs = np.ascontiguousarray(offs, np.uint64)
    ok, tk = C.c_uint64(0), C.c_uint64(0)
    s = L.ora_bench(text.ctypes.data, offs.ctypes.data, len(offs) - 1, nthreads, reps, C.byref(ok), C.byref(tk))
    return s, ok.value, tk.value
