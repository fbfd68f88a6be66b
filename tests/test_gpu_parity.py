"""Parity of the CUDA path (through the C ABI of libnutdb_gpu.so) with the oracle: bit-exact
statement records, flat ASTs, error records and pulled tokens.  Needs a B200: run with -m gpu."""
import numpy as np
import pytest

import fuzz
import oracle_lib as O
import parity as P
from nutdb_b200 import workload as W

pytestmark = pytest.mark.gpu

CORPUS = W.corpus_statements()
APP_D = ["SELECT * FROM table WHERE 1 = 1", "", "  ", ";", "-- c", "SELECT 1d", "SELECT a FROM t ORDER BY a ASC",
         "SELECT $0", "SELECT 'abc", "select 1; 'oops", "\t1d", "select `你 好`, 'he''llo' -- x"]


@pytest.fixture(scope="module")
def ctx():
    from nutdb_b200 import gpu
    c = gpu.Context(0)
    yield c
    c.close()


def check(ctx, stmts):
    text, offs = P.make_batch(stmts)
    got = ctx.parse_batch(text, offs)
    bad = P.compare_with_oracle(got, text, offs)
    assert not bad, "\n".join(bad)
    return got


def test_corpus_and_known_vectors(ctx):
    got = check(ctx, CORPUS + APP_D)
    assert (got.stmt["status"][:len(CORPUS)] == 0).all()  # tests/parser_test.rs:19-34


def test_single_statement_entry(ctx):
    # nutdb_gpu_parse == Parser::parse(sql) (mod.rs:27)
    import ctypes as C
    from nutdb_b200 import gpu
    for s in [b"SELECT * FROM table WHERE 1 = 1", b"", b"SELECT 1d"]:
        raw = gpu.NutdbBatch()
        rc = gpu.lib().nutdb_gpu_parse(ctx._h, s, len(s), C.byref(raw))
        assert rc == 0
        b = gpu.Batch(ctx, raw, True)
        want = O.parse(s)
        assert int(b.stmt["status"][0]) == want.status
        assert np.array_equal(b.node, want.nodes)


def test_empty_and_degenerate_batches(ctx):
    b = ctx.parse_batch(np.zeros(16, np.uint8), np.zeros(1, np.uint64))
    assert b.n_stmt == 0 and b.n_node == 0 and b.n_err == 0
    check(ctx, ["", "", ""])
    check(ctx, ["", "select 1", "", "", "select 2", ""])
    check(ctx, ["x" * 8192, "select 1"])             # statement = exactly one tile
    check(ctx, ["select '" + "a" * 20000 + "'", "select 1 /* " + "c" * 9000 + " */, 2"])  # tokens spanning tiles


@pytest.mark.parametrize("config", [2, 3, 4])
def test_synthetic_config(ctx, config):
    text, offs = W.generate(config, 4 << 20)
    got = ctx.parse_batch(text, offs)
    bad = P.compare_with_oracle(got, text, offs)
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_mutation_fuzz(ctx, seed):
    stmts = fuzz.fuzz_statements(CORPUS + fuzz.EXTRA_SEEDS, 8000, seed=seed, max_mut=5)
    check(ctx, stmts)


def test_fast_path_fuzz(ctx):
    text, offs = W.generate(2, 64 << 10, seed=99)
    seeds = [bytes(text[int(offs[i]):int(offs[i + 1])]) for i in range(len(offs) - 1)][:300] + fuzz.SIMPLE_SEEDS
    for seed in (100, 101, 102):
        stmts = fuzz.fuzz_statements(seeds, 8000, seed=seed, max_mut=3)
        check(ctx, stmts)
        assert 0 < ctx.slow_statements() < len(stmts)
    text, offs = W.generate(2, 1 << 20)
    ctx.parse_batch(text, offs)
    assert ctx.slow_statements() == 0     # config 2 is parsed entirely by the straight-line parser


def test_warp_lexer_exact_walker_split(ctx):
    for config in (2, 4):
        text, offs = W.generate(config, 1 << 20)
        ctx.parse_batch(text, offs)
        assert ctx.exact_lexed_statements() == 0
    text, offs = W.generate(3, 1 << 20)
    got = ctx.parse_batch(text, offs)
    assert 0 < ctx.exact_lexed_statements() < 0.06 * (len(offs) - 1)
    bad = P.compare_with_oracle(got, text, offs)
    assert not bad, "\n".join(bad)
    # hex literals, $n, @name, long tokens: valid text that only the exact walker lexes
    stmts = ["select 0x1F, 0xg, $1 2, a" + "b" * 70 + ", 1", "set @cfg = 1", "select 1 <<= 2", "select 'x\\u\\\\'"]
    got = check(ctx, stmts)
    assert ctx.exact_lexed_statements() >= 3
    assert got.stmt["status"].tolist()[-1] == 4


@pytest.mark.parametrize("seed", [300, 301, 302])
def test_mutation_fuzz_lexer_stress(ctx, seed):
    text, offs = W.generate(3, 64 << 10, seed=98)
    pool = [bytes(text[int(offs[i]):int(offs[i + 1])]) for i in range(len(offs) - 1)][:300]
    check(ctx, fuzz.fuzz_statements(pool, 8000, seed=seed, max_mut=4))


def test_full_token_stream_incl_whitespace_and_comments(ctx):
    from nutdb_b200 import gpu
    stmts = [s for s in CORPUS + [x.encode() for x in APP_D] + fuzz.fuzz_statements(CORPUS, 400, seed=9) if len(s)]
    text, offs = P.make_batch(stmts)
    got = ctx.parse_batch(text, offs, flags=gpu.F_ALL_TOKENS)
    for i, s in enumerate(stmts):
        want, err = O.tokenize(s)
        b = int(got.stmt["tok_begin"][i])
        e = b + int(got.stmt["tok_count"][i])
        g = list(zip(got.tok_type[b:e].tolist(), got.tok_start[b:e].tolist(), got.tok_end[b:e].tolist()))
        assert g[:len(want)] == want, s
        if err is not None:
            assert g[len(want)] == (40, err["pos"], err["site"]), s


def test_full_token_stream_on_the_lexer_stress_set(ctx):
    """NUTDB_F_ALL_TOKENS over 4 MiB of config 3 (strings, quoted identifiers, comments, 5 % malformed): every token
    of every statement incl. Whitespace / Comment, and the poison token at the reference's error position / site."""
    from nutdb_b200 import gpu
    text, offs = W.generate(3, 4 << 20, seed=21)
    got = ctx.parse_batch(text, offs, flags=gpu.F_ALL_TOKENS)
    n_err = 0
    for i in range(len(offs) - 1):
        s = bytes(text[int(offs[i]):int(offs[i + 1])])
        ty, st, en, err = O.tokenize_arrays(s)
        b = int(got.stmt["tok_begin"][i])
        k = len(ty)
        assert np.array_equal(got.tok_type[b:b + k], ty) and np.array_equal(got.tok_start[b:b + k], st) and \
            np.array_equal(got.tok_end[b:b + k], en), s
        if err is not None:
            n_err += 1
            assert (int(got.tok_type[b + k]), int(got.tok_start[b + k]), int(got.tok_end[b + k])) == (40, err["pos"], err["site"]), s
        else:
            assert int(got.stmt["tok_count"][i]) >= k, s
    assert n_err > 100


def test_deep_nesting_takes_the_retry_path(ctx):
    d = 256
    stmts = ["select " + "(" * d + "1" + ")" * d, "select " + "(select " * d + "1" + ")" * d,
             "select " + "[" * d + "1" + "]" * d + " from t", "select " + "f(" * d + "x" + ")" * d,
             "select " + "not " * 300 + "x", "select " + "CASE WHEN a THEN " * 100 + "1" + " END" * 100]
    got = check(ctx, stmts)
    assert (got.stmt["status"] == 0).all()


def test_device_input_and_no_host_copy(ctx):
    import torch
    from nutdb_b200 import gpu
    text, offs = W.generate(2, 1 << 20)
    dt = torch.from_numpy(text).cuda()
    do = torch.from_numpy(offs.astype(np.int64)).cuda()
    torch.cuda.synchronize()
    got = ctx.parse_batch_raw(dt.data_ptr(), do.data_ptr(), len(offs) - 1, gpu.F_DEVICE_INPUT)
    bad = P.compare_with_oracle(got, text, offs)
    assert not bad, "\n".join(bad)
    # misaligned device text pointer
    dt2 = torch.zeros(len(text) + 3, dtype=torch.uint8, device="cuda")
    dt2[3:] = dt
    torch.cuda.synchronize()
    got2 = ctx.parse_batch_raw(dt2.data_ptr() + 3, do.data_ptr(), len(offs) - 1, gpu.F_DEVICE_INPUT)
    assert np.array_equal(got2.node, got.node) and np.array_equal(got2.stmt, got.stmt)
    got3 = ctx.parse_batch_raw(dt.data_ptr(), do.data_ptr(), len(offs) - 1, gpu.F_DEVICE_INPUT | gpu.F_NO_HOST_COPY)
    assert got3.n_node == got.n_node and len(got3.node) == 0
    ptrs = got3.device_pointers()
    assert ptrs["node"] and ptrs["stmt"]


def test_bad_arguments(ctx):
    from nutdb_b200 import gpu
    with pytest.raises(gpu.NutdbGpuError):
        ctx.parse_batch(b"select 1", np.array([0, 8, 4], np.uint64))
    # an interior offset far beyond the batch: ascending pairs, but outside [off[0], off[n]] (must be NUTDB_E_ARG,
    # not an out-of-bounds write into the statement-start bitmap)
    with pytest.raises(gpu.NutdbGpuError):
        ctx.parse_batch(b"select 1, 2", np.array([0, 1000000000, 1000000001, 10], np.uint64))
    check(ctx, ["select 1", "select 2"])   # the context is still usable


def test_parser_parse_mirrors_the_reference_entry(ctx):
    # tests/parser_test.rs:3-34: every corpus file parses Ok; plus the Display / Debug texts
    from nutdb_b200.parser import Parser, ParseError
    for s in CORPUS:
        st = Parser.parse(s)
        assert repr(st) == O.parse(s).debug
    with pytest.raises(ParseError) as ei:
        Parser.parse("SELECT 1d")
    assert str(ei.value) == O.parse("SELECT 1d").error == \
        "Lex Error: Unexpected Char: 'd' cannot be a part of integer literal near line 1 col 9"
    with pytest.raises(ParseError) as ei:
        Parser.parse("")
    assert str(ei.value) == "Syntax Error: empty query"
    res = Parser.parse_many(fuzz.EXTRA_SEEDS + [b"select $0", b"select 'abc"])
    for s, r in zip(fuzz.EXTRA_SEEDS + [b"select $0", b"select 'abc"], res):
        want = O.parse(s)
        assert (repr(r) == want.debug) if want.ok else (str(r) == want.error)


def test_stream_parser_chunks_match_one_batch(ctx):
    from nutdb_b200 import gpu, stream
    text, offs = W.generate(3, 6 << 20)
    whole = ctx.parse_batch(text, offs, flags=gpu.F_NO_TOKENS)
    sp = stream.StreamParser(0, workers=3)
    got = {}

    def on_batch(first, b):
        got[first] = (b.stmt.copy(), b.node.copy(), b.err.copy())

    n = sp.parse(text, offs, on_batch, chunk_bytes=1 << 20)
    sp.close()
    assert n >= 5 and len(got) == n
    nodes = np.concatenate([got[k][1] for k in sorted(got)])
    status = np.concatenate([got[k][0]["status"] for k in sorted(got)])
    ncount = np.concatenate([got[k][0]["node_count"] for k in sorted(got)])
    assert np.array_equal(nodes, whole.node)
    assert np.array_equal(status, whole.stmt["status"]) and np.array_equal(ncount, whole.stmt["node_count"])
    assert sum(len(got[k][2]) for k in got) == whole.n_err


def check_split(ctx, buf):
    """Splitter output vs the offsets the reference tokenizer defines (oracle_lib.split_statements): identical when
    the buffer lexes cleanly, identical up to the tokenizer's first error otherwise."""
    got = ctx.split_statements(buf)
    want, upto = O.split_statements(buf)
    if upto > len(buf):
        assert np.array_equal(got, want), bytes(buf[:200])
    else:   # the reference tokenizer stops at its first error: the offsets in front of it are pinned
        k = int(np.searchsorted(got, upto, side="right"))
        kw = int(np.searchsorted(want, upto, side="right"))
        assert np.array_equal(got[:k], want[:kw]), bytes(buf[:200])
    return got


def test_statement_splitter(ctx):
    import emul_lib as E
    # a query log as one buffer: the generators end every statement with ";\n"
    for config in (2, 3, 4):
        text, offs = W.generate(config, 3 << 20)
        buf = text[:int(offs[-1])]
        got = check_split(ctx, buf)
        assert np.array_equal(got, E.split(buf))   # (sequential run of the kernels' own context automaton)
        if config != 3:   # valid statements: one ';' each, at the end (malformed ones of config 3 may hide or add some)
            assert len(got) == len(offs) and np.array_equal(got[1:], offs[1:] - 1)
        # and the split batch parses like the original one
        b1 = ctx.parse_batch(buf, got)
        b0 = ctx.parse_batch(text, offs)
        if config != 3:
            assert np.array_equal(b1.stmt["status"], b0.stmt["status"])
            assert np.array_equal(b1.node["kind"], b0.node["kind"])
    # config 3 without its malformed statements lexes cleanly end to end: the whole log is pinned by the tokenizer
    text, offs = W.generate(3, 3 << 20, seed=11)
    st = O.parse_batch(text, offs).stmt["status"]
    keep = [bytes(text[int(offs[i]):int(offs[i + 1])]) for i in np.nonzero(st == 0)[0]]
    buf = np.frombuffer(b"".join(keep), np.uint8)
    want, upto = O.split_statements(buf)
    assert upto > len(buf) and len(want) > 1000
    assert np.array_equal(ctx.split_statements(buf), want)
    cases = [b"", b";", b" ;; ", b"select 1", b"select ';' , \";\" , `;` ; select 2 -- ; x\n; /* ; */ select 3;  \n",
             b"select 'it''s; ok'; select 'a\\'; b'; select 4", b"select 1 /* open ; comment", b"select 'open ; string",
             b"a;" * 5000, b"-- only a comment; really\n", b"select 1;\n\n\t "]
    for c in cases + fuzz.fuzz_statements([b"; ".join(CORPUS[:4])], 200, seed=5, max_mut=6):
        check_split(ctx, np.frombuffer(c, np.uint8))
        assert np.array_equal(ctx.split_statements(c), E.split(c)), c


def test_device_hash_matches_host_arrays_and_is_shard_invariant(ctx):
    """nutdb_gpu_batch_hash: equal to the same sum computed from the host copies; the sum of the chunk checksums of a
    log does not depend on how the chunks are dealt to contexts (config 5's cross-GPU-count check)."""
    from nutdb_b200 import gpu
    for config in (2, 3):
        text, offs = W.generate(config, 1 << 20)
        got = ctx.parse_batch(text, offs)
        h = got.device_hash()
        assert h == gpu.host_hash(got)
        # outputs are reproducible bit for bit (config 3 sends a third of its statements through the exact lexer,
        # whose region of the token arrays is laid out in statement order, not in discovery order)
        again = ctx.parse_batch(text, offs)
        assert again.device_hash() == h and np.array_equal(again.stmt, got.stmt)
    # ... and do not depend on what the context parsed before (no stale bytes in any output array)
    from nutdb_b200 import gpu as G
    text, offs = W.generate(3, 1 << 20, seed=5)
    fresh = G.Context(0)
    try:
        h_fresh = fresh.parse_batch(text, offs).device_hash()
    finally:
        fresh.close()
    ctx.parse_batch(*W.generate(4, 1 << 20, seed=6))
    ctx.parse_batch(*W.generate(3, 2 << 20, seed=7))
    assert ctx.parse_batch(text, offs).device_hash() == h_fresh
    # the same two chunks on one context in either order, and a changed byte changes the checksum
    chunks = [W.generate(2, 256 << 10, seed=0x5EED0005 + c) for c in range(2)]
    h = []
    for order in ((0, 1), (1, 0)):
        tot = 0
        for c in order:
            tot = (tot + ctx.parse_batch(*chunks[c]).device_hash()) & 0xFFFFFFFFFFFFFFFF
        h.append(tot)
    assert h[0] == h[1]
    t2 = chunks[0][0].copy()
    t2[10] = ord("x") if t2[10] != ord("x") else ord("y")
    assert ctx.parse_batch(t2, chunks[0][1]).device_hash() != ctx.parse_batch(*chunks[0]).device_hash()


def test_long_batch_uses_the_two_pass_scans(ctx):
    """Batches beyond 16 MiB switch the per-tile scans to their two-pass grid-wide form: same results as the oracle."""
    text, offs = W.generate(3, 20 << 20, seed=77)
    got = ctx.parse_batch(text, offs)
    bad = P.compare_with_oracle(got, text, offs)
    assert not bad, "\n".join(bad[:5])


def test_predicates_and_hex_literals(ctx):
    """Constructs added to the native paths late in round 1: IS [NOT] NULL, [NOT] IN/LIKE/ILIKE/BETWEEN, prefix NOT
    (table-driven parser) and plain hex literals (mask lexer), unmutated and mutated."""
    import test_emul_parity as T
    hexs = [b"select 0x1F, 0X2a, 0x, 0x0 from t where a = 0xdeadBEEF limit 0x10",
            b"select 0x1G, 0xzz, 00x1, 0x1.5, .0x1, 1.0x2, 0x1_2, x0x1, 0x1x", b"select a from t limit 0x, 0xFFFFFFFFFFFFFFFFF"]
    seeds = T.PREDICATES + T.PREDICATES_AUTOMATON + T.JOINS + T.JOINS_AUTOMATON + T.CASES + T.CASES_AUTOMATON + T.QUALIFIED + T.QUALIFIED_AUTOMATON + hexs
    check(ctx, seeds)
    for seed in (81, 82):
        check(ctx, fuzz.fuzz_statements(seeds, 20000, seed=seed, max_mut=3))


def _collect(devices, text, offs, flags, chunk_bytes, workers=2):
    """Runs the C dispatcher and returns its chunks' arrays re-assembled in statement order."""
    from nutdb_b200 import gpu
    m = gpu.MultiContext(devices, workers)
    got = {}

    def on_chunk(c):
        if c.on_device:   # gathered into device-0 memory: read it back through the library's CUDA runtime
            r = c.raw
            got[c.first_stmt] = dict(stmt=gpu.copy_to_host(r.stmt, r.n_stmt, gpu.STMT_DT),
                                     pnode=gpu.copy_to_host(r.pnode, r.n_node, gpu.PNODE_DT),
                                     err=gpu.copy_to_host(r.err, r.n_err, gpu.ERR_DT), device=c.device, node=None, tok=None)
        else:
            b = c.batch
            got[c.first_stmt] = dict(stmt=b.stmt.copy(), pnode=b.pnode.copy(), err=b.err.copy(), device=c.device,
                                     node=b.node.copy(),
                                     tok=None if flags & gpu.F_NO_TOKENS else
                                     (b.tok_type.copy(), b.tok_start.copy(), b.tok_end.copy(), b.tok_kw.copy()))
    try:
        m.parse_stream(text, offs, on_chunk, chunk_bytes, flags)
    finally:
        m.close()
    return got


def _check_dispatcher(ctx, devices, gather_flags):
    from nutdb_b200 import gpu
    text, offs = W.generate(3, 6 << 20, seed=31)
    whole = ctx.parse_batch(text, offs)
    assert not P.compare_with_oracle(whole, text, offs)
    got = _collect(devices, text, offs, gather_flags, 1 << 20)
    keys = sorted(got)
    assert len(keys) >= 5 and keys[0] == 0
    assert len({got[k]["device"] for k in keys}) == len(set(devices))
    # chunk-local records + first_stmt == the one-batch result
    stmt = np.concatenate([got[k]["stmt"] for k in keys])
    assert len(stmt) == whole.n_stmt
    for f in ("status", "tok_count", "node_count", "tok_used"):
        assert np.array_equal(stmt[f], whole.stmt[f]), f
    assert np.array_equal(np.concatenate([got[k]["pnode"] for k in keys]), whole.pnode)
    err = np.concatenate([got[k]["err"] for k in keys])
    err_stmt = np.concatenate([got[k]["err"]["stmt"].astype(np.int64) + k for k in keys])
    assert np.array_equal(err_stmt, whole.err["stmt"].astype(np.int64))
    for f in ("cls", "code", "line", "col", "pos", "a", "b", "c"):
        assert np.array_equal(err[f], whole.err[f]), f
    if got[keys[0]]["node"] is not None:
        assert np.array_equal(np.concatenate([got[k]["node"] for k in keys]), whole.node)
    if got[keys[0]]["tok"] is not None:   # token arrays: each chunk's pulled tokens against the one-batch arrays
        for k in keys:
            s, (ty, st, en, kw) = got[k]["stmt"], got[k]["tok"]
            for j in (0, len(s) // 2, len(s) - 1):
                a, n = int(s["tok_begin"][j]), int(s["tok_count"][j])
                wa = int(whole.stmt["tok_begin"][k + j])
                assert np.array_equal(ty[a:a + n], whole.tok_type[wa:wa + n]) and np.array_equal(st[a:a + n], whole.tok_start[wa:wa + n])
                assert np.array_equal(en[a:a + n], whole.tok_end[wa:wa + n]) and np.array_equal(kw[a:a + n], whole.tok_kw[wa:wa + n])


def test_dispatcher_two_pipelined_device_slots_on_one_gpu(ctx):
    """nutdb_gpu_mctx_*: the dispatcher's logic (cuts, per-device queues, gather slots, callbacks) with the one GPU
    listed twice -- host gather with and without tokens, and the gather into device memory."""
    from nutdb_b200 import gpu
    _check_dispatcher(ctx, (0, 0), gpu.F_NO_TOKENS)
    _check_dispatcher(ctx, (0, 0), 0)
    _check_dispatcher(ctx, (0,), gpu.F_NO_TOKENS | gpu.MF_GATHER_DEVICE0 | gpu.MF_SERIAL_CALLBACKS)


def test_dispatcher_two_gpus(ctx):
    """Two GPUs of one box driven by one process: host gather over each GPU's PCIe link, and the gather into GPU 0's
    memory over NVLink (cudaMemcpyPeerAsync); both against the one-GPU result (itself checked against the oracle)."""
    import torch
    from nutdb_b200 import gpu
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _check_dispatcher(ctx, (0, 1), gpu.F_NO_TOKENS)
    _check_dispatcher(ctx, (0, 1), 0)
    _check_dispatcher(ctx, (0, 1), gpu.F_NO_TOKENS | gpu.MF_GATHER_DEVICE0)


def test_dispatcher_errors(ctx):
    from nutdb_b200 import gpu
    m = gpu.MultiContext((0,), 1)
    try:
        with pytest.raises(gpu.NutdbGpuError):
            m.parse_stream(b"select 1", np.array([0, 8, 4], np.uint64), lambda c: None)
        seen = []
        m.parse_stream(b"select 1;select 2", np.array([0, 9, 17], np.uint64), lambda c: seen.append(int(c.batch.n_stmt)))
        assert sum(seen) == 2
        with pytest.raises(ZeroDivisionError):   # an exception in the consumer surfaces in the caller
            m.parse_stream(b"select 1", np.array([0, 8], np.uint64), lambda c: 1 // 0)
    finally:
        m.close()


def test_wide_table_driven_pass(ctx):
    """The second table-driven pass (k_parse_wide): arrays, maps, index access, prefix ~, IF, parenthesised subqueries,
    nesting to any depth (operator stack at the top of the statement's node range) and constant folding -- same
    records as the oracle, and the automaton only sees what both passes declined."""
    import test_emul_parity as T
    got = check(ctx, T.WIDE + T.DEEP)
    assert (got.stmt["status"] == 0).all()
    assert ctx.slow_statements() == 0 and ctx.wide_statements() >= len(T.WIDE) + len(T.DEEP) - 3
    check(ctx, T.WIDE_AUTOMATON)
    check(ctx, T.fold_statements())
    assert ctx.wide_statements() > 1000
    for seed in (93, 94):
        check(ctx, fuzz.fuzz_statements(T.WIDE + T.WIDE_AUTOMATON + [d[:400] for d in T.DEEP], 20000, seed=seed, max_mut=3))
    text, offs = W.generate(4, 8 << 20)
    got = ctx.parse_batch(text, offs)
    assert ctx.slow_statements() < 0.01 * (len(offs) - 1)
    bad = P.compare_with_oracle(got, text, offs)
    assert not bad, "\n".join(bad)


def test_wire_statement_records(ctx):
    """NUTDB_F_WIRE_STMT: 8-byte statement records on the wire; expanded on the host they equal the 24-byte records
    (token fields aside), directly and through the dispatcher."""
    from nutdb_b200 import gpu
    text, offs = W.generate(3, 2 << 20, seed=41)
    plain = ctx.parse_batch(text, offs, flags=gpu.F_NO_TOKENS)
    wire = ctx.parse_batch(text, offs, flags=gpu.F_NO_TOKENS | gpu.F_WIRE_STMT)
    assert wire.wstmt is not None and len(wire.wstmt) == plain.n_stmt
    for f in ("status", "node_begin", "node_count", "tok_used"):
        assert np.array_equal(wire.stmt[f], plain.stmt[f]), f
    assert np.array_equal(wire.node, plain.node) and np.array_equal(wire.err, plain.err)
    # NUTDB_F_OFFSETS32: the same batch described by 32-bit offsets, directly and through the dispatcher
    o32 = ctx.parse_batch(text, offs.astype(np.uint32), flags=gpu.F_NO_TOKENS | gpu.F_OFFSETS32)
    assert np.array_equal(o32.stmt, plain.stmt) and np.array_equal(o32.pnode, plain.pnode) and np.array_equal(o32.err, plain.err)
    got32 = _collect((0, 0), text, offs.astype(np.uint32), gpu.F_NO_TOKENS | gpu.F_WIRE_STMT | gpu.F_OFFSETS32, 512 << 10)
    k32 = sorted(got32)
    assert np.array_equal(np.concatenate([got32[k]["pnode"] for k in k32]), plain.pnode)
    assert np.array_equal(np.concatenate([got32[k]["stmt"]["status"] for k in k32]), plain.stmt["status"])
    with pytest.raises(gpu.NutdbGpuError):   # (descending 32-bit offsets are refused like 64-bit ones)
        ctx.parse_batch(b"select 1", np.array([0, 8, 4], np.uint32), flags=gpu.F_OFFSETS32)
    # without NUTDB_F_NO_TOKENS the flag is ignored (the token fields are needed)
    assert ctx.parse_batch(text, offs, flags=gpu.F_WIRE_STMT).wstmt is None
    got = _collect((0, 0), text, offs, gpu.F_NO_TOKENS | gpu.F_WIRE_STMT, 512 << 10)
    keys = sorted(got)
    assert np.array_equal(np.concatenate([got[k]["stmt"]["status"] for k in keys]), plain.stmt["status"])
    assert np.array_equal(np.concatenate([got[k]["stmt"]["node_count"] for k in keys]), plain.stmt["node_count"])
    assert np.array_equal(np.concatenate([got[k]["node"] for k in keys]), plain.node)


def test_dispatcher_shards_with_device_resident_text(ctx):
    """nutdb_gpu_mctx_parse_shards: statement ranges whose text already sits in the GPU's memory (config 5's path),
    gathered to pinned host memory in the 8-byte / 4-byte wire forms; against the one-batch result."""
    import torch
    from nutdb_b200 import gpu, dispatch
    text, offs = W.generate(3, 4 << 20, seed=51)
    whole = ctx.parse_batch(text, offs, flags=gpu.F_NO_TOKENS)
    parts = [r for r in dispatch.split_statements(offs, 5) if r[1] > r[0]]
    keep, shards = [], []
    for lo, hi in parts:
        o = offs[lo:hi + 1] - offs[lo]
        t = np.concatenate([text[int(offs[lo]):int(offs[hi])], np.zeros(3, np.uint8)])   # (ragged end: no padding promised)
        dt, do = torch.from_numpy(t).cuda(), torch.from_numpy(o.astype(np.int64)).cuda()
        keep += [dt, do]
        shards.append((0, dt.data_ptr(), do.data_ptr(), hi - lo, gpu.F_DEVICE_INPUT, lo))
    torch.cuda.synchronize()
    m = gpu.MultiContext((0,), 2)
    got = {}
    try:
        m.parse_shards(shards, lambda c: got.__setitem__(c.first_stmt, (c.shard, c.batch.stmt.copy(), c.batch.pnode.copy(), c.batch.err.copy())),
                       gpu.F_NO_TOKENS | gpu.F_WIRE_STMT)
    finally:
        m.close()
    keys = sorted(got)
    assert keys == [lo for lo, _ in parts] and [got[k][0] for k in keys] == list(range(len(parts)))
    assert np.array_equal(np.concatenate([got[k][1]["status"] for k in keys]), whole.stmt["status"])
    assert np.array_equal(np.concatenate([got[k][1]["tok_used"] for k in keys]), whole.stmt["tok_used"])
    assert np.array_equal(np.concatenate([got[k][2] for k in keys]), whole.pnode)
    assert sum(len(got[k][3]) for k in keys) == whole.n_err


def test_lookback_lexer_fallback(ctx):
    """k_lex3 -- one token stream, chained look-back scans between tiles -- takes over when a statement is too long to cut
    the batch around it.  Forced on ordinary workloads, and reached on its own by a batch with a giant statement."""
    from nutdb_b200 import gpu
    c2 = gpu.Context(0)
    try:
        c2.force_lookback(True)
        for config in (2, 3, 4):
            text, offs = W.generate(config, 2 << 20, seed=61 + config)
            got = c2.parse_batch(text, offs)
            assert c2.last_lookback()
            bad = P.compare_with_oracle(got, text, offs)
            assert not bad, "\n".join(bad)
        check(c2, fuzz.fuzz_statements(CORPUS + fuzz.EXTRA_SEEDS, 8000, seed=65, max_mut=4))
    finally:
        c2.close()
    # a 40 MiB statement (one string literal) between ordinary ones: the range lexer gives up, the look-back lexer runs
    text, offs = W.generate(2, 1 << 20, seed=66)
    stmts = [bytes(text[int(offs[i]):int(offs[i + 1])]) for i in range(len(offs) - 1)]
    giant = b"select '" + b"x" * (40 << 20) + b"', 1 from t"
    batch = stmts[:2000] + [giant] + stmts[2000:4000]
    t2, o2 = P.make_batch(batch)
    got = ctx.parse_batch(t2, o2)
    assert ctx.last_lookback()
    bad = P.compare_with_oracle(got, t2, o2)
    assert not bad, "\n".join(bad)
    got = ctx.parse_batch(text, offs)      # and the context goes back to the range lexer afterwards
    assert not ctx.last_lookback()
    assert not P.compare_with_oracle(got, text, offs)


def test_deep_malformed_statements_reach_the_automaton_and_its_retry_pass(ctx):
    """Deep nests the table-driven passes decline (errors, constructs outside their grammar) still get the reference's
    first error and position from the automaton -- through its global-memory retry pass where its local stack overflows."""
    d = 300
    stmts = ["select " + "(" * d + "1" + ")" * (d - 1), "select " + "(" * d + "1" + ")" * (d + 1),
             "select " + "(select " * 200 + "1d" + ")" * 200, "select " + "[" * d + "1" + "]" * (d - 1) + " from t",
             "select " + "f(" * d + "x" + ")" * d + " from", "select " + "not " * 300 + "1d",
             "select " + "CASE WHEN a THEN " * 150 + "1" + " END" * 149, "select " + "{1:" * 200 + "x" + "}" * 199,
             "select " + "a[" * 250 + "x" + "]" * 249, "select " + "(" * d + "1 union all select 2" + ")" * d,
             "select " + "(" * d + "$1" + ")" * d, "select " + "(" * 40 + "1" + ")" * 40]
    got = check(ctx, stmts)
    assert ctx.slow_statements() >= 8
    assert int((got.stmt["status"] != 0).sum()) >= 8


def test_batch_near_the_two_gib_limit(ctx):
    """One batch of 1.9 GiB (the ABI's limit is 2^31 bytes): the statement records and the wire nodes equal those of
    the same statements parsed in four separate batches, a sample equals the oracle's, and one byte more than the limit
    is refused."""
    from nutdb_b200 import gpu
    text, offs = W.generate(2, 1900 << 20, seed=71)
    n = len(offs) - 1
    whole = ctx.parse_batch(text, offs, flags=gpu.F_NO_TOKENS | gpu.F_WIRE_STMT, copy=False)
    w_stmt, w_node, n_err = whole.wstmt.copy(), whole.pnode.copy(), whole.n_err
    assert n_err == 0 and len(w_stmt) == n
    cuts = [0, n // 4, n // 2, 3 * n // 4, n]
    parts_s, parts_n = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        part = ctx.parse_batch(text, offs[a:b + 1], flags=gpu.F_NO_TOKENS | gpu.F_WIRE_STMT, copy=False)
        parts_s.append(part.wstmt.copy())
        parts_n.append(part.pnode.copy())
    assert np.array_equal(np.concatenate(parts_s), w_stmt)
    assert np.array_equal(np.concatenate(parts_n), w_node)
    k = int(np.searchsorted(offs, np.uint64(4 << 20)))
    tail = offs[n - k:] - offs[n - k]                       # the LAST statements of the batch against the oracle
    t_tail = np.concatenate([text[int(offs[n - k]):int(offs[n])], np.zeros(16, np.uint8)])
    got = ctx.parse_batch(t_tail, tail)
    assert not P.compare_with_oracle(got, t_tail, tail)
    first_nodes = int((w_stmt[:n - k] >> np.uint64(4) & np.uint64(0x3FFFFFFF)).sum())
    assert np.array_equal(w_node[first_nodes:], got.pnode)
    with pytest.raises(gpu.NutdbGpuError):
        ctx.parse_batch(text[:16], np.array([0, 1 << 31], np.uint64))


def test_escaped_literal_side_byte(ctx):
    """tok_kw of an escaped literal: 1 iff it holds no backslash-u escape (then the parsers skip its validation) --
    the range lexer, the look-back lexer and the exact walker agree, and statuses / errors equal the oracle's."""
    from nutdb_b200 import gpu
    check(ctx, fuzz.escaped_literal_statements(30000, seed=12, unicode=True))
    stmts = fuzz.escaped_literal_statements(60000, seed=11)
    got = check(ctx, stmts)
    assert (got.stmt["status"] != 0).any() and (got.stmt["status"] == 0).any()
    text, offs = P.make_batch(stmts)
    c2 = gpu.Context(0)
    try:
        c2.force_lookback(True)
        got2 = c2.parse_batch(text, offs)
        assert c2.last_lookback()
        assert not P.compare_with_oracle(got2, text, offs)
        assert np.array_equal(got2.tok_kw, got.tok_kw)
    finally:
        c2.close()


def test_wide_pass_group_scheduling_edges(ctx):
    """The wide pass's queue of statement groups (k_wide_order / k_parse_wide): pools that end exactly on, one short of and
    one past the 256-statement boundary, a single statement, statements far over the token budget of a group (alone in
    theirs), many equal statements (one shape: a full group whatever its tokens) -- all against the oracle."""
    def nest(depth, width):
        return "select " + "(" * depth + " + ".join(f"a{i} * (b{i} - {i})" for i in range(width)) + ")" * depth + " from t where x[1] = {'k': [1, 2]}"
    giant = nest(40, 3000)        # ~27 K tokens: over the budget on its own
    big = nest(200, 300)
    mid = [nest(3 + i % 7, 12 + i % 5) for i in range(40)]
    same = "select a[1], ~b, if c then 1 else 2 end from t where (x + 1) * (y - 2) in [1, 2, 3] and z = {'k': 1}"
    for n in (1, 255, 256, 257, 513):
        stmts = [mid[i % len(mid)] for i in range(n)]
        got = check(ctx, stmts)
        assert ctx.wide_statements() == n
    stmts = [giant, big] + [same] * 300 + mid + [big, giant] + [same] * 211
    got = check(ctx, stmts)
    assert (got.stmt["status"] == 0).all()
    assert ctx.wide_statements() == len(stmts) and ctx.slow_statements() == 0
    assert int(got.stmt["tok_count"].max()) > 20000
