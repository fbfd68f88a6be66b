"""Pins the oracle parser against what the reference's own tests pin (tests/parser_test.rs:19-34: the 14
fixture files parse Ok; benches/parser_bench.rs:5-48: both bench strings parse) and against the
hand-derived vectors of SURVEY.md App. D (derived by reading src/parser/mod.rs, never executed)."""
import os

import pytest

import oracle_lib as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def corpus():
    return [open(os.path.join(GOLDEN, "sql", f"{i}.sql"), encoding="utf-8").read() for i in range(1, 15)]


@pytest.mark.parametrize("i", range(1, 15))
def test_parse_sql_file(i):  # tests/parser_test.rs:19-34
    r = O.parse(corpus()[i - 1])
    assert r.ok, r.error
    assert len(r.nodes) > 0 and r.nodes[-1]["parent"] == 0xFFFFFFFF


def test_bench_strings_parse():  # benches/parser_bench.rs:5,8-47
    short = "SELECT * FROM table WHERE 1 = 1"
    r = O.parse(short)
    assert r.ok
    assert r.debug == (
        'Select(SelectStmt { query: Single(QueryBody { with: None, distinct: None, columns: [QueryExpr { inner: '
        'Identifier(Identifier { name: Wildcard, qualifier: None }), alias: None }], from: Some(FromClause { source: '
        'QuerySource { inner: Table("table"), alias: None } }), joins: [], where: Some(WhereClause { condition: '
        'Literal(Boolean(true)) }), group_by: None, having: None, order_by: None, limit: None }) })')
    long_sql = open(os.path.join(GOLDEN, "bench_long.sql"), encoding="utf-8").read()
    assert len(long_sql.encode()) == 1131
    assert O.parse(long_sql).ok


def test_app_d_vectors():
    r = O.parse(corpus()[12])
    assert r.debug == (
        'Insert(InsertStmt { table_name: "test", column_list: Some(["c1", "c2", "c3"]), data: Rows { column_size: 3, '
        'data: [Literal(Integer(1, true)), Literal(Integer(2, true)), Literal(Integer(3, true)), '
        'Literal(String("1")), Literal(String("2")), Literal(String("3"))] } })')
    d = O.parse(corpus()[13]).debug
    cols = [seg.split(", alias")[0] for seg in d.split("columns: [QueryExpr { inner: ")[1:]]
    assert cols[0] == "Literal(Boolean(true))" and cols[1] == "Literal(Boolean(false))"
    assert cols[2] == "Literal(Boolean(true))" and cols[3] == "Literal(Boolean(false))"
    assert cols[4] == "Literal(Boolean(true))"
    assert cols[5].startswith("UnaryOp(UnaryOp { op: IsNull, operand: Identifier(")
    assert cols[6] == "Literal(Boolean(false))"
    assert cols[7].startswith('UnaryOp(UnaryOp { op: Not, operand: FnCall(FnCall { callee: Others("random"), arguments: [] })')
    assert cols[8] == "Literal(Boolean(false))"
    assert d.count("Union { typ: UnionAll") == 8
    for q in ["", "  ", ";", "-- c"]:
        assert O.parse(q).error == "Syntax Error: empty query"
    assert O.parse("SELECT 1d").error == ("Lex Error: Unexpected Char: 'd' cannot be a part of integer literal near "
                                          "line 1 col 9")
    assert O.parse("SELECT a FROM t ORDER BY a ASC").error == ("Syntax Error: fail to parse (more than one statement) "
                                                               "at line 1 col 28")
    assert O.parse("SELECT $0").error == ("Syntax Error: expected token (IntegerLiteral, HexLiteral) but found token "
                                          "EOF at line 1 col 10")
    assert O.parse("SELECT 'abc").error == "Lex Error: Unexpected EOF: string literal is not complete near line 1 col 12"
    assert O.parse("select 1; 'oops").ok


def test_quirks():  # SURVEY.md App. B quirks, derived from the cited lines of src/parser/mod.rs
    assert "QueryParameter(QueryParameter { index: 5 })" in O.parse("SELECT $0 5").debug  # :1311
    assert "Compound(Map(Scalar(Int8), Scalar(Date)))" in O.parse("CREATE TABLE t (m Map(Date, Int8))").debug  # :1780
    assert 'Table("t")' in O.parse("SELECT 1 FROM db.t").debug  # :552-555
    assert not O.parse("SELECT -x").ok and not O.parse("SELECT - -1").ok  # :1259-1269
    d = O.parse("SELECT NOT a = b").debug  # :1294-1296: NOT binds to the bare prefix
    assert d.count("op: Eq, left: UnaryOp(UnaryOp { op: Not") == 1
    assert 'Others("exists")' in O.parse("SELECT exists(1)").debug  # quirk 7
    assert "LimitClause { size: 5, offset: 2, with_ties: false }" in O.parse("SELECT 1 LIMIT 2, 5").debug  # :513-543
    assert "LimitClause { size: 2, offset: 5, with_ties: true }" in O.parse("SELECT 1 LIMIT 2 OFFSET 5 WITH TIES").debug
    assert not O.parse("ALTER TABLE t ADD COLUMN c Int8 FIRST").ok  # quirk 9
    assert O.parse("ALTER TABLE t ADD INDEX i f(x) FIRST").ok
    e = O.parse("CREATE VIEW v AS SELECT 1").error  # quirk 10
    assert e == "Syntax Error: expected keyword (update) but found token as at line 1 col 15"
    e = O.parse("INSERT INTO t VALUES (1, 2), (3)").error  # quirk 11
    assert e == "Syntax Error: (row has 1 column(s)) conflicts with (previous rows have 2 column(s)) near line 1 col 32"
    assert "Literal(Boolean(true))" in O.parse("SELECT 16 = 0x10").debug  # quirk 13
    assert "Literal(Boolean(false))" in O.parse("SELECT 1 = 1.0").debug
    assert "Literal(Boolean(true))" in O.parse("SELECT 1.0 = 1.00").debug
    assert "Literal(Boolean(true))" in O.parse("SELECT 'a''b' = \"a'b\"").debug
    assert "Literal(Boolean(false))" in O.parse("SELECT -0 = 0").debug
    assert 'Float(BigDecimal("0.0001000000"))' in O.parse("SELECT 0.0001000000").debug
    assert "EnumBind { id: 4, literal: \"b\" }" in O.parse("CREATE TABLE t (e Enum('a' = 3, 'b'))").debug  # quirk 16
    assert O.parse("SELECT 340282366920938463463374607431768211456").error == (
        "Syntax Error: invalid integer '340282366920938463463374607431768211456'")
    assert O.parse("SELECT 1 LIMIT 0x").error == "Syntax Error: invalid hex '0x'"
    assert O.parse("SELECT '\\u{110000}'").error == "Syntax Error: invalid escaped unicode '\\u{110000}' in string literal"
    assert O.parse("SELECT 1 +").error.startswith("Syntax Error: expected token (RawStringLiteral, EscapedSingleQuoted")
    assert O.parse("foo").error == "Syntax Error: fail to parse (cannot recognize statement) at line 1 col 1"
    assert O.parse("1").error == "Syntax Error: fail to parse (statements should start with a keyword) at line 1 col 1"


def test_oracle_reproduces_the_committed_outputs_for_the_rust_differential_harness():
    """tests/golden/oracle_outputs.tsv holds `{:?}` / Display text of the oracle for ~1500 inputs (corpus, App. D,
    seeds, fold / predicate / join / case / wide lists, fuzz, synthetic samples).  tests/rust_diff (cargo; no toolchain
    here) compares the REAL reference against the same file; this test keeps the file and the oracle from drifting."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_oracle_outputs as M
    n = 0
    with open(os.path.join(here, "golden", "oracle_outputs.tsv"), encoding="utf-8", newline="\n") as f:
        for ln in f.read().split("\n"):
            if not ln:
                continue
            a, kind, b = ln.split("\t")
            r = O.parse(M.unesc(a).encode("utf-8"))
            assert (kind == "OK") == r.ok, a
            assert M.unesc(b) == (r.debug if r.ok else r.error), a
            n += 1
    assert n > 1000
