"""Mutation fuzzer shared by the CPU (host-emulated device code) and GPU parity tests."""
import numpy as np

ALPHABET = [b"'", b'"', b"`", b"\\", b"-", b"--", b"/*", b"*/", b"*", b"/", b"\n", b"\r", b"\r\n", b"\t", b" ", b";",
            b"(", b")", b"[", b"]", b"{", b"}", b",", b".", b":", b"!", b"=", b"<", b">", b"<>", b"<<", b">=", b"@",
            b"$", b"#", b"?", b"0", b"1", b"9", b"0x", b"1.", b".5", b"a", b"x", b"_", b"e", b"\xc3\xa9",
            b"\xe4\xbd\xa0", b"\xf0\x9f\x98\x80", b"~", b"+", b"%", b"&", b"|", b"^", b" not ", b" and ", b" or ",
            b" is ", b" null ", b" in ", b" between ", b" like ", b" select ", b" from ", b" as ", b" case ",
            b" when ", b" then ", b" else ", b" end ", b" if ", b" interval ", b" union ", b" all ", b" with ",
            b" desc ", b" asc ", b" limit ", b" join ", b" on ", b" using ", b" true ", b" false ", b"\\u{41}",
            b"\\u{D800}", b"\\u", b"''", b'""', b" $1 2 ", b" exists ", b" distinct ", b" order by ", b" group by "]


def mutate(stmt, rng, n_mut):
    s = bytearray(stmt)
    for _ in range(n_mut):
        k = rng.integers(0, 4)
        pos = int(rng.integers(0, len(s) + 1))
        if k == 0 and len(s) > 0:       # delete a short range
            ln = int(rng.integers(1, 6))
            del s[pos:pos + ln]
        elif k == 1:                    # insert
            s[pos:pos] = ALPHABET[int(rng.integers(0, len(ALPHABET)))]
        elif k == 2 and len(s) > 0:     # replace
            ln = int(rng.integers(1, 4))
            s[pos:pos + ln] = ALPHABET[int(rng.integers(0, len(ALPHABET)))]
        elif len(s) > 0:                # truncate
            if rng.integers(0, 4) == 0:
                del s[pos:]
    # keep the &str precondition of the reference: valid UTF-8 only
    try:
        bytes(s).decode("utf-8")
    except UnicodeDecodeError:
        s = bytearray(bytes(s).decode("utf-8", "ignore").encode("utf-8"))
    return bytes(s)


def fuzz_statements(seeds, count, seed=1234, max_mut=4):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        base = seeds[int(rng.integers(0, len(seeds)))]
        out.append(mutate(base, rng, int(rng.integers(1, max_mut + 1))))
    return out


EXTRA_SEEDS = [
    b"SELECT a, b AS c FROM db.t WHERE x = 1 AND y != 'abc' ORDER BY a DESC LIMIT 10, 20",
    b"SELECT 1 = 1, 0x10 = 16, 1.0 = 1.00, 'a''b' = \"a'b\", null is null, -1 = -1, 1 != 2",
    b"WITH a AS (SELECT 1) SELECT * FROM a UNION ALL SELECT 2 INTERSECT SELECT 3 EXCEPT SELECT 4",
    b"INSERT INTO t (a, b) VALUES (1, 'x'), (2, 'y')",
    b"INSERT INTO t FROM f(1, 2)",
    b"INSERT INTO t SELECT * FROM u",
    b"CREATE TABLE IF NOT EXISTS t (a Int8 DEFAULT 1 COMMENT 'c', b Array(Nullable(String)), INDEX i f(a), "
    b"CONSTRAINT c CHECK a > 0, e Enum('a' = 1, 'b'), m Map(String, Int8), d Decimal32(3), t Tuple(Int8, Chars(4))) "
    b"PRIMARY KEY a ORDER BY a, b PARTITION BY a COMMENT 'tbl'",
    b"CREATE VIEW v UPDATE BY s PRIMARY KEY a ORDER BY a PARTITION BY b COMMENT 'v' AS SELECT 1",
    b"ALTER TABLE t ADD IF NOT EXISTS INDEX i f(x) AFTER j",
    b"ALTER TABLE t ADD COLUMN c Int8",
    b"ALTER TABLE t DROP IF EXISTS PARTITION 'p'",
    b"ALTER TABLE t RENAME COLUMN a b",
    b"ALTER TABLE t RENAME TABLE u",
    b"DESCRIBE TABLE t", b"DESCRIBE DATABASE", b"DROP VIEW IF EXISTS v", b"TRUNCATE TABLE t",
    b"OPTIMIZE TABLE t ON PARTITION 'p'", b"SET @cfg = 1 + 2", b"EXPLAIN (SELECT 1)",
    b"SELECT CASE x WHEN 1 THEN 'a' WHEN 2 THEN 'b' ELSE 'c' END, CASE WHEN a THEN b END, IF a THEN b ELSE c END",
    b"SELECT a NOT IN (1, 2), b NOT LIKE 'x', c NOT BETWEEN 1 AND 2, d BETWEEN 1 AND 2 AND e, NOT EXISTS(SELECT 1), "
    b"x IS NOT NULL, y IS NULL, NOT true, ~a[1], -5, +x, INTERVAL 5 DAY, $1 2, [1, 2], {1: 2, 3: 4}, (1, 2), t.*, `q`.`r`",
    b"SELECT a FROM t LEFT SEMI JOIN u ON a = b FULL OUTER JOIN v USING (x, y) RIGHT ANTI JOIN w ON true JOIN z ON 1",
    b"SELECT DISTINCT ON (a, b) a, b FROM t GROUP BY a, b HAVING count(a) > 1 LIMIT 5 OFFSET 2 WITH TIES",
    b"SELECT 'a\\u{767D}b', 'x\\ty', \"q\\\"r\", 'it''s', '\\u{+41}', '\\ux'",
    b"select 1 -- trailing comment\n , 2 /* block ** comment **/ from t",
    b"SELECT true AND a, false AND a, a AND true, a OR false, true XOR a, a XOR false, 1 = 1 AND b, NOT (1 = 2)",
    b"select a + 1 not exists (select 1), b not exists (1, 2) and c, (a, b) not exists ()",
    b"SELECT 340282366920938463463374607431768211455, 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFF, 1e5",
]


# statements inside (and just outside) the subset the straight-line parser (parse_fast.cuh) accepts
SIMPLE_SEEDS = [
    b"SELECT a, b.c, t.*, f(x, y + 1), g(), (a + b) * c, (1, 2), x IN (1, 2, 3), -5, -1.5, +7, 'it''s', \"d\\\"q\", `q`.`r`, "
    b"0x1F AS h FROM t AS u WHERE a = 1 AND (b < 2 OR c LIKE 'x%') XOR d >= 3 GROUP BY a, b HAVING count(a) > 1 "
    b"ORDER BY a DESC, b LIMIT 5, 10 WITH TIES;",
    b"SELECT a FROM `tbl` WHERE x ILIKE 'y' ORDER BY f(a) DESC LIMIT 3 OFFSET 0x10",
    b"INSERT INTO t VALUES (1, 'a', -2.5, f(3)), (2, 'b', 0.5, g(4, 5))",
    b"INSERT INTO `t` (`a`, b) VALUES ((1 + 2) * 3, (4, 5))",
    b"CREATE TABLE IF NOT EXISTS t (a Int8 DEFAULT 1 + 2 COMMENT 'x', b Array(Nullable(String(10))), c Decimal32(3), "
    b"d Chars(4), e String, `f` Dictionary(UInt64)) PRIMARY KEY a, b ORDER BY f(a) PARTITION BY a % 4 COMMENT 'tbl'",
    b"SELECT 1 = 1, a = 1, true AND a, a OR false, 1 < 2 = 3 != 4, a | b ^ c & d << 2 >> 1 + 3 * 4 / 5 % 6 - 7",
    b"SELECT a b", b"SELECT a, FROM t", b"SELECT f(a, FROM t", b"SELECT (a, b FROM t", b"SELECT a FROM t WHERE",
    b"SELECT a FROM t LIMIT 99999999999999999999", b"SELECT a FROM t ORDER BY a ASC",
    b"select date, type, value, level from table where date = 'x'", b"select f (a) , g( b )from t",
]


def escaped_literal_statements(count, seed=7, unicode=False):
    """Statements built around escaped string literals: backslash-u escapes (valid, invalid, not an escape after all),
    doubled quotes, plain escapes, literals from a few bytes to several lexer windows long -- the cases that decide the
    lexer's side byte of an escaped literal and whether the parser has to validate it (literal.rs:45-102).
    unicode=True: the plain runs are random scalar values of every UTF-8 length (control characters included)."""
    import random
    rng = random.Random(seed)

    def lit():
        q = rng.choice("'\"")
        parts = []
        for _ in range(rng.choice([1, 2, 3, 5, 20, 60])):
            r = rng.random()
            if r < 0.15: parts.append("\\u{41}")
            elif r < 0.25: parts.append("\\\\u")
            elif r < 0.35: parts.append("\\" + rng.choice("ntu'\"x\\"))
            elif r < 0.45: parts.append(q + q)
            elif r < 0.50: parts.append("\\u{zz}")
            elif r < 0.55: parts.append("u")
            elif unicode: parts.append("".join(uchar(q) for _ in range(rng.randint(0, 12))))
            else: parts.append("".join(rng.choice("abcu \n;-/*") for _ in range(rng.randint(0, 40))))
        return q + "".join(parts) + q

    def uchar(q):  # any scalar value, 1-4 bytes of UTF-8; the neighbours of a backslash instead of an escape of its own
        while True:
            r = rng.random()
            cp = (rng.randint(1, 0x7F) if r < 0.5 else rng.randint(0x80, 0x7FF) if r < 0.7
                  else rng.randint(0x800, 0xFFFF) if r < 0.9 else rng.randint(0x10000, 0x10FFFF))
            if 0xD800 <= cp <= 0xDFFF:
                continue
            c = chr(cp)
            return rng.choice("[]") if c in ("\\", q) else c

    out = []
    for _ in range(count):
        k = rng.random()
        if k < 0.5:
            out.append("select " + ", ".join(lit() for _ in range(rng.randint(1, 4))) + " from t")
        elif k < 0.8:
            out.append("select a from t where b = " + lit() + " and c in (" + lit() + ")")
        else:
            out.append("create table t (a enum(" + ", ".join(lit() for _ in range(rng.randint(1, 3))) + "))")
    return out
