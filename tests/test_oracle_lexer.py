"""Pins the oracle tokenizer against the reference's OWN unit-test vectors.

Every test restates one #[test] of /root/reference/src/parser/tokenizer/mod.rs:576-782,
tokenizer/utf8_iter.rs:284-307 or literal.rs:126-151 (same inputs, same expectations).
"""
import oracle_lib as O

(KW, DELIM, CONFIG, QPARAM, RAW, ESQ, EDQ, INT, FLOAT, HEX, COMMA, DOT, COLON, SEMI, PLUS, MINUS, MUL, DIV, MOD,
 EQ, NOTEQ, LT, GT, LTEQ, GTEQ, LPAREN, RPAREN, LBRACKET, RBRACKET, LBRACE, RBRACE, BITAND, BITOR, BITXOR, BITNOT,
 BITLSHIFT, BITRSHIFT, COMMENT, WS, EOF) = range(40)


def first(sql):
    toks, err = O.tokenize(sql)
    return (toks[0] if toks else None), err


def collect(sql):
    """collect_tokens (mod.rs:551-560): until terminator or error."""
    toks, _ = O.tokenize(sql)
    out = []
    for t in toks:
        if t[0] in (EOF, SEMI):
            break
        out.append(t)
    return out


def text(sql, tok):
    return sql.encode()[tok[1]:tok[2]].decode()


def first_fails(sql):
    toks, err = O.tokenize(sql)
    return err is not None and len(toks) == 0


def test_tokenize_whitespaces():  # mod.rs:576
    case = " ".join(["    ", "\t\t", "\n", "\r\n", "\r"])
    toks = collect(case)
    assert len(toks) == 1 and toks[0][0] == WS


def test_tokenize_numerics():  # mod.rs:585
    for s, t in [("510", INT), ("0.123", FLOAT), (".123", FLOAT), ("1.", FLOAT), ("0x123", HEX)]:
        assert first(s)[0][0] == t


def test_tokenize_numerics_fail():  # mod.rs:600
    for s in ["1d", "1好", "1.d"]:
        assert first_fails(s)


def test_tokenize_strings():  # mod.rs:608
    cases = [('"hello"', "hello", RAW), ("'hello'", "hello", RAW), ("'he''llo'", "he''llo", ESQ),
             ('"he""llo"', 'he""llo', EDQ), ("'h\\t i\\r\\n'", "h\\t i\\r\\n", ESQ), ('"\\\n"', "\\\n", EDQ)]
    for s, payload, t in cases:
        tok, err = first(s)
        assert err is None and tok[0] == t and text(s, tok) == payload


def test_tokenize_strings_fail():  # mod.rs:626
    for s in ['"hello\'', '"\n"', '"\r"']:
        assert first_fails(s)


def test_tokenize_identifiers():  # mod.rs:634
    for s, payload, t in [("hello_world", "hello_world", KW), ("`select`", "select", DELIM),
                          ("`你 好`", "你 好", DELIM), ("@a", "a", CONFIG)]:
        tok, err = first(s)
        assert err is None and tok[0] == t and text(s, tok) == payload


def test_tokenize_identifiers_fail():  # mod.rs:650
    for s in ["``", "@", "你好", "@你好", "hello_你好"]:
        assert first_fails(s)


def test_tokenize_query_parameter():  # mod.rs:662
    for s, payload in [("$0", "0"), ("$01", "01"), ("$9", "9")]:
        tok, err = first(s)
        assert err is None and tok[0] == QPARAM and text(s, tok) == payload


def test_tokenize_query_parameter_fail():  # mod.rs:672
    for s in ["$", "$a", "$0a", "$_0"]:
        assert first_fails(s)


def test_tokenize_comment():  # mod.rs:680
    for s, idx, payload in [("hello -- world", 2, "world"), ("/* hello */", 0, " hello "),
                            ("hello /* \n */world", 2, " \n ")]:
        toks = collect(s)
        assert toks[idx][0] == COMMENT and text(s, toks[idx]) == payload


def test_tokenize_comment_fail():  # mod.rs:695
    for s in ["/*", "/* /"]:
        assert first_fails(s)


def test_tokenize_symbols():  # mod.rs:701
    cases = [(".", DOT), ("+", PLUS), ("-", MINUS), ("*", MUL), ("/", DIV), ("%", MOD), ("&", BITAND), ("|", BITOR),
             ("^", BITXOR), (">>", BITRSHIFT), ("<<", BITLSHIFT), ("=", EQ), ("!=", NOTEQ), ("<>", NOTEQ), (">", GT),
             (">=", GTEQ), ("<", LT), ("<=", LTEQ), (":", COLON), (",", COMMA), (";", SEMI), ("[", LBRACKET),
             ("]", RBRACKET), ("{", LBRACE), ("}", RBRACE), ("(", LPAREN), (")", RPAREN)]
    for s, t in cases:
        assert first(s)[0][0] == t


def test_tokenize_symbol_fail():  # mod.rs:739
    assert first_fails("!")


def test_tokenize_simple_query():  # mod.rs:744
    sql = "\nSELECT *\nFROM\n(\n    SELECT count() AS `c`\n    FROM events\n    WHERE event_type = $0\n    GROUP BY name\n)"
    got = [t[0] for t in collect(sql) if t[0] != WS]
    assert got == [KW, MUL, KW, LPAREN, KW, KW, LPAREN, RPAREN, KW, DELIM, KW, KW, KW, KW, EQ, QPARAM, KW, KW, KW,
                   RPAREN]


def test_utf8iter_pos():  # utf8_iter.rs:284
    raw = "select * \n\t你好 ❤\r\n1"
    b = raw.encode()
    # cursor after: "select" | +" * " | +"\n" | +"\t" | +"你好" | +" ❤" | +"\r\n"
    cursors = [6, 9, 10, 11, 17, 21, 23]
    want = [(1, 7), (1, 10), (2, 1), (2, 5), (2, 7), (2, 9), (3, 1)]
    assert b[cursors[-1]:] == b"1"
    assert [O.get_pos(raw, c) for c in cursors] == want


def test_unescape_string():  # literal.rs:126
    dq = [("'", "'"), ("'hello'", "'hello'"), ('h""i', 'h"i'), ("\\r\\n\\t\\\\hello 你好", "\r\n\t\\hello 你好"),
          ("\\u{767D}", "白"), ("\\\r\\\n", "\r\n")]
    for raw, want in dq:
        assert O.unescape(raw, '"') == (0, want)
    sq = [('"', '"'), ('"hello"', '"hello"'), ("h''i", "h'i"), ("\\r\\n\\t\\\\hello 你好", "\r\n\t\\hello 你好"),
          ("\\u{767D}", "白"), ("\\\r\\\n", "\r\n")]
    for raw, want in sq:
        assert O.unescape(raw, "'") == (0, want)


def test_hand_derived_lexer_edges():  # SURVEY.md App. D.11 (hand-traced through tokenizer/mod.rs)
    def toks(s):
        t, e = O.tokenize(s)
        return [x for x in t if x[0] != EOF], e

    assert toks("a--b\nc")[0] == [(KW, 0, 1), (COMMENT, 3, 4), (WS, 4, 5), (KW, 5, 6)]
    assert toks("--   ")[0] == [(COMMENT, 5, 5)]
    assert toks("/* a **/x")[0] == [(COMMENT, 2, 5), (KW, 8, 9)]
    assert toks("/***/")[0] == [(COMMENT, 2, 2)]
    assert toks("/*/")[1]["col"] == 4
    assert toks("1.")[0] == [(FLOAT, 0, 2)]
    assert toks(".")[0] == [(DOT, 0, 1)]
    assert toks(".a")[0] == [(DOT, 0, 1), (KW, 1, 2)]
    assert toks("0123")[0] == [(INT, 0, 4)]
    assert toks("0x")[0] == [(HEX, 2, 2)]
    assert toks("0x1G")[0] == [(HEX, 2, 3), (KW, 3, 4)]
    assert toks("00x1")[1]["ctx"] == "'x' cannot be a part of integer literal" and toks("00x1")[1]["col"] == 3
    assert toks("0q")[1]["ctx"] == "'q' is invalid in numeric literal"
    assert toks("1.5.2")[1]["ctx"] == "'.' cannot be a part of float literal" and toks("1.5.2")[1]["col"] == 4
    assert toks("12(")[1]["ctx"] == "'(' cannot be a part of integer literal"
    assert toks("<<=")[0] == [(BITLSHIFT, 0, 2), (EQ, 2, 3)]
    assert toks("<>=")[0] == [(NOTEQ, 0, 2), (EQ, 2, 3)]
    assert toks(">>>")[0] == [(BITRSHIFT, 0, 2), (GT, 2, 3)]
    assert toks("''")[0] == [(RAW, 1, 1)]
    assert toks("''''")[0] == [(ESQ, 1, 3)]
    assert toks("'a\\'b'")[0] == [(ESQ, 1, 5)]
    assert toks("'a\"b'")[0] == [(RAW, 1, 4)]
    assert toks("'a''")[1]["site"] == 4 and toks("'a''")[1]["col"] == 5
    assert toks("'ab\\")[1]["site"] == 4 and toks("'ab\\")[1]["col"] == 5
    e = toks("'a\n'")[1]
    assert (e["line"], e["col"], e["site"]) == (1, 3, 3)
    assert toks("@1")[1]["ctx"] == "config identifier cannot starts with numbers"
    assert toks("@ x")[1]["ctx"] == "identifier should have name" and toks("@ x")[1]["col"] == 2
    assert toks("$1.")[1]["ctx"] == "'.' cannot be a part of query parameter"
    assert toks("``")[1]["site"] == 12 and toks("`a")[1]["site"] == 14 and toks("`a")[1]["col"] == 3
    assert toks("a:b")[1]["ctx"] == "':' cannot be a part of identifier or keyword"
    assert toks("\t1d")[1]["col"] == 6
