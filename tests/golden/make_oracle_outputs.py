#!/usr/bin/env python3
"""Writes tests/golden/oracle_outputs.tsv: inputs and what the C++ oracle says the reference prints for them.

One line per statement:  esc(sql) <TAB> OK|ERR <TAB> esc(text)   -- text = format!("{:?}", statement) for OK,
format!("{}", parse_error) for ERR; esc() writes backslash, tab, newline and carriage return as \\ \t \n \r.  tests/rust_diff (a cargo crate; needs a Rust toolchain, which this image lacks)
runs the REAL nutdb::parser::Parser::parse over the same inputs and reports every difference, which pins the oracle's
AST / folding / error text to the Rust reference in one command.  tests/test_oracle_parser.py checks that the oracle
still reproduces this file.

    python tests/golden/make_oracle_outputs.py            (rewrites the file)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def statements():
    import fuzz
    import test_emul_parity as T
    from nutdb_b200 import workload as W
    corpus = W.corpus_statements()
    out = list(corpus) + [s.encode() for s in T.APP_D] + list(fuzz.EXTRA_SEEDS) + list(fuzz.SIMPLE_SEEDS)
    for name in ("PREDICATES", "PREDICATES_AUTOMATON", "JOINS", "JOINS_AUTOMATON", "CASES", "CASES_AUTOMATON", "QUALIFIED",
                 "QUALIFIED_AUTOMATON", "WIDE", "WIDE_AUTOMATON"):
        out += list(getattr(T, name))
    out += [s.encode() for s in T.fold_statements(n_random=300)[::7]]
    out += fuzz.fuzz_statements(corpus + fuzz.EXTRA_SEEDS, 500, seed=2026, max_mut=4)
    for config, seed in ((2, 1), (3, 2), (4, 3)):
        text, offs = W.generate(config, 24 << 10, seed=0x60D0 + seed)
        out += [bytes(text[int(offs[i]):int(offs[i + 1])]) for i in range(len(offs) - 1) if int(offs[i + 1] - offs[i]) < 600]
    seen, uniq = set(), []
    for s in out:
        s = s.encode() if isinstance(s, str) else bytes(s)
        try:
            s.decode("utf-8")          # Parser::parse takes &str
        except UnicodeDecodeError:
            continue
        if s not in seen:
            seen.add(s)
            uniq.append(s)
    return uniq


def esc(t):
    return t.replace("\\", "\\\\").replace("\t", "\\t").replace("\n", "\\n").replace("\r", "\\r")


def unesc(t):
    out, i = [], 0
    while i < len(t):
        if t[i] == "\\" and i + 1 < len(t):
            out.append({"\\": "\\", "t": "\t", "n": "\n", "r": "\r"}[t[i + 1]])
            i += 2
        else:
            out.append(t[i])
            i += 1
    return "".join(out)


def render(stmts):
    import oracle_lib as O
    lines = []
    for s in stmts:
        r = O.parse(s)
        if r.status == 4:              # the reference panics on these (literal.rs:63): nothing to compare
            continue
        text = r.debug if r.ok else r.error
        lines.append("%s\t%s\t%s" % (esc(s.decode("utf-8")), "OK" if r.ok else "ERR", esc(text)))
    return lines


if __name__ == "__main__":
    lines = render(statements())
    with open(os.path.join(HERE, "oracle_outputs.tsv"), "w", encoding="utf-8", newline="\n") as f:
        f.write("\n".join(lines) + "\n")
    print(len(lines), "statements")
