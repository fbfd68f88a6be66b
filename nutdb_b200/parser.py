"""Host-side mirror of the reference's public entry `nutdb::parser::Parser::parse`
(reference src/parser/mod.rs:26-29, exported through src/lib.rs:3-4).

    Parser.parse(sql) -> Statement            # Ok(Statement)
    raises ParseError                         # Err(ParseError::{LexError, SyntaxError})

Same argument meaning and error behaviour as the reference: one statement per call, parsing stops
at the first `;`/EOF in statement-final position, first error wins.  Everything runs on the GPU
through libnutdb_gpu.so; `Parser.parse_many` is the batch form the hardware wants.
"""
import ctypes as C

import numpy as np

from . import gpu

LEX_ERROR, SYNTAX_ERROR, LIMIT = 1, 2, 3


class ParseError(Exception):
    """ParseError (reference src/parser/error.rs:8-15).  str(e) is the reference's Display text."""

    def __init__(self, message, record):
        super().__init__(message)
        self.record = record                    # the NutdbError fields
        self.is_lex_error = record["cls"] == LEX_ERROR
        self.line, self.col = record["line"], record["col"]


class Statement:
    """A successfully parsed statement: flat post-order AST nodes + the source they point into."""

    def __init__(self, sql, nodes, debug):
        self.sql, self.nodes, self._debug = sql, nodes, debug

    def __repr__(self):     # == format!("{:?}", statement) of the reference
        return self._debug

    @property
    def kind(self):
        return int(self.nodes["kind"][-1])


class Parser:
    _ctx = {}

    @classmethod
    def context(cls, device=0):
        if device not in cls._ctx:
            cls._ctx[device] = gpu.Context(device)
        return cls._ctx[device]

    @classmethod
    def parse(cls, sql, device=0):
        r = cls.parse_many([sql], device)[0]
        if isinstance(r, ParseError):
            raise r
        return r

    @classmethod
    def parse_many(cls, statements, device=0):
        """-> list with a Statement or a ParseError (not raised) per input, in order."""
        raws = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in statements]
        offs = np.zeros(len(raws) + 1, np.uint64)
        offs[1:] = np.cumsum([len(r) for r in raws])
        text = np.frombuffer(b"".join(raws) + b"\0" * 16, np.uint8)
        ctx = cls.context(device)
        b = ctx.parse_batch(text, offs, flags=gpu.F_NO_TOKENS, copy=False)
        L = gpu.lib()
        out = []
        errs = {int(e["stmt"]): e for e in b.err}
        for i, raw in enumerate(raws):
            st = b.stmt[i]
            fn = L.nutdb_fmt_debug if st["status"] == 0 else L.nutdb_fmt_error
            need = fn(C.byref(b.raw), i, raw, len(raw), None, 0)
            buf = C.create_string_buffer(max(int(need), 1))
            fn(C.byref(b.raw), i, raw, len(raw), buf, len(buf))
            text_out = buf.value.decode("utf-8")
            if st["status"] == 0:
                nb, nc = int(st["node_begin"]), int(st["node_count"])
                out.append(Statement(raw, b.node[nb:nb + nc].copy(), text_out))
            else:
                e = errs[i]
                out.append(ParseError(text_out, {k: int(e[k]) for k in e.dtype.names}))
        return out
