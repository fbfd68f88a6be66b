"""ctypes binding of the host emulation of the DEVICE code (tests/emul/*.cpp compile
nutdb_b200/csrc/*_core.cuh for the CPU).  Test harness only; never used by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_EMUL = os.path.join(_HERE, "emul")
_CSRC = os.path.join(os.path.dirname(_HERE), "nutdb_b200", "csrc")
_libs = {}


def _build(name, src):
    so = os.path.join(_EMUL, f"lib{name}.so")
    deps = [os.path.join(_EMUL, src)] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)
                                         if f.endswith((".cuh", ".hpp", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", so,
                               os.path.join(_EMUL, src)])
    return so


def lex_lib():
    if "lex" not in _libs:
        L = C.CDLL(_build("emul_lex", "emul_lex.cpp"))
        L.emul_lex.restype = C.c_int64
        L.emul_lex.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                               C.POINTER(C.c_uint32)]
        L.emul_keyword.restype = C.c_uint8
        L.emul_keyword.argtypes = [C.c_char_p, C.c_uint32]
        _libs["lex"] = L
    return _libs["lex"]


def lex(text, offs, emit_all=False, chunk=32):
    """Run the device lexer logic on the host.  -> dict(type,start,end,kw, seg_begin, seg_end)"""
    L = lex_lib()
    text = np.frombuffer(bytes(text), np.uint8) if not isinstance(text, np.ndarray) else text
    text = np.ascontiguousarray(text)
    offs = np.ascontiguousarray(offs, np.uint64)
    n = int(offs[-1] - offs[0])
    nstmt = len(offs) - 1
    cap = 2 * n + 2 * nstmt + 16
    ty = np.zeros(cap, np.uint8)
    st = np.zeros(cap, np.uint32)
    en = np.zeros(cap, np.uint32)
    kw = np.zeros(cap, np.uint8)
    sb = np.zeros(nstmt + 1, np.uint32)
    se = np.zeros(nstmt + 1, np.uint32)
    nseg = C.c_uint32(0)
    base = text.ctypes.data + int(offs[0])
    nt = L.emul_lex(base, n, offs.ctypes.data, nstmt, int(emit_all), chunk, ty.ctypes.data, st.ctypes.data,
                    en.ctypes.data, kw.ctypes.data, cap, sb.ctypes.data, se.ctypes.data, C.byref(nseg))
    assert nt >= 0, "token capacity overflow"
    return dict(type=ty[:nt], start=st[:nt], end=en[:nt], kw=kw[:nt], seg_begin=sb[:nseg.value],
                seg_end=se[:nseg.value], n_seg=nseg.value)


def keyword(word):
    b = word.encode()
    return lex_lib().emul_keyword(b, len(b))


def parse_lib():
    if "parse" not in _libs:
        so = os.path.join(_EMUL, "libemul_parse.so")
        srcs = [os.path.join(_EMUL, "emul_parse.cpp"), os.path.join(_EMUL, "emul_lex.cpp"),
                os.path.join(_EMUL, "emul_lex2.cpp"), os.path.join(_EMUL, "emul_lex3.cpp")]
        deps = srcs + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".hpp", ".h"))]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", so] + srcs)
        L = C.CDLL(so)
        L.emul_parse_batch.restype = C.c_int
        L.emul_parse_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p,
                                       C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.c_uint64,
                                       C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_uint32, C.POINTER(C.c_uint64)]
        L.emul_fast_hits.restype = C.c_uint64
        L.emul_lex2_punts.restype = C.c_uint64
        L.emul_lex3_punts.restype = C.c_uint64
        L.emul_lex3_check_classes.restype = C.c_int
        L.emul_lex3_check_classes.argtypes = [C.c_void_p]
        L.emul_set_lexer.argtypes = [C.c_int, C.c_uint32]
        _libs["parse"] = L
    return _libs["parse"]


def set_lexer(version, seg_len=1024):
    """1 = thread-per-chunk walker (lex_core.cuh), 2 = three-pass mask lexer (lex2_core.cuh), 3 = single-pass lexer
    (lex3_core.cuh, the product's)."""
    parse_lib().emul_set_lexer(version, seg_len)


def lex2_punts():
    return parse_lib().emul_lex2_punts()


def lex3_punts():
    return parse_lib().emul_lex3_punts()


def lex3_check_classes(bytes32):
    b = np.frombuffer(bytes(bytes32), np.uint8).copy()
    assert len(b) == 32
    return parse_lib().emul_lex3_check_classes(b.ctypes.data)


def set_fast(on):
    """0 / False: the automaton alone; 1: the narrow table-driven pass first; 2 / True: narrow, then wide (the device's order)."""
    parse_lib().emul_set_fast(2 if on is True else int(on))


def wide_hits():
    L = parse_lib()
    L.emul_wide_hits.restype = C.c_uint64
    return L.emul_wide_hits()


def fast_hits(reset=True):
    L = parse_lib()
    h = L.emul_fast_hits()
    if reset:
        L.emul_reset_fast_hits()
    return h


def parse_batch(text, offs, chunk=32, stack_cap=4096):
    """Device lexer+parser logic on the host -> object with the NutdbBatch arrays."""
    from oracle_lib import NODE_DT, STMT_DT, ERR_DT, Batch
    L = parse_lib()
    text = np.frombuffer(bytes(text), np.uint8) if not isinstance(text, np.ndarray) else text
    text = np.ascontiguousarray(text)
    offs = np.ascontiguousarray(offs, np.uint64)
    n = int(offs[-1] - offs[0])
    nstmt = len(offs) - 1
    tcap = n + 2 * nstmt + 16
    b = Batch()
    b.stmt = np.zeros(nstmt, STMT_DT)
    node = np.zeros(2 * tcap + 16, NODE_DT)
    err = np.zeros(nstmt + 1, ERR_DT)
    ty = np.zeros(tcap, np.uint8)
    st = np.zeros(tcap, np.uint32)
    en = np.zeros(tcap, np.uint32)
    kw = np.zeros(tcap, np.uint8)
    nn, ne, nt = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    rc = L.emul_parse_batch(text.ctypes.data, offs.ctypes.data, nstmt, chunk, stack_cap, b.stmt.ctypes.data,
                            node.ctypes.data, len(node), C.byref(nn), err.ctypes.data, len(err), C.byref(ne),
                            ty.ctypes.data, st.ctypes.data, en.ctypes.data, kw.ctypes.data, tcap, C.byref(nt))
    assert rc == 0, f"emul_parse_batch failed rc={rc} (-1 overflow, -2 scanned context states disagree)"
    b.node, b.err = node[:nn.value], err[:ne.value]
    b.tok_type, b.tok_start, b.tok_end, b.tok_kw = ty[:nt.value], st[:nt.value], en[:nt.value], kw[:nt.value]
    return b


def split(text):
    """Sequential reference of the statement splitter -> uint64 offsets (n + 1)."""
    L = parse_lib()
    L.emul_split.restype = C.c_int64
    L.emul_split.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    t = np.frombuffer(bytes(text), np.uint8) if not isinstance(text, np.ndarray) else np.ascontiguousarray(text)
    offs = np.zeros(len(t) + 2, np.uint64)
    n = L.emul_split(t.ctypes.data, len(t), offs.ctypes.data, len(offs))
    assert n >= 0
    return offs[:n + 1].copy()
