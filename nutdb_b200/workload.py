"""Synthetic SQL batches for the BASELINE.json configs (bench/test tooling, not the parser).

ctypes binding of csrc/workload.cpp (libnutdb_workload.so): config 2 = short SELECT/INSERT/CREATE,
3 = string/comment/quoted-identifier stress with 5 % malformed statements, 4 = deep nesting;
`corpus()` tiles the reference's own tests/sql files (config 1).  Deterministic in
(config, seed, target_bytes), independent of the number of generator threads.
"""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.path.join(_HERE, "libnutdb_workload.so")
_lib = None

SEEDS = {2: 0x5EED0002, 3: 0x5EED0003, 4: 0x5EED0004, 5: 0x5EED0005}


def build(force=False):
    src = os.path.join(_CSRC, "workload.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-o", _SO, src])
    return _SO


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.nutdb_workload_create.restype = C.c_void_p
        L.nutdb_workload_create.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int]
        for f in ("bytes", "statements"):
            getattr(L, "nutdb_workload_" + f).restype = C.c_uint64
            getattr(L, "nutdb_workload_" + f).argtypes = [C.c_void_p]
        L.nutdb_workload_text.restype = C.c_void_p
        L.nutdb_workload_text.argtypes = [C.c_void_p]
        L.nutdb_workload_offsets.restype = C.c_void_p
        L.nutdb_workload_offsets.argtypes = [C.c_void_p]
        L.nutdb_workload_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def generate(config, target_bytes, seed=None, nthreads=0, out_text=None):
    """-> (text uint8[bytes + 64 zero padding], stmt_off uint64[n+1]).

    `out_text`, if given, is a writable uint8 array (e.g. pinned memory) to fill instead of a new one."""
    L = lib()
    seed = SEEDS.get(config, 0x5EED0000 + config) if seed is None else seed
    h = L.nutdb_workload_create(config, seed, int(target_bytes), nthreads)
    try:
        nb, ns = L.nutdb_workload_bytes(h), L.nutdb_workload_statements(h)
        if out_text is None:
            text = np.empty(nb + 64, np.uint8)
        else:
            text = out_text[:nb + 64]
            assert len(text) == nb + 64, "out_text too small"
        C.memmove(text.ctypes.data, L.nutdb_workload_text(h), nb + 64)
        offs = np.empty(ns + 1, np.uint64)
        C.memmove(offs.ctypes.data, L.nutdb_workload_offsets(h), 8 * (ns + 1))
        return text, offs
    finally:
        L.nutdb_workload_free(h)


def corpus_statements(golden_dir=None):
    """The reference's tests/sql/1..14.sql + the two bench strings (committed under tests/golden)."""
    g = golden_dir or os.path.join(os.path.dirname(_HERE), "tests", "golden")
    files = sorted(glob.glob(os.path.join(g, "sql", "*.sql")), key=lambda p: int(os.path.basename(p)[:-4]))
    stmts = [open(f, "rb").read() for f in files]
    stmts.append(open(os.path.join(g, "bench_long.sql"), "rb").read())
    stmts.append(b"SELECT * FROM table WHERE 1 = 1")
    return stmts


def corpus(target_bytes, golden_dir=None):
    """Config 1: the corpus tiled round-robin (file order, each file one statement)."""
    stmts = corpus_statements(golden_dir)
    per = sum(len(s) for s in stmts)
    reps = max(1, int(target_bytes) // per)
    lens = np.tile(np.array([len(s) for s in stmts], np.uint64), reps)
    offs = np.zeros(len(lens) + 1, np.uint64)
    offs[1:] = np.cumsum(lens)
    one = np.frombuffer(b"".join(stmts), np.uint8)
    text = np.concatenate([np.tile(one, reps), np.zeros(64, np.uint8)])
    return text, offs
