"""ctypes binding of libnutdb_gpu.so (the C ABI in include/nutdb_gpu.h).

This is the only way the package reaches the parser: there is no Python or CPU implementation
behind it, and loading fails loudly when the CUDA library is missing or no device is present.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

F_NO_TOKENS, F_DEVICE_INPUT, F_NO_HOST_COPY, F_ALL_TOKENS, F_WIRE_STMT, F_OFFSETS32 = 1, 2, 4, 8, 16, 32

NODE_DT = np.dtype([("kind", "u1"), ("sub", "u1"), ("aux", "<u2"), ("parent", "<u4"), ("a", "<u4"), ("b", "<u4")])
PNODE_DT = np.dtype("<u4")   # the wire form of a node: one 32-bit word (NUTDB_PN_* in nutdb_gpu.h); kind = word & 127
STMT_DT = np.dtype([("status", "<u4"), ("tok_begin", "<u4"), ("tok_count", "<u4"), ("node_begin", "<u4"),
                    ("node_count", "<u4"), ("tok_used", "<u4")])
ERR_DT = np.dtype([("stmt", "<u4"), ("cls", "<u2"), ("code", "<u2"), ("line", "<u4"), ("col", "<u4"),
                   ("pos", "<u4"), ("a", "<u4"), ("b", "<u4"), ("c", "<u4")])


class NutdbBatch(C.Structure):
    _fields_ = [("n_stmt", C.c_uint64), ("n_tok", C.c_uint64), ("n_node", C.c_uint64), ("n_err", C.c_uint64),
                ("stmt", C.c_void_p), ("tok_type", C.c_void_p), ("tok_start", C.c_void_p), ("tok_end", C.c_void_p),
                ("tok_kw", C.c_void_p), ("node", C.c_void_p), ("err", C.c_void_p), ("impl", C.c_void_p),
                ("pnode", C.c_void_p), ("n_ext", C.c_uint64), ("ext", C.c_void_p), ("wstmt", C.c_void_p)]


class NutdbMShard(C.Structure):
    _fields_ = [("device_index", C.c_int), ("sql", C.c_void_p), ("stmt_off", C.c_void_p), ("n_stmt", C.c_uint64),
                ("flags", C.c_uint32), ("first_stmt", C.c_uint64)]


class NutdbMChunk(C.Structure):
    _fields_ = [("shard", C.c_uint64), ("first_stmt", C.c_uint64), ("device", C.c_int), ("on_device", C.c_int),
                ("batch", NutdbBatch)]


CHUNK_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(NutdbMChunk))
MF_GATHER_DEVICE0, MF_SERIAL_CALLBACKS = 0x100, 0x200


class NutdbBatchDevice(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("stmt", "tok_type", "tok_start", "tok_end", "tok_kw", "node", "err", "wstmt")]


class NutdbGpuError(RuntimeError):
    pass


_lib = None


def lib():
    """Loads libnutdb_gpu.so (building it in-tree with nvcc if stale).  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    so = _build.GPU_SO
    try:
        so = _build.build_gpu()
    except Exception as e:  # no nvcc on this machine: use the prebuilt library if there is one
        if not os.path.exists(so):
            raise NutdbGpuError(f"libnutdb_gpu.so is missing and cannot be built: {e}") from e
    L = C.CDLL(so)
    L.nutdb_gpu_ctx_create.restype = C.c_void_p
    L.nutdb_gpu_ctx_create.argtypes = [C.c_int]
    L.nutdb_gpu_ctx_destroy.argtypes = [C.c_void_p]
    L.nutdb_gpu_last_error.restype = C.c_char_p
    L.nutdb_gpu_last_error.argtypes = [C.c_void_p]
    L.nutdb_gpu_parse_batch.restype = C.c_int
    L.nutdb_gpu_parse_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32,
                                        C.POINTER(NutdbBatch)]
    L.nutdb_gpu_batch_free.argtypes = [C.c_void_p, C.POINTER(NutdbBatch)]
    L.nutdb_batch_expand_stmts.restype = C.c_int
    L.nutdb_batch_expand_stmts.argtypes = [C.POINTER(NutdbBatch), C.c_void_p]
    L.nutdb_batch_expand_nodes.restype = C.c_int
    L.nutdb_batch_expand_nodes.argtypes = [C.POINTER(NutdbBatch), C.c_void_p]
    L.nutdb_gpu_batch_device.restype = C.c_int
    L.nutdb_gpu_batch_device.argtypes = [C.POINTER(NutdbBatch), C.POINTER(NutdbBatchDevice)]
    L.nutdb_gpu_batch_hash.restype = C.c_int
    L.nutdb_gpu_batch_hash.argtypes = [C.POINTER(NutdbBatch), C.POINTER(C.c_uint64)]
    L.nutdb_gpu_parse.restype = C.c_int
    L.nutdb_gpu_parse.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(NutdbBatch)]
    L.nutdb_gpu_split_statements.restype = C.c_int
    L.nutdb_gpu_split_statements.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32,
                                             C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.nutdb_gpu_last_timing.restype = C.c_int
    L.nutdb_gpu_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float * 5)]
    L.nutdb_gpu_last_launches.restype = C.c_int
    L.nutdb_gpu_last_launches.argtypes = [C.c_void_p]
    L.nutdb_gpu_version.restype = C.c_char_p
    L.nutdb_gpu_set_profiling.argtypes = [C.c_void_p, C.c_int]
    L.nutdb_gpu_kernel_timing.restype = C.c_int
    L.nutdb_gpu_kernel_timing.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]
    L.nutdb_gpu_last_slow_statements.restype = C.c_uint64
    L.nutdb_gpu_last_slow_statements.argtypes = [C.c_void_p]
    L.nutdb_gpu_last_wide_statements.restype = C.c_uint64
    L.nutdb_gpu_last_wide_statements.argtypes = [C.c_void_p]
    L.nutdb_gpu_last_exact_lexed_statements.restype = C.c_uint64
    L.nutdb_gpu_last_exact_lexed_statements.argtypes = [C.c_void_p]
    L.nutdb_gpu_mctx_create.restype = C.c_void_p
    L.nutdb_gpu_mctx_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int]
    L.nutdb_gpu_mctx_destroy.argtypes = [C.c_void_p]
    L.nutdb_gpu_mctx_last_error.restype = C.c_char_p
    L.nutdb_gpu_mctx_last_error.argtypes = [C.c_void_p]
    L.nutdb_gpu_mctx_device_count.argtypes = [C.c_void_p]
    L.nutdb_gpu_mctx_parse_shards.restype = C.c_int
    L.nutdb_gpu_mctx_parse_shards.argtypes = [C.c_void_p, C.POINTER(NutdbMShard), C.c_uint64, C.c_uint32, CHUNK_FN, C.c_void_p]
    L.nutdb_gpu_mctx_parse_stream.restype = C.c_int
    L.nutdb_gpu_mctx_parse_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32,
                                              CHUNK_FN, C.c_void_p]
    L.nutdb_gpu_copy_to_host.restype = C.c_int
    L.nutdb_gpu_copy_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.nutdb_gpu_ctx_stream.restype = C.c_void_p
    L.nutdb_gpu_ctx_stream.argtypes = [C.c_void_p]
    for f in ("nutdb_fmt_debug", "nutdb_fmt_error"):
        getattr(L, f).restype = C.c_size_t
        getattr(L, f).argtypes = [C.POINTER(NutdbBatch), C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    _lib = L
    return L


def _view(ptr, count, dtype):
    if not ptr or count == 0:
        return np.zeros(0, dtype)
    dtype = np.dtype(dtype)
    buf = (C.c_uint8 * (count * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype, count)


class Batch:
    """Result of one parse_batch call: numpy arrays with the layout of NutdbBatch."""

    def __init__(self, ctx, raw, copy):
        self._ctx, self.raw = ctx, raw
        g = (lambda a: a.copy()) if copy else (lambda a: a)
        self.n_stmt, self.n_tok, self.n_node, self.n_err = raw.n_stmt, raw.n_tok, raw.n_node, raw.n_err
        # NUTDB_F_WIRE_STMT: 8-byte records crossed PCIe (status | node_count << 4 | tok_used << 34); the NutdbStmt form
        # is host arithmetic on top of them (the `stmt` property)
        self.wstmt = g(_view(raw.wstmt, raw.n_stmt, np.uint64)) if raw.wstmt else None
        self._stmt = g(_view(raw.stmt, raw.n_stmt, STMT_DT)) if (raw.stmt or not raw.wstmt) else None
        if copy and self._stmt is None:
            self._stmt = self._expand_stmts()
        self.pnode = g(_view(raw.pnode, raw.n_node, PNODE_DT))   # what crossed PCIe: 4 bytes per node
        # the expanded records (spans, child counts, parent links) are host arithmetic on top of them: eager for a
        # copied batch (the context's buffers are still this batch's), on first use for a view (streaming callers
        # that only want the wire form never pay for it)
        self._node = None
        if copy:
            self._node = self._expand()
        self.err = g(_view(raw.err, raw.n_err, ERR_DT))
        self.tok_type = g(_view(raw.tok_type, raw.n_tok, np.uint8))
        self.tok_kw = g(_view(raw.tok_kw, raw.n_tok, np.uint8))
        self.tok_start = g(_view(raw.tok_start, raw.n_tok, np.uint32))
        self.tok_end = g(_view(raw.tok_end, raw.n_tok, np.uint32))

    def _expand_stmts(self):
        out = np.zeros(int(self.n_stmt), STMT_DT)
        if self.n_stmt:
            rc = lib().nutdb_batch_expand_stmts(C.byref(self.raw), out.ctypes.data)
            if rc != 0:
                raise NutdbGpuError(f"nutdb_batch_expand_stmts failed ({rc})")
        return out

    @property
    def stmt(self):
        if self._stmt is None:
            self._stmt = self._expand_stmts()
        return self._stmt

    def _expand(self):
        node = np.zeros(len(self.pnode), NODE_DT)
        if len(self.pnode):
            rc = lib().nutdb_batch_expand_nodes(C.byref(self.raw), node.ctypes.data)
            if rc != 0:
                raise NutdbGpuError(f"nutdb_batch_expand_nodes failed ({rc})")
        return node

    @property
    def node(self):
        if self._node is None:
            self._node = self._expand()
        return self._node

    def device_pointers(self):
        d = NutdbBatchDevice()
        rc = lib().nutdb_gpu_batch_device(C.byref(self.raw), C.byref(d))
        if rc != 0:
            raise NutdbGpuError("batch is no longer live")
        return {n: getattr(d, n) for n, _ in NutdbBatchDevice._fields_}

    def device_hash(self):
        """64-bit checksum of the batch's device-resident outputs (nutdb_gpu_batch_hash)."""
        h = C.c_uint64(0)
        rc = lib().nutdb_gpu_batch_hash(C.byref(self.raw), C.byref(h))
        if rc != 0:
            raise NutdbGpuError(f"nutdb_gpu_batch_hash failed ({rc}): batch is no longer live?")
        return int(h.value)


def host_hash(batch):
    """The checksum of nutdb_gpu_batch_hash computed with numpy from a Batch's host arrays (test helper)."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)

    def mix(z):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))

    total = np.uint64(0)
    arrays = [np.ascontiguousarray(batch.stmt).view(np.uint32), batch.tok_type, batch.tok_start, batch.tok_end,
              batch.tok_kw, np.ascontiguousarray(batch.pnode).view(np.uint32), np.ascontiguousarray(batch.err).view(np.uint32)]
    with np.errstate(over="ignore"):
        for a, w in enumerate(arrays):
            w = np.asarray(w).reshape(-1)
            if w.size == 0:
                continue
            i = np.arange(w.size, dtype=np.uint64) + np.uint64(((a + 1) * 0x9E3779B97F4A7C15) & int(M))
            total = total + np.sum(mix(mix(i) ^ w.astype(np.uint64)), dtype=np.uint64)
    return int(total)


class Context:
    """One CUDA device (one process per GPU).  Not thread-safe; one live batch at a time."""

    def __init__(self, device=0):
        self._h = lib().nutdb_gpu_ctx_create(device)
        if not self._h:
            raise NutdbGpuError(f"cannot create a nutdb GPU context on CUDA device {device}: no usable device. "
                                "There is no CPU fallback.")
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.nutdb_gpu_ctx_destroy(self._h)
        self._h = None

    __del__ = close

    def last_error(self):
        return lib().nutdb_gpu_last_error(self._h).decode()

    def parse_batch_raw(self, sql_ptr, off_ptr, n_stmt, flags=0, copy=True):
        raw = NutdbBatch()
        rc = lib().nutdb_gpu_parse_batch(self._h, sql_ptr, off_ptr, n_stmt, flags, C.byref(raw))
        if rc != 0:
            raise NutdbGpuError(f"nutdb_gpu_parse_batch failed ({rc}): {self.last_error()}")
        return Batch(self, raw, copy)

    def parse_batch(self, text, offs, flags=0, copy=True):
        """text: bytes / uint8 array holding all statements; offs: uint64[n+1] ascending offsets into it."""
        t = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else np.ascontiguousarray(text)
        o = np.ascontiguousarray(offs, np.uint32 if flags & F_OFFSETS32 else np.uint64)
        assert t.dtype == np.uint8 and o.ndim == 1 and len(o) >= 1
        return self.parse_batch_raw(t.ctypes.data, o.ctypes.data, len(o) - 1, flags, copy)

    def split_statements(self, text, flags=0):
        """Raw buffer -> uint64 offsets (n+1) of its statements: every ';' outside literals / comments ends one."""
        if isinstance(text, (bytes, bytearray, memoryview)):
            t = np.frombuffer(text, np.uint8)
        else:
            t = np.ascontiguousarray(text)
        off, n = C.c_void_p(), C.c_uint64(0)
        rc = lib().nutdb_gpu_split_statements(self._h, t.ctypes.data if len(t) else None, len(t), flags, C.byref(off), C.byref(n))
        if rc != 0:
            raise NutdbGpuError(f"nutdb_gpu_split_statements failed ({rc}): {self.last_error()}")
        return _view(off.value, n.value + 1, np.uint64).copy()

    def timing(self):
        ms = (C.c_float * 5)()
        lib().nutdb_gpu_last_timing(self._h, C.byref(ms))
        return dict(h2d=ms[0], lex=ms[1], parse=ms[2], d2h=ms[3], total=ms[4])

    def launches(self):
        return lib().nutdb_gpu_last_launches(self._h)

    def slow_statements(self):
        return lib().nutdb_gpu_last_slow_statements(self._h)

    def force_lookback(self, on):
        """Debug: lex with k_lex3 (one token stream, look-back scans between tiles) instead of the range lexer."""
        L = lib()
        L.nutdb_gpu_debug_force_lookback.argtypes = [C.c_void_p, C.c_int]
        L.nutdb_gpu_debug_force_lookback(self._h, 1 if on else 0)

    def last_lookback(self):
        L = lib()
        L.nutdb_gpu_debug_last_lookback.argtypes = [C.c_void_p]
        return bool(L.nutdb_gpu_debug_last_lookback(self._h))

    def wide_statements(self):
        return lib().nutdb_gpu_last_wide_statements(self._h)

    def exact_lexed_statements(self):
        return lib().nutdb_gpu_last_exact_lexed_statements(self._h)

    def set_profiling(self, on):
        lib().nutdb_gpu_set_profiling(self._h, 1 if on else 0)

    def kernel_timing(self):
        """[(kernel name, ms)] of the last parse_batch call (needs set_profiling(True))."""
        L = lib()
        n = L.nutdb_gpu_kernel_timing(self._h, -1, None, None)
        out = []
        for i in range(n):
            name, ms = C.c_char_p(), C.c_float()
            L.nutdb_gpu_kernel_timing(self._h, i, C.byref(name), C.byref(ms))
            out.append((name.value.decode(), ms.value))
        return out

    def stream(self):
        """cudaStream_t (as int) that all work of this context is issued on."""
        return lib().nutdb_gpu_ctx_stream(self._h)


def copy_to_host(ptr, count, dtype):
    """count elements of `dtype` at device pointer `ptr` -> numpy array (nutdb_gpu_copy_to_host)."""
    a = np.zeros(int(count), dtype)
    if count:
        rc = lib().nutdb_gpu_copy_to_host(a.ctypes.data, ptr, a.nbytes)
        if rc != 0:
            raise NutdbGpuError(f"nutdb_gpu_copy_to_host failed ({rc})")
    return a


class Chunk:
    """One gathered shard handed to a dispatcher callback: `batch` views the gather slot (only valid inside the
    callback; with on_device the array pointers are device-0 memory and only the counts / raw pointers are usable)."""

    def __init__(self, c):
        self.shard, self.first_stmt, self.device, self.on_device = int(c.shard), int(c.first_stmt), int(c.device), bool(c.on_device)
        raw = NutdbBatch()
        C.memmove(C.byref(raw), C.byref(c.batch), C.sizeof(NutdbBatch))
        self.raw = raw
        self.batch = None if self.on_device else Batch(None, raw, False)


class MultiContext:
    """nutdb_gpu_mctx_*: one process driving several GPUs (or several pipelined contexts on one)."""

    def __init__(self, devices=(0,), workers_per_device=3):
        devs = (C.c_int * len(devices))(*devices)
        self._h = lib().nutdb_gpu_mctx_create(devs, len(devices), workers_per_device)
        if not self._h:
            raise NutdbGpuError(f"cannot create a nutdb GPU dispatcher on CUDA devices {list(devices)}. There is no CPU fallback.")
        self.devices = list(devices)

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.nutdb_gpu_mctx_destroy(self._h)
        self._h = None

    __del__ = close

    def _callback(self, on_chunk, errs):
        def cb(_user, cptr):
            try:
                on_chunk(Chunk(cptr.contents))
            except BaseException as e:  # noqa: BLE001 - re-raised by the caller of the dispatcher
                errs.append(e)
        return CHUNK_FN(cb)

    def _finish(self, rc, errs, what):
        if errs:
            raise errs[0]
        if rc != 0:
            raise NutdbGpuError(f"{what} failed ({rc}): {lib().nutdb_gpu_mctx_last_error(self._h).decode()}")

    def parse_stream(self, text, offs, on_chunk, chunk_bytes=64 << 20, flags=F_NO_TOKENS):
        """One host batch -> on_chunk(Chunk) per chunk, in completion order, from the dispatcher's worker threads."""
        t = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else np.ascontiguousarray(text)
        # NUTDB_F_OFFSETS32: the caller's offsets are uint32 (half the upload); otherwise uint64
        o = np.ascontiguousarray(offs, np.uint32 if flags & F_OFFSETS32 else np.uint64)
        errs = []
        cb = self._callback(on_chunk, errs)
        rc = lib().nutdb_gpu_mctx_parse_stream(self._h, t.ctypes.data, o.ctypes.data, len(o) - 1, int(chunk_bytes), flags, cb, None)
        self._finish(rc, errs, "nutdb_gpu_mctx_parse_stream")

    def parse_shards(self, shards, on_chunk, flags=F_NO_TOKENS):
        """shards: [(device_index, sql_ptr, off_ptr, n_stmt, shard_flags, first_stmt)] (raw pointers: host or device)."""
        arr = (NutdbMShard * len(shards))()
        for i, (d, sp, op, n, f, first) in enumerate(shards):
            arr[i] = NutdbMShard(d, sp, op, n, f, first)
        errs = []
        cb = self._callback(on_chunk, errs)
        rc = lib().nutdb_gpu_mctx_parse_shards(self._h, arr, len(shards), flags, cb, None)
        self._finish(rc, errs, "nutdb_gpu_mctx_parse_shards")
