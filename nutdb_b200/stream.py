"""Host-resident batches: chunked, pipelined parsing.

A batch that starts in HOST memory is bounded by PCIe, not by the kernels: the text goes up and ~1.7x as many bytes of
statement records and wire nodes come down.  `StreamParser` is the Python face of the C dispatcher
(`nutdb_gpu_mctx_parse_stream`, csrc/dispatch.cpp): the batch is cut into statement-aligned chunks, `workers` contexts
per device -- each with its own stream, buffers and host thread -- take them in turn; while one context copies results
down, another one's kernels run and a third's text goes up.  Results are delivered chunk by chunk.
"""
import threading

import numpy as np

from . import gpu


class StreamParser:
    def __init__(self, device=0, workers=3, devices=None):
        self.m = gpu.MultiContext(devices if devices is not None else (device,), workers)

    def close(self):
        self.m.close()

    def parse(self, text, offs, on_batch, chunk_bytes=64 << 20, flags=gpu.F_NO_TOKENS):
        """Calls on_batch(first_statement_index, batch) once per chunk (from worker threads, in completion order).
        `batch` views dispatcher-owned pinned memory and is only valid inside the callback.  Pin `text` / `offs`
        (e.g. torch.Tensor.pin_memory) for asynchronous uploads.  Returns the number of chunks."""
        n, lock = [0], threading.Lock()

        def on_chunk(c):
            with lock:
                n[0] += 1
            on_batch(c.first_stmt, c.batch)

        self.m.parse_stream(np.ascontiguousarray(text), np.ascontiguousarray(offs), on_chunk, chunk_bytes, flags)
        return n[0]
