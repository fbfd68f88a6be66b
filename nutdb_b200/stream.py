"""Host-resident batches: chunked, double-buffered parsing.

A batch that starts in HOST memory is bounded by PCIe, not by the kernels: the text goes up and
~3x as many bytes of statement records and AST nodes come down.  `StreamParser` cuts the batch into
statement-aligned chunks and keeps `workers` GPU contexts busy on one device, each with its own
stream and buffers, each driven by its own host thread (the C call releases the GIL).  While one
context copies results down, another one's kernels run and a third's text goes up -- the copy
engines and the SMs overlap without any change to the C ABI.  Results are delivered chunk by chunk.
"""
import threading

import numpy as np

from . import dispatch, gpu


class StreamParser:
    def __init__(self, device=0, workers=3):
        self.ctxs = [gpu.Context(device) for _ in range(workers)]

    def close(self):
        for c in self.ctxs:
            c.close()

    def parse(self, text, offs, on_batch, chunk_bytes=64 << 20, flags=gpu.F_NO_TOKENS):
        """Calls on_batch(first_statement_index, batch) once per chunk (from worker threads, in
        completion order).  `batch` views context-owned pinned memory and is only valid inside the
        callback.  Pin `text` / `offs` (e.g. torch.Tensor.pin_memory) for asynchronous uploads."""
        text = np.ascontiguousarray(text)
        offs = np.ascontiguousarray(offs, np.uint64)
        total = int(offs[-1] - offs[0])
        nchunks = max(1, -(-total // int(chunk_bytes)))
        ranges = [r for r in dispatch.split_statements(offs, nchunks) if r[1] > r[0]]
        lock = threading.Lock()
        state = {"next": 0, "err": None}

        def work(ctx):
            while True:
                with lock:
                    k = state["next"]
                    state["next"] += 1
                if k >= len(ranges) or state["err"] is not None:
                    return
                lo, hi = ranges[k]
                try:
                    b = ctx.parse_batch_raw(text.ctypes.data, offs.ctypes.data + 8 * lo, hi - lo, flags, copy=False)
                    on_batch(lo, b)
                except Exception as e:  # noqa: BLE001 - reported to the caller below
                    state["err"] = e
                    return

        threads = [threading.Thread(target=work, args=(c,)) for c in self.ctxs]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if state["err"] is not None:
            raise state["err"]
        return len(ranges)
