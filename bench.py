#!/usr/bin/env python3
"""bench.py -- BASELINE.json metric: GB/s of SQL text lexed+parsed (and statements/s) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|1] [--bytes B]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...     (N > 1)
  python bench.py --impl reference ...      CPU arm: the restated reference parser on all host cores

A step = one nutdb_gpu_parse_batch() over one batch (default: config 2, the 1 GiB batch of short
SELECT/INSERT/CREATE statements).  `value` is timed with the batch resident in HBM and outputs left
on the device; `e2e` goes through the same C-ABI call with pinned HOST buffers (H2D of the text and
offsets, D2H of statement records, flat AST nodes and error records inside the timed region).
With N > 1 every rank parses its own shard of the statement log (weak scaling, no collective on
the data path; rank 0 only collects the per-rank times).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {1: "config1: tests/sql corpus tiled (reference bench workload)",
             2: "config2: synthetic short SELECT/INSERT/CREATE statements",
             3: "config3: string/quoted-identifier/comment-heavy mix with 5% malformed statements",
             4: "config4: deeply nested expressions and subqueries (depth <= 256)",
             5: "config5: statement log of config-2 chunks (seed 0x5EED0005 + chunk), pre-staged in HBM, chunks dealt to "
                "the GPUs in contiguous ranges"}
METRIC = "GB/s SQL text lexed+parsed"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_workload(config, nbytes, seed_offset=0, pinned=False):
    from nutdb_b200 import workload as W
    if config == 1:
        text, offs = W.corpus(nbytes)
    elif config == 5:  # chunk `seed_offset` of the log
        text, offs = W.generate(2, nbytes, seed=W.SEEDS[5] + seed_offset)
    else:
        text, offs = W.generate(config, nbytes, seed=W.SEEDS[config] + seed_offset)
    return text, offs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(text, offs, budget_s=15.0):
    """Restated reference parser (oracle port) on all host cores over a bounded prefix of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    cores = os.cpu_count() or 1
    # calibrate on 8 MiB, then size the sample for ~budget_s/3 per repetition (3 reps, best taken)
    total = int(offs[-1])

    def prefix(nbytes):
        k = int(np.searchsorted(offs, min(nbytes, total), side="right")) - 1
        k = max(k, 1)
        return k, int(offs[k])

    k, nb = prefix(8 << 20)
    t, _, _ = O.bench(text, offs[:k + 1], cores, reps=1)
    rate = nb / max(t, 1e-6)
    k, nb = prefix(int(rate * budget_s / 3))
    t, ok, toks = O.bench(text, offs[:k + 1], cores, reps=3)
    return {"value": nb / t / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
            "statements_per_s": k / t,
            "sample": f"first {nb} bytes / {k} statements of the same workload, parse+drop per statement, "
                      f"{cores} threads, best of 3", "sample_bytes": nb, "sample_statements": k,
            "ok_statements": int(ok), "tokens": int(toks)}


def alg_counts(text, offs, sample_bytes=16 << 20):
    """Oracle counts (T pulled tokens, M algorithmic nodes) on a sample, for the B_alg formula of SURVEY 8d."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    k = max(1, int(np.searchsorted(offs, min(sample_bytes, int(offs[-1])), side="right")) - 1)
    b = O.parse_batch(text, offs[:k + 1])
    nb = int(offs[k])
    return {"bytes": nb, "stmts": k, "T": int(b.t_alg), "M": int(b.m_alg)}


def workload_config(config, n_in, n_stmt):
    """The `config` object of the JSON line: what the workload is, nothing measured -- identical for both arms."""
    return {"workload": WORKLOADS[config], "bytes_per_gpu": int(n_in), "statements_per_gpu": int(n_stmt),
            "sharding": "statement ranges, one shard per GPU, outputs stay sharded",
            "l2": "input (1 GiB class) and every intermediate array are larger than the 126 MB L2; no flush needed"}


def run_reference(args, rank, world):
    """CPU arm: the restated reference parser (oracle port; the Rust reference cannot be built in this image) on all
    host cores.  Same workload and `config` as our arm; every step parses a bounded prefix of it."""
    if rank != 0:
        return
    text, offs = make_workload(args.config, args.bytes)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    cores = os.cpu_count() or 1
    total = int(offs[-1])
    # each step = a bounded sample sized so that warmup+steps finish within a few minutes
    t, _, _ = O.bench(text, offs[:int(np.searchsorted(offs, 8 << 20, side="right"))], cores, reps=1)
    rate = (8 << 20) / max(t, 1e-6)
    per_step = min(total, int(rate * 8.0))
    k = max(1, int(np.searchsorted(offs, per_step, side="right")) - 1)
    nb = int(offs[k])
    for _ in range(args.warmup):
        O.bench(text, offs[:k + 1], cores, reps=1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.bench(text, offs[:k + 1], cores, reps=1)
    dt = (time.perf_counter() - t0) / args.steps
    v = nb / dt / 1e9
    sample = (f"first {nb} bytes / {k} statements of the workload per step, {cores} threads; C++ restatement of the "
              "reference parser (oracle port: no Rust toolchain in this image, so NOT the Rust parser itself)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "statements_per_s": k / dt,
            "config": workload_config(args.config, total, len(offs) - 1),
            "sample_per_step": {"bytes": nb, "statements": k},
            "cpu_baseline": {"value": v, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_log(args, rank, world, local, barrier):
    """a fixed statement log (default 64 GiB, 1 GiB chunks of the config-2 generator) parsed
    by N GPUs -- STRONG scaling.  Chunks are dealt in contiguous ranges, generated on the host and pre-staged in HBM
    before the timed regions.  Two timed regions: `value` = all kernels of every chunk with the outputs left sharded
    on the GPUs (CUDA events); `gather` = the same log through the C dispatcher (nutdb_gpu_mctx_parse_shards) with
    every chunk's statement records, wire nodes and error records GATHERED to pinned host memory inside the timed
    region, each GPU over its own PCIe link.  After timing, a 64-bit checksum of every chunk's outputs is computed on
    the device and summed over chunks and ranks: it must not depend on N.

    --single-process: one process drives all N GPUs through one dispatcher (the C-ABI multi-GPU path a Rust host
    would use); --gather device0 then gathers into GPU 0's memory over NVLink instead."""
    import concurrent.futures as cf
    import torch
    import torch.distributed as dist
    from nutdb_b200 import gpu, workload as W
    chunk_bytes = min(args.bytes, 1 << 30)
    single = args.single_process
    ndev = args.gpus if single else 1            # devices this process drives
    nparts = args.gpus if single else world      # GPUs sharing the log
    part0 = 0 if single else rank
    n_chunks = max(nparts, args.log_bytes // chunk_bytes)
    devs = list(range(ndev)) if single else [local]
    mine = []                                     # (chunk, device index in devs)
    for d in range(ndev):
        part = part0 + d
        mine += [(c, d) for c in range(part * n_chunks // nparts, (part + 1) * n_chunks // nparts)]
    ctxs = [gpu.Context(dv) for dv in devs]
    staged = []

    def gen(cd):
        return cd, W.generate(2, chunk_bytes, seed=W.SEEDS[5] + cd[0])

    first_host = None
    with cf.ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 1) // max(1, world)))) as ex:
        for (c, d), (text, offs) in ex.map(gen, mine):
            if first_host is None and args.verify:
                k = int(np.searchsorted(offs, np.uint64(2 << 20), side="right")) - 1
                first_host = (text[:int(offs[k])].copy(), offs[:k + 1].copy())
            dv = torch.device("cuda", devs[d])
            staged.append((c, d, torch.from_numpy(text).to(dv), torch.from_numpy(offs.view(np.int64)).to(dv), len(offs) - 1,
                           int(offs[-1])))
    for dv in devs:
        torch.cuda.synchronize(dv)
    my_bytes = sum(x[5] for x in staged)
    my_stmts = sum(x[4] for x in staged)
    flags = gpu.F_DEVICE_INPUT | gpu.F_NO_HOST_COPY

    def one_pass(hashes=None):
        for c, d, dt, do, ns, nb in staged:
            b = ctxs[d].parse_batch_raw(dt.data_ptr(), do.data_ptr(), ns, flags, copy=False)
            if hashes is not None:
                hashes.append((c, b.device_hash()))

    # ---- device-resident: kernels only, outputs stay sharded (one context per device; devices one after the other in
    # the single-process mode, so that number is only meaningful per device there) ----
    streams = [torch.cuda.ExternalStream(cx.stream(), device=torch.device("cuda", dv)) for cx, dv in zip(ctxs, devs)]
    for _ in range(args.warmup):
        one_pass()
    launches = sum(ctxs[d].launches() for c, d, *_ in staged)
    sampler = ClockSampler(devs[0])
    sampler.start()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in devs]
    for (e0, e1), st in zip(ev, streams):
        e0.record(st)
    for _ in range(args.steps):
        one_pass()
    for (e0, e1), st in zip(ev, streams):
        e1.record(st)
    for dv in devs:
        torch.cuda.synchronize(dv)
    barrier()
    dev_ms = max(e0.elapsed_time(e1) for e0, e1 in ev) / args.steps
    clocks = sampler.stop()

    # ---- with the gather: the C dispatcher, outputs to pinned host memory (or GPU 0) inside the timed region ----
    gather = None
    if args.gather != "none":
        m = gpu.MultiContext(devs, args.e2e_workers)
        gflags = gpu.F_NO_TOKENS | gpu.F_WIRE_STMT | (gpu.MF_GATHER_DEVICE0 if args.gather == "device0" else 0)
        shards = [(d, dt.data_ptr(), do.data_ptr(), ns, gpu.F_DEVICE_INPUT, 0) for c, d, dt, do, ns, nb in staged]
        acc = {}
        lock = threading.Lock()

        def consume(ch):   # the consumer's read of a gathered chunk
            r = ch.raw
            with lock:
                acc["bytes"] = acc.get("bytes", 0) + 8 * r.n_stmt + 4 * r.n_node + 32 * r.n_err
                acc["stmts"] = acc.get("stmts", 0) + r.n_stmt
                if not ch.on_device and r.n_stmt:
                    acc["last"] = int(ch.batch.wstmt[-1] & 15) + int(ch.batch.pnode[-1] & 127)

        for _ in range(2):
            acc.clear()
            m.parse_shards(shards, consume, gflags)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            acc.clear()
            m.parse_shards(shards, consume, gflags)
        for dv in devs:
            torch.cuda.synchronize(dv)
        barrier()
        g_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        assert acc["stmts"] == my_stmts
        gather = {"ms": g_ms, "bytes": acc["bytes"]}
        m.close()

    # ---- checksum of all outputs (untimed) ----
    hs = []
    one_pass(hs)
    M64 = (1 << 64) - 1

    def mix(z):
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        return z ^ (z >> 31)

    total = 0
    for c, h in hs:
        total = (total + mix(h ^ ((c * 0x9E3779B97F4A7C15) & M64))) & M64
    verified = None
    if args.verify and first_host is not None:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import parity as P
        got = ctxs[0].parse_batch(first_host[0], first_host[1])
        bad = P.compare_with_oracle(got, first_host[0], first_host[1])
        verified = not bad
        if bad:
            print(f"rank {rank}: ORACLE MISMATCH: {bad[:3]}", file=sys.stderr, flush=True)

    def allred(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    dev_ms_max = allred(dev_ms, dist.ReduceOp.MAX)
    tot_bytes = allred(float(my_bytes), dist.ReduceOp.SUM)
    tot_stmts = allred(float(my_stmts), dist.ReduceOp.SUM)
    ok_all = allred(0.0 if verified is False else 1.0, dist.ReduceOp.MIN)
    g_ms_max = allred(gather["ms"], dist.ReduceOp.MAX) if gather else None
    g_bytes = allred(float(gather["bytes"]), dist.ReduceOp.SUM) if gather else None
    if world > 1:  # the checksum: a wrapping 64-bit sum over ranks
        t = torch.tensor([total - (1 << 64) if total >= (1 << 63) else total], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total = int(t.item()) & M64
    if rank == 0:
        peak, _ = peaks()
        line = {"metric": METRIC, "value": tot_bytes / (dev_ms_max * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": nparts,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "statements_per_s": tot_stmts / (dev_ms_max * 1e-3),
                "config": {"workload": WORKLOADS[5], "log_bytes": int(tot_bytes), "chunk_bytes": chunk_bytes,
                           "chunks": n_chunks, "statements": int(tot_stmts),
                           "processes": "one process, one dispatcher over all GPUs" if single else "one process per GPU (torchrun)",
                           "timed_region": "value: all kernels of every chunk, inputs resident in HBM, outputs left sharded on "
                                           "the GPUs; gather: see there",
                           "l2": "each chunk and its intermediates exceed the 126 MB L2; no flush needed"},
                "gpu_launches": launches * args.steps, "clocks": clocks,
                "output_hash": f"{total:016x}",
                "oracle_sample_ok": None if not args.verify else bool(ok_all)}
        if single and ndev > 1:
            line["value_note"] = "single-process mode drives the devices one after the other in the kernel-only region: `value` is per device; the gather region runs them concurrently"
        if gather:
            line["gather"] = {"value": tot_bytes / (g_ms_max * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": g_ms_max,
                              "statements_per_s": tot_stmts / (g_ms_max * 1e-3),
                              "to": "pinned host memory, each GPU over its own PCIe link" if args.gather == "host"
                                    else "GPU 0 memory over NVLink (cudaMemcpyPeerAsync)",
                              "gathered_bytes_per_step": int(g_bytes),
                              "gather_rate_gbs": g_bytes / (g_ms_max * 1e-3) / 1e9,
                              "api": f"nutdb_gpu_mctx_parse_shards, {args.e2e_workers} contexts per GPU, NUTDB_F_NO_TOKENS | NUTDB_F_WIRE_STMT; "
                                     "timed by the host clock between device synchronisations (several streams per GPU)"}
        print(json.dumps(line), flush=True)
    for cx in ctxs:
        cx.close()


def device_run(ctx, text, offs, local, steps, warmup, barrier):
    """Device-resident timing of one workload: input pre-staged in HBM, outputs left on the device.  CUDA events on the
    library's stream around `steps` calls, then a separate pass with events around every launch."""
    import torch
    from nutdb_b200 import gpu
    n_stmt = len(offs) - 1
    d_text = torch.from_numpy(text).cuda()
    d_offs = torch.from_numpy(offs.view(np.int64)).cuda()
    torch.cuda.synchronize()
    # (NUTDB_F_NO_TOKENS: the reference's entry returns the AST only -- mod.rs:27 -- so the token arrays stay an
    # intermediate in the lexer's own segmented layout instead of being compacted for a caller)
    dev_flags = gpu.F_DEVICE_INPUT | gpu.F_NO_HOST_COPY | gpu.F_NO_TOKENS

    def step_device():
        return ctx.parse_batch_raw(d_text.data_ptr(), d_offs.data_ptr(), n_stmt, dev_flags, copy=False)

    lib_stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    for _ in range(warmup):
        b = step_device()
    out = {"n_tok": int(b.n_tok), "n_node": int(b.n_node), "n_err": int(b.n_err), "launches": ctx.launches(),
           "n_slow": int(ctx.slow_statements()), "n_punt": int(ctx.exact_lexed_statements())}
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(lib_stream)
    t0 = time.perf_counter()
    for _ in range(steps):
        step_device()
    e1.record(lib_stream)
    barrier()
    out["wall_ms"] = (time.perf_counter() - t0) * 1e3 / steps
    out["dev_ms"] = e0.elapsed_time(e1) / steps
    # ---- per-kernel timing (events around every launch, separate pass) ----
    ctx.set_profiling(True)
    acc = {}
    reps = max(2, min(steps, 3))
    for _ in range(reps):
        step_device()
        for name, ms in ctx.kernel_timing():
            acc.setdefault(name, []).append(ms)
    ctx.set_profiling(False)
    out["kernel_ms"] = {k: sum(v) / reps for k, v in acc.items()}
    del d_text, d_offs
    return out


def roofline_numbers(text, offs, dev, T_pulled, peak):
    """Algorithmic bytes (DESIGN.md "Roofline accounting") of the pipeline and of its dominant kernel."""
    n_in, n_stmt = int(offs[-1]), len(offs) - 1
    cnt = alg_counts(text, offs)
    scale = n_in / cnt["bytes"]
    T = T_pulled if T_pulled is not None else int(cnt["T"] * scale)
    M = int(cnt["M"] * scale)
    b_alg = n_in + 9 * T + 16 * M + 16 * n_stmt
    parse_b = 9 * T + 16 * M + 16 * n_stmt
    alg = {"k_lex_A": n_in, "k_lex_B": n_in, "k_lex_C": n_in, "k_lex_D": n_in + 9 * T,
           "k_lex4": n_in + 9 * T, "k_lex3": n_in + 9 * T,
           "k_parse_fast": parse_b, "k_parse": parse_b, "k_parse_retry": parse_b, "k_parse_coop": parse_b,
           "k_finalize": 16 * M + 16 * n_stmt}
    kernel_ms = dev["kernel_ms"]
    dom = max((k for k in kernel_ms if k in alg), key=lambda k: kernel_ms[k])
    ksum = sum(kernel_ms.values())
    return {"dominant": dom, "alg": alg, "ksum": ksum, "achieved": alg[dom] / (kernel_ms[dom] * 1e-3) / 1e9,
            "pipeline": {"alg_bytes": b_alg, "B_alg_over_N_in": b_alg / n_in, "T": T, "M": M, "S": n_stmt,
                         "achieved": b_alg / (dev["dev_ms"] * 1e-3) / 1e9,
                         "frac": b_alg / (dev["dev_ms"] * 1e-3) / 1e9 / peak,
                         "sum_kernel_ms": ksum, "kernels_ms": {k: round(v, 4) for k, v in kernel_ms.items()}}}


def extra_config(ctx, config, nbytes, local, barrier, peak):
    """One of the other BASELINE.json configurations on this GPU: device-resident value, pipeline roofline fraction,
    how many statements took the slower paths, and whether a sample of the output equals the oracle's."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity as P
    text, offs = make_workload(config, nbytes)
    n_in, n_stmt = int(offs[-1]), len(offs) - 1
    dev = device_run(ctx, text, offs, local, 5, 3, barrier)
    rf = roofline_numbers(text, offs, dev, None, peak)
    k = max(1, int(np.searchsorted(offs, np.uint64(2 << 20), side="right")) - 1)
    sample_text = np.concatenate([text[:int(offs[k])], np.zeros(64, np.uint8)])
    got = ctx.parse_batch(sample_text, offs[:k + 1])
    bad = P.compare_with_oracle(got, sample_text, offs[:k + 1])
    if bad:
        print(f"config {config}: ORACLE MISMATCH: {bad[:3]}", file=sys.stderr, flush=True)
    return {"workload": WORKLOADS[config], "bytes": n_in, "statements": n_stmt, "steps": 5, "warmup": 3,
            "value": n_in / (dev["dev_ms"] * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": dev["dev_ms"],
            "statements_per_s": n_stmt / (dev["dev_ms"] * 1e-3),
            "tokens": dev["n_tok"], "nodes": dev["n_node"], "error_statements": dev["n_err"],
            "automaton_share": dev["n_slow"] / max(1, n_stmt), "exact_lexed_share": dev["n_punt"] / max(1, n_stmt),
            "roofline_pipeline_frac": rf["pipeline"]["frac"], "dominant_kernel": rf["dominant"],
            "kernels_ms": rf["pipeline"]["kernels_ms"],
            "oracle_sample": {"statements": k, "bytes": int(offs[k]), "bit_exact": not bad}}


def single_call_latency(ctx):
    """The reference's own bench shape (benches/parser_bench.rs:5-6,48: one statement per Parser::parse call, the 31-byte
    and the 1,131-byte string): wall time of one nutdb_gpu_parse call (host text in, host AST out) next to the oracle
    port on one host thread."""
    import ctypes as C
    from nutdb_b200 import gpu, workload as W
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    stmts = W.corpus_statements()
    out = {}
    for name, s in (("short sql", stmts[-1]), ("long sql", stmts[-2])):
        raw = gpu.NutdbBatch()
        L = gpu.lib()
        for _ in range(20):
            L.nutdb_gpu_parse(ctx._h, s, len(s), C.byref(raw))
        ts = []
        for _ in range(200):
            t0 = time.perf_counter()
            rc = L.nutdb_gpu_parse(ctx._h, s, len(s), C.byref(raw))
            ts.append(time.perf_counter() - t0)
        assert rc == 0 and raw.n_stmt == 1
        reps = 20000
        offs = np.arange(reps + 1, dtype=np.uint64) * np.uint64(len(s))
        text = np.frombuffer(s * reps + b"\0" * 64, np.uint8)
        t, _, _ = O.bench(text, offs, 1, reps=3)
        out[name] = {"bytes": len(s), "gpu_call_us_median": statistics.median(ts) * 1e6, "gpu_call_us_min": min(ts) * 1e6,
                     "gpu_launches_per_call": ctx.launches(), "cpu_port_single_thread_ns": t / reps * 1e9}
    out["note"] = ("one nutdb_gpu_parse call per statement is launch-latency bound: the GPU path is a batch engine; "
                   "cpu_port = C++ oracle port, not the Rust parser")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--log-bytes", type=int, default=64 << 30, help="config 5: total bytes of the statement log")
    ap.add_argument("--gather", default="host", choices=["host", "device0", "none"], help="config 5: where the outputs are gathered")
    ap.add_argument("--single-process", action="store_true", help="config 5: one process drives all --gpus devices through one dispatcher")
    ap.add_argument("--verify", action="store_true", help="config 5: check a sample of every rank's first chunk against the oracle")
    ap.add_argument("--bytes", type=int, default=1 << 30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the 256 MiB runs of configs 1, 3, 4 and the latency probe")
    ap.add_argument("--extra-bytes", type=int, default=256 << 20)
    ap.add_argument("--e2e-chunk", type=int, default=64 << 20, help="bytes of SQL per pipelined chunk on the host path")
    ap.add_argument("--e2e-workers", type=int, default=4, help="contexts (threads) the host path pipelines chunks over")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from nutdb_b200 import gpu
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.config == 5:
        run_log(args, rank, world, local, barrier)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- workload: rank r parses shard r of the statement log ----
    text, offs = make_workload(args.config, args.bytes, seed_offset=rank)
    n_in, n_stmt = int(offs[-1]), len(offs) - 1
    ctx = gpu.Context(local)
    sampler = ClockSampler(local)
    sampler.start()
    dev = device_run(ctx, text, offs, local, args.steps, args.warmup, barrier)
    clocks = sampler.stop()
    dev_ms, wall_ms, kernel_ms = dev["dev_ms"], dev["wall_ms"], dev["kernel_ms"]
    T_pulled = None

    # ---- end to end through the C ABI with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        h_text = torch.from_numpy(text).pin_memory()
        h_offs = torch.from_numpy(offs.view(np.int64)).pin_memory()
        # the reference API never exposes tokens (mod.rs:27 returns Statement only): no token arrays, and the statement
        # records in their 8-byte wire form (status, node count, tokens pulled)
        flags = gpu.F_NO_TOKENS
        eflags = gpu.F_NO_TOKENS | gpu.F_WIRE_STMT | gpu.F_OFFSETS32   # (the batch ends below 4 GiB: 32-bit offsets go up)

        # host-resident batch -> chunked, pipelined through --e2e-workers contexts (nutdb_b200.stream):
        # uploads, kernels and downloads of different chunks overlap
        from nutdb_b200 import stream
        sp = stream.StreamParser(local, workers=args.e2e_workers)
        h_offs32 = torch.from_numpy(offs.astype(np.uint32).view(np.int32)).pin_memory()
        h_text_np, h_offs_np = h_text.numpy(), h_offs32.numpy().view(np.uint32)
        acc = {}

        def consume(first, bb):   # the caller's read of the step's result (arrays are in pinned host memory now)
            acc["n_node"] = acc.get("n_node", 0) + int(bb.n_node)
            acc["n_err"] = acc.get("n_err", 0) + int(bb.n_err)
            acc["last_status"] = int(bb.wstmt[-1] & 15) if bb.n_stmt else 0
            acc["last_kind"] = int(bb.pnode[-1] & 127) if bb.n_node else 0

        def step_host():
            acc.clear()
            sp.parse(h_text_np, h_offs_np, consume, chunk_bytes=args.e2e_chunk, flags=eflags)

        # untimed: exact T (tokens the reference pulls) and the ok count from one plain host-buffer call
        hb0 = ctx.parse_batch_raw(h_text.data_ptr(), h_offs.data_ptr(), n_stmt, flags, copy=False)
        T_pulled = int(hb0.stmt["tok_used"].astype(np.int64).sum())
        del hb0
        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        sp.close()
        h2d = n_in + 4 * (n_stmt + 1)
        d2h = 8 * n_stmt + 4 * int(acc["n_node"]) + 32 * int(acc["n_err"])   # wire statement records (64-bit), wire nodes (32-bit), NutdbError
        e2e = {"ms": e2e_ms, "h2d": h2d, "d2h": d2h}
        del h_text, h_offs, h_offs32

    # ---- max over ranks ----
    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    dev_ms_max, wall_ms_max = allmax(dev_ms), allmax(wall_ms)
    tot_bytes, tot_stmts = allsum(float(n_in)), allsum(float(n_stmt))
    e2e_ms_max = allmax(e2e["ms"]) if e2e else None
    tot_h2d = allsum(float(e2e["h2d"])) if e2e else 0
    tot_d2h = allsum(float(e2e["d2h"])) if e2e else 0

    if rank == 0:
        peak, peak_src = peaks()
        rf = roofline_numbers(text, offs, dev, T_pulled, peak)
        dom = rf["dominant"]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):  # dram bytes per input byte from the committed ncu --set full capture, scaled to this launch
            try:
                ent = json.load(open(tp)).get(dom)
                traffic = int(ent["dram_bytes_per_input_byte"] * n_in) if ent else None
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": tot_bytes / (dev_ms_max * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max, "wall_ms_per_step": wall_ms_max,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "statements_per_s": tot_stmts / (dev_ms_max * 1e-3),
            "config": workload_config(args.config, n_in, n_stmt),
            "counts": {"tokens_per_gpu": dev["n_tok"], "nodes_per_gpu": dev["n_node"], "error_statements_per_gpu": dev["n_err"],
                       "automaton_statements_per_gpu": dev["n_slow"], "exact_lexed_statements_per_gpu": dev["n_punt"]},
            "gpu_launches": dev["launches"] * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": rf["achieved"], "peak": peak, "unit": "GB/s",
                         "frac": rf["achieved"] / peak, "traffic": traffic, "peak_source": peak_src,
                         "alg_bytes_per_launch": rf["alg"][dom], "kernel_ms": kernel_ms[dom],
                         "traffic_source": "profiles/traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum per input byte x bytes of this launch)",
                         "kernel_share_of_step": kernel_ms[dom] / rf["ksum"]},
            "roofline_pipeline": rf["pipeline"],
        }
        if e2e:
            line["e2e"] = {"value": tot_bytes / (e2e_ms_max * 1e-3) / 1e9, "unit": "GB/s",
                           "h2d_bytes_per_step": int(tot_h2d), "d2h_bytes_per_step": int(tot_d2h),
                           "ms_per_step": e2e_ms_max, "statements_per_s": tot_stmts / (e2e_ms_max * 1e-3),
                           "api": "nutdb_b200.stream.StreamParser: nutdb_gpu_parse_batch(pinned host text, pinned 32-bit host offsets, "
                                  f"NUTDB_F_NO_TOKENS | NUTDB_F_WIRE_STMT | NUTDB_F_OFFSETS32) per chunk on {args.e2e_workers} contexts -> pinned host wire-stmt (8 B) / wire-node "
                                  "(32-bit words) / err arrays",
                           "chunk_bytes": args.e2e_chunk}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(text, offs)
        del text, offs
        if world == 1 and not args.no_extra_configs:
            # the other single-GPU configurations of BASELINE.json, 256 MiB each (device-resident value, share of the
            # slower paths, parity of a sample with the oracle), and the reference's own criterion bench shape
            line["configs"] = {}
            for c in (1, 3, 4):
                if c != args.config:
                    line["configs"][str(c)] = extra_config(ctx, c, args.extra_bytes, local, barrier, peak)
            line["latency"] = single_call_latency(ctx)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
