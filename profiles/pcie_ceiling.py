"""Plain concurrent pinned-memory copies on N GPUs of one box: the host-side ceiling the gather (D2H) and the upload
(H2D) of the end-to-end path run under.  One thread per GPU and direction, 256 MiB copies for ~2 s.
usage: pcie_ceiling.py N  -> one JSON line"""
import json, sys, threading, time
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
SZ = 256 << 20
res = {}
def run(mode):
    out = [0.0] * n
    def work(d, direction):
        torch.cuda.set_device(d)
        dev = torch.empty(SZ, dtype=torch.uint8, device=f"cuda:{d}")
        host = torch.empty(SZ, dtype=torch.uint8).pin_memory()
        st = torch.cuda.Stream(device=d)
        with torch.cuda.stream(st):
            for _ in range(3):
                (host.copy_(dev, non_blocking=True) if direction == "d2h" else dev.copy_(host, non_blocking=True))
            st.synchronize()
            barrier.wait()
            t0 = time.perf_counter(); k = 0
            while time.perf_counter() - t0 < 2.0:
                for _ in range(4):
                    (host.copy_(dev, non_blocking=True) if direction == "d2h" else dev.copy_(host, non_blocking=True))
                st.synchronize(); k += 4
            dt = time.perf_counter() - t0
        rates[(d, direction)] = k * SZ / dt / 1e9
    dirs = ["d2h"] if mode == "d2h" else ["h2d"] if mode == "h2d" else ["d2h", "h2d"]
    rates = {}
    barrier = threading.Barrier(n * len(dirs))
    th = [threading.Thread(target=work, args=(d, x)) for d in range(n) for x in dirs]
    [t.start() for t in th]; [t.join() for t in th]
    return {x: round(sum(v for (d, y), v in rates.items() if y == x), 1) for x in dirs}
for mode in ("d2h", "h2d", "both"):
    res[mode] = run(mode)
print(json.dumps({"n_gpus": n, "copy_bytes": SZ, "aggregate_gbs": res, "note": "pinned host memory, one stream per GPU and direction, host clock"}))
