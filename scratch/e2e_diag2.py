import sys, time, threading
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
from tools import gdbsyn
from genomicsdb_b200 import capi
ctx, ctx2 = capi.Context(0), capi.Context(0)
ctxs = [ctx, ctx2]
d = gdbsyn.generate(**dict(gdbsyn.CONFIGS['config4'], site_lo=0, site_hi=1250))
keep = []
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); keep.append(t); return t.numpy()
dp = dict(d)
for k in ("coords", "end"): dp[k] = pin(d[k])
for k in ("ref", "alt", "gt", "pl", "ad"): dp[k] = (pin(d[k][0]), pin(d[k][1]))
lo, hi = d["col_range"]
flags = capi.SCAN_BCF_FORMAT | capi.SCAN_PRODUCE_GT
T0 = time.perf_counter()
log = []
def stage(k):
    t = time.perf_counter()
    a = gdbsyn.to_gpu_array(ctxs[k % 2], dp)
    log.append(("stage", k, t - T0, time.perf_counter() - T0))
    return a
def scan(a, k):
    t = time.perf_counter()
    sc = capi.Scan(a, lo, hi, flags=flags)
    ts = []
    while True:
        ta = time.perf_counter()
        b = sc.next_batch(0, 2 << 30, raw=True)
        ts.append(time.perf_counter() - ta)
        if b.n_sites == 0: break
    sc.close(); a.close()
    log.append(("scan", k, t - T0, time.perf_counter() - T0, max(ts), sorted(ts)[len(ts) // 2]))
for k in range(2): scan(stage(k), -1)
for mode in ("pipelined", "serial"):
    log.clear()
    steps = 4
    t0 = time.perf_counter()
    if mode == "pipelined":
        with ThreadPoolExecutor(1) as ex:
            fut = ex.submit(stage, 0)
            for k in range(steps):
                a = fut.result()
                if k + 1 < steps: fut = ex.submit(stage, k + 1)
                scan(a, k)
    else:
        for k in range(steps): scan(stage(k), k)
    ctx.synchronize(); ctx2.synchronize()
    dt = time.perf_counter() - t0
    print(mode, "ms per step %.1f" % (dt * 1e3 / steps))
    for e in sorted(log, key=lambda e: e[2]):
        print("   %-5s step %d  start %.1f  end %.1f  dur %.1f ms" % (e[0], e[1], (e[2] - (t0 - T0)) * 1e3, (e[3] - (t0 - T0)) * 1e3, (e[3] - e[2]) * 1e3) + ("  slowest batch %.1f median %.1f" % (e[4] * 1e3, e[5] * 1e3) if e[0] == "scan" else ""))
