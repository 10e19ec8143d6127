import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from tools import gdbsyn
from genomicsdb_b200 import capi
ctx = capi.Context(0)
d = gdbsyn.generate(**dict(gdbsyn.CONFIGS['config4'], site_lo=0, site_hi=1250))
keep = []
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); keep.append(t); return t.numpy()
dp = dict(d)
for k in ("coords", "end"): dp[k] = pin(d[k])
for k in ("ref", "alt", "gt", "pl", "ad"): dp[k] = (pin(d[k][0]), pin(d[k][1]))
lo, hi = d["col_range"]
arr = gdbsyn.to_gpu_array(ctx, dp)
for bb_gib in (2, 8, 1):
    bb = bb_gib << 30
    for flags, name in ((capi.SCAN_BCF_FORMAT | capi.SCAN_PRODUCE_GT, "bcf"), (0, "int32")):
        for it in range(3):
            t0 = time.perf_counter()
            sc = capi.Scan(arr, lo, hi, flags=flags)
            t1 = time.perf_counter()
            ts = []; by = 0; ns = 0
            while True:
                ta = time.perf_counter()
                b = sc.next_batch(0, bb, raw=True)
                ts.append(time.perf_counter() - ta)
                if b.n_sites == 0: break
                ns += b.n_sites
                by += b.fmt_total_bytes if flags & capi.SCAN_BCF_FORMAT else (b.out_bytes + b.n_sites * b.n_query_rows * 10)
            t2 = time.perf_counter()
            sc.close()
            t3 = time.perf_counter()
            if it == 2:
                print("%-6s batch %d GiB: begin %.1f ms, %d batches: first %.1f ms, median %.1f ms, last %.1f ms, sum %.1f ms, close %.1f ms | %.2f GB -> %.1f GB/s over the batches, %.0f sites/s (array resident)" %
                      (name, bb_gib, (t1 - t0) * 1e3, len(ts) - 1, ts[0] * 1e3, sorted(ts[:-1])[len(ts[:-1]) // 2] * 1e3, ts[-2] * 1e3, sum(ts) * 1e3, (t3 - t2) * 1e3, by / 1e9, by / 1e9 / sum(ts), ns / (t3 - t0)))
arr.close(); ctx.close()
