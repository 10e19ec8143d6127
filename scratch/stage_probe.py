import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from tools import gdbsyn
from genomicsdb_b200 import capi
ctx = capi.Context(0)
d = gdbsyn.generate(**dict(gdbsyn.CONFIGS['config4'], site_lo=0, site_hi=1250))
def pin(a):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
keep = []
dp = dict(d)
for k in ("coords", "end"):
    t = pin(d[k]); keep.append(t); dp[k] = t.numpy()
for k in ("ref", "alt", "gt", "pl", "ad"):
    a, b = pin(d[k][0]), pin(d[k][1]); keep += [a, b]; dp[k] = (a.numpy(), b.numpy())
lo, hi = d["col_range"]
def scan(arr, flags):
    sc = capi.Scan(arr, lo, hi, flags=flags)
    while True:
        b = sc.next_batch(0, 2 << 30, raw=True)
        if b.n_sites == 0: break
    sc.close()
for mode in ("none", "bcf", "int32", "dev"):
    for it in range(4):
        t0 = time.perf_counter(); arr = gdbsyn.to_gpu_array(ctx, dp); t1 = time.perf_counter()
        if mode == "bcf": scan(arr, capi.SCAN_BCF_FORMAT)
        elif mode == "int32": scan(arr, 0)
        elif mode == "dev": scan(arr, capi.SCAN_DEVICE_OUTPUT)
        t2 = time.perf_counter(); arr.close(); t3 = time.perf_counter()
        print("%-6s it%d stage %.1f ms scan %.1f ms close %.2f ms" % (mode, it, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3), flush=True)
ctx.close()
