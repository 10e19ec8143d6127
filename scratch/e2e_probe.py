import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from tools import gdbsyn
from genomicsdb_b200 import capi
n_sites = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
bb = int(float(sys.argv[2]) * (1 << 30)) if len(sys.argv) > 2 else 2 << 30
# raw pinned D2H / H2D bandwidth
x = torch.empty(1 << 30, dtype=torch.uint8, device='cuda'); h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
for name, a, b in (("d2h", h, x), ("h2d", x, h)):
    a.copy_(b); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(4): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); print(name, "GB/s", 4 * 1.0737 / (time.perf_counter() - t))
del x, h
ctx = capi.Context(0)
d = gdbsyn.generate(**dict(gdbsyn.CONFIGS['config4'], site_lo=0, site_hi=n_sites))
def pin(a):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
keep = []
dp = dict(d)
for k in ("coords", "end"):
    t = pin(d[k]); keep.append(t); dp[k] = t.numpy()
for k in ("ref", "alt", "gt", "pl", "ad"):
    a, b = pin(d[k][0]), pin(d[k][1]); keep += [a, b]; dp[k] = (a.numpy(), b.numpy())
lo, hi = d["col_range"]
for flags, name in ((0, "int32"), (capi.SCAN_BCF_FORMAT | capi.SCAN_PRODUCE_GT, "bcf"), (capi.SCAN_NO_PREFETCH, "int32-noprefetch"), (capi.SCAN_BCF_FORMAT | capi.SCAN_NO_PREFETCH, "bcf-noprefetch")):
    for it in range(3):
        t0 = time.perf_counter()
        arr = gdbsyn.to_gpu_array(ctx, dp)
        t1 = time.perf_counter()
        sc = capi.Scan(arr, lo, hi, flags=flags)
        t2 = time.perf_counter()
        ns = nb = 0; by = 0
        while True:
            b = sc.next_batch(0, bb, raw=True)
            if b.n_sites == 0: break
            ns += b.n_sites; nb += 1; by += b.fmt_total_bytes if flags & capi.SCAN_BCF_FORMAT else b.out_bytes
        t3 = time.perf_counter()
        sc.close(); arr.close()
        t4 = time.perf_counter()
        print("%-18s it%d stage %.3f begin %.3f batches %.3f (%d sites, %d batches, %.2f GB) close %.3f total %.3f -> %.0f sites/s" % (name, it, t1 - t0, t2 - t1, t3 - t2, ns, nb, by / 1e9, t4 - t3, t4 - t0, ns / (t4 - t0)))
ctx.close()
