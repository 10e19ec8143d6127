import sys
sys.path.insert(0, '/root/repo')
from genomicsdb_b200 import capi
from tools import gdbsyn
d = gdbsyn.generate(**dict(gdbsyn.CONFIGS['config4']), site_lo=0, site_hi=300)
lo, hi = d['col_range']
ctx = capi.Context(0)
arr = gdbsyn.to_gpu_array(ctx, d)
for a, b in ((lo, hi), (lo + (hi - lo) // 2 + 137, hi)):
    sc = capi.Scan(arr, a, b, flags=capi.SCAN_CELLS_ONLY)
    h = sc.cells_histogram(lo, hi, 1000)
    r = sc.cells_allele_counts()
    print(int(h.sum()), len(r))
    sc.close()
arr.close(); ctx.close()
