import sys
sys.path.insert(0, '/root/repo')
from tools import gdbsyn
from genomicsdb_b200 import capi
ctx = capi.Context(0)
d = gdbsyn.generate(**dict(gdbsyn.CONFIGS['config4'], site_lo=0, site_hi=1250), with_fixed=True)
arr = gdbsyn.to_gpu_array(ctx, d, plain=True)
lo, hi = d["col_range"]
def run():
    sc = capi.Scan(arr, lo, hi, flags=capi.SCAN_DEVICE_OUTPUT | capi.SCAN_BCF_FORMAT, dp_fields=gdbsyn.PLAIN_DP_RULE)
    n = 0
    while True:
        b = sc.next_batch(0, 8 << 30, raw=True)
        if b.n_sites == 0: break
        n += b.n_sites
    sc.close()
    return n
run(); run()
ctx.set_profiling(True); ctx.reset_kernel_stats()
n = run()
for k in ("k_fmt_range", "k_fmt_pack", "k_plain_scan", "k_plain_pack", "k_info_collect", "k_info_reduce", "k_remap", "k_copy_tile", "k_live_walk", "k_site_scan", "k_union", "k_"):
    ms, c = ctx.kernel_stats(k)
    print("%-16s %8.3f ms %4d launches" % (k, ms, c))
print("sites", n)
