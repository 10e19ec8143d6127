import sys, time
sys.path.insert(0,'/root/repo')
from tools import gdbsyn
from genomicsdb_b200 import capi
ctx=capi.Context(0)
d=gdbsyn.generate(**dict(gdbsyn.CONFIGS['config4'], site_lo=0, site_hi=300), with_fixed=True)
arr=gdbsyn.to_gpu_array(ctx,d,plain=True)
lo,hi=d['col_range']
for it in range(2):
    t0=time.perf_counter()
    sc=capi.Scan(arr, lo, hi, flags=capi.SCAN_BCF_FORMAT, dp_fields=gdbsyn.PLAIN_DP_RULE)
    tb=[]
    while True:
        t=time.perf_counter()
        b=sc.next_batch(0,0,raw=True)
        tb.append((time.perf_counter()-t, b.n_sites, b.fmt_total_bytes))
        if b.n_sites==0: break
    sc.close()
    print("pass",it,"total %.3f s"%(time.perf_counter()-t0), [("%.3f"%x[0], x[1], round(x[2]/1e6)) for x in tb[:12]])
