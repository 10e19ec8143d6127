import sys
sys.path.insert(0,'/root/repo')
from tools import gdbsyn
from genomicsdb_b200 import capi
ctx=capi.Context(0)
d=gdbsyn.generate(**dict(gdbsyn.CONFIGS['config4'], site_lo=0, site_hi=int(sys.argv[1]) if len(sys.argv)>1 else 156))
arr=gdbsyn.to_gpu_array(ctx,d)
sc=capi.Scan(arr, flags=capi.SCAN_DEVICE_OUTPUT|capi.SCAN_BCF_FORMAT)
while True:
    b=sc.next_batch(0, 8<<30, raw=True)
    if b.n_sites==0: break
sc.close(); arr.close()
