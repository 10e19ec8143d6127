import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np
from tools import gdbsyn
from genomicsdb_b200 import capi
which = sys.argv[1] if len(sys.argv)>1 else 'config2'
kw = dict(gdbsyn.CONFIGS['config2']) if which=='config2' else dict(gdbsyn.CONFIGS['config4'], site_lo=0, site_hi=int(sys.argv[2]) if len(sys.argv)>2 else 156)
ctx=capi.Context(0)
d=gdbsyn.generate(**kw)
arr=gdbsyn.to_gpu_array(ctx,d)
for it in range(int(sys.argv[3]) if len(sys.argv)>3 else 1):
    t=time.time()
    sc=capi.Scan(arr, flags=capi.SCAN_DEVICE_OUTPUT)
    ns=0
    while True:
        b=sc.next_batch(0, 8<<30, raw=True)
        if b.n_sites==0: break
        ns+=b.n_sites
    print(which, ns, time.time()-t)
    sc.close()
arr.close()
