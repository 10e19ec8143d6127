import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from tools import gdbsyn
from genomicsdb_b200 import capi
ctx=capi.Context(0)
for name, kw in [('config2', dict(gdbsyn.CONFIGS['config2'])), ('config4/64', dict(gdbsyn.CONFIGS['config4'], site_lo=0, site_hi=156))]:
    t=time.time(); d=gdbsyn.generate(**kw); tg=time.time()-t
    t=time.time(); arr=gdbsyn.to_gpu_array(ctx,d); topen=time.time()-t
    info=arr.info()
    for it in range(3):
        l0=ctx.launches
        t=time.time()
        sc=capi.Scan(arr, flags=capi.SCAN_DEVICE_OUTPUT)
        tprep=time.time()-t
        ns=0; ob=0; nb=0
        while True:
            b=sc.next_batch(0, 8<<30, raw=True)
            if b.n_sites==0: break
            ns+=b.n_sites; ob+=b.out_bytes; nb+=1
        tt=time.time()-t
        sc.close()
        print(name, 'cells',info.n_cells,'gen %.2fs open %.3fs prep %.4fs total %.4fs sites %d batches %d out %.1f MB -> %.1f GB/s, %.0f sites/s launches %d'%(tg,topen,tprep,tt,ns,nb,ob/1e6,ob/tt/1e9,ns/tt, ctx.launches-l0))
    arr.close()
