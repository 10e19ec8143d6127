#!/usr/bin/env python3
"""BASELINE.json config #5: scan_and_operate latency over random 1 Mb query intervals.

One GPU holds one column shard of the config-#4 cohort (100 k samples; 1.25 M positions per shard).  For each of N random
intervals [start, start + 1 Mb) inside the shard (seed 4321) it times, on the host clock around the C-ABI calls,
    gdbgpu_scan_begin (left sweep, breakpoints, deletions pre-pass) -> gdbgpu_scan_next_batch until empty -> gdbgpu_scan_end
with the remapped output left in HBM (GDBGPU_SCAN_DEVICE_OUTPUT), and prints one JSON line with p50 / p95 / p99 / max.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--intervals", type=int, default=1000)
    ap.add_argument("--length", type=int, default=1000000)
    ap.add_argument("--samples", type=int, default=100000)
    ap.add_argument("--sites", type=int, default=1250)
    ap.add_argument("--batch-bytes", type=int, default=8 << 30)
    ap.add_argument("--bcf", action="store_true", help="GDBGPU_SCAN_BCF_FORMAT: also collate + BCF-encode on the device")
    ap.add_argument("--records", action="store_true", help="--bcf plus the plain FORMAT fields (GQ, MIN_DP) and the INFO/DP rule: "
                                                          "everything BroadCombinedGVCFOperator needs for a record of this cohort")
    args = ap.parse_args()
    from genomicsdb_b200 import capi
    from tools import gdbsyn
    cfg = dict(gdbsyn.CONFIGS["config4"], n_samples=args.samples)
    if args.records:
        args.bcf = True
    d = gdbsyn.generate(**cfg, site_lo=0, site_hi=args.sites, with_fixed=args.records)
    lo, hi = d["col_range"]
    ctx = capi.Context(0)
    arr = gdbsyn.to_gpu_array(ctx, d, plain=args.records)
    rng = np.random.RandomState(4321)
    span = max(1, hi - lo + 1 - args.length)
    starts = lo + rng.randint(0, span, size=args.intervals)
    flags = capi.SCAN_DEVICE_OUTPUT | (capi.SCAN_BCF_FORMAT if args.bcf else 0)
    lat, sites = [], []

    def one(a, b):
        t0 = time.perf_counter()
        sc = capi.Scan(arr, int(a), int(b), flags=flags, dp_fields=gdbsyn.PLAIN_DP_RULE if args.records else (0, 0, 0))
        n = 0
        while True:
            bt = sc.next_batch(0, args.batch_bytes, raw=True)
            if bt.n_sites == 0:
                break
            n += bt.n_sites
        sc.close()
        ctx.synchronize()
        return time.perf_counter() - t0, n

    for a in starts[:5]:
        one(a, a + args.length - 1)
    t_all = time.perf_counter()
    for a in starts:
        t, n = one(a, a + args.length - 1)
        lat.append(t)
        sites.append(n)
    t_all = time.perf_counter() - t_all
    lat_ms = np.array(lat) * 1e3
    out = dict(config="BASELINE config #5: %d random %d-position intervals over a %d-sample x %d-candidate-site shard, one B200, output left in HBM%s"
               % (args.intervals, args.length, args.samples, args.sites, (", BCF records (FORMAT blocks of GT/AD/PL/GQ/MIN_DP + INFO/DP)" if args.records else ", BCF-typed FORMAT blocks") if args.bcf else ""),
               intervals=args.intervals, sites_per_interval_mean=float(np.mean(sites)),
               latency_ms=dict(p50=float(np.percentile(lat_ms, 50)), p95=float(np.percentile(lat_ms, 95)), p99=float(np.percentile(lat_ms, 99)),
                               max=float(lat_ms.max()), mean=float(lat_ms.mean()), min=float(lat_ms.min())),
               sites_per_sec=float(np.sum(sites) / t_all), wall_s=t_all)
    print(json.dumps(out))
    arr.close()
    ctx.close()


if __name__ == "__main__":
    main()
