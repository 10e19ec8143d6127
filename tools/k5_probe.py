#!/usr/bin/env python3
"""Development probe for the K5 kernels: (1) byte-for-byte A/B of the engine variants selected by GDBGPU_K5 / GDBGPU_K34
on a mid-size cohort (host output, every buffer compared), (2) per-kernel CUDA-event times of the resident pass over the
config-#4 shard.  Usage: python tools/k5_probe.py [--samples N] [--sites N] [--variants tile,tma] [--no-ab]"""
import argparse
import hashlib
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402

from genomicsdb_b200 import capi  # noqa: E402
from tools import gdbsyn  # noqa: E402

KERNELS = ["k_verify_alleles", "k_sig_perm", "k_nt_blocks", "k_remap_mix", "k_remap_tma", "k_remap_tile", "k_copy_tile", "k_remap_nt", "k_remap_generic", "k_live_walk", "k_site_scan", "k_union", "k_"]


_SET = set()


def set_variant(v):
    for k in _SET:
        os.environ.pop(k, None)
    _SET.clear()
    for kv in v.split("+"):
        k, _, val = kv.partition("=")
        os.environ[k] = val
        _SET.add(k)


def digest_pass(arr, q_lo, q_hi, flags, batch_bytes):
    h = hashlib.sha256()
    sc = capi.Scan(arr, q_lo, q_hi, flags=flags)
    sites = 0
    while True:
        b = sc.next_batch(0, batch_bytes)
        if b is None:
            break
        sites += b.n_sites
        for a in (b.site_begin, b.site_end, b.site_flags, b.n_alleles, b.n_live, b.live):
            h.update(np.ascontiguousarray(a).tobytes())
        for s in range(b.n_sites):
            h.update(b"|".join(b.alleles(s)))
        for f in b.fields:
            for k in ("base", "stride", "len", "data"):
                h.update(np.ascontiguousarray(f[k]).tobytes())
    sc.close()
    return sites, h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=100000)
    ap.add_argument("--sites", type=int, default=1250)
    ap.add_argument("--ab-samples", type=int, default=20000)
    ap.add_argument("--ab-sites", type=int, default=60)
    ap.add_argument("--variants", default="GDBGPU_K5=tile,GDBGPU_K5=mix")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--batch-bytes", type=int, default=8 << 30)
    ap.add_argument("--no-ab", action="store_true")
    ap.add_argument("--no-time", action="store_true")
    args = ap.parse_args()
    variants = args.variants.split(",")
    ctx = capi.Context(0)
    out = dict(variants=variants)
    if not args.no_ab:
        d = gdbsyn.generate(n_samples=args.ab_samples, n_sites=10000, spacing=1000, p_var=0.1, a_pool=8, site_lo=100, site_hi=100 + args.ab_sites)
        arr = gdbsyn.to_gpu_array(ctx, d)
        res = {}
        for v in variants:
            set_variant(v)
            res[v] = digest_pass(arr, d["col_range"][0], d["col_range"][1], 0, 1 << 30)
            res[v + " tiny batches"] = digest_pass(arr, d["col_range"][0] + 7000, d["col_range"][1] - 9000, 0, 50 << 20)
        arr.close()
        out["ab"] = {k: dict(sites=s, sha256=h[:16]) for k, (s, h) in res.items()}
        digs = {(k.endswith("tiny batches"), h) for k, (s, h) in res.items()}
        out["ab_identical"] = len(digs) == 2
    if not args.no_time:
        d = gdbsyn.generate(n_samples=args.samples, n_sites=10000, spacing=1000, p_var=0.1, a_pool=8, site_lo=0, site_hi=args.sites)
        arr = gdbsyn.to_gpu_array(ctx, d)
        q_lo, q_hi = d["col_range"]

        def one_pass():
            sc = capi.Scan(arr, q_lo, q_hi, flags=capi.SCAN_DEVICE_OUTPUT)
            n = 0
            while True:
                b = sc.next_batch(0, args.batch_bytes, raw=True)
                if b.n_sites == 0:
                    break
                n += b.n_sites
            sc.close()
            return n
        out["timing"] = {}
        for v in variants:
            set_variant(v)
            for _ in range(3):
                one_pass()
            ctx.synchronize()
            ctx.set_profiling(False)
            t0 = time.perf_counter()
            n = 0
            for _ in range(args.steps):
                n += one_pass()
            ctx.synchronize()
            wall = (time.perf_counter() - t0) / args.steps
            ctx.set_profiling(True)
            ctx.reset_kernel_stats()
            for _ in range(args.steps):
                one_pass()
            ctx.synchronize()
            ks = {k: round(ctx.kernel_stats(k)[0] / args.steps, 3) for k in KERNELS}
            ctx.set_profiling(False)
            out["timing"][v] = dict(sites_per_step=n // args.steps, wall_ms_per_step=round(wall * 1e3, 3), sites_per_s=round(n / args.steps / wall), kernel_ms_per_step=ks)
        arr.close()
    ctx.close()
    print(json.dumps(out, indent=1))
    if not args.no_ab and not out["ab_identical"]:
        sys.exit("A/B MISMATCH")


if __name__ == "__main__":
    main()
