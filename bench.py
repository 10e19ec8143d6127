#!/usr/bin/env python3
"""bench.py -- combined variant sites/s of the B200 variant-combine hot path (contract: see the task statement).

A "step" is one pass of the hot path (interval sweep -> allele union -> GT/AD/PL remap) over this rank's column shard of
the GDBSYN-1 cohort named in BASELINE.json (config #4: 100 k samples x 10 M positions, 10 k candidate sites, max 8 ALT
alleles per site).  One rank holds 1/8 of the columns (1 250 candidate sites = 1.25 M positions); 8 ranks = config #4.
  value : sites/s with the shard already resident in HBM and the output left in HBM (CUDA events on the engine's stream)
  e2e   : the same metric through the C ABI with HOST buffers: gdbgpu_array_from_columns (H2D of the fragment columns,
          pinned) -> scan -> every batch copied back to pinned host memory (D2H) -> close.  The headline e2e asks for the
          output a Broad-GVCF consumer takes (GDBGPU_SCAN_BCF_FORMAT: remap + FORMAT collation + BCF typed encoding on
          the device, the typed blocks copied back); e2e_int32 is the same with the raw 32-bit remapped vectors copied back
  roofline : the remap kernel (k_remap), algorithmic bytes / its CUDA-event time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the UNMODIFIED reference (oracle/_ref/bin/gdb_ref_dump --mode hot: scan_and_operate + handle_deletions
          + GA4GHOperator::operate, no VCF serialisation) on all host cores, on a bounded sample of the same cohort.
--impl reference prints the reference arm's line (same metric/config, CPU only).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

COHORT = dict(n_samples=100000, n_sites=10000, spacing=1000, p_var=0.1, a_pool=8)
SITES_PER_GPU = 1250
# --config: which BASELINE.json configuration the main measurement runs on.  "4shard" (default) is config #4 the way
# it is quoted: one 1/8 column shard of the 100 k-sample cohort per GPU, 8 ranks = the whole cohort (weak scaling).
WORKLOADS = {
    "4shard": dict(cohort=COHORT, sites_per_gpu=SITES_PER_GPU, scaling="weak"),
    "2": dict(cohort=dict(n_samples=1000, n_sites=10000, spacing=100, p_var=0.1, a_pool=3), sites_per_gpu=None, scaling="weak"),       # whole, replicated
    "3": dict(cohort=dict(n_samples=10000, n_sites=10000, spacing=100, p_var=0.1, a_pool=3), sites_per_gpu=None, scaling="strong"),   # whole, split over the ranks
    "4full": dict(cohort=COHORT, sites_per_gpu=None, scaling="strong"),                                                                # whole, split over the ranks
}
METRIC = "combined_variant_sites_per_sec"
REF_BIN = os.path.join(REPO, "oracle", "_ref", "bin", "gdb_ref_dump")


def workload_config(args, n_gpus):
    c = dict(COHORT)
    c["n_samples"] = args.samples
    which = getattr(args, "config", "4shard")
    if which != "4shard":
        w = WORKLOADS[which]
        c = dict(w["cohort"])
        return dict(workload="GDBSYN-1 config#%s whole: %d samples x %d candidate sites (%d positions), a_pool=%d, p_var=%.2f; %s over %d GPU(s)"
                    % (which.replace("full", ""), c["n_samples"], c["n_sites"], c["n_sites"] * c["spacing"], c["a_pool"], c["p_var"],
                       "replicated" if w["scaling"] == "weak" else "column range split", n_gpus),
                    samples=c["n_samples"], shards=n_gpus,
                    l2="inputs and outputs are far larger than the 126 MB L2; no explicit flush",
                    sharding="contig-interval (column) shards, one per GPU, no data-path collective")
    return dict(workload="GDBSYN-1 config#4 column shard: %d samples x %d candidate sites (%d positions) per GPU, a_pool=%d, p_var=%.2f; %d shard(s)"
                % (c["n_samples"], args.sites_per_gpu, args.sites_per_gpu * c["spacing"], c["a_pool"], c["p_var"], n_gpus),
                samples=c["n_samples"], sites_per_gpu=args.sites_per_gpu, shards=n_gpus,
                l2="inputs (GBs of fragment columns) and outputs (tens of GB per step) are far larger than the 126 MB L2; no explicit flush",
                sharding="contig-interval (column) shards, one per GPU, no data-path collective")


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu_index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unavailable"])
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                    reasons=reasons, samples=len(self.rows))


# ------------------------------------------------------------------------------------------------ reference (CPU) arm
def reference_sample(args, n_sites_sample, n_procs, tmp, mode="hot"):
    """Times the unmodified reference on `n_procs` host cores, each combining its own `n_sites_sample`-site slice of the
    same cohort (all samples).  Returns (sites_per_sec over all processes, detail)."""
    from tools import gdbsyn, tiledb_writer
    c = dict(COHORT)
    c["n_samples"] = args.samples
    jobs = []
    for k in range(n_procs):
        lo = (k * n_sites_sample) % max(1, c["n_sites"] - n_sites_sample)
        d = gdbsyn.generate(**c, site_lo=lo, site_hi=lo + n_sites_sample, with_fixed=True)
        ws = os.path.join(tmp, "ws%d" % k)
        info = tiledb_writer.write_array(ws, "syn", d, capacity=1000)
        q = tiledb_writer.write_query_json(os.path.join(ws, "q.json"), info, d["col_range"][0], d["col_range"][1])
        jobs.append(q)
        del d
    t0 = time.time()
    procs = [subprocess.Popen([REF_BIN, "-j", q, "--mode", mode], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for q in jobs]
    outs = [p.communicate()[0] for p in procs]
    wall = time.time() - t0
    sites, secs = 0, 0.0
    for o in outs:
        j = json.loads(o.strip().splitlines()[-1])
        sites += j["sites"]
        secs = max(secs, j["seconds"])
    return sites / secs, dict(sites=sites, max_seconds=secs, wall=wall, procs=n_procs)


def port_sample(args, n_sites_sample):
    """Fallback when oracle/_ref is not present: the plain-C oracle port, single thread."""
    from oracle import oracle
    from tools import gdbsyn
    c = dict(COHORT)
    c["n_samples"] = args.samples
    d = gdbsyn.generate(**c, site_lo=0, site_hi=n_sites_sample)
    cols = gdbsyn.to_oracle_columns(d)
    t0 = time.time()
    _, n = oracle.combine(cols, d["col_range"][0], d["col_range"][1])
    dt = time.time() - t0
    return n / dt, dict(sites=n, max_seconds=dt, procs=1)


def cpu_baseline(args, budget_s=12.0, mode="hot"):
    cores = os.cpu_count() or 1
    tmp = tempfile.mkdtemp(prefix="gdbbench")
    try:
        if os.path.exists(REF_BIN):
            # ~0.9 us per sample*site per core (measured); size the per-core slice for `budget_s` seconds
            per_core = max(2, int(budget_s / (0.9e-6 * args.samples * 2)))
            per_core = min(per_core, 64)
            if mode == "null":
                per_core = max(2, per_core // 2)
            v, det = reference_sample(args, per_core, cores, tmp, mode)
            what = ("(unmodified scan_and_operate + handle_deletions + GA4GHOperator::operate, no VCF output)" if mode == "hot" else
                    "(unmodified scan_and_operate + the whole BroadCombinedGVCFOperator::operate incl. htslib bcf_update_*; the finished bcf1_t is dropped)")
            return dict(value=v, unit="sites/s", cores=cores, kind="reference",
                        sample="%d processes x %d candidate sites x %d samples of the same cohort through oracle/_ref/bin/gdb_ref_dump --mode %s "
                               "%s; %d sites in %.1f s (max over processes)"
                               % (cores, per_core, args.samples, mode, what, det["sites"], det["max_seconds"]))
        per = max(2, int(budget_s / (1.0e-6 * args.samples * 2)))
        v, det = port_sample(args, min(per, 64))
        return dict(value=v, unit="sites/s", cores=1, kind="port",
                    sample="oracle/combine_oracle.c single thread, %d sites x %d samples in %.1f s" % (det["sites"], args.samples, det["max_seconds"]))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    base = None
    for _ in range(args.warmup):
        pass    # nothing to warm: every sample is a fresh process over fresh files
    for _ in range(max(1, args.steps)):
        base = cpu_baseline(args, budget_s=max(2.0, 60.0 / max(1, args.steps)))
        vals.append(base["value"])
    v = float(np.median(vals))
    n = args.gpus
    line = dict(metric=METRIC, value=v, unit="sites/s", n_gpus=n, steps=args.steps, warmup=args.warmup,
                ms_per_step=None, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int32", data="synthetic",
                config=workload_config(args, n), impl="reference",
                cpu_baseline=dict(base, value=v),
                e2e=dict(value=v, unit="sites/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
FIELDS = [("AD", "R", "int"), ("PL", "G", "int")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="4shard", choices=sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=COHORT["n_samples"])
    ap.add_argument("--sites-per-gpu", type=int, default=SITES_PER_GPU)
    ap.add_argument("--batch-bytes", type=int, default=16 << 30)
    ap.add_argument("--e2e-steps", type=int, default=8)   # the first step's staging is not hidden behind a scan: enough steps to show the steady state
    ap.add_argument("--e2e-batch-bytes", type=int, default=2 << 30)
    ap.add_argument("--latency-intervals", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e from fragment files, config-#3 strong scaling, gather, latency sweep")
    ap.add_argument("--scratch", default=os.environ.get("GDBBENCH_SCRATCH", "/tmp"))
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from genomicsdb_b200 import capi, sharding
    from tools import gdbsyn, tiledb_writer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py needs a CUDA device: the combine path has no CPU fallback")
    torch.cuda.set_device(local)
    # pinned host buffers are first-touched by this process: keep it (and them) on the GPU's own NUMA node
    numa = sharding.bind_to_gpu_numa_node(local)
    # stdout carries ONE JSON line: whatever libraries write to file descriptor 1 on the way (NCCL prints its version banner
    # there) goes to stderr; the descriptor is handed back right before the line is printed
    sys.stdout.flush()
    saved_stdout_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    w = WORKLOADS[args.config]
    cohort = dict(w["cohort"])
    if args.config == "4shard":
        cohort["n_samples"] = args.samples
        lo = (rank * args.sites_per_gpu) % cohort["n_sites"]
        site_lo, site_hi = lo, min(cohort["n_sites"], lo + args.sites_per_gpu)
    elif w["scaling"] == "weak":
        site_lo, site_hi = 0, cohort["n_sites"]
    else:
        site_lo, site_hi = cohort["n_sites"] * rank // world, cohort["n_sites"] * (rank + 1) // world
    d = gdbsyn.generate(**cohort, site_lo=site_lo, site_hi=site_hi)
    q_lo, q_hi = d["col_range"]
    manifest = gdbsyn.algorithmic_bytes(d)

    ctx = capi.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def one_pass(arr, flags, batch_bytes=None, lo=None, hi=None):
        sc = capi.Scan(arr, q_lo if lo is None else lo, q_hi if hi is None else hi, flags=flags)
        sites = out_bytes = live_words = fmt_bytes = 0
        while True:
            b = sc.next_batch(0, batch_bytes or args.batch_bytes, raw=True)
            if b.n_sites == 0:
                break
            sites += b.n_sites
            out_bytes += b.out_bytes
            fmt_bytes += b.fmt_total_bytes
            live_words += b.n_sites * b.n_query_rows
        sc.close()
        return sites, out_bytes, live_words, fmt_bytes

    # ---------------- HBM-resident metric
    arr = gdbsyn.to_gpu_array(ctx, d)
    info = arr.info()
    for _ in range(max(3, args.warmup)):
        one_pass(arr, capi.SCAN_DEVICE_OUTPUT)
    ctx.set_profiling(True)
    ctx.reset_kernel_stats()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launches
    ev0.record(stream)
    sites = out_bytes = 0
    for _ in range(args.steps):
        s, ob, _, _ = one_pass(arr, capi.SCAN_DEVICE_OUTPUT)
        sites += s
        out_bytes += ob
    ev1.record(stream)
    ctx.synchronize()
    torch.cuda.synchronize()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = ctx.launches - launches0
    ctx.set_profiling(False)      # kernel statistics stay readable; nothing below adds to them
    # a short timed region may end before nvidia-smi has answered twice: keep the same load running (untimed) until the
    # sampler holds a few readings taken under it
    t_extra = time.time()
    while len(sampler.rows) < 4 and time.time() - t_extra < 3.0:
        one_pass(arr, capi.SCAN_DEVICE_OUTPUT)
    sampler.stop_flag = True
    kms = {k: ctx.kernel_stats(k)[0] / args.steps for k in ("k_remap_mix", "k_remap_tile", "k_remap_tma", "k_remap_nt", "k_remap_generic", "k_copy_tile", "k_site_scan",
                                                           "k_union", "k_verify_alleles", "k_sig_perm", "k_live_walk")}
    remap_n = sum(ctx.kernel_stats(k)[1] for k in ("k_remap_mix", "k_remap_tile", "k_remap_tma", "k_remap_nt", "k_remap_generic"))
    remap_ms = (kms["k_remap_mix"] + kms["k_remap_tile"] + kms["k_remap_tma"] + kms["k_remap_nt"] + kms["k_remap_generic"] + kms["k_copy_tile"]) * args.steps
    all_ms, _ = ctx.kernel_stats("k_")
    sites_per_step = sites // args.steps
    total_sites = sum_over_ranks(float(sites))
    value = total_sites / (ms * 1e-3)

    # ---------------- BASELINE config #5: latency of scan_and_operate over random 1 Mb intervals of the resident shard (rank 0)
    latency = None
    if rank == 0 and not args.no_extras and args.latency_intervals > 0 and q_hi - q_lo > 1000000:
        rng = np.random.RandomState(4321)
        starts = q_lo + rng.randint(0, q_hi - q_lo + 1 - 1000000, size=args.latency_intervals)
        lat, nsite = [], []
        for k, a0 in enumerate(list(starts[:5]) + list(starts)):
            t0 = time.perf_counter()
            n, _, _, _ = one_pass(arr, capi.SCAN_DEVICE_OUTPUT, lo=int(a0), hi=int(a0) + 999999)
            ctx.synchronize()
            if k >= 5:
                lat.append((time.perf_counter() - t0) * 1e3)
                nsite.append(n)
        lat = np.array(lat)
        latency = dict(intervals=int(args.latency_intervals), interval_positions=1000000, sites_per_interval=float(np.mean(nsite)),
                       p50_ms=float(np.percentile(lat, 50)), p95_ms=float(np.percentile(lat, 95)), p99_ms=float(np.percentile(lat, 99)), max_ms=float(lat.max()),
                       note="gdbgpu_scan_begin (left sweep, breakpoints) -> next_batch until empty -> scan_end per interval, host clock, output left in HBM")

    # ---------------- the other consumer of the staged shard: the columnar iterator's cell list + the two
    # operators that run as kernels, over this rank's whole shard; only bins / distinct allele records cross PCIe
    columnar = None
    if rank == 0 and not args.no_extras:
        def cells_pass(op):
            sc = capi.Scan(arr, q_lo, q_hi, flags=capi.SCAN_CELLS_ONLY)
            r = op(sc)
            sc.close()
            ctx.synchronize()
            return r
        hist = lambda sc: int(sc.cells_histogram(q_lo, q_hi, 1000).sum())
        ac = lambda sc: len(sc.cells_allele_counts())
        cells_pass(hist), cells_pass(ac)      # warm-up (buffers)
        t0 = time.perf_counter(); listed = cells_pass(hist); t_hist = time.perf_counter() - t0
        t0 = time.perf_counter(); n_rec = cells_pass(ac); t_ac = time.perf_counter() - t0
        columnar = dict(listed_cells=listed, list_plus_histogram_ms=t_hist * 1e3, listed_cells_per_s=listed / t_hist,
                        allele_count_ms=t_ac * 1e3, allele_count_cells_per_s=listed / t_ac, allele_count_records=n_rec,
                        note="gdbgpu_scan_begin(GDBGPU_SCAN_CELLS_ONLY) + gdbgpu_scan_cells_histogram / _cells_allele_counts over the resident shard, "
                             "wall clock incl. the call's host side; reference tool: tools/cells_bench.py")

    # ---------------- NCCL: the one collective of the design -- rank-ordered concatenation of a per-site summary stream
    gather = None
    if world > 1 and not args.no_extras:
        sc = capi.Scan(arr, q_lo, q_hi, flags=capi.SCAN_SUMMARY_ONLY)
        parts = []
        while True:
            b = sc.next_batch(0, 1 << 30)
            if b is None:
                break
            rec = np.zeros((b.n_sites, 3), dtype=np.int64)
            rec[:, 0], rec[:, 1], rec[:, 2] = b.site_begin, b.site_end, b.n_alleles
            parts.append(rec)
        sc.close()
        mine = torch.from_numpy(np.concatenate(parts).view(np.uint8).reshape(-1)).to(dev)
        sharding.gather_concat(mine, device=dev)      # warm-up (communicator, buffers)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cat = sharding.gather_concat(mine, device=dev)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        gather = dict(bytes_per_rank=int(mine.numel()), ms=dt * 1e3, bytes_on_rank0=len(cat) if rank == 0 else None, backend="nccl",
                      what="per-site summary (begin, end, merged allele count) of every rank's shard, concatenated in rank order on rank 0: all-gather of "
                           "the sizes + grouped ncclSend / ncclRecv of exactly those sizes (genomicsdb_b200/sharding.py: gather_concat; MPI_Gatherv in gt_mpi_gather.cc:165-263)")
        if rank == 0:
            allrec = np.frombuffer(cat, dtype=np.int64).reshape(-1, 3)
            gather["records"] = int(len(allrec))
    arr.close()

    # ---------------- roofline of the dominant kernel (remap): algorithmic bytes per launch / its event-timed duration
    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    b_in_step = manifest["b_in"]
    b_out_step = out_bytes / args.steps
    achieved = ((b_in_step + b_out_step) * args.steps / 1e9) / (remap_ms * 1e-3) if remap_ms > 0 else 0.0
    # DRAM traffic of the same kernels from the committed ncu capture of this round (tools/ncu_traffic.sh), per launch
    traffic, traffic_src = None, None
    tpath = os.path.join(REPO, "profiles", "r02_k5_traffic.json")
    if os.path.exists(tpath) and args.config == "4shard" and args.samples == COHORT["n_samples"] and remap_n:
        tj = json.load(open(tpath))
        traffic = tj["k5_dram_bytes_per_site"] * sites / remap_n
        traffic_src = "profiles/r02_k5_traffic.json (ncu --set full of the remap kernels at this round's HEAD, dram bytes per site of the same shard) x sites per launch"
    roofline = dict(bound="hbm", kernel="gdb::k_remap_mix: K5 remap (band blocks: 256 samples x 8 sites, remapped and copy-through; nt blocks: the non-trivial calls) -- with GDBGPU_K5=tile: gdb::k_remap_tile + k_copy_tile + k_remap_nt", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                    traffic_source=traffic_src, algorithmic_bytes_per_launch=(b_in_step + b_out_step) * args.steps / max(1, remap_n),
                    peak_source=peak_src, algorithmic_bytes_per_step=dict(b_in=b_in_step, b_out=b_out_step),
                    launches=int(remap_n), avg_launch_ms=remap_ms / max(1, remap_n),
                    kernel_ms_per_step=dict(kms, k_remap=remap_ms / args.steps, all_kernels=all_ms / args.steps),
                    step_gbs=((b_in_step + b_out_step) * args.steps / 1e9) / (ms * 1e-3),
                    step_frac=((b_in_step + b_out_step) * args.steps / 1e9) / (ms * 1e-3) / peak)

    # ---------------- end to end through the C ABI with host buffers
    e2e = e2e_int32 = e2e_files = e2e_files_gzip = None
    e2e_note = None
    if not args.no_e2e and 2.3 * info.hbm_bytes > torch.cuda.get_device_properties(dev).total_memory:
        # the pipelined end-to-end loop holds two staged copies of the step's array (one scanned, one being staged) beside the
        # resident one of the measurement above: a whole-cohort array (config #4 whole: 79 GB with its indexes) does not fit three times
        args.no_e2e = True
        e2e_note = "skipped: three staged copies of a %.0f GB array (resident + scanned + being staged) exceed this GPU's memory" % (info.hbm_bytes / 1e9)
    ctx2 = capi.Context(local) if not args.no_e2e else None
    if not args.no_e2e:
        from concurrent.futures import ThreadPoolExecutor
        pinned = {}

        def pin(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            pinned[id(t)] = t
            return t.numpy()

        dp = dict(d)
        dp["coords"], dp["end"] = pin(d["coords"]), pin(d["end"])
        for k in ("ref", "alt", "gt", "pl", "ad"):
            dp[k] = (pin(d[k][0]), pin(d[k][1]))
        ctxs = [ctx, ctx2]

        def e2e_run(flags, path, stage, steps):
            """`stage(k)` puts step k's inputs into HBM through the C ABI (on context k % 2) and returns the array; the scan of
            step k copies every batch back to pinned host memory.  Step k + 1 is staged (host -> device) by a second host thread
            while step k's results travel device -> host: PCIe is full duplex.  Every step pays both copies."""
            def scan(a):
                i = a.info()
                s, ob, lw, fb = one_pass(a, flags, args.e2e_batch_bytes)
                a.close()
                d2h_bytes = fb + s * 24 if flags & capi.SCAN_BCF_FORMAT else ob + lw * 4 + lw * 2 * 3 + s * 24
                return s, i.staged_bytes, d2h_bytes

            for k in range(2):
                scan(stage(k))
            barrier()
            t0 = time.perf_counter()
            es = h2d = d2h = 0
            with ThreadPoolExecutor(1) as ex:
                fut = ex.submit(stage, 0)
                for k in range(steps):
                    a = fut.result()
                    if k + 1 < steps:
                        fut = ex.submit(stage, k + 1)
                    s, hb, db = scan(a)
                    es += s
                    h2d += hb
                    d2h += db
            ctx.synchronize()
            ctx2.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0)
            return dict(value=sum_over_ranks(float(es)) / dt, unit="sites/s", h2d_bytes_per_step=int(h2d // steps),
                        d2h_bytes_per_step=int(d2h // steps), steps=steps, ms_per_step=dt * 1e3 / steps,
                        batch_bytes=args.e2e_batch_bytes, path=path)

        stage_cols = lambda k: gdbsyn.to_gpu_array(ctxs[k % 2], dp)
        e2e = e2e_run(capi.SCAN_BCF_FORMAT | capi.SCAN_PRODUCE_GT,
                      "gdbgpu_array_from_columns(pinned host columns) -> gdbgpu_scan_next_batch(GDBGPU_SCAN_BCF_FORMAT, host output: "
                      "remapped GT/AD/PL collated and BCF-typed on the device, blocks copied to pinned host memory) -> close; "
                      "step k+1 staged by a second host thread / context while step k's blocks are copied back", stage_cols, args.e2e_steps)
        e2e_int32 = e2e_run(0, "gdbgpu_array_from_columns(pinned host columns) -> gdbgpu_scan_next_batch(host output: 32-bit remapped "
                               "vectors + lengths + live matrix copied to pinned host memory) -> close", stage_cols, max(2, args.e2e_steps // 2))
        if not args.no_extras:
            # the same metric from the fragment FILES: TileDB reader (book-keeping, pread of the attribute files: page cache),
            # H2D of the file bytes, K0 inflate when GZIP, index build, scan, D2H
            ws = os.path.join(args.scratch, "gdbbench_%d_r%d" % (os.getpid(), rank))
            try:
                tiledb_writer.write_array(ws, "syn", d, capacity=1000)
                stage_files = lambda k: capi.Array.open(ctxs[k % 2], ws, "syn", FIELDS)
                e2e_files = e2e_run(capi.SCAN_BCF_FORMAT | capi.SCAN_PRODUCE_GT,
                                    "fragment files on disk (uncompressed, page cache) -> gdbgpu_array_open_ex -> gdbgpu_scan_next_batch(GDBGPU_SCAN_BCF_FORMAT, "
                                    "host output) -> close", stage_files, args.e2e_steps)
                shutil.rmtree(ws, ignore_errors=True)
                if args.config == "4shard":
                    # GZIP fragment (the reference's default): a 1/8 slice of the shard, so that writing it with zlib stays short
                    n8 = max(8, (site_hi - site_lo) // 8)
                    d8 = gdbsyn.generate(**cohort, site_lo=site_lo, site_hi=site_lo + n8)
                    tiledb_writer.write_array(ws, "syn", d8, capacity=1000, compress=True)
                    lo8, hi8 = d8["col_range"]
                    saved = (q_lo, q_hi)
                    q_lo, q_hi = lo8, hi8
                    e2e_files_gzip = e2e_run(capi.SCAN_BCF_FORMAT | capi.SCAN_PRODUCE_GT,
                                             "GZIP-compressed fragment files of a 1/8 slice of the shard (%d candidate sites) -> gdbgpu_array_open_ex (tiles "
                                             "inflated on the device, K0) -> scan (BCF blocks to the host) -> close" % n8, stage_files, args.e2e_steps)
                    q_lo, q_hi = saved
                    del d8
            finally:
                shutil.rmtree(ws, ignore_errors=True)
    del d

    # ---------------- strong scaling on ONE on-disk array: BASELINE config #3 (10 k samples x 1 M positions), every rank opens
    # its own tile range (cuts from the book-keeping) of the same fragment
    strong = None
    if not args.no_extras and args.config == "4shard":
        c3 = WORKLOADS["3"]["cohort"]
        ws3 = os.path.join(args.scratch, "gdbbench_c3_%s" % os.environ.get("MASTER_PORT", str(os.getpid())))
        try:
            if rank == 0:
                d3 = gdbsyn.generate(**c3)
                tiledb_writer.write_array(ws3, "syn", d3, capacity=1000)
                del d3
            barrier()
            cuts = sharding.tile_balanced_cuts(ws3, "syn", world)
            lo3, hi3 = cuts[rank]
            t0 = time.perf_counter()
            a3 = capi.Array.open(ctx, ws3, "syn", FIELDS, col_range=(lo3, hi3))
            ctx.synchronize()
            open_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
            i3 = a3.info()
            for _ in range(3):
                one_pass(a3, capi.SCAN_DEVICE_OUTPUT, lo=lo3, hi=hi3)
            barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            n3 = 0
            for _ in range(args.steps):
                n3 += one_pass(a3, capi.SCAN_DEVICE_OUTPUT, lo=lo3, hi=hi3)[0]
            e1.record(stream)
            ctx.synchronize()
            torch.cuda.synchronize()
            ms3 = max_over_ranks(e0.elapsed_time(e1))
            a3.close()
            tot3 = sum_over_ranks(float(n3))
            # end to end from the files, every rank its range: open (read + H2D + index) + scan with BCF blocks to the host + close
            barrier()
            t0 = time.perf_counter()
            a3 = capi.Array.open(ctx, ws3, "syn", FIELDS, col_range=(lo3, hi3))
            n3e = one_pass(a3, capi.SCAN_BCF_FORMAT | capi.SCAN_PRODUCE_GT, args.e2e_batch_bytes, lo=lo3, hi=hi3)[0]
            a3.close()
            ctx.synchronize()
            dt3 = max_over_ranks(time.perf_counter() - t0)
            strong = dict(workload="GDBSYN-1 config#3 whole: %d samples x %d candidate sites, ONE on-disk fragment; rank r opens the tiles of its column "
                                   "range (gdbgpu_array_column_cuts -> gdbgpu_array_open_range)" % (c3["n_samples"], c3["n_sites"]),
                          scaling="strong", n_gpus=world, sites_per_pass=tot3 / args.steps, ms_per_pass=ms3 / args.steps,
                          value=tot3 / (ms3 * 1e-3), unit="sites/s", open_ms=open_ms, cells_this_rank=int(i3.n_cells),
                          cells_max_over_ranks=int(max_over_ranks(float(i3.n_cells))), staged_bytes_this_rank=int(i3.staged_bytes),
                          e2e_from_files=dict(value=sum_over_ranks(float(n3e)) / dt3, unit="sites/s", ms=dt3 * 1e3))
            barrier()
        finally:
            if rank == 0:
                shutil.rmtree(ws3, ignore_errors=True)
    ctx.close()
    if ctx2:
        ctx2.close()

    # ---------------- the box's host path: pinned D2H / H2D of this rank alone and of all ranks at once (what caps e2e at N > 1)
    host_path = None
    if not args.no_extras:
        nbytes = 1 << 30
        hbuf = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        dbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)

        def copy_gbs(d2h, reps=4):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                if d2h:
                    hbuf.copy_(dbuf, non_blocking=True)
                else:
                    dbuf.copy_(hbuf, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
        copy_gbs(True, 1)
        alone = {}
        for r in range(world):          # one rank at a time
            barrier()
            if r == rank:
                alone = dict(d2h=copy_gbs(True), h2d=copy_gbs(False))
        barrier()
        conc_d2h = copy_gbs(True)       # every rank at once
        barrier()
        conc_h2d = copy_gbs(False)
        barrier()
        host_path = dict(bytes=nbytes, d2h_alone_gbs_min_over_ranks=-max_over_ranks(-alone["d2h"]), h2d_alone_gbs_min_over_ranks=-max_over_ranks(-alone["h2d"]),
                         d2h_all_ranks_at_once_gbs_sum=sum_over_ranks(conc_d2h), d2h_all_ranks_at_once_gbs_min=-max_over_ranks(-conc_d2h),
                         h2d_all_ranks_at_once_gbs_sum=sum_over_ranks(conc_h2d), h2d_all_ranks_at_once_gbs_min=-max_over_ranks(-conc_h2d),
                         note="1 GiB pinned copies, CUDA events; e2e moves h2d_bytes_per_step + d2h_bytes_per_step per rank and step over this path")
        del hbuf, dbuf

    base = base_bcf = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.config == "4shard":
        base = cpu_baseline(args)
        if os.path.exists(REF_BIN):   # like-for-like for `e2e` (which also collates + BCF-encodes): the reference's whole Broad operator
            base_bcf = cpu_baseline(args, budget_s=8.0, mode="null")

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit="sites/s", n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling=w["scaling"], vs_baseline=None, dtype="int32",
                    data="synthetic", config=workload_config(args, world), clocks=sampler.summary(), e2e=e2e, e2e_int32=e2e_int32,
                    e2e_from_fragment_files=e2e_files, e2e_from_gzip_fragment_files=e2e_files_gzip,
                    gpu_launches=int(launches), roofline=roofline, cpu_baseline=base, cpu_baseline_broad_gvcf_operator=base_bcf,
                    strong_scaling_config3=strong, concatenation=gather, latency_config5=latency, columnar_iterator=columnar,
                    numa_binding=dict(node=numa[0], cores=numa[1]) if numa else None, host_path=host_path, e2e_note=e2e_note,
                    sites_per_step_per_gpu=int(sites_per_step), cells_per_gpu=int(info.n_cells),
                    staged_bytes_per_gpu=int(info.staged_bytes), hbm_bytes_per_gpu=int(info.hbm_bytes))
        sys.stdout.flush()
        os.dup2(saved_stdout_fd, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
