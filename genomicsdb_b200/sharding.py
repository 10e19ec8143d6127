"""Column-interval (genomic range) sharding across GPUs and the rank-ordered concatenation of per-shard results.

Mirrors how the reference scales the combine path: one MPI rank per column partition, no communication during the
combine, rank-ordered gather only where one ordered stream is wanted (tools/src/gt_mpi_gather.cc:161-263; partition
clipping src/main/cpp/src/config/genomicsdb_config_base.cc:205-223).  Here: one process per GPU under torch.distributed
(NCCL on GPUs over NVLink/NVSwitch, gloo in the CPU tests); the only collective is the final concatenation.
Host-side plumbing only -- no compute.
"""
import torch
import torch.distributed as dist


def split_interval(col_lo, col_hi, world):
    """Cut [col_lo, col_hi] into `world` contiguous column intervals that tile it (one per rank)."""
    n = col_hi - col_lo + 1
    cuts = [col_lo + (n * r) // world for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1] - 1) for r in range(world)]


def equal_cell_cuts(cols, world):
    """Column cuts with ~equal cell counts per shard, always at a column boundary (the reference's
    ColumnHistogramOperator::equi_partition_and_print_bins idea, variant_operations.cc:765-788).
    cols: sorted 1-D tensor / array of cell columns.  Returns list of (lo, hi)."""
    cols = torch.as_tensor(cols)
    n = len(cols)
    out, lo = [], int(cols[0])
    for r in range(1, world):
        i = (n * r) // world
        c = int(cols[i])
        # first column >= c that starts a new column group
        while i < n and int(cols[i]) == c and i > 0 and int(cols[i - 1]) == c:
            i += 1
        hi = (int(cols[i]) if i < n else int(cols[-1]) + 1) - 1
        hi = max(hi, lo)
        out.append((lo, hi))
        lo = hi + 1
    out.append((lo, int(cols[-1]) if n else lo))
    return out


def equal_load_cuts_from_histogram(counts, begin, bin_size, world, col_hi=None):
    """Column cuts from a histogram of cell begin columns (ColumnHistogramOperator bins): bins are handed to shards left to
    right, a shard closes once it holds its share of what is left (equi_partition_and_print_bins, variant_operations.cc:767-790,
    with the share re-computed per shard so the last one is not starved).  Cuts fall on bin boundaries.  Returns `world`
    contiguous (lo, hi) intervals tiling [begin, col_hi]."""
    counts = [int(c) for c in counts]
    n_bins = len(counts)
    col_hi = begin + n_bins * bin_size - 1 if col_hi is None else col_hi
    out, b, left = [], 0, sum(counts)
    for r in range(world):
        if r == world - 1:
            e = n_bins
        else:
            share, cur, e = left / (world - r), 0, b
            while e < n_bins - (world - 1 - r) and (cur < share or e == b):
                cur += counts[e]
                e += 1
            left -= cur
        lo = begin + b * bin_size
        hi = col_hi if r == world - 1 else begin + e * bin_size - 1
        out.append((lo, max(hi, lo - 1)))
        b = e
    return out


def device_equal_load_cuts(arr, world, bin_size, col_lo=None, col_hi=None, query_rows=None):
    """Shard boundaries for `world` GPUs from the DEVICE histogram of the staged array's cells (gdbgpu_scan_cells_histogram:
    only the bins cross PCIe): SURVEY 8e's "shards with ~equal cell counts".  arr: genomicsdb_b200.capi.Array."""
    from genomicsdb_b200 import capi
    info = arr.info()
    lo = info.min_col if col_lo is None else col_lo
    hi = max(info.max_col, info.max_end) if col_hi is None else col_hi
    sc = capi.Scan(arr, lo, hi, query_rows=query_rows, flags=capi.SCAN_CELLS_ONLY)
    try:
        counts = sc.cells_histogram(lo, hi, bin_size)
    finally:
        sc.close()
    return equal_load_cuts_from_histogram(counts.tolist(), lo, bin_size, world, col_hi=hi)


def gather_concat(local_bytes, device=None, group=None, dst=0):
    """Rank-ordered concatenation of per-rank byte strings on rank `dst` (MPI_Gatherv in gt_mpi_gather.cc:165-263):
    all-gather of sizes, then one padded all-gather of the payloads.  Returns bytes on dst, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = device or torch.device("cpu")
    size = torch.tensor([len(local_bytes)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes + [1])
    buf = torch.zeros(m, dtype=torch.uint8, device=dev)
    if local_bytes:
        buf[:len(local_bytes)] = torch.frombuffer(bytearray(local_bytes), dtype=torch.uint8).to(dev)
    parts = [torch.zeros(m, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    if rank != dst:
        return None
    return b"".join(bytes(parts[r][:sizes[r]].cpu().numpy().tobytes()) for r in range(world))
