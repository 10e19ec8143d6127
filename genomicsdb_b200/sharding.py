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
    """Rank-ordered concatenation of per-rank byte strings on rank `dst`: MPI_Gather of the sizes + MPI_Gatherv of the
    payloads in gt_mpi_gather.cc:165-263.  Here: an all-gather of the sizes (8 bytes per rank), then every other rank sends
    exactly its bytes to `dst`, which posts one receive of exactly that size per sender (grouped isend / irecv: ncclSend /
    ncclRecv inside one group on NCCL).  Nothing is padded and nobody but `dst` holds more than its own payload.
    local_bytes: bytes-like, or a uint8 tensor (a tensor already on `device` is sent as it is).  Returns bytes on dst, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = device or torch.device("cpu")
    if torch.is_tensor(local_bytes):
        mine = local_bytes.to(dev).contiguous().view(torch.uint8).reshape(-1)
    else:
        mine = torch.frombuffer(bytearray(local_bytes), dtype=torch.uint8).to(dev) if len(local_bytes) else torch.zeros(0, dtype=torch.uint8, device=dev)
    size = torch.tensor([mine.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(s.item()) for s in sizes]
    ops, parts = [], {}
    if rank == dst:
        for r in range(world):
            if r != dst and sizes[r]:
                parts[r] = torch.empty(sizes[r], dtype=torch.uint8, device=dev)
                ops.append(dist.P2POp(dist.irecv, parts[r], r, group=group))
    elif mine.numel():
        ops.append(dist.P2POp(dist.isend, mine, dst, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if rank != dst:
        return None
    parts[dst] = mine
    return b"".join(parts[r].cpu().numpy().tobytes() for r in range(world) if sizes[r])


def bind_to_gpu_numa_node(local_rank):
    """Pins this process (and the pinned host buffers it is about to allocate: first touch) to the CPU cores of the NUMA node
    the rank's GPU hangs off -- with 8 ranks on a two-socket box, staging and result copies that cross the socket
    interconnect share one link.  Returns (numa node, number of cores) or None when the topology cannot be read."""
    import os
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node, len(allowed)
    except Exception:
        return None


def tile_balanced_cuts(workspace, array_name, world):
    """Column cuts for `world` ranks with ~equal cell counts, taken from the fragments' book-keeping alone (tile bounding
    coordinates): no cell is read.  Returns `world` (lo, hi) intervals tiling the array's columns."""
    import ctypes as C
    from genomicsdb_b200 import capi
    cuts = (C.c_int64 * (world + 1))()
    rc = capi.lib().gdbgpu_array_column_cuts(workspace.encode(), array_name.encode(), world, cuts)
    if rc != capi.OK:
        raise capi.GdbGpuError(rc, "cannot read the book-keeping of %s/%s" % (workspace, array_name))
    return [(int(cuts[r]), int(cuts[r + 1]) - 1) for r in range(world)]
