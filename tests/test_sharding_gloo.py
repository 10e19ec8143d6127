"""CPU, world_size 2, gloo: the N>1 path's host logic -- shard assignment and rank-ordered concatenation.  Each rank
combines its own column shard (with the oracle here: there is no GPU in this container; the -m gpu tests prove the
kernels equal the oracle) and the concatenation must equal the single-partition run except that the interval straddling
the shard boundary is split in two records, exactly as two reference ranks with adjacent column partitions would."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from genomicsdb_b200 import sharding
    from oracle import oracle
    from tools import gdbsyn
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shape = dict(n_samples=40, n_sites=60, spacing=100, p_var=0.15, a_pool=8)
    per = shape["n_sites"] // world
    d = gdbsyn.generate(**shape, site_lo=rank * per, site_hi=(rank + 1) * per)
    cols = gdbsyn.to_oracle_columns(d)
    lo, hi = d["col_range"]
    body, n = oracle.combine(cols, lo, hi)
    cat = sharding.gather_concat(body)
    if rank == 0:
        open(os.path.join(out_dir, "cat.bin"), "wb").write(oracle.dump_header(cols) + cat)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_concatenate_to_single_partition_result(tmp_path):
    sys.path.insert(0, REPO)
    from oracle import oracle
    from tools import gdbsyn
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    cat = open(tmp_path / "cat.bin", "rb").read()
    shape = dict(n_samples=40, n_sites=60, spacing=100, p_var=0.15, a_pool=8)
    full = gdbsyn.generate(**shape)
    want, n_full = oracle.combine_query(gdbsyn.to_oracle_columns(full), [list(full["col_range"])])
    _, sa = oracle.parse_dump(want)
    _, sb = oracle.parse_dump(cat)
    assert len(sb) == len(sa) + (world - 1)
    # merge the records split at the shard boundary, then everything must be identical
    merged, i = [], 0
    while i < len(sb):
        s = sb[i]
        if i + 1 < len(sb) and sb[i + 1]["begin"] == s["end"] + 1 and not any(x["begin"] == sb[i + 1]["begin"] for x in sa):
            t = sb[i + 1]
            assert s["alleles"] == t["alleles"] and len(s["calls"]) == len(t["calls"])
            s = dict(s, end=t["end"])
            i += 1
        merged.append(s)
        i += 1
    assert len(merged) == len(sa)
    for x, y in zip(sa, merged):
        assert (x["begin"], x["end"], x["alleles"], x["remap"]) == (y["begin"], y["end"], y["alleles"], y["remap"])
        assert len(x["calls"]) == len(y["calls"])
        for cx, cy in zip(x["calls"], y["calls"]):
            assert cx["row"] == cy["row"] and cx["in2m"] == cy["in2m"]
            for k in cx["fields"]:
                assert (cx["fields"][k] == cy["fields"][k]).all()


def test_split_interval_tiles():
    sys.path.insert(0, REPO)
    from genomicsdb_b200 import sharding
    for world in (1, 2, 3, 8):
        parts = sharding.split_interval(1000, 10_001_001, world)
        assert parts[0][0] == 1000 and parts[-1][1] == 10_001_001
        for a, b in zip(parts, parts[1:]):
            assert b[0] == a[1] + 1


def test_equal_load_cuts_from_histogram():
    """cuts on bin boundaries, contiguous, tiling the range, each shard near its share (the device histogram of
    tests/test_gpu_cells.py feeds this on a GPU box)"""
    from genomicsdb_b200.sharding import equal_load_cuts_from_histogram
    import random
    rng = random.Random(3)
    for world in (1, 2, 3, 8):
        counts = [rng.choice((0, 0, 1, 5, 40, 300)) for _ in range(257)]
        cuts = equal_load_cuts_from_histogram(counts, 1000, 50, world, col_hi=1000 + 257 * 50 - 7)
        assert len(cuts) == world and cuts[0][0] == 1000 and cuts[-1][1] == 1000 + 257 * 50 - 7
        for (a, b), (c, d) in zip(cuts, cuts[1:]):
            assert c == b + 1 and (c - 1000) % 50 == 0
        loads = [sum(counts[(a - 1000) // 50:(b - 1000) // 50 + 1]) for a, b in cuts]
        assert sum(loads) == sum(counts)
        assert max(loads) <= sum(counts) / world + max(counts)      # never more than one bin over the share
    assert equal_load_cuts_from_histogram([0, 0, 0], 0, 10, 2) == [(0, 9), (10, 29)]
