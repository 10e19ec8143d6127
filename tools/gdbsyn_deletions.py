"""GDBSYN-1 cohorts with deletions (tests only): a share of the variant calls of a generated cohort become calls whose REF spans
2-4 bases -- every shorter ALT allele is then a deletion (VariantUtils::contains_deletion), the call's END moves to begin + L - 1,
the call is stored a second time at its END (the loader's END duplicate) and the sample's next reference block begins one base
after it.  What this buys: handle_deletions, the `*` allele, the 1-bp stepping while a deletion is live (query_variants.cc:310)
and the reduced PL / AD / GT of spanning deletions over thousands of samples, held against the unmodified reference."""
import numpy as np

from tools import tiledb_writer


def _mix(x):
    x = np.asarray(x, dtype=np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94d049bb133111eb)
    return x ^ (x >> np.uint64(31))


def with_deletions(d, share=0.3, seed=7):
    coords, end = d["coords"].copy(), d["end"].copy()
    n = len(end)
    aoff, aval = d["alt"]
    alen = np.diff(np.concatenate([aoff.astype(np.int64), [aval.view(np.uint8).nbytes]]))
    is_variant = (alen > 1) & (coords[:, 1] == end)
    with np.errstate(over="ignore"):
        h = _mix(coords[:, 0].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + coords[:, 1].astype(np.uint64) + np.uint64(seed))
    pick = np.nonzero(is_variant & ((h >> np.uint64(11)).astype(np.float64) / 2.0 ** 53 < share))[0]
    L = 2 + (h[pick] % np.uint64(3)).astype(np.int64)
    where = {(int(coords[i, 0]), int(coords[i, 1])): i for i in range(n)}
    # the call: longer REF, END at its last base
    roff, rval = d["ref"]
    refs = [bytes(rval.view(np.uint8)[int(roff[i]):int(roff[i + 1]) if i + 1 < n else None]) for i in range(n)]
    tail = b"ACG"
    for i, l in zip(pick, L):
        refs[i] = refs[i][:1] + tail[:int(l) - 1]
        end[i] = coords[i, 1] + int(l) - 1
    # the sample's next block begins after the deletion; its END duplicate names the new begin
    for i, l in zip(pick, L):
        s, c = int(coords[i, 0]), int(coords[i, 1])
        b = where.get((s, c + 1))
        if b is None:
            continue
        e = where.get((s, int(end[b]))) if end[b] != c + 1 else None   # (the END duplicate lies outside a column shard that ends before it)
        coords[b, 1] = c + int(l)
        if e is not None:
            end[e] = c + int(l)
    # END duplicates of the deletion calls
    dup_coords = np.stack([coords[pick, 0], end[pick]], axis=1)
    dup_end = coords[pick, 1].copy()
    off = np.zeros(n, dtype=np.uint64)
    off[1:] = np.cumsum([len(r) for r in refs])[:-1]
    base = dict(d)
    base["coords"], base["end"] = coords, end
    base["ref"] = (off, np.frombuffer(b"".join(refs), dtype=np.uint8).copy())
    # one gather builds "all cells + duplicates" in the new global order
    src = np.concatenate([np.arange(n), pick])
    all_coords = np.concatenate([coords, dup_coords])
    all_end = np.concatenate([end, dup_end])
    order = np.lexsort((all_coords[:, 0], all_coords[:, 1]))
    out = tiledb_writer.subset_cells(base, src[order])
    out["coords"], out["end"] = all_coords[order], all_end[order]
    out["col_range"] = (int(out["coords"][:, 1].min()), int(max(out["coords"][:, 1].max(), out["end"].max())))
    out["n_deletion_calls"] = int(len(pick))
    return out
