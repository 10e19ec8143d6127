"""K0 (k_inflate, one warp per zlib stream) through gdbgpu_inflate_streams against Python's zlib on streams chosen to reach every
branch of the decoder: stored / fixed / dynamic blocks, several blocks per stream, codes longer than the primary tables, overlapping
matches, matches whose source has left the shared-memory ring (read back from HBM), every output alignment, sizes from 0 bytes to
megabytes; corrupt streams are reported (utils.cc:466-512: gunzip fails -> the read fails), never silently accepted."""
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from genomicsdb_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _payloads():
    rng = np.random.RandomState(7)
    out = []
    out.append(b"")                                                     # empty stream
    out.append(b"A")                                                    # one literal (fixed block)
    out.append(b"abc" * 5)                                              # overlapping match, distance 3
    out.append(b"\0" * 100000)                                          # distance-1 run, maximal match lengths
    out.append(bytes(rng.randint(0, 256, size=70000, dtype=np.uint8)))   # incompressible: stored blocks (> 65535 bytes: two of them)
    out.append(np.arange(50000, dtype="<i8").tobytes())                  # coordinates-like: short distances, dynamic codes
    r = bytes(rng.randint(0, 256, size=3000, dtype=np.uint8))
    out.append(r + bytes(6000) + r + bytes(rng.randint(0, 4, size=5000, dtype=np.uint8)) + r)   # matches 9 KB and 14 KB back: past the ring
    out.append(r * 40)                                                  # 3000-byte period: every match source is beyond a 2 KB ring
    skew = rng.zipf(1.3, size=200000) % 256                            # skewed symbol distribution: long Huffman codes
    out.append(bytes(skew.astype(np.uint8)))
    pl = np.where(rng.rand(40000, 55) < 0.9, 0, rng.randint(0, 3000, size=(40000, 55))).astype("<i4")
    out.append(pl.tobytes())                                            # PL-like, 8.8 MB, many dynamic blocks
    for n in (1, 2, 15, 16, 17, 511, 512, 513, 1023, 4097):             # sizes around the flush granularity
        out.append(bytes(rng.randint(0, 3, size=n, dtype=np.uint8)))
    return out


@pytest.mark.parametrize("level", [1, 6, 9, 0])
def test_streams_inflate_to_what_zlib_gives(ctx, level):
    payloads = _payloads()
    streams = [zlib.compress(p, level) for p in payloads]
    got, ms = ctx.inflate_streams(streams, [len(p) for p in payloads])
    for i, (g, p) in enumerate(zip(got, payloads)):
        assert g == p, "stream %d (level %d, %d bytes)" % (i, level, len(p))


def test_fixed_and_raw_strategies(ctx):
    """Z_FIXED (fixed Huffman codes only), Z_HUFFMAN_ONLY (no matches), Z_RLE (distance-1 matches only), small windows"""
    payloads = _payloads()
    for strategy, wbits in ((zlib.Z_FIXED, 15), (zlib.Z_HUFFMAN_ONLY, 15), (zlib.Z_RLE, 15), (zlib.Z_DEFAULT_STRATEGY, 9)):
        streams = []
        for p in payloads:
            c = zlib.compressobj(6, zlib.DEFLATED, wbits, 8, strategy)
            streams.append(c.compress(p) + c.flush())
        got, _ = ctx.inflate_streams(streams, [len(p) for p in payloads])
        assert got == payloads, strategy


def test_every_output_alignment(ctx):
    """tiles follow each other byte by byte in a column: every misalignment of the output (and input) address"""
    rng = np.random.RandomState(3)
    payloads = [bytes(rng.randint(0, 5, size=600 + 37 * k + k % 16, dtype=np.uint8)) for k in range(64)]
    streams = [zlib.compress(p, 6) for p in payloads]
    got, _ = ctx.inflate_streams(streams, [len(p) for p in payloads])
    assert got == payloads


def test_corrupt_streams_are_reported(ctx):
    from genomicsdb_b200 import capi
    p = np.arange(20000, dtype="<i4").tobytes()
    good = zlib.compress(p, 6)
    bad_variants = {
        "flipped data byte": good[:len(good) // 2] + bytes([good[len(good) // 2] ^ 0x5A]) + good[len(good) // 2 + 1:],
        "flipped Adler-32": good[:-1] + bytes([good[-1] ^ 1]),
        "bad header": b"\x79" + good[1:],
        "truncated": good[:len(good) // 2],
    }
    for what, b in bad_variants.items():
        with pytest.raises(capi.GdbGpuError) as e:
            ctx.inflate_streams([good, b], [len(p), len(p)])
        assert e.value.code == capi.ERR_IO, what
    with pytest.raises(capi.GdbGpuError):                       # right stream, wrong announced size
        ctx.inflate_streams([good], [len(p) - 1])
    got, _ = ctx.inflate_streams([good], [len(p)])               # the context stays usable
    assert got[0] == p
