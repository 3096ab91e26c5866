"""On-disk index format (fmb_index_save / fmb_index_load; replaces saveIndex / loadIndex of fmindex/diskStorage.h:13-27).

CPU part: the loader validates a file completely (magic, version, section sizes, checksums, trailing bytes) before it touches the
device -- checked with files written HERE from the documented layout (csrc/fmb_io.cu), which also pins the format.
GPU part: save -> load round trip; the loaded index holds the same BWT / samples and answers searches like the original."""
import struct

import numpy as np
import pytest

from helpers import hits_equal, locs_equal, make_index_pair

FMB_EINVAL, FMB_ENODEVICE = -1, -2


def _checksum(buf):
    """fmb_checksum64: h = (h ^ w) * 0x100000001B3 over little-endian 8-byte words, tail zero padded"""
    b = bytes(buf)
    b += b"\0" * (-len(b) % 8)
    h = 0xCBF29CE484222325
    for (w,) in struct.iter_unpack("<Q", b):
        h = ((h ^ w) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def _write(path, sigma, bwt, bwt_rev, bitmap, seq, pos, version=2, magic=b"FMB200IX", tamper=None, extra=b"", tamper_head=None):
    secs = [np.asarray(bwt, np.uint8).tobytes(), b"" if bwt_rev is None else np.asarray(bwt_rev, np.uint8).tobytes(),
            np.asarray(bitmap, np.uint64).tobytes(), np.asarray(seq, np.uint32).tobytes(), np.asarray(pos, np.uint32).tobytes()]
    head = magic + struct.pack("<IIQIIQ", version, sigma, len(bwt), 0 if bwt_rev is None else 1, 0, len(seq))
    head += struct.pack("<5Q", *[len(s) for s in secs]) + struct.pack("<5Q", *[_checksum(s) for s in secs])
    assert len(head) == 120
    if version >= 2:        # header checksum: low 32 bits of the checksum of the header with the field zero
        head = head[:28] + struct.pack("<I", _checksum(head) & 0xFFFFFFFF) + head[32:]
    if tamper_head is not None:
        head = head[:tamper_head] + bytes([head[tamper_head] ^ 1]) + head[tamper_head + 1:]
    body = b"".join(secs)
    if tamper is not None:
        body = body[:tamper] + bytes([body[tamper] ^ 1]) + body[tamper + 1:]
    with open(path, "wb") as f:
        f.write(head + body + extra)


@pytest.fixture(scope="module")
def small():
    from fmb200 import synth
    from oracle.pyoracle import Oracle
    text = synth.multi_text([300, 120], 5, 3)
    o = Oracle.build(text, 5, 4)
    bm, sq, sp = o.samples
    return text, o, (o.bwt, o.bwt_rev, bm, sq, sp)


def test_checksum_matches_the_library(fmb):
    from fmb200.capi import lib
    rng = np.random.default_rng(1)
    for n in (0, 1, 7, 8, 9, 1000, 1003):
        a = rng.integers(0, 256, n).astype(np.uint8)
        assert lib().fmb_checksum64(a.ctypes.data if n else None, n) == _checksum(a)


def test_loader_rejects_damaged_files_before_touching_the_device(fmb, small, tmp_path):
    _, _, (bwt, bwt_rev, bm, sq, sp) = small
    p = str(tmp_path / "ix.fmb")

    def load_error():
        with pytest.raises(fmb.FmbError) as e:
            fmb.Index.load(p)
        return e.value

    _write(p, 5, bwt, bwt_rev, bm, sq, sp, magic=b"NOTANIDX")
    assert load_error().code == FMB_EINVAL and "magic" in str(load_error())
    _write(p, 5, bwt, bwt_rev, bm, sq, sp, version=3)
    assert load_error().code == FMB_EINVAL and "version" in str(load_error())
    _write(p, 5, bwt, bwt_rev, bm, sq, sp, tamper_head=33)       # n_samples: covered by the header checksum
    assert load_error().code == FMB_EINVAL and "header checksum" in str(load_error())
    _write(p, 5, bwt, bwt_rev, bm, sq, sp, tamper=17)
    assert load_error().code == FMB_EINVAL and "checksum" in str(load_error())
    _write(p, 5, bwt, bwt_rev, bm, sq, sp, tamper=2 * len(bwt) + 3)
    assert "checksum mismatch in section 2" in str(load_error())
    _write(p, 5, bwt, bwt_rev, bm, sq, sp, extra=b"x")
    assert "trailing" in str(load_error())
    _write(p, 5, bwt, bwt_rev, bm, sq, sp)
    data = open(p, "rb").read()
    open(p, "wb").write(data[:-5])
    assert "truncated" in str(load_error())
    open(p, "wb").write(data[:60])
    assert "truncated header" in str(load_error())
    # a header that claims 2^31 rows (consistent, checksum included): refused from the file size, nothing is allocated
    def lying_header(rows):
        head = bytearray(data[:120])
        head[16:24] = struct.pack("<Q", rows)
        head[40:48] = struct.pack("<Q", rows)
        head[48:56] = struct.pack("<Q", rows)
        head[56:64] = struct.pack("<Q", ((rows + 63) // 64) * 8)
        head[28:32] = b"\0" * 4
        head[28:32] = struct.pack("<I", _checksum(head) & 0xFFFFFFFF)
        return bytes(head)
    open(p, "wb").write(lying_header(1 << 31) + data[120:])
    assert load_error().code == FMB_EINVAL and "truncated" in str(load_error())
    # ... and 2^60 rows are outside what this build supports
    open(p, "wb").write(lying_header(1 << 60) + data[120:])
    assert load_error().code == -5 and "2^32" in str(load_error())
    # version 1 files (no header checksum) are still read
    _write(p, 5, bwt, bwt_rev, bm, sq, sp, version=1, tamper=17)
    assert "checksum mismatch in section 0" in str(load_error())
    _write(p, 5, bwt, bwt_rev[:-1], bm, sq, sp)                     # bwtRev shorter than bwt
    assert "section 1" in str(load_error())
    # a well-formed file passes validation: without a device the call then fails with ENODEVICE (no CPU fallback), with one it loads
    _write(p, 5, bwt, bwt_rev, bm, sq, sp)
    if fmb.device_count() == 0:
        assert load_error().code == FMB_ENODEVICE
    else:
        assert fmb.Index.load(p).info.n == len(bwt)


@pytest.mark.gpu
@pytest.mark.parametrize("bidirectional", [True, False])
def test_save_load_round_trip(gpu, small, tmp_path, bidirectional):
    from fmb200 import schemes, synth
    text, _, _ = small
    o, g = make_index_pair(gpu, text, 5, 4, bidirectional=bidirectional)
    p = str(tmp_path / "ix.fmb")
    g.save(p)
    g2 = gpu.Index.load(p)
    a, b = g.export(), g2.export()
    assert all(np.array_equal(x, y) for x, y in zip(a, b) if x is not None)
    i1, i2 = g.info, g2.info
    assert (i1.sigma, i1.n, i1.bidirectional, i1.n_samples, i1.tables) == (i2.sigma, i2.n, i2.bidirectional, i2.n_samples, i2.tables)
    # the file is exactly what this test's writer produces from the oracle's data (pins the layout)
    bm, sq, sp = o.samples
    q = str(tmp_path / "expected.fmb")
    _write(q, 5, o.bwt, o.bwt_rev if bidirectional else None, bm, sq, sp)
    assert open(p, "rb").read() == open(q, "rb").read()
    reads, _ = synth.reads_from_text(text[:300], 80, 24, 5)
    reads = synth.plant_errors(reads, 5, 1, True, 2)
    sym, off = synth.flatten(reads)
    h = g2.search_exact(g2.upload(sym, off))
    assert hits_equal(h.hits(), g.search_exact(g.upload(sym, off)).hits())
    assert locs_equal(g2.locate(h).locs(), o.locate(o.search_exact(sym, off)))
    if bidirectional:
        sch = schemes.optimum(0, 2)
        part = schemes.uniform_partition(4, 24)
        r = g2.search_scheme(g2.upload(sym, off), sch, part, True)
        exp = o.search_ng26(sym, off, sch, part, True)
        assert hits_equal(r.hits(), exp) and locs_equal(g2.locate(r).locs(), o.locate(exp))


@pytest.mark.gpu
def test_protein_round_trip(gpu, tmp_path):
    from fmb200 import synth
    text = synth.multi_text([800, 200], 21, 4)
    o, g = make_index_pair(gpu, text, 21, 8)
    p = str(tmp_path / "prot.fmb")
    g.save(p)
    g2 = gpu.Index.load(p)
    reads, _ = synth.reads_from_text(text[:800], 50, 12, 1)
    sym, off = synth.flatten(reads)
    exp = o.search_exact(sym, off)
    assert hits_equal(g2.search_exact(g2.upload(sym, off)).hits(), exp)
