"""CPU: the bit-level format of the on-GPU BGZF writer. The per-lane encoder of
panagram_b200/csrc/pk_deflate.cuh (fixed-Huffman pieces joined by sync markers, lane-wise CRC-32 combined with
GF(2) shift operators) is compiled for the host by tests/native/bgzf_host_check.cpp and driven exactly like the
kernels drive it; zlib's inflate, the gzip module and the .gzi reader of layout.py are the judges."""
import gzip
import struct
import subprocess
import zlib

import numpy as np
import pytest

from panagram_b200 import layout

from conftest import ROOT


@pytest.fixture(scope="module")
def host_check(tmp_path_factory):
    exe = tmp_path_factory.mktemp("native") / "bgzf_host_check"
    subprocess.check_call(["g++", "-O2", "-Wall", "-Werror", "-o", str(exe), str(ROOT / "tests" / "native" / "bgzf_host_check.cpp")])
    return exe


def check_bgzf_image(raw: bytes, gzi: bytes, data: bytes):
    """Every property the reference's reader relies on (index.py:793-845) plus gzip validity."""
    assert gzip.decompress(raw) == data
    assert raw[-28:] == layout.BGZF_EOF
    (n,) = struct.unpack_from("<Q", gzi, 0)
    assert len(gzi) == 8 + 16 * n
    pairs = np.frombuffer(gzi, dtype="<u8", offset=8).reshape(n, 2)
    o, u, starts = 0, 0, []
    while o < len(raw):
        assert raw[o:o + 4] == b"\x1f\x8b\x08\x04" and raw[o + 12:o + 16] == b"BC\x02\x00"
        bsize = struct.unpack_from("<H", raw, o + 16)[0] + 1
        assert bsize <= 0x10000
        d = zlib.decompress(raw[o + 18:o + bsize - 8], -15)
        crc, isize = struct.unpack_from("<II", raw, o + bsize - 8)
        assert zlib.crc32(d) == crc and isize == len(d) <= layout.BGZF_PAYLOAD
        starts.append((o, u))
        o += bsize
        u += len(d)
    assert u == len(data)
    # .gzi = the start of every member after the first, EOF member excluded
    assert [tuple(map(int, p)) for p in pairs] == starts[1:-1]


def cases():
    rng = np.random.default_rng(5)
    runs1 = np.repeat(rng.integers(0, 256, 30000, dtype=np.uint8), rng.integers(1, 60, 30000))
    rows4 = np.repeat(rng.integers(0, 256, (20000, 4), dtype=np.uint8), rng.integers(1, 40, 20000), axis=0)
    rows16 = np.repeat(rng.integers(0, 256, (5000, 16), dtype=np.uint8), rng.integers(1, 40, 5000), axis=0)
    return [("empty", b"", 1), ("one byte", b"A", 1), ("random (stored members)", rng.integers(0, 256, 200000, dtype=np.uint8).tobytes(), 1),
            ("9-bit literals", rng.integers(144, 256, 70000, dtype=np.uint8).tobytes(), 3),
            ("constant", b"\xff" * 300000, 1), ("runs, 1-byte rows", runs1.tobytes(), 1), ("runs, 4-byte rows", rows4.tobytes(), 4),
            ("runs, 16-byte rows", rows16.tobytes(), 16), ("exactly one member", rows16.tobytes()[:0xff00], 16),
            ("one member + 1", rows16.tobytes()[:0xff00 + 1], 16), ("3 lanes + 7", rows16.tobytes()[:2040 * 3 + 7], 16),
            ("distance 9", rows4.tobytes()[:50000], 9), ("distance 300", rows16.tobytes()[:200000], 300)]


@pytest.mark.parametrize("label,data,dist", cases(), ids=[c[0] for c in cases()])
def test_host_compiled_encoder_produces_valid_bgzf(host_check, tmp_path, label, data, dist):
    (tmp_path / "in").write_bytes(data)
    subprocess.check_call([str(host_check), str(tmp_path / "in"), str(tmp_path / "o.gz"), str(tmp_path / "o.gzi"), str(dist)])
    raw, gzi = (tmp_path / "o.gz").read_bytes(), (tmp_path / "o.gzi").read_bytes()
    check_bgzf_image(raw, gzi, data)
    if len(data) > 3000:
        off = len(data) // 3
        assert layout.query_bytes(tmp_path / "o.gz", tmp_path / "o.gzi", off, 777) == data[off:off + 777]
    if label.startswith("runs") or label == "constant":
        assert len(raw) < len(data) / 5
    assert len(raw) <= len(data) + 31 * ((len(data) + 0xfeff) // 0xff00) + 28      # never worse than stored members
