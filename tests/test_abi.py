"""CPU: the C-ABI library loads, exports every symbol include/pk_anchor.h declares, and its
host-only entry points (KMC header parsing, bin rule) agree with the oracle. No compute calls."""
import ctypes as C
import re
from pathlib import Path

import pytest

from oracle import oracle
from panagram_b200 import _lib, kmc_api

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    src = (ROOT / "include" / "pk_anchor.h").read_text()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = C.CDLL(str(_lib.LIB_PATH))
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in pk_anchor.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes binding table and header are out of sync"
    assert _lib.lib().pk_abi_version() == 1


def test_struct_sizes():
    assert C.sizeof(_lib.PkConfig) == 44
    assert C.sizeof(_lib.PkKmcdbInfo) == 56
    assert C.sizeof(_lib.PkStats) == 80


@pytest.mark.parametrize("name", ["kmc_db", "kmc_db_sorted"])
def test_kmcdb_header_matches_oracle(kat, name):
    info = kmc_api.read_info(kat["dir"] / name)
    o = oracle.OracleDB.open(kat["dir"] / name)
    assert (info.kmer_length, info.counter_size, info.lut_prefix_length, info.signature_len, info.min_count,
            info.max_count, info.total_kmers, info.both_strands, info.mode) == \
           (o.k, o.counter_size, o.lut_prefix_len, o.signature_len, o.min_count, o.max_count, o.total_kmers,
            o.both_strands, o.mode)


def test_kmcdb_headers_of_pipeline_dbs(pan3):
    for stem in ("bitvec0", "g0.count", "g1.onehot"):
        info = kmc_api.read_info(pan3["dir"] / "kmc" / stem)
        o = oracle.OracleDB.open(pan3["dir"] / "kmc" / stem)
        assert info.total_kmers == o.total_kmers and info.kmer_length == 21
        assert info.max_count == o.max_count and info.both_strands


def test_open_errors_are_reported_not_raised_in_c(tmp_path):
    L = _lib.lib()
    h = C.c_void_p()
    assert L.pk_kmcdb_open(str(tmp_path / "nope").encode(), C.byref(h)) == -2       # PK_EIO
    assert b"nope" in L.pk_last_error()
    (tmp_path / "bad.kmc_pre").write_bytes(b"XXXX" + b"\0" * 100 + b"KMCP")
    assert L.pk_kmcdb_open(str(tmp_path / "bad").encode(), C.byref(h)) == -2
    with pytest.raises(_lib.PkError):
        kmc_api.read_info(tmp_path / "nope")
    assert kmc_api.KMCFile().OpenForRA(str(tmp_path / "nope")) is False             # OpenForRA -> false


def test_bin_len_rule():
    L = _lib.lib()
    for nk in (0, 99, 100, 199, 4321, 19_999_999, 20_000_000, 135_000_000):
        assert L.pk_bin_len(None, nk) == oracle.binlen(nk)
    cfg = _lib.PkConfig(21, 1, 0, 1, 0, 100, 50_000, 10, 0.5, 0, 0)
    assert L.pk_bin_len(C.byref(cfg), 1_000_000) == 50_000
    assert L.pk_bin_len(C.byref(cfg), 1000) == 100


def test_no_gpu_means_loud_failure():
    L = _lib.lib()
    if L.pk_device_count() > 0:
        pytest.skip("a GPU is visible")
    from panagram_b200.engine import Engine
    with pytest.raises(_lib.PkError) as ei:
        Engine(21, 8)
    assert ei.value.code == -3 and "no CPU path" in str(ei.value)
