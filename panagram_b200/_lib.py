"""ctypes loader for libpkanchor.so (include/pk_anchor.h).

There is no CPU implementation behind this module: if the CUDA library has not
been built (``python -c 'import __graft_entry__ as g; g.build()'`` or
``make -C panagram_b200/csrc``) importing the engine fails loudly.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libpkanchor.so"

PK_OK = 0
STATUS = {0: "PK_OK", -1: "PK_EINVAL", -2: "PK_EIO", -3: "PK_ECUDA", -4: "PK_ENOMEM",
          -5: "PK_EUNSUPPORTED", -6: "PK_ESTATE"}


class PkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


class PkConfig(C.Structure):
    _fields_ = [("k", C.c_uint32), ("n_genomes", C.c_uint32), ("genome_begin", C.c_uint32),
                ("genome_end", C.c_uint32), ("device", C.c_int32), ("lowres_step", C.c_uint32),
                ("max_bin_len", C.c_uint32), ("min_bin_count", C.c_uint32),
                ("load_factor", C.c_float), ("chunk_positions", C.c_uint32), ("probe_mode", C.c_uint32)]


class PkKmcdbInfo(C.Structure):
    _fields_ = [("kmer_length", C.c_uint32), ("mode", C.c_uint32), ("counter_size", C.c_uint32),
                ("lut_prefix_length", C.c_uint32), ("signature_len", C.c_uint32),
                ("kmc_version", C.c_uint32), ("both_strands", C.c_uint32), ("_pad", C.c_uint32),
                ("min_count", C.c_uint64), ("max_count", C.c_uint64), ("total_kmers", C.c_uint64)]


class PkTableStats(C.Structure):
    _fields_ = [("n_keys", C.c_uint64), ("n_buckets", C.c_uint64), ("n_overflow", C.c_uint64),
                ("bytes", C.c_uint64)]


class PkStats(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("pack_ms", C.c_float), ("probe_ms", C.c_float),
                ("reduce_ms", C.c_float), ("d2h_ms", C.c_float), ("total_ms", C.c_float),
                ("k_partition_ms", C.c_float), ("k_fine_ms", C.c_float), ("k_probe_ms", C.c_float),
                ("k_spill_ms", C.c_float), ("k_unpermute_ms", C.c_float), ("k_probe_window", C.c_float),
                ("positions", C.c_uint64), ("probes", C.c_uint64), ("probe_launches", C.c_uint64),
                ("kernel_launches", C.c_uint64)]


class PkSegment(C.Structure):
    _fields_ = [("src_row", C.c_uint64), ("n_rows", C.c_uint64), ("dst_row", C.c_uint64)]


# every symbol include/pk_anchor.h declares: name -> (restype, argtypes)
_vp, _cp, _u32, _u64, _sz = C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint64, C.c_size_t
_pp = C.POINTER(C.c_void_p)
_pu64 = C.POINTER(C.c_uint64)
SIGNATURES = {
    "pk_abi_version": (C.c_int, []),
    "pk_last_error": (_cp, []),
    "pk_device_count": (C.c_int, []),
    "pk_kmcdb_open": (C.c_int, [_cp, _pp]),
    "pk_kmcdb_info_get": (C.c_int, [_vp, C.POINTER(PkKmcdbInfo)]),
    "pk_kmcdb_close": (None, [_vp]),
    "pk_fasta_open": (C.c_int, [_cp, C.c_int, _pp]),
    "pk_fasta_n_records": (_u32, [_vp]),
    "pk_fasta_record": (C.c_int, [_vp, _u32, C.POINTER(_cp), _pp, _pu64]),
    "pk_fasta_close": (None, [_vp]),
    "pk_engine_create": (C.c_int, [C.POINTER(PkConfig), _pp]),
    "pk_engine_destroy": (None, [_vp]),
    "pk_engine_reserve": (C.c_int, [_vp, _u32, _u64]),
    "pk_engine_add_kmc": (C.c_int, [_vp, _u32, _cp]),
    "pk_engine_add_bitvec": (C.c_int, [_vp, _u32, _cp]),
    "pk_engine_add_keys": (C.c_int, [_vp, _u32, _vp, _u64]),
    "pk_engine_add_sequence": (C.c_int, [_vp, _u32, _vp, _u64]),
    "pk_engine_add_sequence_device": (C.c_int, [_vp, _u32, _vp, _u64]),
    "pk_engine_finalize": (C.c_int, [_vp]),
    "pk_engine_seal_group": (C.c_int, [_vp, _u32]),
    "pk_engine_table_stats": (C.c_int, [_vp, _u32, C.POINTER(PkTableStats)]),
    "pk_engine_group_stats": (C.c_int, [_vp, _u32, C.POINTER(PkTableStats)]),
    "pk_engine_sample_kmers": (C.c_int, [_vp, _u32, _vp, _vp, _u64, _pu64]),
    "pk_bin_len": (_u64, [C.POINTER(PkConfig), _u64]),
    "pk_anchor_genome": (C.c_int, [_vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pk_anchor_chrom": (C.c_int, [_vp, _vp, _u64, _vp, _vp, _vp, _vp, _pu64]),
    "pk_anchor_layout": (_u64, [_u32, _vp, _vp]),
    "pk_anchor_genome_plane": (C.c_int, [_vp, _u32, _vp, _vp, _vp, _u64, _u32, _vp]),
    "pk_get_counters_for_read": (C.c_int, [_vp, _u32, _vp, _u64, _vp, _pu64]),
    "pk_host_alloc": (C.c_int, [_pp, _sz]),
    "pk_host_free": (C.c_int, [_vp]),
    "pk_packed_words": (_u64, [_u64]),
    "pk_pack_device": (C.c_int, [_vp, _vp, _u64, _vp, _vp, _vp]),
    "pk_probe_device": (C.c_int, [_vp, _vp, _vp, _u64, _u64, _vp, _u32, _u32, _vp]),
    "pk_reduce_device": (C.c_int, [_vp, _vp, _u32, _u32, _u64, _u64, _u64, _vp, _vp, _vp, _u32, _vp]),
    "pk_interleave_device": (C.c_int, [_vp, _vp, _u32, _u64, _u32, _vp, _u32, _vp]),
    "pk_paircount_bins_device": (C.c_int, [_vp, _vp, _u32, _u32, _u64, _u32, _vp, _vp]),
    "pk_anchor_paircount_bins": (C.c_int, [_vp, _u32, _vp, _u32, _vp]),
    "pk_engine_stats": (C.c_int, [_vp, C.POINTER(PkStats)]),
    "pk_engine_tune": (C.c_int, [_vp, _cp, C.c_int]),
    "pk_bgzf_bound": (_u64, [_u64]),
    "pk_bgzf_gzi_bound": (_u64, [_u64]),
    "pk_bgzf_compress_device": (C.c_int, [_vp, _vp, _u64, _u32, _vp, _vp, _vp, _vp]),
    "pk_anchor_genome_bgzf": (C.c_int, [_vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pk_device_alloc": (C.c_int, [_vp, _pp, _sz]),
    "pk_device_free": (C.c_int, [_vp, _vp]),
    "pk_ipc_export": (C.c_int, [_vp, _vp, _vp]),
    "pk_ipc_open": (C.c_int, [_vp, _vp, _pp]),
    "pk_ipc_close": (C.c_int, [_vp, _vp]),
    "pk_gather_interleave_device": (C.c_int, [_vp, _vp, _u32, _u64, _u32, _vp, _u32, _vp]),
    "pk_gather_slice_device": (C.c_int, [_vp, _vp, _u32, _u64, _u32, _vp, _u32, _vp, _u32, _u32, _vp]),
}

_LIB = None


def build(verbose: bool = False) -> Path:
    """Compile libpkanchor.so in-tree with nvcc for sm_100a."""
    r = subprocess.run(["make", "-j6", "-C", str(PKG_DIR / "csrc")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"building libpkanchor.so failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stdout)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} is missing: the CUDA library has not been built "
                              "(run `make -C panagram_b200/csrc`); panagram_b200 has no CPU fallback")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.pk_abi_version() != 1:
            raise ImportError("libpkanchor.so ABI version mismatch")
        _LIB = L
    return _LIB


def check(rc: int) -> None:
    if rc != PK_OK:
        raise PkError(rc, lib().pk_last_error().decode(errors="replace"))
