"""Drop-in for the slice of ``py_kmc_api`` that panagram's anchoring path uses
(``KMC/py_kmc_api/py_kmc_api.cpp:24-108``, consumed at ``panagram/index.py:847-863,
932-947``): ``KMCFile().OpenForRA(prefix)``, ``.GetCountersForRead(seq, CountVec)``,
``.Info()``, ``.KmerLength()``, ``.Close()`` and ``CountVec``.

Same names, argument meaning and error behaviour (bool returns, never exceptions;
a read shorter than k clears the vector and returns False, kmc_file.cpp:878-882).
The database is interpreted the way the anchoring path uses it: the counter of a
"bitvec" database is a presence bit-vector (bit j <-> genome 32*i + j,
workflow/Snakefile:26-28), and lookups run on the GPU against per-genome tables.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PkError, PkKmcdbInfo, check
from .engine import Engine


class CountVec:
    """``py_kmc_api.CountVec``: a uint32 vector exposing the buffer protocol
    (py_kmc_api.cpp:31-43), so ``np.array(vec, dtype="uint32")`` works."""

    def __init__(self):
        self._a = np.empty(0, dtype=np.uint32)

    @property
    def value(self):
        return self._a.tolist()

    def __len__(self):
        return self._a.size

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype, copy=False)


class KMCFileInfo:
    """``CKMCFileInfo`` fields (kmc_file.h; py_kmc_api.cpp:66-79)."""

    def __init__(self, i: PkKmcdbInfo):
        self.kmer_length = i.kmer_length
        self.mode = i.mode
        self.counter_size = i.counter_size
        self.lut_prefix_length = i.lut_prefix_length
        self.signature_len = i.signature_len
        self.min_count = i.min_count
        self.max_count = i.max_count
        self.both_strands = bool(i.both_strands)
        self.total_kmers = i.total_kmers


def read_info(prefix) -> KMCFileInfo:
    L = _lib.lib()
    h = C.c_void_p()
    check(L.pk_kmcdb_open(str(prefix).encode(), C.byref(h)))
    try:
        info = PkKmcdbInfo()
        check(L.pk_kmcdb_info_get(h, C.byref(info)))
    finally:
        L.pk_kmcdb_close(h)
    return KMCFileInfo(info)


class KMCFile:
    def __init__(self, device: int = 0):
        self._engine = None
        self._info = None
        self._device = device

    def OpenForRA(self, prefix: str) -> bool:
        if self._engine is not None:
            return False                       # kmc_file.cpp:30-31: already open
        try:
            info = read_info(prefix)
            nbits = min(32, 8 * info.counter_size) if info.counter_size else 1
            eng = Engine(info.kmer_length, nbits, device=self._device)
            eng.add_bitvec(0, prefix)
            eng.finalize()
        except (PkError, OSError):
            return False
        self._engine, self._info = eng, info
        return True

    def GetCountersForRead(self, read, counters: CountVec) -> bool:
        if self._engine is None:
            return False
        out = self._engine.get_counters_for_read(0, read)
        if out is None:
            counters._a = np.empty(0, dtype=np.uint32)
            return False
        counters._a = out
        return True

    def KmerLength(self) -> int:
        return self._info.kmer_length if self._info else 0

    def Info(self) -> KMCFileInfo | None:
        return self._info

    def Close(self) -> bool:
        if self._engine is None:
            return False
        self._engine.close()
        self._engine = self._info = None
        return True
