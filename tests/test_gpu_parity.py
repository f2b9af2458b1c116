"""GPU: the CUDA path, called through the C ABI, against the reference's golden outputs and the
oracle. Bit-exact throughout (integer/byte work)."""
import numpy as np
import pytest

from oracle import oracle
from panagram_b200 import _lib, anchor, kmc_api, layout
from panagram_b200.engine import Engine, pinned_empty

pytestmark = pytest.mark.gpu


def anchor_outputs(eng, fasta, n_genomes, batch=False):
    """What cpp/anchor.cpp writes for one anchor, assembled from Engine.anchor_chrom (one call per
    chromosome) or Engine.anchor_genome (one batch)."""
    b1, lo, chroms, bins = [], [], [], []
    col = np.zeros(eng.n_local, dtype=np.uint64)
    recs = anchor.parse_fasta(fasta)
    if batch:
        res = eng.anchor_genome([s for _, s in recs])
        rs, col = res["chroms"], res["col_sums"]
    else:
        rs = [eng.anchor_chrom(seq) for _, seq in recs]
        for r in rs:
            col += r["col_sums"]
    for (name, _), r in zip(recs, rs):
        b1.append(r["bitmap1"].tobytes()); lo.append(r["low"].tobytes())
        chroms.append((name, r["nkmers"])); bins.append((r["binlen"], r["bin_hist"]))
    return {"bitmap.1": b"".join(b1), "bitmap.100": b"".join(lo), "chrs.tsv": layout.chrs_tsv(chroms),
            "bitsum.bins.tsv": layout.bins_tsv(n_genomes, bins), "col_sums": col}


def check_against_golden(eng, pan):
    for ai, a in enumerate(pan["anchors"]):
        got = anchor_outputs(eng, pan["fasta"][a], pan["n_genomes"], batch=bool(ai % 2))
        for key, want in pan["expected"][a].items():
            assert got[key] == want, f"{a}/{key}"
        rows = np.frombuffer(got["bitmap.1"], dtype=np.uint8).reshape(-1, eng.row_bytes)
        bits = np.unpackbits(rows, axis=1, bitorder="little")[:, :pan["n_genomes"]]
        assert (bits.sum(axis=0) == got["col_sums"]).all()       # index.py:1051 paircount sums


@pytest.mark.parametrize("which", ["pan3", "pan35"])
@pytest.mark.parametrize("mode,chunk", [("auto", 0), ("direct", 777), ("partitioned", 0), ("partitioned", 1000)])
def test_bitvec_db_engine_equals_reference(which, mode, chunk, request):
    pan = request.getfixturevalue(which)
    eng = Engine(pan["k"], pan["n_genomes"], chunk_positions=chunk, probe_mode=mode)
    for i in range(pan["ndb"]):
        eng.add_bitvec(32 * i, pan["dir"] / "kmc" / f"bitvec{i}")
    eng.finalize()
    check_against_golden(eng, pan)
    bk, bc = oracle.OracleDB.open(pan["dir"] / "kmc" / "bitvec0").list()
    for g in range(min(3, pan["n_genomes"])):
        assert eng.table_stats(g)["n_keys"] == int(((bc >> g) & 1).sum())


@pytest.mark.parametrize("kind", ["count", "onehot"])
def test_per_genome_kmc_dbs_equal_reference(pan3, kind):
    """K_g ingested from kmc/{s}.count (KMC2, written by kmc) or .onehot (KMC1, kmc_tools)."""
    eng = Engine(pan3["k"], 3, load_factor=0.85)
    for g, name in enumerate(pan3["names"]):
        eng.add_kmc(g, pan3["dir"] / "kmc" / f"{name}.{kind}")
    eng.finalize()
    check_against_golden(eng, pan3)


def test_tables_built_from_fasta_equal_reference(pan3):
    """On-GPU k-mer set construction straight from the sequences == kmc -ci1 (FASTA input).
    kmc drops control bytes (the fixture's CRLF) when reading, so the set is built that way;
    the anchor side keeps cpp/anchor.cpp's verbatim lines."""
    eng = Engine(pan3["k"], 3)
    for g, name in enumerate(pan3["names"]):
        recs = anchor.parse_fasta(pan3["fasta"][name], strip_cr=True)
        eng.reserve(g, sum(max(s.size - pan3["k"] + 1, 0) for _, s in recs))
        for _, s in recs:
            eng.add_sequence(g, s)
    eng.finalize()
    check_against_golden(eng, pan3)
    km, _ = oracle.OracleDB.open(pan3["dir"] / "kmc" / "g1.count").list()
    assert eng.table_stats(1)["n_keys"] == km.size


def test_kmcfile_dropin_known_answer(kat):
    """test_py_kmc_file.py::test_get_counters_for_read through the py_kmc_api mirror."""
    for name in ("kmc_db", "kmc_db_sorted"):
        f = kmc_api.KMCFile()
        assert f.OpenForRA(str(kat["dir"] / name))
        assert not f.OpenForRA(str(kat["dir"] / name))
        assert f.KmerLength() == 17 and f.Info().total_kmers > 0
        v = kmc_api.CountVec()
        assert f.GetCountersForRead(kat["query"], v)
        assert v.value == kat["counters"]
        assert np.array(v, dtype="uint32").tolist() == kat["counters"]
        assert not f.GetCountersForRead("ACGTACGT", v) and v.value == []
        assert f.Close() and not f.Close()


def random_case(rng, n_genomes, k, length, p_member=0.6):
    seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=length)
    for _ in range(3):
        o = int(rng.integers(0, max(length - 40, 1)))
        seq[o:o + int(rng.integers(1, 40))] = ord("N")
    seq[rng.integers(0, length, size=5)] = np.frombuffer(b"acgtn", dtype=np.uint8)
    sb = seq.tobytes()
    canon = sorted(oracle.kmer_set([sb], k))
    ints = np.array([oracle.kmer_int(s) for s in canon], dtype=np.uint64)
    extra = rng.integers(0, 1 << min(2 * k, 62), size=len(ints) // 2 + 1, dtype=np.uint64)
    member = rng.random((len(ints), n_genomes)) < p_member
    return sb, ints, member, extra


@pytest.mark.parametrize("n_genomes,k,load", [(1, 21, 0.5), (2, 5, 0.5), (8, 31, 0.9), (9, 32, 0.75), (33, 21, 0.9),
                                              (64, 17, 0.5), (70, 6, 0.5), (8, 24, 0.6), (3, 25, 0.6), (4, 14, 0.5),
                                              (2, 15, 0.9)])
def test_random_sets_vs_oracle(n_genomes, k, load):
    rng = np.random.default_rng(1000 * n_genomes + k)
    sb, ints, member, extra = random_case(rng, n_genomes, k, 6000)
    eng = Engine(k, n_genomes, load_factor=load, chunk_positions=1500)
    dbs = []
    for d in range((n_genomes + 31) // 32):
        cnt = np.zeros(len(ints), dtype=np.uint32)
        for g in range(32 * d, min(32 * d + 32, n_genomes)):
            cnt |= member[:, g].astype(np.uint32) << np.uint32(g - 32 * d)
        keep = cnt != 0
        dbs.append(oracle.OracleDB.from_kmers(k, ints[keep], cnt[keep]))
    for g in range(n_genomes):
        keys = ints[member[:, g]]
        eng.add_keys(g, keys)
        eng.add_keys(g, keys[: len(keys) // 3])       # duplicates are absorbed
    eng.finalize()
    for g in range(n_genomes):
        st = eng.table_stats(g)
        assert st["n_keys"] == int(member[:, g].sum())
    want = oracle.anchor_chrom(dbs, n_genomes, sb)
    got = eng.anchor_chrom(sb)
    assert (got["bitmap1"] == want["bitmap1"]).all()
    assert (got["low"] == want["bitmap100"]).all()
    assert (got["bin_hist"] == want["bin_hist"]).all()
    assert (got["col_sums"] == want["paircounts"]).all()
    for d, db in enumerate(dbs):
        assert (eng.get_counters_for_read(d, sb) == db.get_counters_for_read(sb)).all()


def test_edge_cases(pan3):
    eng = Engine(21, 3)
    eng.add_bitvec(0, pan3["dir"] / "kmc" / "bitvec0")
    eng.finalize()
    assert eng.anchor_chrom(b"ACGT")["nkmers"] == 0                        # len < k
    assert eng.get_counters_for_read(0, b"ACGTACGT") is None
    r = eng.anchor_chrom(b"N" * 500)                                         # nothing valid
    assert r["nkmers"] == 480 and not r["bitmap1"].any() and r["bin_hist"][:, 0].sum() == 480
    r = eng.anchor_chrom(b"ACGT" * 10, hist=False)                           # 20 k-mers: below min_bin_count
    assert r["nkmers"] == 20 and r["bin_hist"] is None and r["low"].shape == (1, 1)
    with pytest.raises(_lib.PkError):                                        # ... and asking for bins is an error
        import ctypes as C
        buf = np.zeros(64, dtype=np.uint64)
        seq = np.frombuffer(b"ACGT" * 10, dtype=np.uint8)
        _lib.check(eng._L.pk_anchor_chrom(eng._h, seq.ctypes.data, seq.size, None, None, buf.ctypes.data, None,
                                          C.byref(C.c_uint64())))
    # an engine with an empty genome still answers
    e2 = Engine(21, 2)
    e2.add_keys(0, np.array([5], dtype=np.uint64))
    e2.finalize()
    assert e2.table_stats(1)["n_keys"] == 0
    assert not e2.anchor_chrom(b"ACGTTGCA" * 20)["bitmap1"].any()
    # probing before finalize is a state error, not a crash
    e3 = Engine(21, 1)
    with pytest.raises(_lib.PkError) as ei:
        e3.anchor_chrom(b"A" * 200)
    assert ei.value.code == -6
    with pytest.raises(_lib.PkError):
        Engine(33, 1)                                                       # k > 32 unsupported
    with pytest.raises(_lib.PkError):
        eng2 = Engine(31, 3); eng2.add_bitvec(0, pan3["dir"] / "kmc" / "bitvec0")   # k mismatch


def test_pinned_buffers_and_anchor_dir(pan3, tmp_path):
    eng = Engine(pan3["k"], 3)
    eng.add_bitvec(0, pan3["dir"] / "kmc" / "bitvec0")
    eng.finalize()
    name, seq = anchor.parse_fasta(pan3["fasta"]["g0"])[0]
    nk = seq.size - 21 + 1
    out = {"bitmap1": pinned_empty((nk, 1)), "low": pinned_empty(((nk + 99) // 100, 1))}
    r = eng.anchor_chrom(seq, out=out)
    assert r["bitmap1"].tobytes() == pan3["expected"]["g0"]["bitmap.1"][:nk]
    s = anchor.anchor_fasta(eng, "g0", pan3["fasta"]["g0"], tmp_path / "anchor" / "g0", genome_names=pan3["names"])
    exp = pan3["expected"]["g0"]
    d = tmp_path / "anchor" / "g0"
    assert layout.read_bgzf(d / "bitmap.1.gz") == exp["bitmap.1"]
    assert layout.read_bgzf(d / "bitmap.100.gz") == exp["bitmap.100"]
    assert (d / "chrs.tsv").read_text() == exp["chrs.tsv"]
    assert (d / "bitsum.bins.tsv").read_text() == exp["bitsum.bins.tsv"]
    pc = (d / "total_paircounts.csv").read_text().splitlines()
    assert pc[0] == "name,count,frac" and pc[1].endswith(",1.0") and len(pc) == 4
    assert layout.query_bytes(d / "bitmap.1.gz", d / "bitmap.1.gzi", 1234, 50) == exp["bitmap.1"][1234:1284]
    assert s["positions"] == len(exp["bitmap.1"])


@pytest.mark.parametrize("bgzf", ["gpu", "zlib"])
def test_anchor_dir_both_bgzf_writers(pan3, tmp_path, bgzf):
    """The GPU BGZF writer and the host zlib writer give files that decompress to the reference's bytes and
    are seekable through their .gzi the way Genome._query_bytes seeks (index.py:827-845)."""
    from test_bgzf_format import check_bgzf_image
    eng = Engine(pan3["k"], 3)
    eng.add_bitvec(0, pan3["dir"] / "kmc" / "bitvec0")
    eng.finalize()
    for a in pan3["anchors"]:
        d = tmp_path / a
        anchor.anchor_fasta(eng, a, pan3["fasta"][a], d, genome_names=pan3["names"], bgzf=bgzf)
        exp = pan3["expected"][a]
        for step in (1, 100):
            raw, gzi = (d / f"bitmap.{step}.gz").read_bytes(), (d / f"bitmap.{step}.gzi").read_bytes()
            if bgzf == "gpu":
                check_bgzf_image(raw, gzi, exp[f"bitmap.{step}"])
            assert layout.read_bgzf(d / f"bitmap.{step}.gz") == exp[f"bitmap.{step}"]
        n = len(exp["bitmap.1"])
        for off in (0, 1, n // 2, n - 60):
            assert layout.query_bytes(d / "bitmap.1.gz", d / "bitmap.1.gzi", off, 60) == exp["bitmap.1"][off:off + 60]
        assert (d / "chrs.tsv").read_text() == exp["chrs.tsv"]
        assert (d / "bitsum.bins.tsv").read_text() == exp["bitsum.bins.tsv"]


def test_bgzf_compress_device_formats():
    """pk_bgzf_compress_device on device buffers: compressible rows of several widths, incompressible bytes
    (stored members), empty and ragged sizes; images validated member by member against zlib."""
    import torch
    from test_bgzf_format import cases, check_bgzf_image
    eng = Engine(21, 1)
    eng.add_keys(0, np.array([1], dtype=np.uint64))
    eng.finalize()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    big = np.repeat(np.random.default_rng(9).integers(0, 256, (400000, 2), dtype=np.uint8),
                    np.random.default_rng(10).integers(1, 30, 400000), axis=0).tobytes()
    for label, data, dist in cases() + [("12 MB of 2-byte rows", big, 2)]:
        n = len(data)
        cap_gz, cap_gzi = eng.bgzf_bound(n)
        d_in = torch.frombuffer(bytearray(data) or bytearray(1), dtype=torch.uint8).to(dev)
        d_gz = torch.zeros(cap_gz, dtype=torch.uint8, device=dev)
        d_gzi = torch.zeros(cap_gzi // 8, dtype=torch.int64, device=dev)
        d_tot = torch.zeros(2, dtype=torch.int64, device=dev)
        eng.bgzf_compress_device(d_in.data_ptr(), n, dist, d_gz.data_ptr(), d_gzi.data_ptr(), d_tot.data_ptr(), st)
        torch.cuda.synchronize()
        tot = d_tot.cpu().numpy()
        assert tot[0] <= cap_gz and tot[1] <= cap_gzi, label
        raw = d_gz[: int(tot[0])].cpu().numpy().tobytes()
        gzi = d_gzi.cpu().numpy().tobytes()[: int(tot[1])]
        check_bgzf_image(raw, gzi, data)
        if label.startswith("runs") or label.startswith("12 MB"):
            assert len(raw) < n / 5, label


@pytest.mark.parametrize("n_cols,row_stride", [(9, 2), (16, 2), (20, 4), (32, 4), (33, 8), (64, 8), (70, 16), (128, 16), (12, 4)])
def test_reduce_device_wide_rows(n_cols, row_stride):
    """pk_reduce_device on contiguous rows of 2..16 bytes (the vector kernel; the generic kernel where the bins are
    short) against numpy: popcount histograms per bin, column sums, low-res rows; aligned and unaligned starts,
    slices that begin in the middle of a bin, bits beyond n_cols set in the padding (must be ignored)."""
    import torch
    eng = Engine(21, 1)
    eng.add_keys(0, np.array([1], dtype=np.uint64))
    eng.finalize()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(n_cols * 100 + row_stride)
    nbytes = (n_cols + 7) // 8
    n_all = 700_001
    base = rng.integers(0, 256, (n_all // 5, row_stride), dtype=np.uint8)
    rows = np.repeat(base, rng.integers(1, 14, base.shape[0]), axis=0)[:n_all].copy()      # runs, like a real bitmap
    assert rows.shape[0] == n_all
    d_all = torch.from_numpy(rows).to(dev)
    bits_all = np.unpackbits(rows, axis=1, bitorder="little")[:, :n_cols]
    step = eng.lowres_step
    for first_row, p_first, n, binlen in ((0, 0, n_all, 200_000), (3, 150_003, 400_000, 200_000), (1, 65_537, 250_001, 65_536),
                                          (16, 1_000, 300_000, 1_000), (5, 7, 50, 200_000), (2, 99_999, 131_072, 100_000)):
        r = rows[first_row:first_row + n]
        bits = bits_all[first_row:first_row + n]
        nb_total = (p_first + n + binlen - 1) // binlen
        d_hist = torch.zeros(nb_total * (n_cols + 1), dtype=torch.int64, device=dev)
        d_col = torch.zeros(n_cols, dtype=torch.int64, device=dev)
        l0, l1 = (p_first + step - 1) // step, (p_first + n + step - 1) // step
        d_low = torch.zeros((max(l1 - l0, 1), row_stride), dtype=torch.uint8, device=dev)
        eng.reduce_device(d_all.data_ptr() + first_row * row_stride, row_stride, n_cols, p_first, n, binlen,
                          d_hist.data_ptr(), d_col.data_ptr(), d_low.data_ptr(), st)
        torch.cuda.synchronize()
        pc = bits.sum(axis=1).astype(np.int64)
        want = np.bincount(((p_first + np.arange(n, dtype=np.int64)) // binlen) * (n_cols + 1) + pc, minlength=nb_total * (n_cols + 1))
        assert (d_hist.cpu().numpy() == want).all(), (first_row, p_first, n, binlen)
        assert (d_col.cpu().numpy() == bits.sum(axis=0)).all(), (first_row, p_first, n, binlen)
        low_idx = np.arange(l0, l1) * step - p_first
        assert (d_low.cpu().numpy()[: l1 - l0, :nbytes] == r[low_idx][:, :nbytes]).all(), (first_row, p_first, n, binlen)
    eng.close()


def test_device_level_sharded_gather_equals_single_engine():
    """Two genome shards on one GPU stand in for two ranks: per-shard rows, gathered as planes,
    interleaved, reduced == one engine holding all 16 genomes."""
    import torch
    rng = np.random.default_rng(7)
    k, n = 23, 16
    sb, ints, member, _ = random_case(rng, n, k, 5000)
    full = Engine(k, n)
    shards = [Engine(k, n, 0, 8), Engine(k, n, 8, 16)]
    for g in range(n):
        for e in [full] + shards:
            e.add_keys(g, ints[member[:, g]])            # non-local genomes are ignored by a shard
    for e in [full] + shards:
        e.finalize()
    want = full.anchor_chrom(sb)
    nk = want["nkmers"]
    dev = torch.device("cuda:0")
    asc = torch.frombuffer(bytearray(sb), dtype=torch.uint8).to(dev)
    nw = full.packed_words(len(sb))
    words = torch.empty(nw, dtype=torch.int64, device=dev)
    mask = torch.empty(nw, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    planes = torch.empty((2, nk, 1), dtype=torch.uint8, device=dev)
    for r, e in enumerate(shards):
        e.pack_device(asc.data_ptr(), len(sb), words.data_ptr(), mask.data_ptr(), st)
        e.probe_device(words.data_ptr(), mask.data_ptr(), 0, nk, planes[r].data_ptr(), 1, 0, st)
    rows = torch.zeros((nk, 2), dtype=torch.uint8, device=dev)
    full.interleave_device(planes.data_ptr(), 2, nk, 1, rows.data_ptr(), 2, st)
    binlen = full.bin_len(nk)
    nbins = (nk + binlen - 1) // binlen
    hist = torch.zeros((nbins, n + 1), dtype=torch.int64, device=dev)
    cols = torch.zeros(n, dtype=torch.int64, device=dev)
    low = torch.zeros(((nk + 99) // 100, 2), dtype=torch.uint8, device=dev)
    full.reduce_device(rows.data_ptr(), 2, n, 0, nk, binlen, hist.data_ptr(), cols.data_ptr(), low.data_ptr(), st)
    torch.cuda.synchronize()
    assert (rows.cpu().numpy() == want["bitmap1"]).all()
    assert (low.cpu().numpy() == want["low"]).all()
    assert (hist.cpu().numpy().astype(np.uint64) == want["bin_hist"]).all()
    assert (cols.cpu().numpy().astype(np.uint64) == want["col_sums"]).all()
    # direct strided write (col_offset) gives the same rows without the interleave
    rows2 = torch.zeros((nk, 2), dtype=torch.uint8, device=dev)
    for r, e in enumerate(shards):
        e.probe_device(words.data_ptr(), mask.data_ptr(), 0, nk, rows2.data_ptr(), 2, r, st)
    torch.cuda.synchronize()
    assert torch.equal(rows, rows2)


@pytest.mark.parametrize("n_ranks,w", [(2, 1), (4, 1), (8, 1), (16, 1), (3, 1), (2, 2), (8, 2), (5, 3), (2, 4)])
def test_fused_gather_interleave_kernel(n_ranks, w):
    """pk_gather_interleave_device (the peer-memory exchange of the genome-sharded path): rows[i][r*w:(r+1)*w] =
    planes[r][i]. The planes are ordinary device buffers here — the kernel does not care whether a pointer is
    peer-mapped — so every layout branch (1 byte x 2 / 4k ranks, general) runs on one GPU, ragged sizes included."""
    import torch
    eng = Engine(21, 1)
    eng.add_keys(0, np.array([1], dtype=np.uint64))
    eng.finalize()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(n_ranks * 10 + w)
    for n in (1, 5, 4096, 100_003, 1_000_000):
        planes = [torch.from_numpy(rng.integers(0, 256, size=(n, w), dtype=np.uint8)).to(dev) for _ in range(n_ranks)]
        rows = torch.zeros((n, n_ranks * w), dtype=torch.uint8, device=dev)
        eng.gather_interleave_device([p.data_ptr() for p in planes], n, w, rows.data_ptr(), n_ranks * w, st)
        torch.cuda.synchronize()
        want = np.concatenate([p.cpu().numpy() for p in planes], axis=1)
        assert (rows.cpu().numpy() == want).all(), (n_ranks, w, n)


@pytest.mark.parametrize("n_ranks,w", [(1, 1), (2, 1), (4, 1), (8, 1), (16, 1), (3, 1), (2, 2), (8, 2), (5, 3), (2, 4), (4, 8), (2, 16)])
def test_gather_slice_kernel(n_ranks, w):
    """pk_gather_slice_device (the position-split exchange): rows[dst + i][q*w:(q+1)*w] = planes[q][src + i] for every
    segment. The planes are ordinary device buffers here — the kernel does not care whether a pointer is peer-mapped
    — so every store shape runs on one GPU: ragged segments, padded strides, a narrow last shard, the planes' last
    chunk. (The same per-chunk code runs on the host in tests/test_sharding.py.)"""
    import torch
    eng = Engine(21, 1)
    eng.add_keys(0, np.array([1], dtype=np.uint64))
    eng.finalize()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(n_ranks * 100 + w)
    for n in (1, 17, 4096, 100_003, 1_000_000):
        planes_h = [rng.integers(0, 256, size=(n, w), dtype=np.uint8) for _ in range(n_ranks)]
        planes = [torch.from_numpy(p).to(dev) for p in planes_h]
        for variant in range(3):
            nseg = int(rng.integers(1, 5))
            cuts = np.sort(rng.integers(0, n + 1, size=2 * nseg))
            segs, dst = [], 0
            for j in range(nseg):
                a, b = int(cuts[2 * j]), int(cuts[2 * j + 1])
                if variant == 2 and j == nseg - 1:
                    b = n                                  # reach the planes' last row
                segs.append((a, b - a, dst))
                dst += b - a
            rb = n_ranks * w
            if variant == 1 and w > 1:
                rb -= 1                                    # narrow last shard: the row is shorter than n_ranks * w
            stride = rb + (3 if variant == 1 else 0)
            rows = torch.full((max(dst, 1), stride), 0xAB, dtype=torch.uint8, device=dev)
            eng.gather_slice_device([p.data_ptr() for p in planes], n, w, segs, rows.data_ptr(), stride, rb, st)
            torch.cuda.synchronize()
            want = np.full((max(dst, 1), stride), 0xAB, dtype=np.uint8)
            cat = np.concatenate(planes_h, axis=1)[:, :rb]
            for a, m, d in segs:
                want[d:d + m, :rb] = cat[a:a + m]
            assert (rows.cpu().numpy() == want).all(), (n_ranks, w, n, variant, segs)


def _shard_engines(genomes, k, n, bounds, **kw):
    engs = [Engine(k, n, b, e, **kw) for b, e in bounds]
    for g, chroms in enumerate(genomes):
        for e in engs:
            if e.genome_begin <= g < e.genome_end:
                e.reserve(g, sum(s.size for _, s in chroms))
                for _, s in chroms:
                    e.add_sequence(g, s)
    for e in engs:
        e.finalize()
    return engs


@pytest.mark.parametrize("n,k,world", [(16, 21, 2), (35, 31, 2), (64, 21, 4)])
def test_position_split_exchange_on_one_gpu_equals_single_engine(n, k, world):
    """The genome-sharded product path with the ranks played by shard engines on ONE GPU (no NCCL, no IPC — those
    are covered by tests/test_multigpu.py on >= 2 GPUs): every shard probes the whole anchor into its plane
    (pk_anchor_genome_plane), every 'rank' assembles ITS slice out of all planes (pk_gather_slice_device), reduces the
    chromosome pieces inside it (pk_reduce_device), and deflates it (pk_bgzf_compress_device); the parts are merged
    with the host logic of sharded.py. Equal to what one engine holding all genomes delivers — the .gz byte for byte."""
    import torch
    from panagram_b200 import sharded, synth
    anc = synth.ancestor_codes(1_500_000 if n <= 35 else 700_000, 77)
    genomes = [synth.genome_chroms(anc, g, 77, n_chroms=3, n_run=300, lower_run=1000) for g in range(n)]
    full = _shard_engines(genomes, k, n, [(0, n)])[0]
    seqs = [s for _, s in genomes[1]]
    want = full.anchor_genome_bgzf(seqs)
    want_raw = full.anchor_genome(seqs)
    bounds = sharded.shard_bounds(n, world)
    shards = _shard_engines(genomes, k, n, bounds)
    w, rb, step = sharded.plane_width(n, world), (n + 7) // 8, 100
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    lens = [s.size for s in seqs]
    cat_off, plane_rows = full.anchor_layout(lens)
    nks = [l - k + 1 for l in lens]
    planes = [torch.zeros((plane_rows, w), dtype=torch.uint8, device=dev) for _ in range(world)]
    for e, pl in zip(shards, planes):
        assert e.anchor_genome_plane(seqs, pl.data_ptr(), plane_rows, w) == nks
    sb = sharded.slice_bounds(sum(nks), rb, world)
    binlen = [full.bin_len(nk) for nk in nks]
    nbins = [(nk + b - 1) // b for nk, b in zip(nks, binlen)]
    hoff = np.concatenate(([0], np.cumsum([nb * (n + 1) for nb in nbins])))
    nlow = [(nk + step - 1) // step for nk in nks]
    loff = np.concatenate(([0], np.cumsum(nlow)))
    red = torch.zeros(int(hoff[-1]) + n, dtype=torch.int64, device=dev)
    low = torch.zeros((int(loff[-1]), rb), dtype=torch.uint8, device=dev)
    parts = []
    with torch.cuda.stream(stream):
        st = stream.cuda_stream
        for r, e in enumerate(shards):
            s0, s1 = sb[r], sb[r + 1]
            rows = torch.zeros((max(s1 - s0, 1), rb), dtype=torch.uint8, device=dev)
            segs = sharded.stream_segments(cat_off, nks, s0, s1)
            if segs:
                e.gather_slice_device([p.data_ptr() for p in planes], plane_rows, w, segs, rows.data_ptr(), rb, rb, st)
            for c, p_first, m, r0 in sharded.slice_pieces(nks, s0, s1):
                l0 = (p_first + step - 1) // step
                e.reduce_device(rows.data_ptr() + r0 * rb, rb, n, p_first, m, binlen[c], red.data_ptr() + 8 * int(hoff[c]),
                                red.data_ptr() + 8 * int(hoff[-1]), low.data_ptr() + (int(loff[c]) + l0) * rb, st)
            nb = (s1 - s0) * rb
            cap_gz, cap_gzi = e.bgzf_bound(nb)
            gz = torch.empty(cap_gz, dtype=torch.uint8, device=dev)
            gzi = torch.empty(cap_gzi // 8 + 1, dtype=torch.int64, device=dev)
            tot = torch.zeros(2, dtype=torch.int64, device=dev)
            e.bgzf_compress_device(rows.data_ptr(), nb, rb, gz.data_ptr(), gzi.data_ptr(), tot.data_ptr(), st)
            t = tot.cpu()
            parts.append((gz[: int(t[0])].cpu().numpy(), gzi.view(torch.uint8)[: int(t[1])].cpu().numpy(), s0 * rb, nb,
                          rows[: s1 - s0].cpu().numpy()))
    torch.cuda.synchronize()
    rows_all = np.concatenate([p[4] for p in parts])
    assert rows_all.tobytes() == b"".join(c["bitmap1"].tobytes() for c in want_raw["chroms"])
    live = [r for r in range(world) if parts[r][3] > 0]
    offs, total = sharded.merge_bgzf_parts([(parts[r][0].size, parts[r][1].size) for r in live], [parts[r][2] for r in live])
    out = bytearray(total)
    for j, r in enumerate(live):
        d = parts[r][0] if j == len(live) - 1 else parts[r][0][:-sharded.BGZF_EOF_LEN]
        out[offs[j]:offs[j] + d.size] = d.tobytes()
    assert bytes(out) == want["gz"].tobytes()
    gzi = sharded.merge_gzi([parts[r][1].tobytes() for r in live], offs, [parts[r][2] for r in live], [parts[r][3] for r in live])
    assert gzi == want["gzi"].tobytes()
    red_h = red.cpu().numpy().astype(np.uint64)
    for c in range(len(nks)):
        assert (red_h[int(hoff[c]):int(hoff[c + 1])].reshape(nbins[c], n + 1) == want_raw["chroms"][c]["bin_hist"]).all()
    assert (red_h[int(hoff[-1]):] == want_raw["col_sums"]).all()
    assert low.cpu().numpy().tobytes() == b"".join(c["low"].tobytes() for c in want_raw["chroms"])


def test_sharded_anchorer_world1_directory_equals_anchor_fasta(pan3, tmp_path):
    """sharded.anchor_fasta_sharded with a world of one rank (plane -> slice gather -> piece-wise reduce -> BGZF of
    the slice -> collective file assembly) writes the directory anchor.anchor_fasta writes, which equals the
    reference's (golden) outputs."""
    from panagram_b200 import sharded
    sh = sharded.ShardedAnchorer(pan3["k"], 3, 0, 1, 0)
    sh.engine.add_bitvec(0, pan3["dir"] / "kmc" / "bitvec0")
    sh.engine.finalize()
    eng = Engine(pan3["k"], 3)
    eng.add_bitvec(0, pan3["dir"] / "kmc" / "bitvec0")
    eng.finalize()
    for a in pan3["anchors"]:
        sharded.anchor_fasta_sharded(sh, a, pan3["fasta"][a], tmp_path / "sh" / a, genome_names=pan3["names"])
        anchor.anchor_fasta(eng, a, pan3["fasta"][a], tmp_path / "one" / a, genome_names=pan3["names"])
        files = sorted(p.name for p in (tmp_path / "one" / a).iterdir())
        assert sorted(p.name for p in (tmp_path / "sh" / a).iterdir()) == files
        for f in files:
            assert (tmp_path / "sh" / a / f).read_bytes() == (tmp_path / "one" / a / f).read_bytes(), f
        exp = pan3["expected"][a]
        assert layout.read_bgzf(tmp_path / "sh" / a / "bitmap.1.gz") == exp["bitmap.1"]
        assert layout.read_bgzf(tmp_path / "sh" / a / "bitmap.100.gz") == exp["bitmap.100"]
        assert (tmp_path / "sh" / a / "bitsum.bins.tsv").read_text() == exp["bitsum.bins.tsv"]
    sh.close_p2p()


def big_case(n_genomes, k, length, seed, repeats=False):
    """A seeded pan-genome too large for the Python oracle loops but fine for the C oracle."""
    from panagram_b200 import synth
    anc = synth.ancestor_codes(length, seed)
    if repeats:      # skew: long low-complexity stretches hash to a handful of partitions
        anc[length // 4: length // 2] = np.tile(np.array([0, 1], dtype=np.uint8), length // 8 + 1)[: length // 2 - length // 4]
        anc[length // 2: length // 2 + length // 8] = 0
    return [synth.genome_chroms(anc, g, seed, n_chroms=3, n_run=700, lower_run=3000) for g in range(n_genomes)]


@pytest.mark.parametrize("n_genomes,k,repeats,load", [(8, 21, False, 0.5), (5, 31, True, 0.8), (40, 25, False, 0.6)])
def test_partitioned_path_large_vs_oracle_and_direct(n_genomes, k, repeats, load):
    """>= 1 Mi positions so the auto mode takes the partitioned path; the repeat-rich case overflows
    partition capacity and exercises the spill list."""
    length = 2_600_000 if n_genomes <= 8 else 1_300_000
    genomes = big_case(n_genomes, k, length, 31 + n_genomes, repeats)
    engs = {m: Engine(k, n_genomes, load_factor=load, probe_mode=m) for m in ("partitioned", "direct")}
    for g, chroms in enumerate(genomes):
        for e in engs.values():
            e.reserve(g, sum(s.size for _, s in chroms))
            for _, s in chroms:
                e.add_sequence(g, s)
    for e in engs.values():
        e.finalize()
    anchor_seqs = [s for _, s in genomes[1]]
    rp = engs["partitioned"].anchor_genome(anchor_seqs)
    rd = engs["direct"].anchor_genome(anchor_seqs)
    assert (rp["col_sums"] == rd["col_sums"]).all()
    for a, b in zip(rp["chroms"], rd["chroms"]):
        assert a["nkmers"] == b["nkmers"]
        assert (a["bitmap1"] == b["bitmap1"]).all()
        assert (a["low"] == b["low"]).all()
        assert (a["bin_hist"] == b["bin_hist"]).all()
    # the anchor's own column is set exactly where the window is valid (structural invariant)
    own = (rp["chroms"][0]["bitmap1"][:, 0] >> 1) & 1
    seq = anchor_seqs[0]
    bad = ~np.isin(seq, np.frombuffer(b"ACGTacgt", dtype=np.uint8))
    badwin = np.convolve(bad.astype(np.int32), np.ones(k, dtype=np.int32))[k - 1: seq.size] > 0
    assert (own == (~badwin).astype(np.uint8)).all()
    # C oracle on the first chromosome (keys listed from the sequences by the oracle's own code path)
    kk = {}
    for g, chroms in enumerate(genomes):
        for _, s in chroms:
            codes = np.full(256, 255, dtype=np.uint8)
            for i, c in enumerate(b"ACGT"):
                codes[c] = codes[c | 0x20] = i
            v = codes[s]
            ok = v < 4
            okwin = np.convolve((~ok).astype(np.int32), np.ones(k, dtype=np.int32))[k - 1: s.size] == 0
            f = np.zeros(s.size - k + 1, dtype=np.uint64)
            r = np.zeros(s.size - k + 1, dtype=np.uint64)
            vv = np.where(ok, v, 0).astype(np.uint64)
            for j in range(k):
                f = (f << np.uint64(2)) | vv[j: j + f.size]
                r = r | ((np.uint64(3) - vv[j: j + f.size]) << np.uint64(2 * j))
            can = np.minimum(f, r)[okwin]
            kk[g] = np.unique(np.concatenate([kk.get(g, np.zeros(0, dtype=np.uint64)), can]))
    allk = np.unique(np.concatenate(list(kk.values())))
    dbs = []
    for d in range((n_genomes + 31) // 32):
        cnt = np.zeros(allk.size, dtype=np.uint32)
        for g in range(32 * d, min(32 * d + 32, n_genomes)):
            cnt[np.searchsorted(allk, kk[g])] |= np.uint32(1 << (g - 32 * d))
        keep = cnt != 0
        dbs.append(oracle.OracleDB.from_kmers(k, allk[keep], cnt[keep]))
    for g in range(n_genomes):
        assert engs["partitioned"].table_stats(g)["n_keys"] == kk[g].size
    # group tables (8 genomes per table, built at finalize): exactly the distinct k-mers of each group
    for u in range((n_genomes + 7) // 8):
        gs = engs["partitioned"].group_stats(u)
        assert gs is not None and gs["n_keys"] == np.unique(np.concatenate([kk[g] for g in range(8 * u, min(8 * u + 8, n_genomes))])).size
    want = oracle.anchor_chrom(dbs, n_genomes, anchor_seqs[0].tobytes())
    assert (rp["chroms"][0]["bitmap1"] == want["bitmap1"]).all()
    assert (rp["chroms"][0]["bin_hist"] == want["bin_hist"]).all()


@pytest.mark.parametrize("n_genomes,k,load", [(8, 21, 0.5), (20, 31, 0.7)])
def test_k3_tuning_knobs_do_not_change_results(n_genomes, k, load):
    """The TMA-staged window kernel (every variant / stage count), the L1/L2 kernel and the direct kernel
    give identical rows; so does a batch whose table windows are too large for a shared-memory stage."""
    # k = 21: 6 M positions = 2^13 partitions, so that a partition's window of the (forced, 2^22-bucket) 32-bit-slot
    # group table fits one shared-memory stage and the lean window kernel takes the launch
    genomes = big_case(n_genomes, k, 6_000_000 if k == 21 else 1_400_000, 77 + n_genomes)
    eng = Engine(k, n_genomes, load_factor=load, probe_mode="partitioned")
    engd = Engine(k, n_genomes, load_factor=load, probe_mode="direct")
    for g, chroms in enumerate(genomes):
        for e in (eng, engd):
            e.reserve(g, sum(s.size for _, s in chroms))
            for _, s in chroms:
                e.add_sequence(g, s)
    eng.finalize(); engd.finalize()
    seqs = [s for _, s in genomes[2]]
    want = engd.anchor_genome(seqs)
    try:
        for knobs in (dict(k3_window=1, k3w_variant=-1, k3w_group=4), dict(k3w_variant=0, k3w_group=1),
                      dict(k3w_variant=1, k3w_group=2), dict(k3w_variant=2, k3w_group=4), dict(k3w_variant=3),
                      dict(k3_window=0), dict(k3_window=1, k3w_variant=-1, k3w_group=4, unpermute=0),
                      dict(unpermute=1, fine_out=0), dict(fine_out=0, k3_rank_atomic=0), dict(fine_out=1, k3_rank_atomic=1, k3w_variant=3),
                      dict(k3_variant=6, k3w_variant=4), dict(k3_variant=-1, k3w_variant=-1), dict(k1_roll=0), dict(k1_roll=1), dict(compact_items=0), dict(compact_items=1, k3w_big=5)):
            eng.tune(**knobs)
            got = eng.anchor_genome(seqs)
            assert (got["col_sums"] == want["col_sums"]).all(), knobs
            for a, b in zip(got["chroms"], want["chroms"]):
                assert (a["bitmap1"] == b["bitmap1"]).all(), knobs
                assert (a["bin_hist"] == b["bin_hist"]).all(), knobs
        # per-genome tables only (one probe per genome) vs group tables (one probe per 8 genomes)
        eng.tune(k3_window=1, k3w_variant=-1, k3w_group=0, unpermute=1, group_tables=0)
        eng.finalize()
        assert eng.group_stats(0) is None
        for knobs in (dict(k3_window=1), dict(k3_window=0)):
            eng.tune(**knobs)
            got = eng.anchor_genome(seqs)
            assert (got["col_sums"] == want["col_sums"]).all(), knobs
            for a, b in zip(got["chroms"], want["chroms"]):
                assert (a["bitmap1"] == b["bitmap1"]).all() and (a["bin_hist"] == b["bin_hist"]).all(), knobs
        eng.tune(k3_window=1, group_tables=1)
        eng.finalize()
        assert eng.group_stats(0)["n_keys"] >= max(eng.table_stats(g)["n_keys"] for g in range(min(8, n_genomes)))
        bytes64 = eng.group_stats(0)["bytes"]
        # the compact 32-bit slot format of the group tables, forced at this (small) table size: where k allows it
        # (k <= 23) the group table's own hash partitions the positions; window kernel, L1/L2 kernel, direct kernel
        eng.tune(group_tables=0)
        eng.tune(group_tables=1, group_g32=2)
        eng.finalize()
        if k <= 23:
            assert eng.group_stats(0)["n_buckets"] >= 1 << (2 * k - 20)
        else:
            assert eng.group_stats(0)["bytes"] == bytes64
        # k3_lean: the lean K3 (one 32-bit-slot group table, compact items, fine bins), every variant, and the general
        # window kernel (0) on the same launch; on other launches (k = 31: 64-bit slots) the knob changes nothing
        # k3_l2: the form without K2 (coarse regions probed through L2; the default), every block shape; 0: K2 + window kernels
        state = dict(k3_window=1, k3_lean=1, k3_l2=7)
        for knobs in (dict(k3_window=1), dict(k3_l2=0), dict(k3_window=0), dict(k3_window=1, k3_lean=0), dict(k3_lean=2), dict(k3_lean=3),
                      dict(k3_lean=4), dict(k3_lean=5), dict(k3_lean=6), dict(k3_lean=1), dict(k3_l2=1), dict(k3_l2=2), dict(k3_l2=3),
                      dict(k3_l2=4), dict(k3_l2=5), dict(k3_l2=6), dict(k3_l2=8), dict(k3_l2=9), dict(k3_l2=10), dict(k3_l2=7)):
            eng.tune(**knobs)
            state.update(knobs)
            got = eng.anchor_genome(seqs)
            if k == 21:
                want_kernel = 5 if state["k3_l2"] else 3 if not state["k3_window"] else 2 if state["k3_lean"] == 0 else 4
                assert int(eng.stats()["k_probe_window"]) == want_kernel, (knobs, state)
            gz = eng.anchor_genome_bgzf(seqs)
            assert (got["col_sums"] == want["col_sums"]).all() and (gz["col_sums"] == want["col_sums"]).all(), knobs
            for a, b in zip(got["chroms"], want["chroms"]):
                assert (a["bitmap1"] == b["bitmap1"]).all() and (a["bin_hist"] == b["bin_hist"]).all(), knobs
        short = seqs[0][:30_000]
        assert (eng.anchor_chrom(short, hist=False)["bitmap1"] == engd.anchor_chrom(short, hist=False)["bitmap1"]).all()
        for dbi in range((n_genomes + 31) // 32):
            assert (eng.get_counters_for_read(dbi, short) == engd.get_counters_for_read(dbi, short)).all()
        # ... and with the per-genome tables freed once the group tables hold their keys (group_only)
        eng.tune(k3_window=1, group_tables=0)
        eng.tune(group_tables=1, group_only=1)
        eng.finalize()
        assert all(eng.table_stats(g)["bytes"] == 0 for g in range(n_genomes))
        assert eng.table_stats(1)["n_keys"] == engd.table_stats(1)["n_keys"]
        for knobs in (dict(k3_window=1), dict(k3_window=0)):
            eng.tune(**knobs)
            got = eng.anchor_genome(seqs)
            for a, b in zip(got["chroms"], want["chroms"]):
                assert (a["bitmap1"] == b["bitmap1"]).all() and (a["bin_hist"] == b["bin_hist"]).all(), knobs
        assert (eng.anchor_chrom(short, hist=False)["bitmap1"] == engd.anchor_chrom(short, hist=False)["bitmap1"]).all()
        with pytest.raises(_lib.PkError):
            eng.add_sequence(0, seqs[0])             # sealed: its k-mers live in the group table only
        eng.tune(k3_window=1)
        # the same genome as 2 and 3 batches of whole chromosomes (copies of one batch under the kernels of the other)
        eng.tune(k3_window=1, k3w_variant=-1, k3w_group=0, unpermute=1, e2e_batch_min=0)
        for nb in (2, 3, 8):
            eng.tune(e2e_batches=nb)
            got = eng.anchor_genome(seqs)
            gz = eng.anchor_genome_bgzf(seqs)
            assert (got["col_sums"] == want["col_sums"]).all() and (gz["col_sums"] == want["col_sums"]).all(), nb
            for a, b, c in zip(got["chroms"], want["chroms"], gz["chroms"]):
                assert (a["bitmap1"] == b["bitmap1"]).all() and (a["low"] == b["low"]).all(), nb
                assert (a["bin_hist"] == b["bin_hist"]).all() and (c["bin_hist"] == b["bin_hist"]).all(), nb
            import gzip
            assert gzip.decompress(gz["gz"].tobytes()) == b"".join(c["bitmap1"].tobytes() for c in want["chroms"]), nb
        eng.tune(e2e_batches=2, e2e_batch_min=32 << 20)
        # a short anchor against the same (large) tables: few partitions -> windows of > 16 KB -> L1/L2 kernel
        eng.tune(k3_window=1, k3w_variant=-1, k3w_group=0, unpermute=1)
        short = seqs[0][:30_000]
        a = eng.anchor_chrom(short, hist=False)["bitmap1"]
        b = engd.anchor_chrom(short, hist=False)["bitmap1"]
        assert (a == b).all()
    finally:
        eng.tune(k3_window=1, k3w_variant=-1, k3w_group=0, unpermute=1, e2e_batches=2, e2e_batch_min=32 << 20)
    with pytest.raises(_lib.PkError):
        eng.tune(no_such_knob=1)


def test_panagram_index_cli_end_to_end(pan3, tmp_path, capsys):
    """`panagram index` from a samples TSV: k-mer sets built on the GPU from the FASTAs, every anchor
    directory byte-compatible with the reference's run_anchor output; bitdump reads it back."""
    from panagram_b200.cli import main
    from panagram_b200.index import make_bins_bits
    tsv = tmp_path / "samples.tsv"
    tsv.write_text("name\tfasta\n" + "".join(f"{n}\t{p}\n" for n, p in pan3["fasta"].items()))
    assert main(["index", str(tsv), "-k", "21", "-o", str(tmp_path / "idx"), "-c", "2"]) == 0
    for a in pan3["anchors"]:
        d = tmp_path / "idx" / "anchor" / a
        exp = pan3["expected"][a]
        assert layout.read_bgzf(d / "bitmap.1.gz") == exp["bitmap.1"]
        assert layout.read_bgzf(d / "bitmap.100.gz") == exp["bitmap.100"]
        assert (d / "chrs.tsv").read_text() == exp["chrs.tsv"]
        assert (d / "bitsum.bins.tsv").read_text() == exp["bitsum.bins.tsv"]
        assert (d / "total_paircounts.csv").exists() and (d / "bitmap.1.gzi").exists()
    # genome_dist.tsv (rule mash_triangle): exact Jaccard of the k-mer sets at this size -> Mash distance, lower triangle
    sets = {n: set(oracle.canonical_kmers([q for _, q in anchor.parse_fasta(p, strip_cr=True)], 21).tolist()) for n, p in pan3["fasta"].items()}
    lines = [l.split("\t") for l in (tmp_path / "idx" / "genome_dist.tsv").read_text().splitlines()]
    names = list(pan3["fasta"])
    assert [(l[0], l[1]) for l in lines] == [(names[i], names[j]) for i in range(1, 3) for j in range(i)]
    for f, t, d, p, x in lines:
        inter, union = len(sets[f] & sets[t]), len(sets[f] | sets[t])
        assert x == f"{inter}/{union}"
        assert abs(float(d) - layout.mash_distance(inter / union, 21)) < 1e-6 and 0.0 <= float(p) <= 1.0
    capsys.readouterr()
    assert main(["bitdump", str(tmp_path / "idx"), "g1", "chr2:100-110"]) == 0
    out = capsys.readouterr().out.split()
    rows = np.frombuffer(pan3["expected"]["g1"]["bitmap.1"], dtype=np.uint8)
    nk1 = int(pan3["expected"]["g1"]["chrs.tsv"].splitlines()[1].split("\t")[2])
    want = ["".join(str((int(rows[nk1 + p]) >> g) & 1) for g in range(3)) for p in range(100, 110)]
    assert out == want
    # scripts/make_bins_bits.py numbers == columns 1 and N of a popcount histogram over bitmap.100
    mb = make_bins_bits(tmp_path / "idx" / "anchor" / "g0", 3)
    low = np.frombuffer(pan3["expected"]["g0"]["bitmap.100"], dtype=np.uint8)
    pc = np.unpackbits(low.reshape(-1, 1), axis=1).sum(axis=1)
    assert mb[:, 0].sum() == (pc == 1).sum() and mb[:, 1].sum() == (pc == 3).sum()
    # an existing reference index (bitvec DBs under kmc/) is picked up instead of the FASTAs
    import shutil
    (tmp_path / "idx2" / "kmc").mkdir(parents=True)
    for f in (pan3["dir"] / "kmc").glob("bitvec0.*"):
        shutil.copy(f, tmp_path / "idx2" / "kmc" / f.name)
    assert main(["index", str(tsv), "-k", "21", "-o", str(tmp_path / "idx2"), "--anchor_genomes", "g2"]) == 0
    assert layout.read_bgzf(tmp_path / "idx2" / "anchor" / "g2" / "bitmap.1.gz") == pan3["expected"]["g2"]["bitmap.1"]
    assert not (tmp_path / "idx2" / "anchor" / "g0").exists()


# ---- S32 slot format: quotienting must stay exact -------------------------------------------------
@pytest.mark.parametrize("n_genomes,k,g32", [(3, 21, 0), (11, 13, 1), (20, 31, 0), (9, 21, 2), (8, 27, 0)])
def test_kmer_sample_and_paircount_bins(n_genomes, k, g32):
    """pk_engine_sample_kmers lists the k-mers of the group tables back — decoded from their slots, the key bits a slot
    does not store recovered from the home bucket (64-bit slots at k <= 26 and k >= 27, 32-bit slots natural at k = 13
    and forced at k = 21): all of them (== the sets the engine was given, with the right genome bits) and the
    hash-threshold sample (== the subset below the threshold). pk_anchor_paircount_bins == the host restatement of
    Index.bitmap_to_paircount_bins on the low-res rows."""
    rng = np.random.default_rng(100 + n_genomes)
    sb, ints, member, _ = random_case(rng, n_genomes, k, 60_000)
    eng = Engine(k, n_genomes)
    eng.tune(group_g32=g32)
    for g in range(n_genomes):
        eng.add_keys(g, ints[member[:, g]])
    eng.finalize()
    keys, tags = eng.sample_kmers(1.0)
    got = {}
    for x, t in zip(keys.tolist(), tags.tolist()):
        for b in range(8):
            if (t >> b) & 1:
                got.setdefault(8 * (t >> 8) + b, set()).add(x)
    for g in range(n_genomes):
        assert got.get(g, set()) == set(ints[member[:, g]].tolist()), g
    inter = layout.pair_counts([(keys, tags, 0)], n_genomes)
    for i, j in ((0, 0), (1, 0), (n_genomes - 1, 1)):
        assert inter[i, j] == int((member[:, i] & member[:, j]).sum())
    frac = 0.3
    k2, t2 = eng.sample_kmers(frac)
    hmax = int(frac * 2.0 ** 32)
    want = set(x for x in keys.tolist() if int(layout.kmer_sample_hash(np.array([x], dtype=np.uint64))[0]) < hmax)
    assert set(k2.tolist()) == want and 0.2 * len(set(keys.tolist())) < len(want) < 0.4 * len(set(keys.tolist()))
    # pair-count bins of the last anchored genome, reduced on the device
    res = eng.anchor_genome([sb[:41_000], sb[41_000:]])
    nks = [c["nkmers"] for c in res["chroms"]]
    for bin_size in (100000, 1000):
        pcs = eng.anchor_paircount_bins(nks, bin_size)
        for c, pc in zip(res["chroms"], pcs):
            starts, frac_h = layout.paircount_bins(c["low"], n_genomes, 100, bin_size)
            s2, frac_d = layout.paircount_frac(pc, bin_size)
            assert (starts == s2).all() and np.allclose(frac_h, frac_d)
            bits = np.unpackbits(c["low"], axis=1, bitorder="little")[:, :n_genomes]
            assert (pc.sum(axis=0) == bits.sum(axis=0)).all()


@pytest.mark.parametrize("n_genomes,k,length", [(32, 21, 20_000_000), (64, 31, 20_000_000)])
def test_many_genomes_at_20mbp_vs_c_oracle(n_genomes, k, length):
    """The default path (group tables — four of them at N=32, eight at N=64; partitioned probe; BGZF on the GPU) at
    a size where every table holds ~20 M k-mers, against the C oracle: per genome the k-mer set `kmc -ci1` would
    count (pko_kmers_of_seq) as an in-memory KMC1 database, queried with the restatement of
    CKMCFile::GetCountersForRead (kmc_file.cpp:954-1027) for EVERY position of the anchor. Bit-exact rows, column
    sums and popcount histograms."""
    import gzip
    from concurrent.futures import ThreadPoolExecutor
    from panagram_b200 import synth
    anc = synth.ancestor_codes(length, 4242 + n_genomes)
    anchor_g = 5
    eng = Engine(k, n_genomes)
    eng.tune(group_only=1)
    anchor_chroms = None
    dbs = [None] * n_genomes

    def build(g):
        chroms = synth.genome_chroms(anc, g, 4242 + n_genomes, n_chroms=3, n_run=700, lower_run=3000)
        keys = oracle.canonical_kmers([s for _, s in chroms], k)
        dbs[g] = oracle.OracleDB.from_kmers(k, keys, lut_prefix_len=11 if k == 31 else 9)     # (k - lut) % 4 == 0
        return chroms

    with ThreadPoolExecutor(8) as pool:
        for g0 in range(0, n_genomes, 8):
            for g, chroms in zip(range(g0, g0 + 8), pool.map(build, range(g0, min(g0 + 8, n_genomes)))):
                if g == anchor_g:
                    anchor_chroms = chroms
                eng.reserve(g, sum(s.size for _, s in chroms))
                for _, s in chroms:
                    eng.add_sequence(g, s)
                assert eng.table_stats(g) is not None
            eng.seal_group(g0 // 8)
    eng.finalize()
    for g in (0, 7, n_genomes - 1):
        assert eng.table_stats(g)["n_keys"] == dbs[g].total_kmers and eng.table_stats(g)["bytes"] == 0
    seqs = [s for _, s in anchor_chroms]
    got = eng.anchor_genome(seqs)
    gz = eng.anchor_genome_bgzf(seqs)
    rb = (n_genomes + 7) // 8
    col = np.zeros(n_genomes, dtype=np.uint64)
    for ci, s in enumerate(seqs):
        b = s.tobytes()
        with ThreadPoolExecutor(16) as pool:
            cols = list(pool.map(lambda g: dbs[g].get_counters_for_read(b), range(n_genomes)))
        bits = np.stack([(c != 0) for c in cols], axis=1)
        want = np.packbits(bits, axis=1, bitorder="little")
        assert want.shape == (s.size - k + 1, rb)
        rows = got["chroms"][ci]["bitmap1"]
        assert (rows == want).all(), f"chromosome {ci}: {int((rows != want).any(axis=1).sum())} rows differ from the C oracle"
        assert (got["chroms"][ci]["low"] == want[::100]).all()
        pc = bits.sum(axis=1)
        bl = got["chroms"][ci]["binlen"]
        nb = (pc.size + bl - 1) // bl
        hist = np.bincount((np.arange(pc.size) // bl) * (n_genomes + 1) + pc, minlength=nb * (n_genomes + 1)).reshape(nb, n_genomes + 1)
        assert (got["chroms"][ci]["bin_hist"] == hist.astype(np.uint64)).all()
        col += bits.sum(axis=0).astype(np.uint64)
        del bits, want, cols
    assert (got["col_sums"] == col).all() and (gz["col_sums"] == col).all()
    assert gzip.decompress(gz["gz"].tobytes()) == b"".join(c["bitmap1"].tobytes() for c in got["chroms"])


def _kmer_str(v: int, k: int) -> bytes:
    return bytes(b"ACGT"[(v >> (2 * (k - 1 - i))) & 3] for i in range(k))


def _s32_hash(canon: np.ndarray, k: int) -> np.ndarray:
    """numpy restatement of pk_key_hash (S32) for building adversarial key sets."""
    eb = max(0, 2 * k - 28)
    lo = (canon & np.uint64(0x0FFFFFFF)).astype(np.uint32)
    m = lo.copy()
    m ^= m >> np.uint32(16); m *= np.uint32(0x7FEB352D); m ^= m >> np.uint32(15); m *= np.uint32(0x846CA68B); m ^= m >> np.uint32(16)
    if eb == 0:
        return m
    hi = (canon >> np.uint64(28)).astype(np.uint32)
    f = (m * np.uint32(0x9E3779B1)) >> np.uint32(32 - eb)
    return ((hi ^ f) << np.uint32(32 - eb)) | (m >> np.uint32(eb))


def _canonical_keys(rng, k, n):
    """random k-mers that start and end with A: fwd < revcomp, so the integer is its own canonical form"""
    v = rng.integers(0, 1 << (2 * k - 4), size=n, dtype=np.uint64) << np.uint64(2)     # ...A at the end, A.. at the top
    return np.unique(v)


@pytest.mark.parametrize("k", [21, 24, 16])
def test_s32_same_low_bits_never_alias(k):
    """k-mers that agree in the 28 stored bits and differ only in the bits implied by the home bucket:
    present ones are found, absent ones are not (no false positives from quotienting)."""
    rng = np.random.default_rng(k)
    lo = _canonical_keys(rng, 14, 300) & np.uint64(0x0FFFFFFC)
    hi_bits = 2 * k - 28
    his = rng.integers(0, 1 << (hi_bits - 2), size=(lo.size, 6), dtype=np.uint64)       # top base A
    keys = np.unique((his << np.uint64(28)) | lo[:, None])
    member = rng.random(keys.size) < 0.5
    eng = Engine(k, 1, probe_mode="direct")
    eng.add_keys(0, keys[member])
    eng.finalize()
    assert eng.table_stats(0)["n_keys"] == int(member.sum())
    got = np.array([int(eng.get_counters_for_read(0, _kmer_str(int(v), k))[0]) for v in keys])
    assert (got == member.astype(int)).all()


def test_s32_overfull_neighbourhood_goes_to_the_stash():
    """> 15 buckets' worth of k-mers with one home bucket: the surplus lands in the stash and is still found,
    while colliding k-mers that were never inserted are not."""
    k = 21
    rng = np.random.default_rng(3)
    cand = _canonical_keys(rng, k, 3_000_000)
    nb = (1 << 14)                                    # the minimum table: n_buckets = 2^eb for k=21
    h = _s32_hash(cand, k)
    bucket = ((h.astype(np.uint64) * np.uint64(nb)) >> np.uint64(32)).astype(np.int64)
    target = np.bincount(bucket, minlength=nb).argmax()
    coll = cand[bucket == target]
    assert coll.size >= 170, coll.size
    ins, absent = coll[:150], coll[150:170]
    eng = Engine(k, 1, probe_mode="direct")
    eng.reserve(0, 1000)                              # -> n_buckets = 2^14 exactly
    eng.add_keys(0, ins)
    eng.finalize()
    st = eng.table_stats(0)
    assert st["n_buckets"] == nb and st["n_keys"] == ins.size
    for v in ins:
        assert eng.get_counters_for_read(0, _kmer_str(int(v), k))[0] == 1
    for v in absent:
        assert eng.get_counters_for_read(0, _kmer_str(int(v), k))[0] == 0
    # the same through the partitioned path: one long sequence holding all of them, N-separated
    seq = b"N".join(_kmer_str(int(v), k) for v in np.concatenate([ins, absent]))
    engp = Engine(k, 1, probe_mode="partitioned")
    engp.reserve(0, 1000)
    engp.add_keys(0, ins)
    engp.finalize()
    rows = engp.anchor_chrom(seq, hist=False)["bitmap1"][:, 0]
    starts = np.arange(ins.size + absent.size) * (k + 1)
    assert (rows[starts[:ins.size]] == 1).all() and (rows[starts[ins.size:]] == 0).all()


def test_full_size_configs1_properties():
    """BASELINE configs[1] at full size (8 x 135 Mbp, k=21, anchor = genome 0) through the public call:
    size-independent properties of the reference's output (SURVEY §4 [probed]) on all 135 M rows — the anchor's own
    column, zero rows exactly under invalid windows, low-res phase, every bin's popcount histogram, column sums — the
    BGZF images of the same rows, and bit-exact agreement of the default path (K1 -> probe_g32l2_kernel -> K4) with the
    direct kernel on a 3 M-position slice. The byte-for-byte comparison of this configuration with the UNMODIFIED
    reference (all 8 columns, all rows) is `bench.py --index-e2e configs1`
    (profiles/r2ag_index_configs1_full_vs_reference.json: it needs ~90 s of kmc + run_anchor, too long for this suite);
    32- and 64-genome cases are compared with the C oracle on every position in test_many_genomes_at_20mbp_vs_c_oracle."""
    from panagram_b200 import synth
    k, n, length, seed = 21, 8, 135_000_000, 20260001
    anc = synth.ancestor_codes(length, seed)
    eng = Engine(k, n)
    anchor_seqs = None
    for g in range(n):
        chroms = [s for _, s in synth.genome_chroms(anc, g, seed)]
        if g == 0:
            anchor_seqs = chroms
        eng.reserve(g, sum(c.size for c in chroms))
        for c in chroms:
            eng.add_sequence(g, c)
    eng.finalize()
    res = eng.anchor_genome(anchor_seqs)
    total = 0
    col = np.zeros(n, dtype=np.uint64)
    for seq, r in zip(anchor_seqs, res["chroms"]):
        rows = r["bitmap1"][:, 0]
        nk = seq.size - k + 1
        assert r["nkmers"] == nk and rows.size == nk
        bad = ~np.isin(seq, np.frombuffer(b"ACGTacgt", dtype=np.uint8))
        c = np.concatenate(([0], np.cumsum(bad, dtype=np.int64)))
        valid = (c[k:] - c[:-k]) == 0                                  # window holds only ACGTacgt
        assert ((rows & 1) == valid).all()                              # the anchor's own column
        assert ((rows == 0) == ~valid).all()                            # all-zero iff a non-ACGT byte in the window
        assert (r["low"][:, 0] == rows[::100]).all()                    # bitmap.100 = every 100th row per chromosome
        pc = np.unpackbits(rows[:, None], axis=1).sum(axis=1)
        binlen = r["binlen"]
        assert binlen == 200000 and r["bin_hist"].shape == ((nk + binlen - 1) // binlen, n + 1)
        assert r["bin_hist"].sum() == nk
        b3 = np.bincount(pc[3 * binlen:4 * binlen], minlength=n + 1)
        assert (r["bin_hist"][3] == b3).all()
        nbins = (nk + binlen - 1) // binlen                              # every bin, incl. the ragged last one
        full = np.bincount((np.arange(nk, dtype=np.int64) // binlen) * (n + 1) + pc.astype(np.int64), minlength=nbins * (n + 1))
        assert (r["bin_hist"].ravel() == full.astype(np.uint64)).all()
        col += np.unpackbits(rows[:, None], axis=1, bitorder="little").sum(axis=0).astype(np.uint64)
        total += nk
    assert (res["col_sums"] == col).all() and res["col_sums"][0] == col.max()
    assert total == 134_999_900
    # the same genome through the GPU BGZF writer: the file images decompress to exactly these rows
    import gzip
    rz = eng.anchor_genome_bgzf(anchor_seqs)
    assert (rz["col_sums"] == res["col_sums"]).all()
    assert gzip.decompress(rz["gz"].tobytes()) == b"".join(r["bitmap1"].tobytes() for r in res["chroms"])
    assert gzip.decompress(rz["gz_low"].tobytes()) == b"".join(r["low"].tobytes() for r in res["chroms"])
    nblk = (total + 0xff00 - 1) // 0xff00
    assert rz["gzi"].size == 8 + 16 * (nblk - 1) and int(rz["gzi"][:8].view(np.uint64)[0]) == nblk - 1
    for a, b in zip(rz["chroms"], res["chroms"]):
        assert (a["bin_hist"] == b["bin_hist"]).all()
    print("configs[1] bitmap.1: %d bytes -> %d bytes of BGZF (GPU)" % (total, rz["gz"].size))
    # partitioned (default for this size) == direct kernel on a 3 M-position slice
    engd = Engine(k, n, probe_mode="direct")
    for g in range(n):
        chroms = [s for _, s in synth.genome_chroms(anc, g, seed)]
        engd.reserve(g, sum(c.size for c in chroms))
        for c in chroms:
            engd.add_sequence(g, c)
    engd.finalize()
    sl = anchor_seqs[1][:3_000_000]
    rd = engd.anchor_chrom(sl, hist=False, low=False, colsums=False)["bitmap1"]
    assert (rd[:, 0] == res["chroms"][1]["bitmap1"][: rd.shape[0], 0]).all()
