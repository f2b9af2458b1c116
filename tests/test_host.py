"""CPU: host-side logic around the kernel — FASTA parsing, BGZF/.gzi layout, text writers,
synthetic genomes, genome sharding."""
import gzip
import struct

import numpy as np
import pytest

from oracle import oracle
from panagram_b200 import anchor, layout, synth


def test_parse_fasta_matches_reference_rules(pan3):
    for g, p in pan3["fasta"].items():
        want = oracle.parse_fasta(p)
        got = anchor.parse_fasta(p)
        assert [n for n, _ in got] == [n for n, _ in want]
        for (_, a), (_, b) in zip(got, want):
            assert a.tobytes() == b
        assert any(b"\r" in s for _, s in want), "fixture carries a CRLF line"
        stripped = anchor.parse_fasta(p, strip_cr=True)
        assert all(13 not in s for _, s in stripped)
        assert [n for n, _ in got][1] == "chr2"      # header "chr2 description text" cut at the space


def test_parse_fasta_gz_and_no_trailing_newline(tmp_path):
    txt = b">a b c\nACGT\nAC\n>b\n\nGG\nTT"
    (tmp_path / "x.fa").write_bytes(txt)
    with gzip.open(tmp_path / "x.fa.gz", "wb") as fh:
        fh.write(txt)
    for p in ("x.fa", "x.fa.gz"):
        recs = anchor.parse_fasta(tmp_path / p)
        assert [(n, s.tobytes()) for n, s in recs] == [("a", b"ACGTAC"), ("b", b"GGTT")]


@pytest.mark.parametrize("n", [0, 1, 0xFF00 - 1, 0xFF00, 0xFF00 + 1, 300_000])
def test_bgzf_roundtrip_and_gzi(tmp_path, n):
    rng = np.random.default_rng(n)
    data = (rng.integers(0, 4, size=n, dtype=np.uint8) * 85).tobytes()
    w = layout.BgzfWriter(tmp_path / "b.gz", threads=3)
    for o in range(0, n, 70_001):            # ragged writes
        w.write(data[o:o + 70_001])
    w.close(tmp_path / "b.gzi")
    raw = (tmp_path / "b.gz").read_bytes()
    assert raw.endswith(layout.BGZF_EOF)
    assert gzip.decompress(raw) == data
    gzi = (tmp_path / "b.gzi").read_bytes()
    (cnt,) = struct.unpack_from("<Q", gzi)
    # htslib's write-mode index (cpp/anchor.cpp:47,102; pk_bgzf.cu): one entry per member AFTER the first, none for EOF
    assert cnt == max((n + 0xFF00 - 1) // 0xFF00 - 1, 0) and len(gzi) == 8 + 16 * cnt
    blocks = layout.load_bgz_blocks(tmp_path / "b.gzi")
    assert blocks[0].tolist() == [0, 0]
    assert (np.diff(blocks[:, 1]) <= 0xFF00).all()
    for start, ln in ((0, 10), (0xFF00 - 3, 7), (n // 2, 1000), (max(n - 5, 0), 5)):
        if start + ln <= n:
            assert layout.query_bytes(tmp_path / "b.gz", tmp_path / "b.gzi", start, ln) == data[start:start + ln]


def test_text_writers_match_reference_text(pan3):
    exp = pan3["expected"]["g0"]
    chroms = [(l.split("\t")[0], int(l.split("\t")[2])) for l in exp["chrs.tsv"].splitlines()[1:]]
    assert layout.chrs_tsv(chroms) == exp["chrs.tsv"]
    rows = [l.split("\t") for l in exp["bitsum.bins.tsv"].splitlines()[1:]]
    per = []
    for cid in range(len(chroms)):
        rr = [r for r in rows if int(r[0]) == cid]
        binlen = int(rr[1][1]) - int(rr[0][1])
        per.append((binlen, np.array([[int(x) for x in r[2:]] for r in rr], dtype=np.uint64)))
    assert layout.bins_tsv(3, per) == exp["bitsum.bins.tsv"]


def test_paircounts_csv():
    txt = layout.paircounts_csv(["a", "b", "c"], np.array([10, 5, 3], dtype=np.uint64), "a")
    assert txt == "name,count,frac\na,10,1.0\nb,5,0.5\nc,3,0.3\n"


def test_synth_is_deterministic_and_shaped():
    anc = synth.ancestor_codes(50_000, 20260001)
    a = synth.genome_chroms(anc, 3, 20260001)
    b = synth.genome_chroms(synth.ancestor_codes(50_000, 20260001), 3, 20260001)
    assert all((x[1] == y[1]).all() for x, y in zip(a, b))
    assert [n for n, _ in a] == ["chr1", "chr2", "chr3", "chr4", "chr5"]
    assert (a[0][1][5000:6000] == ord("N")).all()
    assert bytes(a[1][1][:100]).islower()
    g0 = synth.genome_codes(anc, 0, 1)
    assert 0.001 < (g0 != anc).mean() < 0.003
    # bench.py's weak-scaling shards: with mu_period = shard width, genome g of every shard has the divergence RATE of
    # genome g of shard 0 (same table size, same work per GPU) but its own mutations
    big = synth.ancestor_codes(400_000, 7)
    g8_default, g8_shard = synth.genome_codes(big, 8, 7), synth.genome_codes(big, 8, 7, mu_period=8)
    assert 0.016 < (g8_default != big).mean() < 0.020 and 0.0015 < (g8_shard != big).mean() < 0.0025
    assert (g8_shard != synth.genome_codes(big, 0, 7, mu_period=8)).any()
    assert (synth.genome_codes(big, 3, 7) == synth.genome_codes(big, 3, 7, mu_period=8)).all()


def test_index_prepare_writes_reference_config_files(pan3, tmp_path):
    """`panagram index samples.tsv -k 21 --prepare` (index.py:269-295,347-353): samples.tsv with
    name/fasta/gff/id/anchor and a config.yaml carrying the keys the reference's reader needs."""
    import yaml
    from panagram_b200.cli import main
    tsv = tmp_path / "samples.tsv"
    tsv.write_text("name\tfasta\n" + "".join(f"{n}\t{p}\n" for n, p in pan3["fasta"].items()))
    assert main(["index", str(tsv), "-k", "21", "-o", str(tmp_path / "idx"), "--prepare",
                 "--anchor_genomes", "g0", "g2"]) == 0
    rows = [l.split("\t") for l in (tmp_path / "idx" / "samples.tsv").read_text().splitlines()]
    assert rows[0] == ["name", "fasta", "gff", "id", "anchor"]
    assert [(r[0], r[3], r[4]) for r in rows[1:]] == [("g0", "0", "True"), ("g1", "1", "False"), ("g2", "2", "True")]
    cfg = yaml.safe_load((tmp_path / "idx" / "config.yaml").read_text())
    for key in ("k", "lowres_step", "anchor_genomes", "gff_anno_types", "gff_gene_types", "gff_name", "max_bin_kbp",
                "min_bin_count"):
        assert key in cfg
    assert cfg["k"] == 21 and cfg["anchor_genomes"] == ["g0", "g2"]
    assert not (tmp_path / "idx" / "anchor").exists()
    bad = tmp_path / "bad.tsv"
    bad.write_text("name\tfasta\nbad name!\tx.fa\n")
    with pytest.raises(ValueError):
        main(["index", str(bad), "--prepare"])


def test_native_fasta_reader_equals_the_numpy_parser(tmp_path, pan3):
    """pk_fasta_open (csrc/pk_fasta.cpp) follows the same rules as anchor.parse_fasta_numpy — the C++ reference's
    (cpp/anchor.cpp:74-100: name to the first space, lines verbatim) and, with strip_cr, kmc's/Biopython's."""
    from panagram_b200 import anchor
    cases = {
        "plain": b">chr1 desc here\nACGT\nNNAC\n>chr2\nacgtn\n",
        "crlf": b">chr1 x\r\nACGT\r\nAC\rGT\r\n>c2\r\nTT\r\n",
        "no_trailing_newline": b">a\nACGT\n>b\nGG",
        "leading_garbage_and_blank_lines": b"junk\n\n>a\n\nAC\n\nGT\n>b\n",
        "header_only_and_empty_name": b">\nACGT\n> spaced\nTT\n>\tTabbed name\nGG\n",
        "control_bytes": b">a\nAC\x01GT\x0bTT\n>b\tx y\nA\tC\n",
        "empty": b"",
        "no_records": b"ACGT\nACGT\n",
        "gt_inside_line": b">a\nAC>GT\n>b\nTT\n",
    }
    for label, data in cases.items():
        p = tmp_path / f"{label}.fa"
        p.write_bytes(data)
        for strip in (False, True):
            got = anchor.parse_fasta(p, strip_cr=strip)
            want = anchor.parse_fasta_numpy(p, strip_cr=strip)
            assert [n for n, _ in got] == [n for n, _ in want], (label, strip)
            assert [s.tobytes() for _, s in got] == [s.tobytes() for _, s in want], (label, strip)
    for name, path in pan3["fasta"].items():          # the golden fixtures (CRLF, IUPAC, lowercase, N runs)
        for strip in (False, True):
            got, want = anchor.parse_fasta(path, strip_cr=strip), anchor.parse_fasta_numpy(path, strip_cr=strip)
            assert [(n, s.tobytes()) for n, s in got] == [(n, s.tobytes()) for n, s in want]
    import gzip
    gz = tmp_path / "x.fa.gz"
    gz.write_bytes(gzip.compress(cases["plain"]))
    assert [(n, s.tobytes()) for n, s in anchor.parse_fasta(gz)] == [("chr1", b"ACGTNNAC"), ("chr2", b"acgtn")]
    disguised = tmp_path / "y.fa"                      # gzip magic without the suffix: handed to the numpy path
    disguised.write_bytes(gzip.compress(cases["plain"]))
    assert [(n, s.tobytes()) for n, s in anchor.parse_fasta(disguised)] == [("chr1", b"ACGTNNAC"), ("chr2", b"acgtn")]


def test_umap_inputs_and_csv_match_the_reference_restated_in_pandas():
    """layout.paircount_bins == Index.bitmap_to_paircount_bins (index.py:454-459) restated with the reference's own
    pandas calls; umap CSV text = DataFrame(...).to_csv() of Genome.run_umap's zero-fill fallback (index.py:1149-1156)."""
    import pandas as pd
    from panagram_b200 import layout
    rng = np.random.default_rng(11)
    for n_genomes, nrows, step, bin_size in ((8, 5000, 100, 100000), (3, 777, 100, 100000), (35, 2500, 100, 50000), (9, 1, 100, 100000)):
        nb = (n_genomes + 7) // 8
        rows = rng.integers(0, 256, size=(nrows, nb), dtype=np.uint8)
        rows[rng.random(nrows) < 0.1] = 0
        starts, frac = layout.paircount_bins(rows, n_genomes, step, bin_size)
        bits = np.unpackbits(rows, axis=1, bitorder="little")[:, :n_genomes]
        bitmap = pd.DataFrame(bits, index=np.arange(0, nrows * step, step))
        df = bitmap.set_index(bitmap.index // bin_size)
        pc = df.groupby(level=0).sum()
        pc = pc.set_index(pc.index * bin_size).T
        pc = pc.div(pc.max(axis=0), axis=1).T.fillna(0)
        assert list(pc.index) == list(starts)
        assert np.allclose(pc.to_numpy(dtype=float), frac)
        # the fallback rows and their CSV text
        urows = layout.umap_rows("chrA", starts, frac, bin_size)
        want = pd.DataFrame({"umap1": 0, "umap2": 0, "cluster": 0}, index=pd.MultiIndex.from_product([["chrA"], starts], names=["chrom", "start"])).reset_index()
        want["end"] = want["start"] + bin_size
        want = want[["chrom", "start", "end", "umap1", "umap2", "cluster"]]
        try:
            import umap  # noqa: F401
        except ImportError:
            assert layout.umaps_csv(urows) == want.to_csv(index=False)
            assert layout.umaps_csv(urows) == want.set_index("chrom").to_csv()


def test_s32_hash_is_invertible_from_slot_and_home_bucket():
    """The group-table merge (csrc/pk_kernels.cu::union_merge_kernel) re-derives a k-mer from an S32 slot: low 28 bits
    from the slot, the other 2k-28 bits from the slot's HOME bucket by inverting pk_key_hash — the hash is
    (X << (32-eb)) | (mix32(lo) >> eb) for the one X whose hash maps to that bucket. Restated in numpy and checked for
    the table sizes the engine produces (>= 2^eb buckets)."""
    rng = np.random.default_rng(1)

    def mix32(x):
        x = x.astype(np.uint32).copy()
        x ^= x >> np.uint32(16); x *= np.uint32(0x7FEB352D); x ^= x >> np.uint32(15); x *= np.uint32(0x846CA68B); x ^= x >> np.uint32(16)
        return x

    for k, nb in [(21, 33748694), (21, 1 << 14), (21, 16385), (24, 1 << 20), (24, 40000000), (16, 1000), (15, 77)]:
        eb = max(0, 2 * k - 28)
        canon = rng.integers(0, 1 << (2 * k), size=100000, dtype=np.uint64)
        lo = (canon & np.uint64(0x0FFFFFFF)).astype(np.uint32)
        m = mix32(lo)
        hi = (canon >> np.uint64(28)).astype(np.uint32)
        f = (m * np.uint32(0x9E3779B1)) >> np.uint32(32 - eb)
        h = ((hi ^ f) << np.uint32(32 - eb)) | (m >> np.uint32(eb))                      # pk_key_hash, S32, eb > 0
        home = (h.astype(np.uint64) * np.uint64(nb)) >> np.uint64(32)
        sh, low = 32 - eb, m >> np.uint32(eb)
        hmin = ((home << np.uint64(32)) + np.uint64(nb - 1)) // np.uint64(nb)            # smallest hash of the home bucket
        X = (hmin >> np.uint64(sh)).astype(np.uint32)
        h2 = (X << np.uint32(sh)) | low
        miss = ((h2.astype(np.uint64) * np.uint64(nb)) >> np.uint64(32)) != home
        X = np.where(miss, (X + np.uint32(1)) & np.uint32((1 << eb) - 1), X)
        h3 = (X << np.uint32(sh)) | low
        assert (((h3.astype(np.uint64) * np.uint64(nb)) >> np.uint64(32)) == home).all()
        back = ((X ^ f).astype(np.uint64) << np.uint64(28)) | lo.astype(np.uint64)
        assert (back == canon).all() and (h3 == h).all(), (k, nb)


def test_genome_dist_from_kmer_samples():
    """layout.pair_counts / genome_dist_tsv (genome_dist.tsv, rule mash_triangle, workflow/Snakefile:139-149; consumer
    figs.py:50-59) against brute-force set arithmetic, records from two engines (genome shards [0,16) and [16,19))."""
    rng = np.random.default_rng(1)
    n = 19
    univ = np.unique(rng.integers(0, 2 ** 42, size=5000, dtype=np.uint64))
    sets = [set(univ[rng.random(univ.size) < p].tolist()) for p in rng.random(n)]
    recs = []
    for begin, end in ((0, 16), (16, 19)):
        keys, tags = [], []
        for grp in range((end - begin + 7) // 8):
            gs = list(range(begin + 8 * grp, min(begin + 8 * grp + 8, end)))
            for x in set().union(*[sets[g] for g in gs]):
                keys.append(x)
                tags.append((grp << 8) | sum(1 << b for b, g in enumerate(gs) if x in sets[g]))
        recs.append((np.array(keys, dtype=np.uint64), np.array(tags, dtype=np.uint32), begin))
    inter = layout.pair_counts(recs, n)
    for i in range(n):
        for j in range(n):
            assert inter[i, j] == len(sets[i] & sets[j])
    names = [f"g{i}" for i in range(n)]
    lines = [l.split("\t") for l in layout.genome_dist_tsv(names, inter, 21).splitlines()]
    assert len(lines) == n * (n - 1) // 2 and all(len(l) == 5 for l in lines)
    assert [(l[0], l[1]) for l in lines[:3]] == [("g1", "g0"), ("g2", "g0"), ("g2", "g1")]       # mash triangle order
    for f, t, d, p, x in lines:
        i, j = names.index(f), names.index(t)
        u = len(sets[i] | sets[j])
        assert x == f"{len(sets[i] & sets[j])}/{u}"
        jac = len(sets[i] & sets[j]) / u if u else 0.0
        want = 1.0 if jac == 0 else -np.log(2 * jac / (1 + jac)) / 21
        assert abs(float(d) - want) < 1e-5 and 0.0 <= float(p) <= 1.0
    assert layout.mash_distance(1.0, 21) == 0.0 and layout.mash_distance(0.0, 21) == 1.0
    # what make_all_genome_dend does with the file (figs.py:50-59)
    dm = np.zeros((n, n))
    for f, t, d, p, x in lines:
        dm[names.index(f)][names.index(t)] = d
        dm[names.index(t)][names.index(f)] = d
    assert (dm == dm.T).all() and (np.diag(dm) == 0).all()


def test_paircount_frac_matches_bitmap_to_paircount_bins():
    rng = np.random.default_rng(9)
    n_genomes, step, bin_size = 11, 100, 100000
    rows = rng.integers(0, 256, size=(2503, 2), dtype=np.uint8)
    starts, frac = layout.paircount_bins(rows, n_genomes, step, bin_size)
    bits = np.unpackbits(rows, axis=1, bitorder="little")[:, :n_genomes]
    rpb = bin_size // step
    counts = np.stack([bits[b:b + rpb].sum(axis=0) for b in range(0, len(rows), rpb)])
    s2, f2 = layout.paircount_frac(counts, bin_size)
    assert (starts == s2).all() and np.allclose(frac, f2)
