"""CPU: pins the oracle (oracle/anchor_oracle.c) against the reference's own
known-answer test and against outputs of the unmodified reference binaries."""
import numpy as np
import pytest

from oracle import oracle, refpipe
from panagram_b200 import synth


def brute_pattern(reads, query, k):
    """test_py_kmc_file.py:51-66 (_cout_kmers) + :174-190 (pattern)."""
    kmers = {}
    for r in reads:
        for s in range(len(r) - k + 1):
            w = r[s:s + k].encode()
            if b"N" in w:
                continue
            c = oracle.canon_str(w)
            kmers[c] = kmers.get(c, 0) + 1
    pat = []
    for i in range(len(query) - k + 1):
        w = query[i:i + k].encode()
        pat.append(0 if b"N" in w else kmers.get(oracle.canon_str(w), 0))
    return kmers, pat


def test_kat_get_counters_for_read(kat):
    k = kat["k"]
    kmers, pat = brute_pattern(kat["reads"], kat["query"], k)
    assert pat == kat["counters"], "golden (reference py_kmc_api) == the reference test's brute force"
    db = oracle.OracleDB.open(kat["dir"] / "kmc_db_sorted")
    assert (db.k, db.version, db.both_strands, db.min_count) == (17, 0, True, 1)
    got = db.get_counters_for_read(kat["query"])
    assert got.tolist() == kat["counters"]
    assert db.get_counters_for_read("ACGT") is None          # len < k: false + cleared vector


def test_kat_listing_kmc2_and_kmc1(kat):
    kmers, _ = brute_pattern(kat["reads"], kat["query"], kat["k"])
    want = sorted((oracle.kmer_int(s), c) for s, c in kmers.items())
    for name, ver in (("kmc_db", 0x200), ("kmc_db_sorted", 0)):
        db = oracle.OracleDB.open(kat["dir"] / name)
        assert db.version == ver and db.total_kmers == len(kmers)
        if ver == 0x200:
            assert db.signature_len == 9 and db.counter_size == 1      # test_info, :147-161
        km, ct = db.list()
        assert sorted(zip(km.tolist(), ct.tolist())) == want
    dump = [l.split("\t") for l in (kat["dir"] / "dump.txt").read_text().splitlines()]
    assert sorted((oracle.kmer_int(a.encode()), int(b)) for a, b in dump) == want


def test_kat_check_kmer(kat):
    """test_check_kmer (:199-218): present k-mers found with their counts, absent not found."""
    k = kat["k"]
    kmers, _ = brute_pattern(kat["reads"], kat["query"], k)
    db = oracle.OracleDB.open(kat["dir"] / "kmc_db_sorted")
    for s, c in kmers.items():
        assert db.get_counters_for_read(s).tolist() == [c]
    rng = np.random.default_rng(5)
    n_abs = 0
    for _ in range(300):
        s = bytes(rng.choice(list(b"ACGT"), size=k).tolist())
        if oracle.canon_str(s) not in kmers:
            assert db.get_counters_for_read(s).tolist() == [0]
            n_abs += 1
    assert n_abs > 250


@pytest.mark.parametrize("which", ["pan3", "pan35"])
def test_oracle_equals_reference_run_anchor(which, request):
    pan = request.getfixturevalue(which)
    dbs = [oracle.OracleDB.open(pan["dir"] / "kmc" / f"bitvec{i}") for i in range(pan["ndb"])]
    for a in pan["anchors"]:
        got = oracle.anchor_fasta(dbs, pan["n_genomes"], pan["fasta"][a])
        for key, want in pan["expected"][a].items():
            assert got[key] == want, f"{which}/{a}/{key}"


def test_structural_invariants(pan3):
    """SURVEY §4 [probed]: anchor's own column is 1 wherever the window is all-ACGT, rows are
    all-zero iff the window holds a non-ACGT byte; bitmap.100 == bitmap.1[::100] per chromosome."""
    k, n = pan3["k"], pan3["n_genomes"]
    for ai, a in enumerate(pan3["anchors"]):
        rows = np.frombuffer(pan3["expected"][a]["bitmap.1"], dtype=np.uint8)
        low = np.frombuffer(pan3["expected"][a]["bitmap.100"], dtype=np.uint8)
        off = loff = 0
        for name, seq in oracle.parse_fasta(pan3["fasta"][a]):
            nk = len(seq) - k + 1
            r = rows[off:off + nk]
            valid = np.array([all(c in b"ACGTacgt" for c in seq[p:p + k]) for p in range(nk)])
            assert ((r >> ai) & 1 == 1).tolist() == valid.tolist()
            assert ((r == 0) == ~valid).all()
            nl = (nk + 99) // 100
            assert (low[loff:loff + nl] == r[::100]).all()
            off += nk
            loff += nl
        assert off == rows.size and loff == low.size


def test_bruteforce_matches_c_oracle(pan3):
    k = pan3["k"]
    recs = {g: oracle.parse_fasta(p) for g, p in pan3["fasta"].items()}
    sets = [oracle.kmer_set([s for _, s in recs[g]], k) for g in pan3["names"]]
    dbs = [oracle.OracleDB.open(pan3["dir"] / "kmc" / "bitvec0")]
    name, seq = recs["g1"][2]
    want = oracle.brute_rows(seq, k, sets)
    got = oracle.anchor_chrom(dbs, 3, seq)["bitmap1"]
    assert (want == got).all()
    # an in-memory DB built from the sets answers like the kmc_tools-written one
    allk = sorted(set().union(*sets))
    ints = np.array([oracle.kmer_int(s) for s in allk], dtype=np.uint64)
    cnt = np.array([sum(1 << g for g in range(3) if s in sets[g]) for s in allk], dtype=np.uint32)
    mem = oracle.OracleDB.from_kmers(k, ints, cnt)
    assert (oracle.anchor_chrom([mem], 3, seq)["bitmap1"] == got).all()


def test_per_genome_dbs_list_to_the_bitvec_union(pan3):
    """K_g from kmc/{s}.count (KMC2) and .onehot (KMC1) equals bit g of the bitvec DB."""
    bk, bc = oracle.OracleDB.open(pan3["dir"] / "kmc" / "bitvec0").list()
    for g, name in enumerate(pan3["names"]):
        want = set(bk[(bc >> g) & 1 == 1].tolist())
        for kind in ("count", "onehot"):
            db = oracle.OracleDB.open(pan3["dir"] / "kmc" / f"{name}.{kind}")
            km, ct = db.list()
            assert set(km.tolist()) == want
            if kind == "onehot":
                assert (ct == 1 << g).all()


@pytest.mark.skipif(not refpipe.have_ref(), reason="oracle/_ref not built")
def test_live_reference_run(tmp_path):
    """Fresh random pan-genome through the real reference binaries vs the oracle."""
    samples = synth.make_pangenome(tmp_path / "fa", 9, 7000, 99, n_chroms=2, n_run=30, lower_run=90)
    refpipe.build_index(tmp_path / "idx", samples, 25, anchors=["g0", "g8"], threads=2)
    dbs = [oracle.OracleDB.open(tmp_path / "idx" / "kmc" / "bitvec0")]
    for a in ("g0", "g8"):
        want = refpipe.read_anchor_dir(tmp_path / "idx" / "anchor" / a)
        got = oracle.anchor_fasta(dbs, 9, dict(samples)[a])
        for key in want:
            assert got[key] == want[key], key
