import json
import sys
import tarfile
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden(tmp_path_factory):
    """Unpacked golden fixtures (made from the reference by tests/golden/make_golden.py)."""
    d = tmp_path_factory.mktemp("golden")
    for t in sorted(GOLDEN.glob("*.tar.gz")):
        with tarfile.open(t) as tf:
            tf.extractall(d, filter="data")
    return d


def load_pan(golden_dir: Path, name: str) -> dict:
    d = golden_dir / name
    meta = json.loads((d / "meta.json").read_text())
    meta["dir"] = d
    meta["fasta"] = {n: d / "fasta" / f"{n}.fa" for n in meta["names"]}
    exp = {}
    for a in meta["anchors"]:
        e = d / "expected" / a
        exp[a] = {"bitmap.1": (e / "bitmap.1").read_bytes(), "bitmap.100": (e / "bitmap.100").read_bytes(),
                  "chrs.tsv": (e / "chrs.tsv").read_text(), "bitsum.bins.tsv": (e / "bitsum.bins.tsv").read_text()}
    meta["expected"] = exp
    meta["ndb"] = (meta["n_genomes"] + 31) // 32
    return meta


@pytest.fixture(scope="session")
def pan3(golden):
    return load_pan(golden, "pan3_k21")


@pytest.fixture(scope="session")
def pan35(golden):
    return load_pan(golden, "pan35_k31")


@pytest.fixture(scope="session")
def kat(golden):
    d = golden / "kat_k17"
    m = json.loads((d / "kat.json").read_text())
    m["dir"] = d
    return m
