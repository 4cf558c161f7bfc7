import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


def read_fasta(path):
    """sw_solve_small.cpp:25-30: skip the header line, concatenate the rest."""
    with open(path) as f:
        return "".join(f.read().split("\n")[1:])


def read_truth(path):
    """sw_solve_small.cpp:56-67: CSV index,QNAME,SEQ,POS with one header line."""
    rows = []
    with open(path) as f:
        for i, line in enumerate(f):
            if i == 0 or not line.strip():
                continue
            r = line.rstrip("\n").split(",")
            rows.append((int(r[0]), r[1], r[2], int(r[3])))
    return rows


def read_golden_csv(name):
    import csv
    with open(os.path.join(GOLDEN, name)) as f:
        rows = list(csv.reader(f))[1:]
    return [dict(index=int(r[0]), score=int(r[1]), pos=int(r[2]), cx=r[3], cy=r[4]) for r in rows]


@pytest.fixture(scope="session")
def data_small():
    ref = read_fasta(os.path.join(GOLDEN, "data_small", "genome.chr22.5K.fa"))
    truth = read_truth(os.path.join(GOLDEN, "data_small", "data_small_ground_truth.csv"))
    return ref, truth


@pytest.fixture(scope="session")
def random_pairs():
    with open(os.path.join(GOLDEN, "random_pairs.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def c3_sample():
    with open(os.path.join(GOLDEN, "c3_sample.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def c4_sample():
    with open(os.path.join(GOLDEN, "c4_sample.json")) as f:
        return json.load(f)


def load_pkg():
    """import the hyphen-named product package (parallel-genomeseq_b200)."""
    import importlib
    return importlib.import_module("parallel-genomeseq_b200")


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def engine(pkg):
    """A swb_ctx on cuda:0.  Fails loudly (no skip, no fallback) when the CUDA path is unavailable."""
    e = pkg.Engine(0)
    yield e
    e.close()
