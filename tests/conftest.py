import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oraclelib
    oraclelib.build_oracle()
    return oraclelib.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The compiled, unmodified reference (oracle/_ref).  Built here when /root/reference
    exists; on the GPU box the prebuilt .so travels with the snapshot."""
    import oraclelib
    if not oraclelib.Ref.available():
        pytest.skip("oracle/_ref/libpqref.so not available")
    return oraclelib.Ref()


@pytest.fixture(scope="session")
def pq():
    import pqb200
    if not os.path.exists(pqb200.LIB_PATH):
        pqb200.build()
    pqb200.lib()
    return pqb200


@pytest.fixture(scope="session")
def files(tmp_path_factory, oracle):
    """name -> path of reference-written parquet files: generated through the reference's
    ParquetWriter when oracle/_ref is present, plus the committed golden files."""
    import fixtures
    import oraclelib
    out = {}
    gold = os.path.join(HERE, "golden")
    for f in sorted(os.listdir(gold)):
        if f.endswith(".parquet"):
            out["golden_" + f[:-8]] = os.path.join(gold, f)
    if oraclelib.Ref.available():
        r = oraclelib.Ref()
        d = tmp_path_factory.mktemp("pq")
        for name in fixtures.standard_files():
            out[name] = fixtures.make_file(r, oracle, name, str(d / (name + ".parquet")))
    return out


def to_values(d):
    """dict of numpy arrays (pqb200 dump) -> oraclelib.Values"""
    import oraclelib
    return oraclelib.Values(d["is_null"], d["vidx"], d["fixed"], d["str_off"], d["chars"])
