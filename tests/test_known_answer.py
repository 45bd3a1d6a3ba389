"""The chunk-index known answer (reference src/main.cpp:7-35): 2 x 300 000
`l_comment` rows -> `Total tuples: 600000 / Total chunks: N`, N recorded from the unmodified
reference by tests/golden/make_lcomment.py (tests/golden/lcomment.json).  The file is regenerated
by the workload generator and must be byte-identical to the reference-written one (sha256)."""
import hashlib
import json
import os
import subprocess
import sys
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = json.load(open(os.path.join(HERE, "golden", "lcomment.json")))


@pytest.fixture(scope="module")
def lcomment_file(pq, tmp_path_factory):
    import make_lcomment as m
    rgs = m.lcomment_row_groups()
    cols = [dict(fixed=np.concatenate([rg[0]["fixed"] for rg in rgs])),
            dict(str_off=np.concatenate([[0]] + [rg[1]["str_off"][1:] + sum(int(r[1]["str_off"][-1]) for r in rgs[:i])
                                                 for i, rg in enumerate(rgs)]).astype(np.uint64),
                 chars=np.concatenate([rg[1]["chars"] for rg in rgs]))]
    g = pq.generate(m.lcomment_specs(), cols, [m.ROWS_PER_GROUP] * m.GROUPS)
    path = str(tmp_path_factory.mktemp("lc") / "lineitem.parquet")
    g.write(path)
    g.free()
    return path


def test_generated_file_is_the_reference_written_one(lcomment_file):
    data = open(lcomment_file, "rb").read()
    assert len(data) == GOLD["file_size"]
    assert hashlib.sha256(data).hexdigest() == GOLD["file_sha256"]


def test_oracle_reproduces_the_known_answer(oracle, lcomment_file):
    h = oracle.open(lcomment_file)
    try:
        t2c, n = oracle.chunk_index(h, "l_comment", GOLD["chunk_size"])
        assert oracle.num_rows(h) == GOLD["rows"]
    finally:
        oracle.close(h)
    assert n == GOLD["total_chunks"]
    assert zlib.crc32(np.ascontiguousarray(t2c, dtype=np.uint64).tobytes()) == GOLD["t2c_crc32"]


@pytest.mark.gpu
def test_gpu_reproduces_the_known_answer(pq, lcomment_file):
    r = pq.Reader(lcomment_file)
    try:
        t2c, n = r.chunk_index("l_comment", GOLD["chunk_size"])
        assert r.num_rows == GOLD["rows"]
    finally:
        r.close()
    assert n == GOLD["total_chunks"]
    assert zlib.crc32(np.ascontiguousarray(t2c, dtype=np.uint64).tobytes()) == GOLD["t2c_crc32"]
    # and the CLI prints what the reference's parser prints
    out = subprocess.run([os.path.join(pq.PKG_DIR, "bin", "parser"), lcomment_file, "--chunk-index", "l_comment"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert f"Total tuples: {GOLD['rows']}" in out.stdout and f"Total chunks: {GOLD['total_chunks']}" in out.stdout
