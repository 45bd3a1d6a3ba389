"""Known answer of the chunk-index prototype (reference src/main.cpp:7-35), the recipe of
SURVEY.md section 4: a file with a BYTE_ARRAY column `l_comment`, 2 row groups x 300 000 rows,
written by the REFERENCE's ParquetWriter, then the unmodified loop of main.cpp over the
REFERENCE's StringColumnIterator (oracle/ref_shim.cpp:ref_chunk_index -- main.cpp itself
hard-codes a path outside this container).  The survey's own run printed `Total chunks: 4710`
but did not record its input text, so that number cannot be regenerated; this script pins the
same experiment on a committed, seeded input instead and records what the reference prints.

    python tests/golden/make_lcomment.py      # needs /root/reference (oracle/_ref)

Output: tests/golden/lcomment.json {rows, total_chunks, t2c_crc32, file_sha256, file_size}.
The tests regenerate the file with the workload generator (byte-identical to the reference's
writer: file_sha256 is checked) and compare the oracle / the GPU path with these numbers."""
import hashlib
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

ROWS_PER_GROUP, GROUPS, SEED = 300_000, 2, 4710
WORDS = ("furiously carefully quickly slyly blithely regular final ironic express special pending bold even silent unusual "
         "deposits requests accounts packages foxes ideas theodolites pinto beans instructions dependencies excuses platelets "
         "asymptotes courts dolphins multipliers sauternes warthogs frets dinos attainments somas tithes sheaves gifts "
         "sleep wake nag haggle cajole boost detect integrate use are among above according to the across after along").split()


def lcomment_column(n, seed):
    """TPC-H-like comment text: words joined by blanks, cut to a length in [10, 43]"""
    rng = np.random.default_rng(seed)
    lens = rng.integers(10, 44, size=n)
    picks = rng.integers(0, len(WORDS), size=(n, 9))
    out = []
    for i in range(n):
        s = " ".join(WORDS[k] for k in picks[i])
        out.append(s[: int(lens[i])].encode())
    return out


def lcomment_specs():
    from oraclelib import BYTE_ARRAY, INT64, REQUIRED, UTF8
    return [("l_orderkey", INT64, REQUIRED, -1), ("l_comment", BYTE_ARRAY, REQUIRED, UTF8)]


def lcomment_row_groups():
    from oraclelib import fixed_col, strings_to_col
    rgs = []
    for g in range(GROUPS):
        strs = lcomment_column(ROWS_PER_GROUP, SEED + g)
        rgs.append([fixed_col(np.arange(ROWS_PER_GROUP, dtype=np.int64) + g * ROWS_PER_GROUP), strings_to_col(strs)])
    return rgs


def main():
    from oraclelib import Ref
    ref = Ref()
    path = os.path.join(HERE, "_lcomment_tmp.parquet")
    ref.write_file(path, lcomment_specs(), lcomment_row_groups())
    try:
        h = ref.open(path)
        t2c, n = ref.chunk_index(h, "l_comment", 4096)
        rows = ref.num_rows(h)
        ref.close(h)
        data = open(path, "rb").read()
        out = {"rows": int(rows), "total_chunks": int(n), "t2c_crc32": int(zlib.crc32(np.ascontiguousarray(t2c, dtype=np.uint64).tobytes())),
               "file_sha256": hashlib.sha256(data).hexdigest(), "file_size": len(data), "chunk_size": 4096,
               "how": "reference ParquetWriter -> reference StringColumnIterator -> loop of src/main.cpp:21-32 (oracle/ref_shim.cpp:ref_chunk_index)"}
    finally:
        os.unlink(path)
    json.dump(out, open(os.path.join(HERE, "lcomment.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
