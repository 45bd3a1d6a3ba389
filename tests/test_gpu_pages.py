"""Hand-built pages through the C-ABI (pqg_plan_create / run / finish / download): dictionary
index streams at EVERY bit width 1..32 (reference encoder output, mixed RLE + bit-packed runs,
multi-group literal runs as foreign writers emit them), definition-level streams, truncated
pages and bad runs -> per-page error codes.  Expected values come from the oracle's restatement
of RleDecoder::get_batch (include/reader/rle_decoder.hpp:17-95)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def varint(x):
    out = bytearray()
    while x >= 0x80:
        out.append((x & 0x7F) | 0x80)
        x >>= 7
    out.append(x)
    return bytes(out)


def bitpack(vals, bw):
    acc, bits, out = 0, 0, bytearray()
    for v in vals:
        acc |= int(v) << bits
        bits += bw
        while bits >= 8:
            out.append(acc & 0xFF)
            acc >>= 8
            bits -= 8
    if bits:
        out.append(acc & 0xFF)
    return bytes(out)


def foreign_stream(vals, bw, rng):
    """RLE / bit-packed hybrid the way pyarrow-style writers emit it: literal runs of several
    groups of 8, RLE runs of any length"""
    out, i, n = bytearray(), 0, len(vals)
    while i < n:
        if rng.random() < 0.4:
            run = int(rng.integers(1, 70))
            run = min(run, n - i)
            v = int(vals[i])
            out += varint(run << 1) + int(v).to_bytes((bw + 7) // 8, "little")
            vals[i:i + run] = v
            i += run
        else:
            groups = int(rng.integers(1, 6))
            cnt = min(groups * 8, n - i)
            groups = (cnt + 7) // 8
            chunk = list(vals[i:i + cnt]) + [0] * (groups * 8 - cnt)
            out += varint((groups << 1) | 1) + bitpack(chunk, bw)
            i += cnt
    return bytes(out)


class Builder:
    """one image = [dictionary payload][page payloads...]; one chunk; INT64 values"""

    def __init__(self, pq, dict_vals, max_def=0):
        self.pq, self.max_def = pq, max_def
        self.img = bytearray(np.asarray(dict_vals, dtype=np.int64).tobytes()) if dict_vals is not None else bytearray()
        self.dict_n = 0 if dict_vals is None else len(dict_vals)
        self.dict_size = len(self.img)
        self.pages = []
        self.rows = 0

    def add_page(self, payload, num_values, dict_page=True, misalign=0, flags=None):
        self.img += b"\xEE" * misalign
        self.pages.append((len(self.img), len(payload), num_values, (1 if dict_page else 0) if flags is None else flags, self.rows))
        self.img += payload
        self.rows += num_values

    def run(self, expect_error=None):
        pq = self.pq
        ctx = pq.Context(0)
        img = np.frombuffer(bytes(self.img) + b"\0" * 64, dtype=np.uint8)
        buf = ctx.upload(img.ctypes.data, len(self.img))
        ck = (pq.ChunkDesc * 1)()
        ck[0] = pq.ChunkDesc(0, 0, self.rows, self.dict_size, self.dict_n, 0, len(self.pages), 0, 0, self.max_def, 0, pq.INT64,
                             1 if self.dict_n else 0, (C.c_uint8 * 2)(0, 0))
        pg = (pq.PageDesc * max(len(self.pages), 1))()
        for i, (off, size, nv, dp, row) in enumerate(self.pages):
            pg[i] = pq.PageDesc(off, row, size, nv, 0, dp)
        if expect_error is not None and expect_error[0] == "create":
            with pytest.raises(pq.PqgError, match=expect_error[1]):
                ctx.plan(buf, (ck, 1, pg, len(self.pages), self.rows))
            ctx.buf_free(buf)
            ctx.close()
            return None, None, None
        plan = ctx.plan(buf, (ck, 1, pg, len(self.pages), self.rows))
        plan.run()
        try:
            if expect_error is not None:
                pe = pq.PageError()
                rc = pq.lib().pqg_plan_finish(ctx.h, plan.h, C.byref(pe))
                assert rc == pq.PQG_ERR_PAGE, ctx.err()
                assert pe.code == expect_error[0] and pe.page == expect_error[1], (pe.code, pe.page, ctx.err())
                return None, None, ctx.err()
            plan.finish()
            vals = np.zeros(self.rows, dtype=np.int64)
            valid = np.zeros((self.rows + 31) // 32 + 1, dtype=np.uint32)
            has_validity = bool(plan.validity_ptr)
            assert has_validity or not self.max_def
            plan.download(values=vals.ctypes.data, validity=valid.ctypes.data if has_validity else None)
            ctx.sync()
            v = None
            if has_validity:
                v = ((valid[np.arange(self.rows) >> 5] >> (np.arange(self.rows) & 31).astype(np.uint32)) & 1).astype(bool)
            return vals, v, None
        finally:
            plan.destroy()
            ctx.buf_free(buf)
            ctx.close()


@pytest.mark.parametrize("bw", list(range(1, 33)))
def test_dictionary_index_streams_every_bit_width(pq, oracle, ref, bw):
    rng = np.random.default_rng(bw)
    dict_n = min(1 << min(bw, 14), 5000)
    dict_vals = rng.integers(-2**62, 2**62, size=dict_n, dtype=np.int64)
    b = Builder(pq, dict_vals)
    expected = []
    hi = min(1 << bw, dict_n)
    for page, n in enumerate((1, 7, 8, 9, 341, 512, 1024, 1500, 3000)):
        idx = rng.integers(0, hi, size=n).astype(np.uint32)
        if page % 3 == 1:
            idx[: n // 2] = idx[0]  # long RLE run at the front
        if page % 2 == 0:
            stream = bytes(ref.rle_encode(idx, bw))  # the reference's own encoder: single groups of 8
        else:
            stream = foreign_stream(idx, bw, rng)  # (mutates idx for its RLE runs)
        dec = oracle.rle_decode_i32(np.frombuffer(stream, dtype=np.uint8), bw, n)
        assert np.array_equal(dec.astype(np.uint32), idx)
        b.add_page(bytes([bw]) + stream, n, misalign=page % 5)
        expected.append(dict_vals[idx])
    vals, _, _ = b.run()
    assert np.array_equal(vals, np.concatenate(expected)), bw


def test_optional_levels_and_out_of_range_indices_become_nulls(pq, oracle):
    rng = np.random.default_rng(5)
    dict_vals = np.arange(100, dtype=np.int64) * 11 + 1
    b = Builder(pq, dict_vals, max_def=1)
    exp_vals, exp_valid = [], []
    for page, n in enumerate((5, 64, 1000, 1024, 2500)):
        present = rng.random(n) < (0.7 if page != 1 else 1.0)
        if page == 2:
            present[:] = False  # an all-null page
        # definition levels: bit-packed literal runs (foreign style) on odd pages, RLE runs otherwise
        lv = present.astype(np.uint32)
        if page % 2:
            groups = (n + 7) // 8
            def_stream = varint((groups << 1) | 1) + bitpack(list(lv) + [0] * (groups * 8 - n), 1)
        else:
            def_stream, i = bytearray(), 0
            while i < n:
                j = i
                while j < n and lv[j] == lv[i]:
                    j += 1
                def_stream += varint((j - i) << 1) + bytes([int(lv[i])])
                i = j
            def_stream = bytes(def_stream)
        nn = int(present.sum())
        idx = rng.integers(0, 100, size=nn).astype(np.uint32)
        if nn > 3:
            idx[1] = 100  # out of range -> null (column_reader.cpp:190-194)
            idx[3] = 127
        stream = foreign_stream(idx.copy(), 7, rng) if page % 2 else bytes(_enc(idx, 7))
        dec = oracle.rle_decode_i32(np.frombuffer(stream, dtype=np.uint8), 7, nn).astype(np.uint32)
        payload = len(def_stream).to_bytes(4, "little") + def_stream + bytes([7]) + stream
        b.add_page(payload, n, misalign=page)
        v = np.zeros(n, dtype=np.int64)
        ok = present.copy()
        slots = np.nonzero(present)[0]
        good = dec < 100
        v[slots[good]] = dict_vals[dec[good]]
        ok[slots[~good]] = False
        exp_vals.append(v)
        exp_valid.append(ok)
    vals, valid, _ = b.run()
    assert np.array_equal(valid, np.concatenate(exp_valid))
    assert np.array_equal(vals, np.concatenate(exp_vals))


def _level_stream(present, rng, style):
    """definition levels (bit width 1) of a max_def 1 column: 'mixed' = RLE runs and literal groups in turns (pyarrow with
    scattered nulls), 'literal' = one bit-packed run per 504 values, 'rle' = one RLE run per stretch of equal levels"""
    lv = present.astype(np.uint32)
    n = len(lv)
    if style == "mixed":
        s = foreign_stream(lv, 1, rng)  # (mutates lv: its RLE runs repeat the first level)
        present[:] = lv.astype(bool)
        return s
    if style == "literal":
        out = bytearray()
        for i in range(0, n, 504):
            part = list(lv[i:i + 504])
            groups = (len(part) + 7) // 8
            out += varint((groups << 1) | 1) + bitpack(part + [0] * (groups * 8 - len(part)), 1)
        return bytes(out)
    out, i = bytearray(), 0
    while i < n:
        j = i
        while j < n and lv[j] == lv[i]:
            j += 1
        out += varint((j - i) << 1) + bytes([int(lv[i])])
        i = j
    return bytes(out)


@pytest.mark.parametrize("bw", [0, 1, 5, 12, 16, 17, 20, 32])
def test_big_optional_dictionary_pages_through_the_block_decode(pq, oracle, bw):
    """pages of more than 1024 slots of an OPTIONAL dictionary column go through k_flat_scan / k_flat_ranks / k_flat_emit:
    every level-stream style, index streams as foreign writers emit them (long literal runs, RLE runs with 1..3-byte
    headers), out-of-range indices -> nulls, short streams -> nulls / index 0, page sizes around the 1024-slot blocks"""
    rng = np.random.default_rng(100 + bw)
    dict_n = 1 if bw == 0 else (70000 if bw == 17 else min(1 << min(bw, 13), 6000))  # (bw 17: beyond 16-bit index buffers)
    dict_vals = rng.integers(-2**62, 2**62, size=dict_n, dtype=np.int64)
    hi = 1 if bw == 0 else min((1 << bw) if bw < 32 else 1 << 32, dict_n + (dict_n // 50 if bw >= 5 else 0))  # a few out of range
    b = Builder(pq, dict_vals, max_def=1)
    exp_vals, exp_valid = [], []
    sizes = (1025, 2048, 2049, 5000, 20000, 33000, 1500, 3071)
    for page, n in enumerate(sizes):
        style = ("mixed", "literal", "rle")[page % 3]
        present = rng.random(n) < (0.75 if page != 3 else 0.02)
        if page == 4:
            present[:] = True
            present[7] = False
        if page == 5:
            present[1000:12000] = True  # an RLE level run with a 3-byte header
        def_stream = _level_stream(present, rng, style)
        if page == 5:  # the level stream ends early: the remaining slots are null (rle_decoder.hpp:21-24)
            cut = n - 4000
            present[cut:] = False
            def_stream = _level_stream(present[:cut].copy(), rng, "rle")
        nn = int(present.sum())
        idx = rng.integers(0, hi, size=nn).astype(np.uint32)
        if page == 1 and nn > 2000:
            idx[100:1900] = idx[100]  # an RLE run across a block boundary (3-byte varint header when long enough)
        if page == 4 and bw:  # an RLE index run with a 3-byte header in front
            idx[:9000] = idx[0]
            stream = varint(9000 << 1) + int(idx[0]).to_bytes((bw + 7) // 8, "little") + foreign_stream(idx[9000:], bw, rng)
        else:
            stream = b"" if bw == 0 and page % 2 else foreign_stream(idx, bw, rng)
        if bw == 0 and page % 2:
            idx[:] = 0  # no stream at all: every index reads 0
        if page == 6 and nn > 600:  # the index stream ends early: the remaining values read index 0
            keep = nn - 500
            stream = foreign_stream(idx[:keep].copy(), bw, rng) if bw else b""
            dec_full = np.zeros(nn, dtype=np.uint32)
            if bw:
                dec_full[:keep] = oracle.rle_decode_i32(np.frombuffer(stream, dtype=np.uint8), bw, keep).astype(np.uint32)
            idx = dec_full
        payload = len(def_stream).to_bytes(4, "little") + def_stream + bytes([bw]) + stream
        b.add_page(payload, n, misalign=page % 7)
        v = np.zeros(n, dtype=np.int64)
        ok = present.copy()
        slots = np.nonzero(present)[0]
        good = idx < dict_n
        v[slots[good]] = dict_vals[idx[good]]
        ok[slots[~good]] = False
        exp_vals.append(v)
        exp_valid.append(ok)
    vals, valid, _ = b.run()
    assert np.array_equal(valid, np.concatenate(exp_valid)), bw
    assert np.array_equal(vals, np.concatenate(exp_vals)), bw


def test_big_optional_plain_pages_through_the_block_decode(pq):
    rng = np.random.default_rng(77)
    b = Builder(pq, None, max_def=1)
    exp_vals, exp_valid = [], []
    for page, n in enumerate((1025, 4096, 10000, 2500, 50000)):
        style = ("mixed", "literal", "rle")[page % 3]
        present = rng.random(n) < (0.6 if page != 3 else 1.0)
        def_stream = _level_stream(present, rng, style)
        nn = int(present.sum())
        data = rng.integers(-2**62, 2**62, size=nn, dtype=np.int64)
        b.add_page(len(def_stream).to_bytes(4, "little") + def_stream + data.tobytes(), n, dict_page=False, misalign=page % 5)
        v = np.zeros(n, dtype=np.int64)
        v[present] = data
        exp_vals.append(v)
        exp_valid.append(present.copy())
    vals, valid, _ = b.run()
    assert np.array_equal(valid, np.concatenate(exp_valid))
    assert np.array_equal(vals, np.concatenate(exp_vals))
    # a page whose values are cut short reports the truncation (the general kernel, after the block decode gave it back)
    b = Builder(pq, None, max_def=1)
    present = rng.random(3000) < 0.5
    def_stream = _level_stream(present, rng, "mixed")
    data = rng.integers(0, 100, size=int(present.sum()) - 3, dtype=np.int64)
    ok_page = np.arange(10, dtype=np.int64)
    b.add_page((2).to_bytes(4, "little") + b"\x14\x01" + ok_page.tobytes(), 10, dict_page=False)
    b.add_page(len(def_stream).to_bytes(4, "little") + def_stream + data.tobytes(), 3000, dict_page=False)
    _, _, msg = b.run(expect_error=(1, 1))
    assert "ByteBuffer: read beyond end" in msg
    # a zero-length run inside the levels of a big page: explicit error (undefined behaviour in the reference)
    b = Builder(pq, None, max_def=1)
    lv = b"\x10\x01" * 40 + b"\x00\x01" + b"\x10\x01" * 400
    b.add_page(len(lv).to_bytes(4, "little") + lv + np.arange(3528, dtype=np.int64).tobytes(), 3528, dict_page=False)
    b.run(expect_error=(3, 0))


def _enc(idx, bw):
    """single bit-packed groups of 8 ("03 <bw bytes>"), zero padded: the reference writer's layout
    for data without 4-fold repeats"""
    out = bytearray()
    for i in range(0, len(idx), 8):
        g = list(idx[i:i + 8]) + [0] * (8 - len(idx[i:i + 8]))
        out += b"\x03" + bitpack(g, bw)
    return bytes(out)


def test_page_errors(pq):
    dict_vals = np.arange(16, dtype=np.int64)
    # truncated PLAIN page: 10 values announced, 5 present -> ByteBuffer-style error on page 1
    b = Builder(pq, None)
    b.add_page(np.arange(8, dtype=np.int64).tobytes(), 8, dict_page=False)
    b.add_page(np.arange(5, dtype=np.int64).tobytes(), 10, dict_page=False)
    _, _, msg = b.run(expect_error=(pq.lib() and 1, 1))
    assert "ByteBuffer: read beyond end" in msg
    # bit width 33
    b = Builder(pq, dict_vals)
    b.add_page(bytes([4]) + _enc(np.arange(8, dtype=np.uint32), 4), 8)
    b.add_page(bytes([33]) + b"\x03" + b"\0" * 33, 8)
    b.run(expect_error=(2, 1))
    # zero-length RLE run: undefined behaviour in the reference -> explicit error
    b = Builder(pq, dict_vals)
    b.add_page(bytes([4]) + b"\x00\x01" + _enc(np.arange(8, dtype=np.uint32), 4), 8)
    b.run(expect_error=(3, 0))
    # definition-level section longer than the page
    b = Builder(pq, dict_vals, max_def=1)
    b.add_page((500).to_bytes(4, "little") + b"\x10\x01", 8)
    b.run(expect_error=(1, 0))


def test_short_streams_zero_fill_like_the_reference(pq, oracle):
    """RleDecoder returns zeros once the data is exhausted (rle_decoder.hpp:21-24)"""
    dict_vals = np.arange(8, dtype=np.int64) + 100
    b = Builder(pq, dict_vals)
    stream = _enc(np.array([1, 2, 3, 4, 5, 6, 7, 0], dtype=np.uint32), 3)
    b.add_page(bytes([3]) + stream, 20)  # 8 values encoded, 20 announced
    vals, _, _ = b.run()
    dec = oracle.rle_decode_i32(np.frombuffer(stream, dtype=np.uint8), 3, 20)
    assert np.array_equal(vals, dict_vals[dec])


def test_oversized_plain_pages_are_split_and_errors_name_the_page(pq):
    """a 100 KB PLAIN REQUIRED page (foreign writers) runs as 1 KB virtual slices on the tile
    kernel; a truncated one still reports the ORIGINAL page index"""
    rng = np.random.default_rng(9)
    big = rng.integers(-2**62, 2**62, size=12_345, dtype=np.int64)
    small = rng.integers(-2**62, 2**62, size=100, dtype=np.int64)
    b = Builder(pq, None)
    b.add_page(small.tobytes(), len(small), dict_page=False, misalign=3)
    b.add_page(big.tobytes(), len(big), dict_page=False, misalign=5)
    b.add_page(small.tobytes(), len(small), dict_page=False, misalign=1)
    vals, _, _ = b.run()
    assert np.array_equal(vals, np.concatenate([small, big, small]))
    b = Builder(pq, None)
    b.add_page(small.tobytes(), len(small), dict_page=False)
    b.add_page(big.tobytes()[:-8], len(big), dict_page=False)  # one value short
    _, _, msg = b.run(expect_error=(1, 1))
    assert "ByteBuffer: read beyond end" in msg


def _string_plan(pq, pages, n_each):
    """pages: list of payload bytes of PLAIN REQUIRED BYTE_ARRAY pages; -> (offsets, chars) of the decode"""
    img = bytearray()
    descs = []
    row = 0
    for pay, n in zip(pages, n_each):
        img += b"\xAA" * (len(descs) % 3)
        descs.append((len(img), len(pay), n, row))
        img += pay
        row += n
    ctx = pq.Context(0)
    arr = np.frombuffer(bytes(img) + b"\0" * 64, dtype=np.uint8)
    buf = ctx.upload(arr.ctypes.data, len(img))
    ck = (pq.ChunkDesc * 1)()
    ck[0] = pq.ChunkDesc(0, 0, row, 0, 0, 0, len(descs), 0, 0, 0, 0, pq.BYTE_ARRAY, 0, (C.c_uint8 * 2)(0, 0))
    pg = (pq.PageDesc * len(descs))()
    for i, (off, size, n, r0) in enumerate(descs):
        pg[i] = pq.PageDesc(off, r0, size, n, 0, 0)
    plan = ctx.plan(buf, (ck, 1, pg, len(descs), row))
    try:
        for _ in range(2):  # the second run reuses whatever mode the first one settled on
            plan.run()
            plan.finish()
            offs = np.zeros(row + 1, dtype=np.uint32)
            chars = np.zeros(max(plan.chars_size, 1), dtype=np.uint8)
            plan.download(offsets=offs.ctypes.data, chars=chars.ctypes.data)
            ctx.sync()
        return offs, chars[:plan.chars_size].tobytes()
    finally:
        plan.destroy()
        ctx.buf_free(buf)
        ctx.close()


def test_plain_string_pages_with_trailing_bytes_fall_back_to_the_size_pass(pq):
    """byte counts normally come from the page headers (payload - 4 * values); a page that
    carries bytes beyond its values (legal for the reference: it reads n values and stops) must
    make the plan fall back to the exact size pass and still decode correctly"""
    def page(strs, trailing=b""):
        return b"".join(len(s).to_bytes(4, "little") + s for s in strs) + trailing
    a = [b"alpha", b"", b"beta-gamma", b"x" * 70]
    b = [b"delta", b"epsilon" * 9, b"z"]
    for trailing in (b"", b"JUNKJUNK"):
        offs, chars = _string_plan(pq, [page(a), page(b, trailing), page(a)], [len(a), len(b), len(a)])
        exp = a + b + a
        assert chars == b"".join(exp)
        assert offs.tolist() == np.concatenate([[0], np.cumsum([len(s) for s in exp])]).tolist()


@pytest.mark.parametrize("ulen", [0, 1, 3, 4, 5, 7, 8, 13, 16, 17, 31, 33, 48, 49, 64, 100, 500])
def test_plain_string_pages_of_one_length(pq, ulen):
    """REQUIRED PLAIN pages whose values all have one length take the copy pass without staging (chars assembled straight
    into aligned vectors): every length class, page sizes from one value to the 2 KB staging limit, so that the pages'
    char ranges start at every alignment; a page whose LAST prefix differs has the same section size only by accident
    and must take the general path"""
    rng = np.random.default_rng(ulen)
    def page(strs):
        return b"".join(len(s).to_bytes(4, "little") + s for s in strs)
    pages, counts, exp = [], [], []
    cap = max(1, min(40, 2040 // (ulen + 4)))
    for i in range(60):
        n = int(rng.integers(1, cap + 1))
        strs = [bytes(rng.integers(33, 127, ulen, dtype=np.uint8)) for _ in range(n)]
        if i % 7 == 3 and n >= 2 and ulen >= 1:  # same section size, two lengths
            strs[-2] = strs[-2] + b"+"
            strs[-1] = strs[-1][:-1]
        pages.append(page(strs)); counts.append(n); exp += strs
    offs, chars = _string_plan(pq, pages, counts)
    assert chars == b"".join(exp)
    assert offs.tolist() == np.concatenate([[0], np.cumsum([len(s) for s in exp])]).tolist()


def test_required_chunk_out_of_range_indices_become_nulls(pq, oracle):
    """Value::null() for an out-of-range dictionary index whatever max_def is (reference
    src/reader/column_reader.cpp:190-194): a REQUIRED-only plan has no validity bitmap, so
    pqg_plan_finish adds one and decodes again.  Writer-shaped streams (tile kernel), foreign
    streams (run-by-run path) and a page that only the general kernel takes."""
    rng = np.random.default_rng(17)
    dict_vals = np.arange(100, dtype=np.int64) * 13 + 5
    b = Builder(pq, dict_vals, max_def=0)
    exp_vals, exp_valid = [], []
    for page, n in enumerate((40, 1024, 777, 64, 1500)):
        idx = rng.integers(0, 100, size=n).astype(np.uint32)
        if page != 3:  # page 3 stays clean
            idx[n // 3] = 100
            idx[n - 1] = 126
        stream = foreign_stream(idx, 7, rng) if page % 2 else bytes(_enc(idx, 7))
        dec = oracle.rle_decode_i32(np.frombuffer(stream, dtype=np.uint8), 7, n).astype(np.uint32)
        b.add_page(bytes([7]) + stream, n, misalign=page)
        good = dec < 100
        v = np.zeros(n, dtype=np.int64)
        v[good] = dict_vals[dec[good]]
        exp_vals.append(v)
        exp_valid.append(good)
    vals, valid, _ = b.run()
    assert valid is not None, "the plan must grow a validity bitmap"
    assert np.array_equal(valid, np.concatenate(exp_valid))
    assert np.array_equal(vals, np.concatenate(exp_vals))
    # a clean REQUIRED plan still carries none
    b = Builder(pq, dict_vals, max_def=0)
    b.add_page(bytes([7]) + bytes(_enc(np.arange(64, dtype=np.uint32), 7)), 64)
    vals, valid, _ = b.run()
    assert valid is None and np.array_equal(vals, dict_vals[:64])


def test_unsupported_pages_are_rejected_at_plan_creation(pq):
    """SURVEY 8(b) "Unsupported inputs": DATA_PAGE_V2 and DELTA_* / BYTE_STREAM_SPLIT data pages fail with
    PQG_ERR_UNSUPPORTED and a message instead of being decoded as PLAIN (what the reference does,
    src/reader/column_reader.cpp:66-67,173-222)"""
    vals = np.arange(8, dtype=np.int64).tobytes()
    for enc, name in ((5, "DELTA_BINARY_PACKED"), (6, "DELTA_LENGTH_BYTE_ARRAY"), (7, "DELTA_BYTE_ARRAY"), (9, "BYTE_STREAM_SPLIT"), (4, "BIT_PACKED")):
        b = Builder(pq, None)
        b.add_page(vals, 8, flags=0)
        b.add_page(vals, 8, flags=enc << 8)
        b.run(expect_error=("create", name))
    b = Builder(pq, None)
    b.add_page(vals, 8, flags=pq.PQG_PAGE_FLAG_V2)
    b.run(expect_error=("create", "DATA_PAGE_V2"))
    # the encodings that ARE decoded, spelled out in the flags
    b = Builder(pq, np.arange(16, dtype=np.int64))
    b.add_page(vals, 8, flags=0 << 8)
    b.add_page(bytes([4]) + _enc(np.arange(8, dtype=np.uint32), 4), 8, flags=1 | (8 << 8))
    b.add_page(bytes([4]) + _enc(np.arange(8, dtype=np.uint32), 4), 8, flags=1 | (2 << 8))
    out, _, _ = b.run()
    assert out.tolist() == list(range(8)) * 3


def test_wrap_device_requires_a_readable_tail(pq):
    ctx = pq.Context(0)
    img = np.zeros(4096, dtype=np.uint8)
    buf = ctx.upload(img.ctypes.data, 1024)
    ptr = pq.lib().pqg_buf_device_ptr(buf)
    with pytest.raises(pq.PqgError, match="capacity"):
        ctx.wrap_device(ptr, 1024, capacity=1024)
    w = ctx.wrap_device(ptr, 1024 - 64, capacity=1024)
    ctx.buf_free(w)
    ctx.buf_free(buf)
    ctx.close()
