"""ctypes bindings for the checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

* `Ref`     -- oracle/_ref/libpqref.so: the UNMODIFIED reference compiled from
               /root/reference (oracle/Makefile `ref` target) behind oracle/ref_shim.cpp.
               Present in this container and shipped to the GPU box as a built file;
               absent => tests that need it skip.
* `Oracle`  -- oracle/liboracle.so: the plain-C restatement (pq_oracle.c, regex_oracle.c).

Both return "value dumps" (oracle/valdump.h) converted to numpy arrays so that parity is a
slot-by-slot array comparison: is_null, variant index, payload bits, string bytes.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libpqref.so")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")

# Parquet enums (reference include/common.hpp:16-64)
BOOLEAN, INT32, INT64, INT96, FLOAT, DOUBLE, BYTE_ARRAY, FLBA = range(8)
REQUIRED, OPTIONAL, REPEATED = range(3)
UTF8 = 0


class ValDump(C.Structure):
    _fields_ = [("n", C.c_int64), ("is_null", C.POINTER(C.c_uint8)), ("vidx", C.POINTER(C.c_uint8)),
                ("fixed", C.POINTER(C.c_uint64)), ("str_off", C.POINTER(C.c_uint64)),
                ("chars", C.POINTER(C.c_uint8)), ("chars_len", C.c_int64)]


class PageDump(C.Structure):
    _fields_ = [("n_pages", C.c_int64), ("page_num", C.POINTER(C.c_int32)),
                ("page_type", C.POINTER(C.c_int32)), ("num_values", C.POINTER(C.c_int32)),
                ("first_value", C.POINTER(C.c_int64)), ("values", ValDump)]


class StrDump(C.Structure):
    _fields_ = [("n", C.c_int64), ("pos", C.POINTER(C.c_uint64)), ("off", C.POINTER(C.c_uint64)),
                ("chars", C.POINTER(C.c_uint8))]


class ColInfo(C.Structure):
    _fields_ = [("name", C.c_char * 256), ("type", C.c_int32), ("column_index", C.c_int32),
                ("max_def_level", C.c_int32), ("max_rep_level", C.c_int32),
                ("repetition", C.c_int32), ("converted", C.c_int32)]


class PageEntry(C.Structure):
    _fields_ = [("data_offset", C.c_uint64), ("data_size", C.c_uint64),
                ("row_group_idx", C.c_uint64), ("column_idx", C.c_uint64)]


class ColSpec(C.Structure):
    _fields_ = [("name", C.c_char_p), ("type", C.c_int32), ("repetition", C.c_int32),
                ("converted", C.c_int32)]


class ColIn(C.Structure):
    _fields_ = [("is_null", C.c_void_p), ("fixed", C.c_void_p), ("str_off", C.c_void_p),
                ("chars", C.c_void_p)]


def _np(ptr, n, dtype):
    if n <= 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class Values:
    """A decoded column in dump form (numpy arrays)."""

    def __init__(self, is_null, vidx, fixed, str_off, chars):
        self.is_null, self.vidx, self.fixed, self.str_off, self.chars = is_null, vidx, fixed, str_off, chars

    @property
    def n(self):
        return len(self.is_null)

    @classmethod
    def from_dump(cls, d):
        n = d.n
        return cls(_np(d.is_null, n, np.uint8), _np(d.vidx, n, np.uint8), _np(d.fixed, n, np.uint64),
                   _np(d.str_off, n + 1, np.uint64), _np(d.chars, d.chars_len, np.uint8))

    def slice(self, a, b):
        base = self.str_off[a]
        return Values(self.is_null[a:b], self.vidx[a:b], self.fixed[a:b],
                      self.str_off[a:b + 1] - base, self.chars[int(base):int(self.str_off[b])])

    def string(self, i):
        return bytes(self.chars[int(self.str_off[i]):int(self.str_off[i + 1])])

    def diff(self, other):
        """None when identical slot by slot, else a short description of the first mismatch."""
        if self.n != other.n:
            return f"length {self.n} != {other.n}"
        for name in ("is_null", "vidx", "fixed"):
            a, b = getattr(self, name), getattr(other, name)
            bad = np.nonzero(a != b)[0]
            if len(bad):
                i = int(bad[0])
                return f"{name}[{i}]: {a[i]} != {b[i]} ({len(bad)} mismatches)"
        la = np.diff(self.str_off.astype(np.int64))
        lb = np.diff(other.str_off.astype(np.int64))
        bad = np.nonzero(la != lb)[0]
        if len(bad):
            i = int(bad[0])
            return f"string length[{i}]: {la[i]} != {lb[i]}"
        if len(self.chars) != len(other.chars) or not np.array_equal(self.chars, other.chars):
            bad = np.nonzero(self.chars != other.chars)[0]
            return f"chars differ at byte {int(bad[0]) if len(bad) else -1}"
        return None


class Pages:
    def __init__(self, page_num, page_type, num_values, first_value, values):
        self.page_num, self.page_type, self.num_values = page_num, page_type, num_values
        self.first_value, self.values = first_value, values

    @classmethod
    def from_dump(cls, d):
        n = d.n_pages
        return cls(_np(d.page_num, n, np.int32), _np(d.page_type, n, np.int32), _np(d.num_values, n, np.int32),
                   _np(d.first_value, n + 1, np.int64), Values.from_dump(d.values))


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def build_ref():
    """Only possible where /root/reference exists (this container)."""
    if os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", ORACLE_DIR, "-s", "ref"], check=True)


class _Lib:
    prefix = ""

    def _fn(self, name, restype, *argtypes):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        f.argtypes = list(argtypes)
        return f

    def err(self):
        return self._fn("last_error", C.c_char_p)().decode()


class _ReaderMixin:
    """Common reader surface of Ref and Oracle (same function names modulo prefix)."""

    def _values(self, call, *args):
        d = ValDump()
        if call(*args, C.byref(d)) != 0:
            raise RuntimeError(self.err())
        v = Values.from_dump(d)
        self._fn("valdump_free", None, C.POINTER(ValDump))(C.byref(d))
        return v

    def read_column_by_idx(self, h, rg, col):
        return self._values(self._fn("read_column_by_idx", C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(ValDump)), h, rg, col)

    def read_column(self, h, name):
        return self._values(self._fn("read_column", C.c_int, C.c_void_p, C.c_char_p, C.POINTER(ValDump)), h, name.encode())

    def read_pages(self, h, rg, col):
        d = PageDump()
        if self._fn("read_pages", C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(PageDump))(h, rg, col, C.byref(d)) != 0:
            raise RuntimeError(self.err())
        p = Pages.from_dump(d)
        self._fn("pagedump_free", None, C.POINTER(PageDump))(C.byref(d))
        return p

    def num_rows(self, h):
        return self._fn("num_rows", C.c_int64, C.c_void_p)(h)

    def num_row_groups(self, h):
        return self._fn("num_row_groups", C.c_int64, C.c_void_p)(h)

    def num_columns(self, h):
        return self._fn("num_columns", C.c_int64, C.c_void_p)(h)

    def num_pages(self, h):
        return self._fn("num_pages", C.c_int64, C.c_void_p)(h)

    def column_info(self, h, col):
        ci = ColInfo()
        if self._fn("column_info", C.c_int, C.c_void_p, C.c_int, C.POINTER(ColInfo))(h, col, C.byref(ci)) != 0:
            raise RuntimeError(self.err())
        return dict(name=ci.name.decode(), type=ci.type, column_index=ci.column_index,
                    max_def_level=ci.max_def_level, max_rep_level=ci.max_rep_level,
                    repetition=ci.repetition, converted=ci.converted)

    def find_column(self, h, name):
        return self._fn("find_column", C.c_int, C.c_void_p, C.c_char_p)(h, name.encode())

    def page_index(self, h):
        n = self.num_pages(h)
        arr = (PageEntry * max(n, 1))()
        self._fn("page_index", C.c_int64, C.c_void_p, C.POINTER(PageEntry), C.c_int64)(h, arr, n)
        out = np.zeros((n, 4), dtype=np.uint64)
        for i in range(n):
            out[i] = (arr[i].data_offset, arr[i].data_size, arr[i].row_group_idx, arr[i].column_idx)
        return out

    def read_page_data(self, h, pid, cap=1 << 22):
        buf = (C.c_uint8 * cap)()
        n = self._fn("read_page_data", C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_uint8), C.c_int64)(h, pid, buf, cap)
        if n < 0:
            raise RuntimeError(self.err())
        return bytes(buf[:n])

    def read_pages_chunk(self, h, s, e, max_bytes, cap=1 << 22):
        buf = (C.c_uint8 * cap)()
        n = self._fn("read_pages_chunk", C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                     C.POINTER(C.c_uint8), C.c_int64)(h, s, e, max_bytes, buf, cap)
        if n < 0:
            raise RuntimeError(self.err())
        return bytes(buf[:n])

    def string_iterator(self, h, name):
        d = StrDump()
        if self._fn("string_iterator_dump", C.c_int, C.c_void_p, C.c_char_p, C.POINTER(StrDump))(h, name.encode(), C.byref(d)) != 0:
            raise RuntimeError(self.err())
        n = d.n
        pos = _np(d.pos, n, np.uint64)
        off = _np(d.off, n + 1, np.uint64)
        chars = _np(d.chars, int(off[-1]) if n else 0, np.uint8)
        self._fn("strdump_free", None, C.POINTER(StrDump))(C.byref(d))
        return pos, off, chars

    def chunk_index(self, h, name, chunk_size=4096):
        nrows = self.num_rows(h)
        t2c = np.zeros(max(nrows, 1), dtype=np.uint64)
        n = self._fn("chunk_index", C.c_int64, C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_int64)(
            h, name.encode(), chunk_size, t2c.ctypes.data, nrows)
        if n < 0:
            raise RuntimeError(self.err())
        return t2c[:nrows], n


class Ref(_Lib, _ReaderMixin):
    prefix = "ref_"

    def __init__(self):
        if not os.path.exists(REF_SO):
            build_ref()
        self.lib = C.CDLL(REF_SO)

    @staticmethod
    def available():
        return os.path.exists(REF_SO) or os.path.isdir("/root/reference/src")

    def open(self, path):
        h = self._fn("reader_open", C.c_void_p, C.c_char_p)(path.encode())
        if not h:
            raise RuntimeError(self.err())
        return h

    def close(self, h):
        self._fn("reader_close", None, C.c_void_p)(h)

    def row_group_num_rows(self, h, rg):
        return self._fn("row_group_num_rows", C.c_int64, C.c_void_p, C.c_int)(h, rg)

    # --- writer -----------------------------------------------------------------
    def write_file(self, path, specs, row_groups):
        """specs: [(name, type, repetition, converted|-1)];
        row_groups: list of row groups, each a list of columns; a column is
        dict(is_null=uint8[n]|None, fixed=uint64[n]) or dict(is_null=..., str_off=uint64[n+1], chars=uint8[])."""
        arr = (ColSpec * len(specs))()
        keep = []
        for i, (name, t, rep, conv) in enumerate(specs):
            b = name.encode()
            keep.append(b)
            arr[i] = ColSpec(b, t, rep, conv)
        w = self._fn("writer_open", C.c_void_p, C.c_char_p, C.c_int, C.POINTER(ColSpec))(path.encode(), len(specs), arr)
        if not w:
            raise RuntimeError(self.err())
        wr = self._fn("writer_write_row_group", C.c_int, C.c_void_p, C.c_int64, C.POINTER(ColIn))
        for rg in row_groups:
            cols = (ColIn * len(specs))()
            hold = []
            nrows = None
            for i, col in enumerate(rg):
                isn = col.get("is_null")
                if isn is not None:
                    isn = np.ascontiguousarray(isn, dtype=np.uint8)
                    hold.append(isn)
                if "fixed" in col:
                    fx = np.ascontiguousarray(col["fixed"], dtype=np.uint64)
                    hold.append(fx)
                    n = len(fx)
                    cols[i] = ColIn(isn.ctypes.data if isn is not None else None, fx.ctypes.data, None, None)
                else:
                    so = np.ascontiguousarray(col["str_off"], dtype=np.uint64)
                    ch = np.ascontiguousarray(col["chars"], dtype=np.uint8)
                    if len(ch) == 0:
                        ch = np.zeros(1, dtype=np.uint8)
                    hold += [so, ch]
                    n = len(so) - 1
                    cols[i] = ColIn(isn.ctypes.data if isn is not None else None, None, so.ctypes.data, ch.ctypes.data)
                nrows = n if nrows is None else nrows
                assert n == nrows
            if wr(w, nrows or 0, cols) != 0:
                raise RuntimeError(self.err())
        if self._fn("writer_close", C.c_int, C.c_void_p)(w) != 0:
            raise RuntimeError(self.err())

    # --- codecs -----------------------------------------------------------------
    def rle_encode(self, values, bw):
        v = np.ascontiguousarray(values, dtype=np.uint32)
        cap = len(v) * 6 + 64
        buf = np.zeros(cap, dtype=np.uint8)
        n = self._fn("rle_encode", C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64)(
            v.ctypes.data, len(v), bw, buf.ctypes.data, cap)
        assert n >= 0
        return buf[:n].copy()

    def rle_decode_i32(self, data, bw, count, size=None):
        d = np.concatenate([np.ascontiguousarray(data, dtype=np.uint8), np.zeros(64, dtype=np.uint8)])
        out = np.zeros(max(count, 1), dtype=np.int32)
        self._fn("rle_decode_i32", None, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32)(
            d.ctypes.data, len(data) if size is None else size, bw, out.ctypes.data, count)
        return out[:count]

    def time_read_chunks(self, path, rgs, cols, threads):
        rg = np.ascontiguousarray(rgs, dtype=np.int32)
        cl = np.ascontiguousarray(cols, dtype=np.int32)
        nv = C.c_int64(0)
        t = self._fn("time_read_chunks", C.c_double, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                     C.POINTER(C.c_int64))(path.encode(), rg.ctypes.data, cl.ctypes.data, len(rg), threads, C.byref(nv))
        if t < 0:
            raise RuntimeError(self.err())
        return t, nv.value


class Oracle(_Lib, _ReaderMixin):
    prefix = "orc_"

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        self.lib = C.CDLL(ORACLE_SO)

    def open(self, path):
        h = self._fn("open", C.c_void_p, C.c_char_p)(path.encode())
        if not h:
            raise RuntimeError(self.err())
        return h

    def close(self, h):
        self._fn("close", None, C.c_void_p)(h)

    def row_group_num_rows(self, h, rg):
        return self._fn("row_group_num_rows", C.c_int64, C.c_void_p, C.c_int)(h, rg)

    def rle_decode_i32(self, data, bw, count, size=None, avail=None):
        d = np.concatenate([np.ascontiguousarray(data, dtype=np.uint8), np.zeros(64, dtype=np.uint8)])
        out = np.zeros(max(count, 1), dtype=np.int32)
        size = len(data) if size is None else size
        avail = len(data) if avail is None else avail
        self._fn("rle_decode_i32", None, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32)(
            d.ctypes.data, size, avail, bw, out.ctypes.data, count)
        return out[:count]

    def regex_search(self, pattern, text):
        if isinstance(pattern, str):
            pattern = pattern.encode()
        t = np.frombuffer(bytes(text) + b"\0", dtype=np.uint8)
        r = self._fn("regex_search", C.c_int, C.c_char_p, C.c_void_p, C.c_int64)(pattern, t.ctypes.data, len(text))
        if r < 0:
            raise ValueError(self._fn("regex_last_error", C.c_char_p)().decode())
        return bool(r)

    def regex_prune(self, h, col, pattern, neg):
        if isinstance(pattern, str):
            pattern = pattern.encode()
        cap = max(self.num_pages(h), 1)
        bits = np.zeros(cap, dtype=np.uint8)
        n = self._fn("regex_prune", C.c_int64, C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_int64)(
            h, col, pattern, int(neg), bits.ctypes.data, cap)
        if n < 0:
            raise ValueError(self._fn("regex_last_error", C.c_char_p)().decode())
        return bits[:n]

    def page_chunk_index(self, h, col, chunk_size=4096):
        cap = max(self.num_pages(h), 1)
        pc = np.zeros(cap, dtype=np.uint32)
        po = np.zeros(cap, dtype=np.uint32)
        cf = np.zeros(cap, dtype=np.uint32)
        first = C.c_int64(0)
        ncol = C.c_int64(0)
        n = self._fn("page_chunk_index", C.c_int64, C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p,
                     C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64))(
            h, col, chunk_size, pc.ctypes.data, po.ctypes.data, cf.ctypes.data, cap, C.byref(first), C.byref(ncol))
        if n < 0:
            raise RuntimeError(self.err())
        return pc[:ncol.value], po[:ncol.value], cf[:n]


# ── synthetic columns (numpy) shared by tests, golden generation and bench ─────────────

def strings_to_col(strs, is_null=None):
    lens = np.array([len(s) for s in strs], dtype=np.uint64)
    off = np.zeros(len(strs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    chars = np.frombuffer(b"".join(strs), dtype=np.uint8) if len(strs) else np.zeros(0, dtype=np.uint8)
    return dict(is_null=is_null, str_off=off, chars=chars)


def fixed_col(arr, is_null=None):
    a = np.ascontiguousarray(arr)
    if a.dtype.itemsize == 8:
        fx = a.view(np.uint64)
    elif a.dtype.itemsize == 4:
        fx = a.view(np.uint32).astype(np.uint64)
    else:
        fx = a.astype(np.uint64)
    return dict(is_null=is_null, fixed=fx)


def re2_page_bits(src, h, col, pattern, neg):
    """the frozen spec of a-19 evaluated with RE2 itself (pyarrow's match_substring_regex) over the per-page
    values of ColumnReader::read_pages as `src` (oracle or compiled reference) returns them"""
    import pyarrow as pa
    import pyarrow.compute as pc
    nrg = src.num_row_groups(h)
    bits = []
    for rg in range(nrg):
        pages = src.read_pages(h, rg, col)
        v = pages.values
        strs = [None if v.is_null[i] else v.chars[int(v.str_off[i]):int(v.str_off[i + 1])].tobytes().decode() for i in range(v.n)]
        m = np.array(pc.match_substring_regex(pa.array(strs, type=pa.string()), pattern).fill_null(False).to_pylist(), dtype=bool)
        isn = v.is_null.astype(bool)
        pred = (~m if neg else m) & ~isn
        for k in range(len(pages.page_num)):
            if pages.page_type[k] != 0:
                continue
            a, b = int(pages.first_value[k]), int(pages.first_value[k + 1])
            bits.append(1 if pred[a:b].any() else 0)
    return np.array(bits, dtype=np.uint8)
