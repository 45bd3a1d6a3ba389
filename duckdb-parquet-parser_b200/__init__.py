"""B200-native Parquet data-page decoder / page-pruning scanner -- Python (ctypes) binding.

The product is the C-ABI shared library `libpqg.so` built from `csrc/` (hand-written
sm_100a CUDA kernels + the C-ABI of include/pqg.h) and `host/` (C++ host reader mirroring the
reference's ParquetReader / ColumnReader interface, C-ABI in include/pqg_reader.h).  This
module only binds it for tests and bench.py; it contains no decode logic and NO fallback:
if the library is missing, or no CUDA device is usable, calls fail loudly.

The directory name contains '-', so import it with
    importlib.import_module("duckdb-parquet-parser_b200")
(the repo root's `pqb200.py` does that and re-exports everything).
"""
import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libpqg.so")

# Parquet physical types / repetition (values fixed by the format)
BOOLEAN, INT32, INT64, INT96, FLOAT, DOUBLE, BYTE_ARRAY, FIXED_LEN_BYTE_ARRAY = range(8)
PQG_OK, PQG_ERR_CUDA, PQG_ERR_ARG, PQG_ERR_UNSUPPORTED, PQG_ERR_PAGE, PQG_ERR_REGEX, PQG_ERR_NOMEM = range(7)
PQG_OPT_PARTITIONED_DICT = 1
PQG_PAGE_FLAG_DICT, PQG_PAGE_FLAG_V2, PQG_PAGE_FLAG_NO_NULLS, PQG_PAGE_FLAG_LEVELS_SEEN = 1, 2, 4, 8  # pqg_page_desc.flags; bits 8..15 = DataPageHeader.encoding


def build(verbose=False):
    """Compile libpqg.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    for target in ([], ["tools"]):
        r = subprocess.run(["make", "-C", PKG_DIR, "-j8"] + target, capture_output=not verbose, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libpqg.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


class ChunkDesc(C.Structure):
    _fields_ = [("dict_off", C.c_uint64), ("out_row_base", C.c_uint64), ("num_values", C.c_uint64),
                ("dict_size", C.c_uint32), ("dict_num_values", C.c_uint32), ("first_page", C.c_uint32),
                ("n_pages", C.c_uint32), ("row_group", C.c_uint32), ("column", C.c_uint32),
                ("max_def", C.c_int16), ("max_rep", C.c_int16), ("phys_type", C.c_uint8),
                ("has_dict", C.c_uint8), ("reserved", C.c_uint8 * 2)]


class PageDesc(C.Structure):
    _fields_ = [("payload_off", C.c_uint64), ("out_row_base", C.c_uint64), ("payload_size", C.c_uint32),
                ("num_values", C.c_uint32), ("chunk_idx", C.c_uint32), ("flags", C.c_uint32)]


class PageError(C.Structure):
    _fields_ = [("count", C.c_uint32), ("page", C.c_uint32), ("code", C.c_uint32), ("pos", C.c_uint32),
                ("need", C.c_uint32), ("size", C.c_uint32)]


class Timings(C.Structure):
    _fields_ = [("dict_ms", C.c_float), ("fixed_ms", C.c_float), ("str_size_ms", C.c_float),
                ("str_copy_ms", C.c_float), ("total_ms", C.c_float), ("launches", C.c_uint32), ("general_ms", C.c_float),
                ("tile_launches", C.c_uint32)]


class ValDump(C.Structure):
    _fields_ = [("n", C.c_int64), ("is_null", C.POINTER(C.c_uint8)), ("vidx", C.POINTER(C.c_uint8)),
                ("fixed", C.POINTER(C.c_uint64)), ("str_off", C.POINTER(C.c_uint64)),
                ("chars", C.POINTER(C.c_uint8)), ("chars_len", C.c_int64)]


class PageDump(C.Structure):
    _fields_ = [("n_pages", C.c_int64), ("page_num", C.POINTER(C.c_int32)), ("page_type", C.POINTER(C.c_int32)),
                ("num_values", C.POINTER(C.c_int32)), ("first_value", C.POINTER(C.c_int64)), ("values", ValDump)]


class StrDump(C.Structure):
    _fields_ = [("n", C.c_int64), ("pos", C.POINTER(C.c_uint64)), ("off", C.POINTER(C.c_uint64)),
                ("chars", C.POINTER(C.c_uint8))]


class ColInfo(C.Structure):
    _fields_ = [("name", C.c_char * 256), ("type", C.c_int32), ("column_index", C.c_int32),
                ("max_def_level", C.c_int32), ("max_rep_level", C.c_int32), ("repetition", C.c_int32),
                ("converted", C.c_int32)]


class PageEntry(C.Structure):
    _fields_ = [("data_offset", C.c_uint64), ("data_size", C.c_uint64), ("row_group_idx", C.c_uint64),
                ("column_idx", C.c_uint64)]


class Columnar(C.Structure):
    _fields_ = [("type", C.c_int32), ("width", C.c_uint32), ("num_slots", C.c_uint64), ("has_validity", C.c_int32),
                ("n_chunks", C.c_uint32), ("values", C.c_void_p), ("validity", C.c_void_p), ("offsets", C.c_void_p),
                ("char_bases", C.c_void_p), ("chars", C.c_void_p), ("chars_size", C.c_uint64),
                ("chunk_row_base", C.c_void_p), ("bytes_in", C.c_uint64), ("bytes_out", C.c_uint64),
                ("kernel_ms", C.c_float), ("owner", C.c_void_p)]


class GenCol(C.Structure):
    _fields_ = [("name", C.c_char_p), ("type", C.c_int32), ("repetition", C.c_int32), ("converted", C.c_int32),
                ("fixed", C.c_void_p), ("str_off", C.c_void_p), ("chars", C.c_void_p), ("is_null", C.c_void_p)]


class Dst(C.Structure):
    _fields_ = [("values", C.c_void_p), ("values_cap", C.c_uint64), ("validity", C.c_void_p), ("validity_cap", C.c_uint64)]


class ReadStats(C.Structure):
    _fields_ = [("num_slots", C.c_uint64), ("width", C.c_uint32), ("has_validity", C.c_int32), ("bytes_in", C.c_uint64),
                ("bytes_out", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


class StringsDst(C.Structure):
    _fields_ = [("offsets", C.c_void_p), ("offsets_cap", C.c_uint64), ("chars", C.c_void_p), ("chars_cap", C.c_uint64),
                ("validity", C.c_void_p), ("validity_cap", C.c_uint64), ("char_bases", C.c_void_p), ("char_bases_cap", C.c_uint64)]


class StringsStats(C.Structure):
    _fields_ = [("num_slots", C.c_uint64), ("n_chunks", C.c_uint64), ("chars_size", C.c_uint64), ("has_validity", C.c_int32),
                ("reserved", C.c_int32), ("bytes_in", C.c_uint64), ("bytes_out", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64)]


class H2dRange(C.Structure):
    _fields_ = [("host", C.c_void_p), ("image_off", C.c_uint64), ("len", C.c_uint64), ("chunk", C.c_uint32), ("reserved", C.c_uint32)]


class Tables(C.Structure):
    _fields_ = [("n_chunks", C.c_uint32), ("n_pages", C.c_uint32), ("chunks", C.POINTER(ChunkDesc)),
                ("pages", C.POINTER(PageDesc)), ("total_slots", C.c_uint64)]


# every symbol include/pqg.h and include/pqg_reader.h declare (checked by the CPU tests)
PQG_SYMBOLS = [
    "pqg_ctx_create", "pqg_ctx_destroy", "pqg_last_error", "pqg_ctx_sync", "pqg_ctx_set_profiling", "pqg_device_count",
    "pqg_kernel_launches", "pqg_upload", "pqg_wrap_device", "pqg_buf_alloc", "pqg_buf_write", "pqg_buf_size",
    "pqg_buf_free", "pqg_buf_device_ptr", "pqg_host_alloc", "pqg_host_free", "pqg_plan_create", "pqg_plan_create_dict_indices", "pqg_plan_destroy",
    "pqg_plan_set_image", "pqg_plan_set_option", "pqg_plan_run", "pqg_plan_run_pipelined", "pqg_plan_finish", "pqg_plan_timings", "pqg_plan_timings_avg", "pqg_plan_num_slots",
    "pqg_plan_value_width", "pqg_plan_values", "pqg_plan_validity", "pqg_plan_offsets", "pqg_plan_chars",
    "pqg_plan_chars_size", "pqg_plan_char_bases", "pqg_plan_bytes_in", "pqg_plan_bytes_out", "pqg_plan_download",
    "pqg_regex_compile", "pqg_dfa_free", "pqg_dfa_num_states", "pqg_dfa_match_host", "pqg_regex_scan",
    "pqg_chunk_index", "pqg_page_chunk_index", "pqg_chunk_index_prepare", "pqg_chunk_index_stitch", "pqg_chunk_index_emit",
    "pqg_chunk_job_ids", "pqg_chunk_job_total_weight", "pqg_chunk_job_free", "pqg_plan_create_ext", "pqg_plan_filter",
]
PQR_SYMBOLS = [
    "pqr_last_error", "pqr_open", "pqr_open_memory", "pqr_close", "pqr_num_rows", "pqr_num_row_groups",
    "pqr_num_columns", "pqr_num_pages", "pqr_row_group_num_rows", "pqr_column_info", "pqr_find_column",
    "pqr_schema_string", "pqr_page_scan_seconds", "pqr_file_size", "pqr_page_index", "pqr_read_page_data",
    "pqr_read_pages_chunk", "pqr_read_column_by_idx", "pqr_read_column", "pqr_read_column_rg", "pqr_read_pages",
    "pqr_string_iterator_dump", "pqr_valdump_free", "pqr_pagedump_free", "pqr_strdump_free", "pqr_read_columnar",
    "pqr_columnar_free", "pqr_read_columns_into", "pqr_read_dictionary_indices_into", "pqr_chunk_dictionary", "pqr_release_plans", "pqr_column_tables", "pqr_columns_tables", "pqr_tables_free", "pqr_chunk_index", "pqr_regex_prune",
    "pqr_page_chunk_index", "pqr_shard_row_groups", "pqr_regex_prune_rgs", "pqr_chunk_index_rgs", "pqr_read_columns_into_rgs",
    "pqr_chunk_index_prepare_rgs", "pqr_chunk_index_stitch", "pqr_chunk_index_emit", "pqr_chunk_job_free", "pqr_column_tables_rgs",
    "pqr_set_extensions", "pqr_read_strings_into",
]

PQGEN_SYMBOLS = ["pqgen_last_error", "pqgen_encode", "pqgen_size", "pqgen_emit", "pqgen_write_file", "pqgen_free", "pqgen_string_len", "pqgen_fill_strings"]
PQGEN_EMAILS, PQGEN_CITY64K = 0, 1
PQG_CMP_EQ, PQG_CMP_NE, PQG_CMP_LT, PQG_CMP_LE, PQG_CMP_GT, PQG_CMP_GE = range(6)

_lib = None


def lib():
    """The loaded libpqg.so.  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (make -C {PKG_DIR}); "
                               "this package has no CPU / PyTorch fallback")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L):
    vp, u64, u32, i64, i32, cp = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64, C.c_int, C.c_char_p

    def d(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    d("pqg_ctx_create", i32, i32, vp, C.POINTER(vp))
    d("pqg_ctx_destroy", None, vp)
    d("pqg_last_error", cp, vp)
    d("pqg_ctx_sync", i32, vp)
    d("pqg_ctx_set_profiling", i32, vp, i32)
    d("pqg_device_count", i32)
    d("pqg_kernel_launches", u64, vp)
    d("pqg_upload", i32, vp, vp, u64, C.POINTER(vp))
    d("pqg_wrap_device", i32, vp, vp, u64, u64, C.POINTER(vp))
    d("pqg_buf_alloc", i32, vp, u64, C.POINTER(vp))
    d("pqg_buf_write", i32, vp, vp, u64, vp, u64)
    d("pqg_buf_size", u64, vp)
    d("pqg_buf_free", None, vp, vp)
    d("pqg_buf_device_ptr", vp, vp)
    d("pqg_host_alloc", vp, u64)
    d("pqg_host_free", None, vp)
    d("pqg_plan_create", i32, vp, vp, C.POINTER(ChunkDesc), u32, C.POINTER(PageDesc), u32, C.POINTER(vp))
    d("pqg_plan_create_dict_indices", i32, vp, vp, C.POINTER(ChunkDesc), u32, C.POINTER(PageDesc), u32, C.POINTER(vp))
    d("pqg_plan_destroy", None, vp, vp)
    d("pqg_plan_set_image", i32, vp, vp, vp)
    d("pqg_plan_set_option", i32, vp, i32, i32)
    d("pqg_plan_run", i32, vp, vp)
    d("pqg_plan_run_pipelined", i32, vp, vp, vp, C.POINTER(H2dRange), u32, vp, vp)
    d("pqg_plan_finish", i32, vp, vp, C.POINTER(PageError))
    d("pqg_plan_timings", i32, vp, C.POINTER(Timings))
    d("pqg_plan_timings_avg", i32, vp, u32, C.POINTER(Timings), C.POINTER(u32))
    d("pqg_plan_num_slots", u64, vp)
    d("pqg_plan_value_width", u32, vp)
    for n in ("pqg_plan_values", "pqg_plan_validity", "pqg_plan_offsets", "pqg_plan_chars"):
        d(n, vp, vp)
    d("pqg_plan_chars_size", u64, vp)
    d("pqg_plan_char_bases", i32, vp, vp, vp, u32)
    d("pqg_plan_bytes_in", u64, vp)
    d("pqg_plan_bytes_out", u64, vp)
    d("pqg_plan_download", i32, vp, vp, vp, vp, vp, vp)
    d("pqg_regex_compile", i32, cp, C.POINTER(vp), C.c_char_p, C.c_size_t)
    d("pqg_dfa_free", None, vp)
    d("pqg_dfa_num_states", u32, vp)
    d("pqg_dfa_match_host", i32, vp, vp, u64)
    d("pqg_regex_scan", i32, vp, vp, vp, i32, vp, C.POINTER(C.c_float))
    d("pqg_plan_filter", i32, vp, vp, i32, i32, vp, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_float))
    d("pqg_chunk_index", i32, vp, vp, u64, u64, u32, vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_float))
    d("pqg_page_chunk_index", i32, vp, vp, u32, u64, vp, vp, vp, u32, C.POINTER(u32))
    d("pqg_chunk_index_prepare", i32, vp, vp, u64, C.POINTER(vp), C.POINTER(C.c_float))
    d("pqg_chunk_index_stitch", i32, vp, vp, u64, C.POINTER(u64), C.POINTER(u64))
    d("pqg_chunk_index_emit", i32, vp, vp, u32, vp, C.POINTER(C.c_float))
    d("pqg_chunk_job_ids", vp, vp)
    d("pqg_chunk_job_total_weight", u64, vp)
    d("pqg_chunk_job_free", None, vp, vp)

    d("pqgen_last_error", cp)
    d("pqgen_encode", vp, C.POINTER(GenCol), i32, vp, i32, i32)
    d("pqgen_size", u64, vp)
    d("pqgen_emit", i32, vp, vp, u64)
    d("pqgen_write_file", i32, vp, cp)
    d("pqgen_free", None, vp)
    d("pqgen_string_len", u32, i32)
    d("pqgen_fill_strings", i32, i32, u64, u64, u64, vp, vp, u64, vp, u32, i32)

    d("pqr_last_error", cp)
    d("pqr_open", vp, cp, i32)
    d("pqr_open_memory", vp, vp, u64, i32)
    d("pqr_close", None, vp)
    d("pqr_set_extensions", None, vp, i32)
    for n in ("pqr_num_rows", "pqr_num_row_groups", "pqr_num_columns", "pqr_num_pages"):
        d(n, i64, vp)
    d("pqr_row_group_num_rows", i64, vp, i32)
    d("pqr_column_info", i32, vp, i32, C.POINTER(ColInfo))
    d("pqr_find_column", i32, vp, cp)
    d("pqr_schema_string", i32, vp, C.c_char_p, i64)
    d("pqr_page_scan_seconds", C.c_double, vp)
    d("pqr_file_size", u64, vp)
    d("pqr_page_index", i64, vp, C.POINTER(PageEntry), i64)
    d("pqr_read_page_data", i64, vp, i64, vp, i64)
    d("pqr_read_pages_chunk", i64, vp, i64, i64, i64, vp, i64)
    d("pqr_read_column_by_idx", i32, vp, i32, i32, C.POINTER(ValDump))
    d("pqr_read_column", i32, vp, cp, C.POINTER(ValDump))
    d("pqr_read_column_rg", i32, vp, cp, i64, C.POINTER(ValDump))
    d("pqr_read_pages", i32, vp, i32, i32, C.POINTER(PageDump))
    d("pqr_string_iterator_dump", i32, vp, cp, C.POINTER(StrDump))
    d("pqr_valdump_free", None, C.POINTER(ValDump))
    d("pqr_pagedump_free", None, C.POINTER(PageDump))
    d("pqr_strdump_free", None, C.POINTER(StrDump))
    d("pqr_read_columnar", i32, vp, i32, i32, C.POINTER(Columnar))
    d("pqr_columnar_free", None, C.POINTER(Columnar))
    d("pqr_read_columns_into", i32, vp, C.POINTER(C.c_int32), i32, i32, C.POINTER(Dst), C.POINTER(ReadStats))
    d("pqr_read_dictionary_indices_into", i32, vp, i32, i64, i64, C.POINTER(Dst), C.POINTER(ReadStats))
    d("pqr_read_strings_into", i32, vp, i32, i64, i64, C.POINTER(StringsDst), C.POINTER(StringsStats))
    d("pqr_chunk_dictionary", i32, vp, i32, i64, vp, i64, vp, i64, C.POINTER(i64), C.POINTER(i64))
    d("pqr_release_plans", None, vp)
    d("pqr_shard_row_groups", i32, vp, i32, i32, C.POINTER(C.c_int32))
    d("pqr_regex_prune_rgs", i64, vp, i32, i64, i64, cp, i32, vp, i64, C.POINTER(C.c_float))
    d("pqr_chunk_index_rgs", i64, vp, cp, i64, i64, u64, u64, u32, vp, i64, C.POINTER(u64))
    d("pqr_read_columns_into_rgs", i32, vp, C.POINTER(C.c_int32), i32, i64, i64, C.POINTER(Dst), C.POINTER(ReadStats))
    d("pqr_chunk_index_prepare_rgs", vp, vp, cp, i64, i64, u64, C.POINTER(u64), C.POINTER(C.c_float), C.POINTER(C.c_float))
    d("pqr_chunk_index_stitch", i64, vp, u64, C.POINTER(u64))
    d("pqr_chunk_index_emit", i32, vp, u32, vp, i64, C.POINTER(C.c_float))
    d("pqr_chunk_job_free", None, vp)
    d("pqr_column_tables_rgs", i32, vp, i32, i64, i64, C.POINTER(Tables))
    d("pqr_column_tables", i32, vp, i32, i32, C.POINTER(Tables))
    d("pqr_columns_tables", i32, vp, C.POINTER(C.c_int), i32, i32, C.POINTER(Tables))
    d("pqr_tables_free", None, C.POINTER(Tables))
    d("pqr_chunk_index", i64, vp, cp, u64, vp, i64)
    d("pqr_regex_prune", i64, vp, i32, cp, i32, vp, i64, C.POINTER(C.c_float))
    d("pqr_page_chunk_index", i64, vp, i32, u64, vp, vp, vp, i64, C.POINTER(i64), C.POINTER(i64))


def _arr(ptr, n, dtype):
    if n <= 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    ct = np.ctypeslib.as_ctypes_type(dtype)
    p = C.cast(ptr, C.POINTER(ct))
    return np.ctypeslib.as_array(p, shape=(n,)).astype(dtype, copy=True)


class PqgError(RuntimeError):
    pass


def _valdump_to_dict(d):
    n = d.n
    return dict(is_null=_arr(d.is_null, n, np.uint8), vidx=_arr(d.vidx, n, np.uint8),
                fixed=_arr(d.fixed, n, np.uint64), str_off=_arr(d.str_off, n + 1, np.uint64),
                chars=_arr(d.chars, d.chars_len, np.uint8))


class Reader:
    """pqg::ParquetReader through include/pqg_reader.h (reference: ParquetReader)."""

    def __init__(self, path=None, data=None, device=-1, extensions=False):
        """extensions=True: SNAPPY-compressed chunks and DATA_PAGE_V2 pages decode (beyond the reference, which refuses /
        skips them); the default keeps the reference's behaviour"""
        L = lib()
        self._keep = data
        if path is not None:
            self.h = L.pqr_open(path.encode(), device)
        else:
            arr = np.ascontiguousarray(data, dtype=np.uint8) if not isinstance(data, int) else None
            if arr is not None:
                self._keep = arr
                self.h = L.pqr_open_memory(arr.ctypes.data, arr.size, device)
            else:
                raise ValueError("data must be a uint8 array")
        if not self.h:
            raise PqgError(L.pqr_last_error().decode())
        if extensions:
            L.pqr_set_extensions(self.h, 1)

    @classmethod
    def from_pointer(cls, ptr, size, device=-1):
        self = cls.__new__(cls)
        self._keep = None
        self.h = lib().pqr_open_memory(ptr, size, device)
        if not self.h:
            raise PqgError(lib().pqr_last_error().decode())
        return self

    def close(self):
        if getattr(self, "h", None):
            lib().pqr_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise PqgError(lib().pqr_last_error().decode())
        return rc

    num_rows = property(lambda s: lib().pqr_num_rows(s.h))
    num_row_groups = property(lambda s: lib().pqr_num_row_groups(s.h))
    num_columns = property(lambda s: lib().pqr_num_columns(s.h))
    num_pages = property(lambda s: lib().pqr_num_pages(s.h))
    page_scan_seconds = property(lambda s: lib().pqr_page_scan_seconds(s.h))
    file_size = property(lambda s: lib().pqr_file_size(s.h))

    def row_group_num_rows(self, rg):
        return self._check(lib().pqr_row_group_num_rows(self.h, rg))

    def column_info(self, col):
        ci = ColInfo()
        self._check(lib().pqr_column_info(self.h, col, C.byref(ci)))
        return dict(name=ci.name.decode(), type=ci.type, column_index=ci.column_index,
                    max_def_level=ci.max_def_level, max_rep_level=ci.max_rep_level,
                    repetition=ci.repetition, converted=ci.converted)

    def find_column(self, name):
        return lib().pqr_find_column(self.h, name.encode())

    def schema_string(self):
        buf = C.create_string_buffer(1 << 16)
        self._check(lib().pqr_schema_string(self.h, buf, len(buf)))
        return buf.value.decode()

    def page_index(self):
        n = self.num_pages
        arr = (PageEntry * max(n, 1))()
        lib().pqr_page_index(self.h, arr, n)
        return np.frombuffer(arr, dtype=np.uint64).reshape(-1, 4)[:n].copy()

    def read_page_data(self, pid, cap=1 << 22):
        buf = np.zeros(cap, dtype=np.uint8)
        n = self._check(lib().pqr_read_page_data(self.h, pid, buf.ctypes.data, cap))
        return buf[:n].tobytes()

    def read_pages_chunk(self, s, e, max_bytes, cap=1 << 22):
        buf = np.zeros(cap, dtype=np.uint8)
        n = self._check(lib().pqr_read_pages_chunk(self.h, s, e, max_bytes, buf.ctypes.data, cap))
        return buf[:n].tobytes()

    def _vals(self, fn, *args):
        d = ValDump()
        self._check(fn(self.h, *args, C.byref(d)))
        out = _valdump_to_dict(d)
        lib().pqr_valdump_free(C.byref(d))
        return out

    def read_column_by_idx(self, rg, col):
        return self._vals(lib().pqr_read_column_by_idx, rg, col)

    def read_column(self, name, rg=None):
        if rg is None:
            return self._vals(lib().pqr_read_column, name.encode())
        return self._vals(lib().pqr_read_column_rg, name.encode(), rg)

    def read_pages(self, rg, col):
        d = PageDump()
        self._check(lib().pqr_read_pages(self.h, rg, col, C.byref(d)))
        n = d.n_pages
        out = dict(page_num=_arr(d.page_num, n, np.int32), page_type=_arr(d.page_type, n, np.int32),
                   num_values=_arr(d.num_values, n, np.int32), first_value=_arr(d.first_value, n + 1, np.int64),
                   values=_valdump_to_dict(d.values))
        lib().pqr_pagedump_free(C.byref(d))
        return out

    def string_iterator(self, name):
        d = StrDump()
        self._check(lib().pqr_string_iterator_dump(self.h, name.encode(), C.byref(d)))
        n = d.n
        pos = _arr(d.pos, n, np.uint64)
        off = _arr(d.off, n + 1, np.uint64)
        chars = _arr(d.chars, int(off[-1]) if n else 0, np.uint8)
        lib().pqr_strdump_free(C.byref(d))
        return pos, off, chars

    def read_columnar(self, col, rg=-1):
        c = Columnar()
        self._check(lib().pqr_read_columnar(self.h, col, rg, C.byref(c)))
        n, nc = c.num_slots, c.n_chunks
        out = dict(type=c.type, width=c.width, num_slots=n, has_validity=bool(c.has_validity), n_chunks=nc,
                   values=_arr(c.values, n * c.width, np.uint8),
                   validity=_arr(c.validity, (n + 31) // 32 if c.has_validity else 0, np.uint32),
                   offsets=_arr(c.offsets, n + nc if c.type == BYTE_ARRAY else 0, np.uint32),
                   char_bases=_arr(c.char_bases, nc + 1 if c.type == BYTE_ARRAY else 0, np.uint64),
                   chars=_arr(c.chars, c.chars_size, np.uint8),
                   chunk_row_base=_arr(c.chunk_row_base, nc, np.uint64),
                   bytes_in=c.bytes_in, bytes_out=c.bytes_out, kernel_ms=c.kernel_ms)
        lib().pqr_columnar_free(C.byref(c))
        return out

    def read_columns_into(self, cols, dsts, rg=-1):
        """Streaming read of fixed-width columns into caller-owned host buffers.
        dsts: per column (values_ptr, values_cap_bytes, validity_ptr | None, validity_cap_words).
        Returns per-column stats dicts."""
        n = len(cols)
        ci = (C.c_int32 * n)(*cols)
        ds = (Dst * n)()
        for i, (vp_, vc, mp, mc) in enumerate(dsts):
            ds[i] = Dst(vp_, vc, mp, mc)
        st = (ReadStats * n)()
        self._check(lib().pqr_read_columns_into(self.h, ci, n, rg, ds, st))
        return [dict(num_slots=s.num_slots, width=s.width, has_validity=bool(s.has_validity), bytes_in=s.bytes_in,
                     bytes_out=s.bytes_out, h2d_bytes=s.h2d_bytes, d2h_bytes=s.d2h_bytes) for s in st]

    def shard_row_groups(self, col, n_shards):
        out = (C.c_int32 * (n_shards + 1))()
        self._check(lib().pqr_shard_row_groups(self.h, col, n_shards, out))
        return list(out)

    def regex_prune_rgs(self, col, rg_begin, rg_end, pattern, neg=False):
        if isinstance(pattern, str):
            pattern = pattern.encode()
        cap = max(self.num_pages, 1)
        bits = np.zeros(cap, dtype=np.uint8)
        ms = C.c_float(0)
        n = self._check(lib().pqr_regex_prune_rgs(self.h, col, rg_begin, rg_end, pattern, int(neg), bits.ctypes.data, cap, C.byref(ms)))
        return bits[:n], ms.value

    def chunk_index_rgs(self, name, rg_begin, rg_end, chunk_size=4096, carry_in=0, id_base=0):
        """-> (chunk ids of the shard's rows: id_base + local id, nulls 0; shard chunk count; carry_out)"""
        rows = sum(self.row_group_num_rows(rg) for rg in range(rg_begin, rg_end))
        ids = np.zeros(rows + 1, dtype=np.uint32)
        carry = C.c_uint64(0)
        n = self._check(lib().pqr_chunk_index_rgs(self.h, name.encode(), rg_begin, rg_end, chunk_size, carry_in, id_base,
                                                  ids.ctypes.data, ids.size, C.byref(carry)))
        return ids[:rows], n, carry.value

    # chunk index of a shard in phases (multi-GPU: prepare everywhere, stitch in shard order, emit everywhere)
    def chunk_index_prepare_rgs(self, name, rg_begin, rg_end, chunk_size=4096):
        """-> job dict(h, num_slots, decode_ms, prepare_ms): upload + decode + everything that needs no carry"""
        slots, dms, pms = C.c_uint64(0), C.c_float(0), C.c_float(0)
        h = lib().pqr_chunk_index_prepare_rgs(self.h, name.encode(), rg_begin, rg_end, chunk_size, C.byref(slots), C.byref(dms), C.byref(pms))
        if not h:
            raise PqgError(lib().pqr_last_error().decode())
        return dict(h=h, num_slots=slots.value, decode_ms=dms.value, prepare_ms=pms.value)

    def chunk_index_stitch(self, job, carry_in):
        """-> (shard chunk count, carry_out); the only step ordered between shards"""
        carry = C.c_uint64(0)
        n = self._check(lib().pqr_chunk_index_stitch(job["h"], carry_in, C.byref(carry)))
        return n, carry.value

    def chunk_index_emit(self, job, id_base):
        """-> chunk ids of the shard's rows (id_base + local id, nulls 0); frees the job"""
        ids = np.zeros(job["num_slots"] + 1, dtype=np.uint32)
        ms = C.c_float(0)
        try:
            self._check(lib().pqr_chunk_index_emit(job["h"], id_base, ids.ctypes.data, ids.size, C.byref(ms)))
        finally:
            lib().pqr_chunk_job_free(job["h"])
            job["h"] = None
        job["emit_ms"] = ms.value
        return ids[:job["num_slots"]]

    def column_tables_rgs(self, col, rg_begin, rg_end):
        """descriptor tables of the row groups [rg_begin, rg_end) (file offsets)"""
        t = Tables()
        self._check(lib().pqr_column_tables_rgs(self.h, col, rg_begin, rg_end, C.byref(t)))
        chunks = (ChunkDesc * max(t.n_chunks, 1))()
        pages = (PageDesc * max(t.n_pages, 1))()
        C.memmove(chunks, t.chunks, C.sizeof(ChunkDesc) * t.n_chunks)
        C.memmove(pages, t.pages, C.sizeof(PageDesc) * t.n_pages)
        out = (chunks, t.n_chunks, pages, t.n_pages, t.total_slots)
        lib().pqr_tables_free(C.byref(t))
        return out

    def read_columns_into_rgs(self, cols, dsts, rg_begin, rg_end):
        n = len(cols)
        ci = (C.c_int32 * n)(*cols)
        ds = (Dst * n)()
        for i, (vp_, vc, mp, mc) in enumerate(dsts):
            ds[i] = Dst(vp_, vc, mp, mc)
        st = (ReadStats * n)()
        self._check(lib().pqr_read_columns_into_rgs(self.h, ci, n, rg_begin, rg_end, ds, st))
        return [dict(num_slots=s.num_slots, width=s.width, has_validity=bool(s.has_validity), bytes_in=s.bytes_in,
                     bytes_out=s.bytes_out, h2d_bytes=s.h2d_bytes, d2h_bytes=s.d2h_bytes) for s in st]

    def read_strings_into(self, col, rg_begin, rg_end, offsets, chars, char_bases, validity=None):
        """pipelined BYTE_ARRAY read into caller-owned numpy arrays / (ptr, capacity) pairs: offsets uint32, chars uint8,
        char_bases uint64, validity uint32 words or None.  Returns the stats dict (chars_size, n_chunks, ...)."""
        def pc(x, item):
            return (x[0], x[1]) if isinstance(x, tuple) else (x.ctypes.data, x.nbytes // item)
        (op, oc), (cp, cc), (bp, bc) = pc(offsets, 4), pc(chars, 1), pc(char_bases, 8)
        vp_, vc = pc(validity, 4) if validity is not None else (None, 0)
        ds = StringsDst(op, oc, cp, cc, vp_, vc, bp, bc)
        st = StringsStats()
        self._check(lib().pqr_read_strings_into(self.h, col, rg_begin, rg_end, C.byref(ds), C.byref(st)))
        return dict(num_slots=st.num_slots, n_chunks=st.n_chunks, chars_size=st.chars_size, has_validity=bool(st.has_validity),
                    bytes_in=st.bytes_in, bytes_out=st.bytes_out, h2d_bytes=st.h2d_bytes, d2h_bytes=st.d2h_bytes)

    def read_dictionary_indices(self, col, rg_begin=0, rg_end=None, out=None, validity=None):
        """dictionary-form read: (uint32 indices per slot, validity words | None, stats)"""
        if rg_end is None:
            rg_end = self.num_row_groups
        rows = sum(self.row_group_num_rows(rg) for rg in range(rg_begin, rg_end))
        idx = out if out is not None else np.zeros(rows, dtype=np.uint32)
        val = validity if validity is not None else np.zeros((rows + 31) // 32 + 1, dtype=np.uint32)
        ds = Dst(idx.ctypes.data, idx.nbytes, val.ctypes.data, val.size)
        st = ReadStats()
        self._check(lib().pqr_read_dictionary_indices_into(self.h, col, rg_begin, rg_end, C.byref(ds), C.byref(st)))
        stats = dict(num_slots=st.num_slots, width=st.width, has_validity=bool(st.has_validity), bytes_in=st.bytes_in,
                     bytes_out=st.bytes_out, h2d_bytes=st.h2d_bytes, d2h_bytes=st.d2h_bytes)
        return idx[:rows], (val if st.has_validity else None), stats

    def chunk_dictionary(self, col, rg):
        """(offsets uint32[n+1], chars bytes) of the dictionary page of (row group, column)"""
        n, nb = C.c_int64(0), C.c_int64(0)
        self._check(lib().pqr_chunk_dictionary(self.h, col, rg, None, 0, None, 0, C.byref(n), C.byref(nb)))
        off = np.zeros(n.value + 1, dtype=np.uint32)
        ch = np.zeros(max(nb.value, 1), dtype=np.uint8)
        self._check(lib().pqr_chunk_dictionary(self.h, col, rg, off.ctypes.data, off.size, ch.ctypes.data, ch.size, C.byref(n), C.byref(nb)))
        return off, ch[:nb.value].tobytes()

    def release_plans(self):
        lib().pqr_release_plans(self.h)

    def column_tables(self, col, rg=-1):
        """(ctypes array of ChunkDesc, ctypes array of PageDesc, total_slots), file offsets."""
        t = Tables()
        self._check(lib().pqr_column_tables(self.h, col, rg, C.byref(t)))
        chunks = (ChunkDesc * max(t.n_chunks, 1))()
        pages = (PageDesc * max(t.n_pages, 1))()
        C.memmove(chunks, t.chunks, C.sizeof(ChunkDesc) * t.n_chunks)
        C.memmove(pages, t.pages, C.sizeof(PageDesc) * t.n_pages)
        out = (chunks, t.n_chunks, pages, t.n_pages, t.total_slots)
        lib().pqr_tables_free(C.byref(t))
        return out

    def columns_tables(self, cols, rg=-1):
        """one table set for several fixed-width columns of the same value width: column k of
        `cols` owns the output slots [k * S, (k + 1) * S), S = total_slots / len(cols)"""
        t = Tables()
        arr = (C.c_int * len(cols))(*cols)
        self._check(lib().pqr_columns_tables(self.h, arr, len(cols), rg, C.byref(t)))
        chunks = (ChunkDesc * max(t.n_chunks, 1))()
        pages = (PageDesc * max(t.n_pages, 1))()
        C.memmove(chunks, t.chunks, C.sizeof(ChunkDesc) * t.n_chunks)
        C.memmove(pages, t.pages, C.sizeof(PageDesc) * t.n_pages)
        out = (chunks, t.n_chunks, pages, t.n_pages, t.total_slots)
        lib().pqr_tables_free(C.byref(t))
        return out

    def chunk_index(self, name, chunk_size=4096):
        nrows = self.num_rows
        t2c = np.zeros(max(nrows, 1), dtype=np.uint64)
        n = self._check(lib().pqr_chunk_index(self.h, name.encode(), chunk_size, t2c.ctypes.data, nrows))
        return t2c[:nrows], n

    def regex_prune(self, col, pattern, neg=False):
        if isinstance(pattern, str):
            pattern = pattern.encode()
        cap = max(self.num_pages, 1)
        bits = np.zeros(cap, dtype=np.uint8)
        ms = C.c_float(0)
        n = self._check(lib().pqr_regex_prune(self.h, col, pattern, int(neg), bits.ctypes.data, cap, C.byref(ms)))
        return bits[:n], ms.value

    def page_chunk_index(self, col, chunk_size=4096):
        cap = max(self.num_pages, 1)
        pc = np.zeros(cap, dtype=np.uint32)
        po = np.zeros(cap, dtype=np.uint32)
        cf = np.zeros(cap, dtype=np.uint32)
        first, ncol = C.c_int64(0), C.c_int64(0)
        n = self._check(lib().pqr_page_chunk_index(self.h, col, chunk_size, pc.ctypes.data, po.ctypes.data,
                                                   cf.ctypes.data, cap, C.byref(first), C.byref(ncol)))
        return pc[:ncol.value], po[:ncol.value], cf[:n]


class Context:
    """pqg_ctx: one GPU + one stream (include/pqg.h)."""

    def __init__(self, device=0, stream=None):
        L = lib()
        h = C.c_void_p()
        rc = L.pqg_ctx_create(device, stream, C.byref(h))
        if rc != PQG_OK:
            raise PqgError(L.pqg_last_error(None).decode())
        self.h = h

    def err(self):
        return lib().pqg_last_error(self.h).decode()

    def check(self, rc):
        if rc != PQG_OK:
            raise PqgError(self.err())

    def sync(self):
        self.check(lib().pqg_ctx_sync(self.h))

    def set_profiling(self, on):
        lib().pqg_ctx_set_profiling(self.h, int(on))

    @property
    def launches(self):
        return lib().pqg_kernel_launches(self.h)

    def upload(self, host_ptr, size):
        b = C.c_void_p()
        self.check(lib().pqg_upload(self.h, host_ptr, size, C.byref(b)))
        return b

    def wrap_device(self, dev_ptr, size, capacity=None):
        """capacity = bytes readable behind dev_ptr (>= size + 64; defaults to exactly that: the caller vouches for the tail)"""
        if capacity is None:
            capacity = size + 64
        b = C.c_void_p()
        self.check(lib().pqg_wrap_device(self.h, dev_ptr, size, capacity, C.byref(b)))
        return b

    def buf_free(self, b):
        lib().pqg_buf_free(self.h, b)

    def plan(self, image, tables):
        chunks, nc, pages, npg, _ = tables
        p = C.c_void_p()
        self.check(lib().pqg_plan_create(self.h, image, chunks, nc, pages, npg, C.byref(p)))
        return Plan(self, p)

    def close(self):
        if self.h:
            lib().pqg_ctx_destroy(self.h)
            self.h = None


class Plan:
    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    def run(self):
        self.ctx.check(lib().pqg_plan_run(self.ctx.h, self.h))

    def finish(self):
        pe = PageError()
        rc = lib().pqg_plan_finish(self.ctx.h, self.h, C.byref(pe))
        if rc != PQG_OK:
            raise PqgError(self.ctx.err())
        return pe

    def timings(self):
        t = Timings()
        lib().pqg_plan_timings(self.h, C.byref(t))
        return dict(dict_ms=t.dict_ms, fixed_ms=t.fixed_ms, str_size_ms=t.str_size_ms, str_copy_ms=t.str_copy_ms,
                    total_ms=t.total_ms, launches=t.launches, general_ms=t.general_ms, tile_launches=t.tile_launches)

    def filter(self, value_type, op, constant, n_slots, want_bits=True):
        """device-side predicate on the decoded column of n_slots slots (pqg_plan_filter): -> (row bitmap words | None, matches, kernel ms)"""
        dt = {INT32: np.int32, INT64: np.int64, FLOAT: np.float32, DOUBLE: np.float64}[value_type]
        c = np.array([constant], dtype=dt)
        bits = np.zeros((n_slots + 31) // 32 + 1, dtype=np.uint32) if want_bits else None
        cnt, ms = C.c_uint64(0), C.c_float(0)
        self.ctx.check(lib().pqg_plan_filter(self.ctx.h, self.h, value_type, op, c.ctypes.data, bits.ctypes.data if want_bits else None,
                                             C.byref(cnt), C.byref(ms)))
        return bits, cnt.value, ms.value

    def timings_avg(self, last_n=0):
        t, n = Timings(), C.c_uint32(0)
        lib().pqg_plan_timings_avg(self.h, last_n, C.byref(t), C.byref(n))
        return dict(dict_ms=t.dict_ms, fixed_ms=t.fixed_ms, str_size_ms=t.str_size_ms, str_copy_ms=t.str_copy_ms,
                    total_ms=t.total_ms, launches=t.launches, general_ms=t.general_ms, tile_launches=t.tile_launches, runs=n.value)

    num_slots = property(lambda s: lib().pqg_plan_num_slots(s.h))
    width = property(lambda s: lib().pqg_plan_value_width(s.h))
    bytes_in = property(lambda s: lib().pqg_plan_bytes_in(s.h))
    bytes_out = property(lambda s: lib().pqg_plan_bytes_out(s.h))
    chars_size = property(lambda s: lib().pqg_plan_chars_size(s.h))
    values_ptr = property(lambda s: lib().pqg_plan_values(s.h))
    validity_ptr = property(lambda s: lib().pqg_plan_validity(s.h))
    offsets_ptr = property(lambda s: lib().pqg_plan_offsets(s.h))
    chars_ptr = property(lambda s: lib().pqg_plan_chars(s.h))

    def set_option(self, option, value):
        self.ctx.check(lib().pqg_plan_set_option(self.h, option, value))

    def set_image(self, image):
        self.ctx.check(lib().pqg_plan_set_image(self.ctx.h, self.h, image))

    def download(self, values=None, validity=None, offsets=None, chars=None):
        self.ctx.check(lib().pqg_plan_download(self.ctx.h, self.h, values, validity, offsets, chars))

    def destroy(self):
        if self.h:
            lib().pqg_plan_destroy(self.ctx.h, self.h)
            self.h = None


def device_count():
    return lib().pqg_device_count()


def regex_compile(pattern):
    if isinstance(pattern, str):
        pattern = pattern.encode()
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib().pqg_regex_compile(pattern, C.byref(h), err, len(err))
    if rc != PQG_OK:
        raise ValueError(err.value.decode())
    return h


def dfa_match_host(dfa, text):
    t = np.frombuffer(bytes(text) + b"\0", dtype=np.uint8)
    return lib().pqg_dfa_match_host(dfa, t.ctypes.data, len(text))


_WIDTH = {BOOLEAN: 1, INT32: 4, FLOAT: 4, INT64: 8, DOUBLE: 8}


class Generated:
    """A parquet file image produced by the workload generator (include/pqg_gen.h)."""

    def __init__(self, job):
        self.job = job
        self.size = lib().pqgen_size(job)

    def emit(self, dst_ptr, cap):
        if lib().pqgen_emit(self.job, dst_ptr, cap) != 0:
            raise PqgError(lib().pqgen_last_error().decode())

    def to_numpy(self):
        out = np.empty(self.size, dtype=np.uint8)
        self.emit(out.ctypes.data, out.size)
        return out

    def write(self, path):
        if lib().pqgen_write_file(self.job, path.encode()) != 0:
            raise PqgError(lib().pqgen_last_error().decode())
        return path

    def free(self):
        if self.job:
            lib().pqgen_free(self.job)
            self.job = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def generate(specs, columns, rg_rows, threads=0):
    """specs: [(name, type, repetition, converted|-1)]; columns: one dict per column over ALL
    rows of the file: dict(fixed=array[, is_null=uint8[]]) or dict(str_off=uint64[n+1],
    chars=uint8[][, is_null]); `fixed` may be the natural dtype (int32/float32/int64/float64/
    uint8) or one uint64 per row holding the payload bits (tests/oraclelib.fixed_col)."""
    n = len(specs)
    arr = (GenCol * n)()
    hold = []
    for i, ((name, t, rep, conv), col) in enumerate(zip(specs, columns)):
        nb = name.encode()
        hold.append(nb)
        isn = col.get("is_null")
        if isn is not None:
            isn = np.ascontiguousarray(isn, dtype=np.uint8)
            hold.append(isn)
        g = GenCol(nb, t, rep, conv, None, None, None, isn.ctypes.data if isn is not None else None)
        if t == BYTE_ARRAY:
            so = np.ascontiguousarray(col["str_off"], dtype=np.uint64)
            ch = np.ascontiguousarray(col["chars"], dtype=np.uint8)
            if ch.size == 0:
                ch = np.zeros(1, dtype=np.uint8)
            hold += [so, ch]
            g.str_off, g.chars = so.ctypes.data, ch.ctypes.data
        else:
            w = _WIDTH[t]
            fx = np.ascontiguousarray(col["fixed"])
            if fx.dtype.itemsize != w:  # one uint64 of payload bits per row -> natural width
                fx = fx.astype(np.uint64).astype({1: np.uint8, 4: np.uint32, 8: np.uint64}[w])
            if fx.size == 0:
                fx = np.zeros(1, dtype=fx.dtype)
            hold.append(fx)
            g.fixed = fx.ctypes.data
        arr[i] = g
    rg = np.ascontiguousarray(rg_rows, dtype=np.int64)
    job = lib().pqgen_encode(arr, n, rg.ctypes.data, len(rg), threads)
    if not job:
        raise PqgError(lib().pqgen_last_error().decode())
    return Generated(job)


def synth_strings(kind, rows, seed, first_row=0, null_permille=0, threads=0, off_base=0, chars=None, str_off=None, is_null=None):
    """native filler of the BASELINE string workloads (include/pqg_gen.h: pqgen_fill_strings) -> column dict for generate().
    Buffers may be passed in (slices of larger arrays: mixed columns are filled row group by row group)."""
    L = lib().pqgen_string_len(kind)
    if chars is None:
        chars = np.empty(max(rows * L, 1), dtype=np.uint8)
    if str_off is None:
        str_off = np.empty(rows + 1, dtype=np.uint64)
    if is_null is None and null_permille:
        is_null = np.empty(max(rows, 1), dtype=np.uint8)
    if lib().pqgen_fill_strings(kind, first_row, rows, seed, chars.ctypes.data, str_off.ctypes.data, off_base,
                                is_null.ctypes.data if is_null is not None else None, null_permille, threads) != 0:
        raise PqgError(lib().pqgen_last_error().decode())
    col = dict(str_off=str_off, chars=chars)
    if is_null is not None:
        col["is_null"] = is_null
    return col
