/*
 * valdump.h -- TEST INFRASTRUCTURE (oracle side).
 *
 * A flat, C-ABI "dump" of a std::vector<Value> (reference include/common.hpp:177-201).
 * Both the compiled reference (oracle/_ref, via ref_shim.cpp) and the CPU restatement
 * (pq_oracle.c) emit this layout so that tests can compare slot by slot:
 *   is_null[i]  Value::is_null
 *   vidx[i]     std::variant index: 0 bool, 1 int32, 2 int64, 3 float, 4 double, 5 string
 *   fixed[i]    payload bits of the non-string alternatives, zero-extended to 64 bit
 *               (float/double as raw IEEE bits so NaNs compare bit-exactly)
 *   str_off[i]..str_off[i+1]  byte range in chars for string alternatives (empty otherwise)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use anything under oracle/.
 */
#ifndef PQ_ORACLE_VALDUMP_H
#define PQ_ORACLE_VALDUMP_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct valdump {
    int64_t n;
    uint8_t* is_null;
    uint8_t* vidx;
    uint64_t* fixed;
    uint64_t* str_off; /* n + 1 entries */
    uint8_t* chars;
    int64_t chars_len;
} valdump;

/* per-page dump used for ColumnReader::read_pages (reference column_reader.hpp:12-17) */
typedef struct pagedump {
    int64_t n_pages;
    int32_t* page_num;
    int32_t* page_type;
    int32_t* num_values;
    int64_t* first_value; /* n_pages + 1: slot range of each page inside values */
    valdump values;       /* all data-page values concatenated */
} pagedump;

#ifdef __cplusplus
}
#endif
#endif
