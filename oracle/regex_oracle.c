/*
 * regex_oracle.c -- TEST INFRASTRUCTURE: checker for the regex page-pruning mode.
 *
 * PARITY UNPINNED.  The reference checkout contains no source for
 * `--regex-column/--regex/--neg-regex` (README.md:54-64 only; re2 is not vendored and
 * not installed), so there are no call sites, tests or golden vectors to pin against.
 * The contract is the frozen spec in SURVEY.md section 8 (a-19):
 *   m(v)   = RE2-style partial match (search) of `pattern` in value v; ^ and $ honoured;
 *   m'(v)  = neg ? !m(v) : m(v); nulls never satisfy m';
 *   bit[p] = OR over the values of data page p of m'(v).
 * Supported syntax: literals, escapes, `.` (any code point but \n), classes / ranges /
 * negation, \d \w \s (+ upper-case complements), * + ? {m,n} (and their lazy forms, which
 * do not change match/no-match), alternation, groups, (?:...), ^ $.  Everything else is
 * rejected with an explicit message, like the product's host DFA compiler does.
 *
 * To stay independent from the product (NFA -> subset-construction DFA in C++), this
 * checker is a backtracking VM (split/jmp program, memoised on (pc, pos)); tests
 * cross-check both against Python's `re` on the common subset.
 */
#include "pq_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* shared error slot lives in pq_oracle.c; declare a setter through a tiny bridge */
static _Thread_local char rx_err[256];

enum { N_SET, N_CAT, N_ALT, N_STAR, N_PLUS, N_QUEST, N_REPEAT, N_BOL, N_EOL, N_EMPTY };

typedef struct node {
    int kind;
    uint8_t mask[32]; /* N_SET: byte membership for bytes < 0x80 (and raw bytes when !mb) */
    int mb;           /* N_SET: also matches any well-formed multi-byte UTF-8 sequence */
    int min, max;     /* N_REPEAT, max < 0 = unbounded */
    struct node *a, *b;
} node;

typedef struct {
    const uint8_t* p;
    const uint8_t* end;
    int failed;
    int depth;
} parser;

static node* new_node(int kind) {
    node* n = (node*)calloc(1, sizeof(node));
    n->kind = kind;
    return n;
}
static void free_node(node* n) {
    if (!n) return;
    free_node(n->a);
    free_node(n->b);
    free(n);
}
static void fail(parser* ps, const char* msg) {
    if (!ps->failed) snprintf(rx_err, sizeof(rx_err), "regex: %s", msg);
    ps->failed = 1;
}
static void set_bit(uint8_t* m, int c) { m[c >> 3] |= (uint8_t)(1u << (c & 7)); }
static int get_bit(const uint8_t* m, int c) { return (m[c >> 3] >> (c & 7)) & 1; }
static void set_range(uint8_t* m, int lo, int hi) { for (int c = lo; c <= hi; c++) set_bit(m, c); }

static void add_perl_class(uint8_t* m, int c) {
    uint8_t t[32];
    memset(t, 0, 32);
    switch (c) {
        case 'd': case 'D': set_range(t, '0', '9'); break;
        case 'w': case 'W': set_range(t, '0', '9'); set_range(t, 'A', 'Z'); set_range(t, 'a', 'z'); set_bit(t, '_'); break;
        case 's': case 'S': set_bit(t, '\t'); set_bit(t, '\n'); set_bit(t, '\f'); set_bit(t, '\r'); set_bit(t, ' '); break;
    }
    if (c == 'D' || c == 'W' || c == 'S')
        for (int i = 0; i < 16; i++) t[i] = (uint8_t)~t[i]; /* complement within ASCII */
    for (int i = 0; i < 32; i++) m[i] |= t[i];
}

static int hexval(int c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}

/* parses one escape after the backslash; returns a byte value 0..255, or -2 when it was a
 * perl class (added to *cls, *cls_mb set for upper-case complements), or -1 on error */
static int parse_escape(parser* ps, uint8_t* cls, int* cls_mb) {
    if (ps->p >= ps->end) { fail(ps, "trailing backslash"); return -1; }
    int c = *ps->p++;
    switch (c) {
        case 'd': case 'w': case 's': add_perl_class(cls, c); return -2;
        case 'D': case 'W': case 'S': add_perl_class(cls, c); *cls_mb = 1; return -2;
        case 'n': return '\n';
        case 't': return '\t';
        case 'r': return '\r';
        case 'f': return '\f';
        case 'v': return '\v';
        case 'a': return 7;
        case 'x': {
            if (ps->end - ps->p >= 2 && hexval(ps->p[0]) >= 0 && hexval(ps->p[1]) >= 0) {
                int v = hexval(ps->p[0]) * 16 + hexval(ps->p[1]);
                ps->p += 2;
                if (v >= 0x80) { fail(ps, "\\x escape above 0x7f is not supported"); return -1; }
                return v;
            }
            fail(ps, "bad \\x escape");
            return -1;
        }
        default:
            if ((c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z')) {
                char msg[64];
                snprintf(msg, sizeof(msg), "unsupported escape \\%c", c);
                fail(ps, msg);
                return -1;
            }
            if (c >= 0x80) { fail(ps, "escaped non-ASCII byte"); return -1; }
            return c; /* escaped punctuation */
    }
}

static node* parse_alt(parser* ps);

static node* parse_class(parser* ps) {
    node* n = new_node(N_SET);
    int negate = 0;
    if (ps->p < ps->end && *ps->p == '^') { negate = 1; ps->p++; }
    int first = 1;
    for (;;) {
        if (ps->p >= ps->end) { fail(ps, "missing ]"); break; }
        int c = *ps->p;
        if (c == ']' && !first) { ps->p++; break; }
        first = 0;
        int lo;
        if (c == '[' && ps->p + 1 < ps->end && ps->p[1] == ':') { fail(ps, "POSIX classes are not supported"); break; }
        if (c >= 0x80) { fail(ps, "non-ASCII in character class is not supported"); break; }
        ps->p++;
        if (c == '\\') {
            int mb = 0;
            lo = parse_escape(ps, n->mask, &mb);
            if (lo == -1) break;
            if (lo == -2) { if (mb) n->mb = 1; continue; }
        } else lo = c;
        if (ps->p + 1 < ps->end && ps->p[0] == '-' && ps->p[1] != ']') {
            ps->p++;
            int hi = *ps->p++;
            if (hi >= 0x80) { fail(ps, "non-ASCII in character class is not supported"); break; }
            if (hi == '\\') {
                uint8_t tmp[32]; int mb = 0;
                memset(tmp, 0, 32);
                hi = parse_escape(ps, tmp, &mb);
                if (hi < 0) { fail(ps, "bad range end"); break; }
            }
            if (hi < lo) { fail(ps, "bad character range"); break; }
            set_range(n->mask, lo, hi);
        } else set_bit(n->mask, lo);
    }
    if (negate) {
        for (int i = 0; i < 16; i++) n->mask[i] = (uint8_t)~n->mask[i];
        for (int i = 16; i < 32; i++) n->mask[i] = 0;
        n->mb = !n->mb;
    }
    return n;
}

static node* parse_atom(parser* ps) {
    if (ps->p >= ps->end) return new_node(N_EMPTY);
    int c = *ps->p;
    if (c == '(') {
        ps->p++;
        if (ps->p < ps->end && *ps->p == '?') {
            if (ps->p + 1 < ps->end && ps->p[1] == ':') ps->p += 2;
            else { fail(ps, "only (?:...) groups are supported (no flags, look-around or named groups)"); return new_node(N_EMPTY); }
        }
        if (++ps->depth > 200) { fail(ps, "nesting too deep"); return new_node(N_EMPTY); }
        node* n = parse_alt(ps);
        ps->depth--;
        if (ps->p >= ps->end || *ps->p != ')') { fail(ps, "missing )"); return n; }
        ps->p++;
        return n;
    }
    if (c == '[') { ps->p++; return parse_class(ps); }
    if (c == '.') {
        ps->p++;
        node* n = new_node(N_SET);
        set_range(n->mask, 0, 127);
        n->mask['\n' >> 3] &= (uint8_t)~(1u << ('\n' & 7));
        n->mb = 1;
        return n;
    }
    if (c == '^') { ps->p++; return new_node(N_BOL); }
    if (c == '$') { ps->p++; return new_node(N_EOL); }
    if (c == '*' || c == '+' || c == '?') { fail(ps, "missing argument to repetition operator"); return new_node(N_EMPTY); }
    if (c == '{') { /* RE2 treats a '{' that does not start a valid repeat as a literal */ }
    if (c == '\\') {
        ps->p++;
        node* n = new_node(N_SET);
        int mb = 0;
        int v = parse_escape(ps, n->mask, &mb);
        if (v == -1) return n;
        if (v == -2) { n->mb = mb; return n; }
        if (v >= '1' && v <= '9' && 0) {}
        set_bit(n->mask, v);
        return n;
    }
    ps->p++;
    if (c < 0x80) {
        node* n = new_node(N_SET);
        set_bit(n->mask, c);
        return n;
    }
    /* a non-ASCII literal: the whole UTF-8 sequence is ONE atom (so that é+ repeats é) */
    int len = (c >= 0xF0) ? 4 : (c >= 0xE0) ? 3 : (c >= 0xC2) ? 2 : 0;
    if (len == 0 || ps->p + (len - 1) > ps->end) { fail(ps, "invalid UTF-8 in pattern"); return new_node(N_EMPTY); }
    node* seq = new_node(N_SET);
    set_bit(seq->mask, c);
    for (int i = 1; i < len; i++) {
        int t = *ps->p++;
        if ((t & 0xC0) != 0x80) { fail(ps, "invalid UTF-8 in pattern"); return seq; }
        node* b = new_node(N_SET);
        set_bit(b->mask, t);
        node* cat = new_node(N_CAT);
        cat->a = seq; cat->b = b;
        seq = cat;
    }
    return seq;
}

/* {m}, {m,}, {m,n}; returns 0 when the brace is not a valid repeat (literal '{') */
static int parse_braces(parser* ps, int* mn, int* mx) {
    const uint8_t* q = ps->p + 1;
    int a = 0, b = -1, nd = 0;
    while (q < ps->end && *q >= '0' && *q <= '9') { a = a * 10 + (*q - '0'); q++; if (++nd > 4) return 0; }
    if (nd == 0) return 0;
    if (q < ps->end && *q == '}') { *mn = a; *mx = a; ps->p = q + 1; return 1; }
    if (q >= ps->end || *q != ',') return 0;
    q++;
    nd = 0;
    if (q < ps->end && *q == '}') { *mn = a; *mx = -1; ps->p = q + 1; return 1; }
    b = 0;
    while (q < ps->end && *q >= '0' && *q <= '9') { b = b * 10 + (*q - '0'); q++; if (++nd > 4) return 0; }
    if (nd == 0 || q >= ps->end || *q != '}') return 0;
    *mn = a; *mx = b; ps->p = q + 1;
    return 1;
}

static node* parse_repeat(parser* ps) {
    node* n = parse_atom(ps);
    while (!ps->failed && ps->p < ps->end) {
        int c = *ps->p;
        node* r = NULL;
        if (c == '*') { ps->p++; r = new_node(N_STAR); }
        else if (c == '+') { ps->p++; r = new_node(N_PLUS); }
        else if (c == '?') { ps->p++; r = new_node(N_QUEST); }
        else if (c == '{') {
            int mn, mx;
            if (!parse_braces(ps, &mn, &mx)) break;
            if (mn > 1000 || mx > 1000) { fail(ps, "bad repetition operator: repeat count above 1000"); break; }
            if (mx >= 0 && mx < mn) { fail(ps, "bad repetition operator: min > max"); break; }
            r = new_node(N_REPEAT);
            r->min = mn; r->max = mx;
        } else break;
        if (n->kind == N_BOL || n->kind == N_EOL || n->kind == N_EMPTY) {
            /* RE2 accepts ^* but it is useless; keep the grammar strict */
            free_node(r);
            fail(ps, "missing argument to repetition operator");
            break;
        }
        r->a = n;
        n = r;
        if (ps->p < ps->end && *ps->p == '?') ps->p++; /* lazy form: same match set */
        else if (ps->p < ps->end && (*ps->p == '*' || *ps->p == '+')) { fail(ps, "bad repetition operator"); break; }
    }
    return n;
}

static node* parse_cat(parser* ps) {
    node* n = NULL;
    while (!ps->failed && ps->p < ps->end && *ps->p != '|' && *ps->p != ')') {
        node* r = parse_repeat(ps);
        if (!n) n = r;
        else { node* c = new_node(N_CAT); c->a = n; c->b = r; n = c; }
    }
    return n ? n : new_node(N_EMPTY);
}

static node* parse_alt(parser* ps) {
    node* n = parse_cat(ps);
    while (!ps->failed && ps->p < ps->end && *ps->p == '|') {
        ps->p++;
        node* r = parse_cat(ps);
        node* a = new_node(N_ALT);
        a->a = n; a->b = r;
        n = a;
    }
    return n;
}

/* ── program ─────────────────────────────────────────────────────────────────────────── */

enum { I_SET, I_SPLIT, I_JMP, I_BOL, I_EOL, I_MATCH };
typedef struct { int op; int x, y; uint8_t mask[32]; int mb; } inst;
typedef struct { inst* code; int n, cap; int failed; } prog;

static int emit(prog* pr, int op) {
    if (pr->n >= 200000) { pr->failed = 1; return 0; }
    if (pr->n == pr->cap) {
        pr->cap = pr->cap ? pr->cap * 2 : 64;
        pr->code = (inst*)realloc(pr->code, sizeof(inst) * (size_t)pr->cap);
    }
    memset(&pr->code[pr->n], 0, sizeof(inst));
    pr->code[pr->n].op = op;
    return pr->n++;
}

static void compile(prog* pr, const node* n) {
    if (pr->failed) return;
    switch (n->kind) {
        case N_EMPTY: break;
        case N_SET: { int i = emit(pr, I_SET); memcpy(pr->code[i].mask, n->mask, 32); pr->code[i].mb = n->mb; break; }
        case N_BOL: emit(pr, I_BOL); break;
        case N_EOL: emit(pr, I_EOL); break;
        case N_CAT: compile(pr, n->a); compile(pr, n->b); break;
        case N_ALT: {
            int s = emit(pr, I_SPLIT);
            pr->code[s].x = pr->n;
            compile(pr, n->a);
            int j = emit(pr, I_JMP);
            pr->code[s].y = pr->n;
            compile(pr, n->b);
            pr->code[j].x = pr->n;
            break;
        }
        case N_QUEST: {
            int s = emit(pr, I_SPLIT);
            pr->code[s].x = pr->n;
            compile(pr, n->a);
            pr->code[s].y = pr->n;
            break;
        }
        case N_STAR: {
            int s = emit(pr, I_SPLIT);
            pr->code[s].x = pr->n;
            compile(pr, n->a);
            int j = emit(pr, I_JMP);
            pr->code[j].x = s;
            pr->code[s].y = pr->n;
            break;
        }
        case N_PLUS: {
            int start = pr->n;
            compile(pr, n->a);
            int s = emit(pr, I_SPLIT);
            pr->code[s].x = start;
            pr->code[s].y = pr->n;
            break;
        }
        case N_REPEAT: {
            for (int i = 0; i < n->min; i++) compile(pr, n->a);
            if (n->max < 0) {
                node star; memset(&star, 0, sizeof(star));
                star.kind = N_STAR; star.a = n->a;
                compile(pr, &star);
            } else {
                for (int i = n->min; i < n->max; i++) {
                    node q; memset(&q, 0, sizeof(q));
                    q.kind = N_QUEST; q.a = n->a;
                    /* nested optionals: (a(a(a)?)?)? is equivalent for matching to a?a?a? */
                    compile(pr, &q);
                }
            }
            break;
        }
    }
}

/* length of the well-formed UTF-8 multi-byte sequence at t[pos], 0 if none */
static int utf8_mb_len(const uint8_t* t, int64_t pos, int64_t len) {
    uint8_t c = t[pos];
    int n;
    uint8_t lo = 0x80, hi = 0xBF;
    if (c >= 0xC2 && c <= 0xDF) n = 2;
    else if (c >= 0xE0 && c <= 0xEF) { n = 3; if (c == 0xE0) lo = 0xA0; if (c == 0xED) hi = 0x9F; }
    else if (c >= 0xF0 && c <= 0xF4) { n = 4; if (c == 0xF0) lo = 0x90; if (c == 0xF4) hi = 0x8F; }
    else return 0;
    if (pos + n > len) return 0;
    if (t[pos + 1] < lo || t[pos + 1] > hi) return 0;
    for (int i = 2; i < n; i++) if ((t[pos + i] & 0xC0) != 0x80) return 0;
    return n;
}

typedef struct {
    const prog* pr;
    const uint8_t* t;
    int64_t len;
    uint8_t* seen; /* (pc, pos) bitmap */
} vm;

static int run(vm* m, int pc, int64_t pos) {
    for (;;) {
        size_t key = (size_t)pc * (size_t)(m->len + 1) + (size_t)pos;
        if (m->seen[key >> 3] & (1u << (key & 7))) return 0;
        m->seen[key >> 3] |= (uint8_t)(1u << (key & 7));
        const inst* in = &m->pr->code[pc];
        switch (in->op) {
            case I_MATCH: return 1;
            case I_BOL: if (pos != 0) return 0; pc++; break;
            case I_EOL: if (pos != m->len) return 0; pc++; break;
            case I_JMP: pc = in->x; break;
            case I_SPLIT:
                if (run(m, in->x, pos)) return 1;
                pc = in->y;
                break;
            case I_SET: {
                if (pos >= m->len) return 0;
                uint8_t c = m->t[pos];
                if (c < 0x80) { if (!get_bit(in->mask, c)) return 0; pos++; }
                else if (get_bit(in->mask, c)) pos++; /* explicit raw byte from a UTF-8 literal */
                else if (in->mb) { int n = utf8_mb_len(m->t, pos, m->len); if (!n) return 0; pos += n; }
                else return 0;
                pc++;
                break;
            }
        }
    }
}

typedef struct orc_regex { prog pr; } orc_regex;

static int rx_compile(const char* pattern, prog* pr) {
    parser ps = {(const uint8_t*)pattern, (const uint8_t*)pattern + strlen(pattern), 0, 0};
    rx_err[0] = 0;
    node* ast = parse_alt(&ps);
    if (!ps.failed && ps.p < ps.end) fail(&ps, *ps.p == ')' ? "unexpected )" : "trailing characters");
    if (ps.failed) { free_node(ast); return -1; }
    memset(pr, 0, sizeof(*pr));
    compile(pr, ast);
    free_node(ast);
    if (pr->failed) { free(pr->code); snprintf(rx_err, sizeof(rx_err), "regex: program too large"); return -1; }
    emit(pr, I_MATCH);
    return 0;
}

static int rx_search(const prog* pr, const uint8_t* text, int64_t len) {
    size_t bits = (size_t)pr->n * (size_t)(len + 1);
    vm m = {pr, text, len, (uint8_t*)calloc(bits / 8 + 1, 1)};
    int hit = 0;
    /* the memo can be shared across start positions: a (pc, pos) that failed once fails always */
    for (int64_t s = 0; s <= len && !hit; s++) hit = run(&m, 0, s);
    free(m.seen);
    return hit;
}

extern const char* orc_last_error(void);
/* pq_oracle.c owns the error slot; regex errors are reported through this accessor */
const char* orc_regex_last_error(void) { return rx_err; }

int orc_regex_search(const char* pattern, const uint8_t* text, int64_t len) {
    prog pr;
    if (rx_compile(pattern, &pr) != 0) return -1;
    int hit = rx_search(&pr, text, len);
    free(pr.code);
    return hit;
}

int64_t orc_regex_prune(orc_file* f, int col, const char* pattern, int neg,
                        uint8_t* bits, int64_t cap) {
    prog pr;
    if (rx_compile(pattern, &pr) != 0) return -1;
    int64_t np = 0;
    int64_t nrg = orc_num_row_groups(f);
    for (int64_t rg = 0; rg < nrg; rg++) {
        pagedump pd;
        if (orc_read_pages(f, (int)rg, col, &pd) != 0) {
            snprintf(rx_err, sizeof(rx_err), "%s", orc_last_error());
            free(pr.code);
            return -1;
        }
        for (int64_t p = 0; p < pd.n_pages; p++) {
            if (pd.page_type[p] != 0) continue;
            uint8_t bit = 0;
            for (int64_t i = pd.first_value[p]; i < pd.first_value[p + 1] && !bit; i++) {
                if (pd.values.is_null[i] || pd.values.vidx[i] != 5) continue;
                int m = rx_search(&pr, pd.values.chars + pd.values.str_off[i],
                                  (int64_t)(pd.values.str_off[i + 1] - pd.values.str_off[i]));
                if (neg ? !m : m) bit = 1;
            }
            if (np < cap) bits[np] = bit;
            np++;
        }
        orc_pagedump_free(&pd);
    }
    free(pr.code);
    return np;
}
