// pq_regex.cpp -- host compiler: RE2-syntax subset -> byte-class DFA for the GPU page scan
// (pqg_regex_compile / pqg_dfa_* in include/pqg.h).
//
// Replaces the re2 dependency of the reference's `parser --regex-column` mode
// (README.md:7-29,54-64; no source in the checkout).  Frozen semantics (SURVEY.md section 8,
// a-19): partial match (search); ^ and $ are honoured; `.` and negated classes match one
// UTF-8 encoded code point (never '\n' for `.`); classes are ASCII; anything outside the
// subset -- back-references, look-around, \b, flags, POSIX classes, non-ASCII inside
// classes -- is rejected with an explicit message instead of being silently mis-compiled.
//
// Pipeline: recursive-descent parser -> Thompson NFA over bytes (multi-byte UTF-8 expanded
// into byte sequences) -> subset construction over byte equivalence classes -> Moore
// minimisation -> flat tables.  Search semantics live in the automaton: the start state
// re-seeds the pattern at every position (unless every alternative is anchored by ^), a
// match that does not depend on $ is absorbing, and acceptance is read after the last byte
// with $-edges enabled.
#include <algorithm>
#include <array>
#include <bitset>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "pq_regex.hpp"
#include "pqg.h"

namespace pqg {
namespace {

using ByteSet = std::bitset<256>;

struct RegexError : std::runtime_error { using std::runtime_error::runtime_error; };

// ── AST ──────────────────────────────────────────────────────────────────────────────────
struct Node {
    enum Kind { SET, CAT, ALT, STAR, PLUS, QUEST, REPEAT, BOL, EOL, EMPTY } kind = EMPTY;
    ByteSet ascii;      // SET: members among bytes < 0x80, or exact raw bytes of a literal
    bool multibyte = false; // SET: also any well-formed multi-byte UTF-8 sequence
    int lo = 0, hi = 0; // REPEAT (hi < 0: unbounded)
    std::unique_ptr<Node> a, b;
};
using NodeP = std::unique_ptr<Node>;

NodeP mk(Node::Kind k) { auto n = std::make_unique<Node>(); n->kind = k; return n; }

class Parser {
public:
    explicit Parser(const std::string& p) : s_(p) {}
    NodeP parse() {
        NodeP n = alt();
        if (i_ < s_.size()) throw RegexError(s_[i_] == ')' ? "unexpected )" : "trailing characters");
        return n;
    }

private:
    const std::string& s_;
    size_t i_ = 0;
    int depth_ = 0;

    bool more() const { return i_ < s_.size(); }
    uint8_t peek() const { return static_cast<uint8_t>(s_[i_]); }

    static void perl_class(ByteSet& out, char c, bool& mb) {
        ByteSet t;
        switch (c | 0x20) {
            case 'd': for (int x = '0'; x <= '9'; x++) t.set(x); break;
            case 'w': for (int x = '0'; x <= '9'; x++) t.set(x);
                      for (int x = 'a'; x <= 'z'; x++) { t.set(x); t.set(x - 32); }
                      t.set('_'); break;
            case 's': t.set('\t'); t.set('\n'); t.set('\f'); t.set('\r'); t.set(' '); break;
        }
        if (c < 'a') { // upper case: complement within ASCII, plus every multi-byte code point
            for (int x = 0; x < 128; x++) t.flip(x);
            mb = true;
        }
        out |= t;
    }
    static int hex(int c) {
        if (c >= '0' && c <= '9') return c - '0';
        c |= 0x20;
        return (c >= 'a' && c <= 'f') ? c - 'a' + 10 : -1;
    }
    // after the backslash; returns a byte, or -1 when a class was merged into `cls`
    int escape(ByteSet& cls, bool& mb) {
        if (!more()) throw RegexError("trailing backslash");
        uint8_t c = peek();
        i_++;
        switch (c) {
            case 'd': case 'w': case 's': case 'D': case 'W': case 'S': perl_class(cls, static_cast<char>(c), mb); return -1;
            case 'n': return '\n';
            case 't': return '\t';
            case 'r': return '\r';
            case 'f': return '\f';
            case 'v': return '\v';
            case 'a': return 7;
            case 'x': {
                if (i_ + 1 < s_.size() + 0 && hex(s_[i_]) >= 0 && i_ + 1 < s_.size() && hex(s_[i_ + 1]) >= 0) {
                    int v = hex(s_[i_]) * 16 + hex(s_[i_ + 1]);
                    i_ += 2;
                    if (v >= 0x80) throw RegexError("\\x escape above 0x7f is not supported");
                    return v;
                }
                throw RegexError("bad \\x escape");
            }
            default:
                if ((c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'))
                    throw RegexError(std::string("unsupported escape \\") + static_cast<char>(c));
                if (c >= 0x80) throw RegexError("escaped non-ASCII byte");
                return c;
        }
    }

    NodeP char_class() {
        NodeP n = mk(Node::SET);
        bool negate = false;
        if (more() && peek() == '^') { negate = true; i_++; }
        bool first = true;
        for (;;) {
            if (!more()) throw RegexError("missing ]");
            uint8_t c = peek();
            if (c == ']' && !first) { i_++; break; }
            first = false;
            if (c == '[' && i_ + 1 < s_.size() && s_[i_ + 1] == ':') throw RegexError("POSIX classes are not supported");
            if (c >= 0x80) throw RegexError("non-ASCII in character class is not supported");
            i_++;
            int lo = c;
            if (c == '\\') {
                bool mb = false;
                lo = escape(n->ascii, mb);
                if (lo < 0) { if (mb) n->multibyte = true; continue; }
            }
            if (i_ + 1 < s_.size() && s_[i_] == '-' && s_[i_ + 1] != ']') {
                i_++;
                int hi = peek();
                i_++;
                if (hi >= 0x80) throw RegexError("non-ASCII in character class is not supported");
                if (hi == '\\') {
                    ByteSet tmp; bool mb = false;
                    hi = escape(tmp, mb);
                    if (hi < 0) throw RegexError("bad range end");
                }
                if (hi < lo) throw RegexError("bad character range");
                for (int x = lo; x <= hi; x++) n->ascii.set(x);
            } else n->ascii.set(lo);
        }
        if (negate) {
            for (int x = 0; x < 128; x++) n->ascii.flip(x);
            for (int x = 128; x < 256; x++) n->ascii.reset(x);
            n->multibyte = !n->multibyte;
        }
        return n;
    }

    NodeP atom() {
        if (!more()) return mk(Node::EMPTY);
        uint8_t c = peek();
        if (c == '(') {
            i_++;
            if (more() && peek() == '?') {
                if (i_ + 1 < s_.size() && s_[i_ + 1] == ':') i_ += 2;
                else throw RegexError("only (?:...) groups are supported (no flags, look-around or named groups)");
            }
            if (++depth_ > 200) throw RegexError("nesting too deep");
            NodeP n = alt();
            depth_--;
            if (!more() || peek() != ')') throw RegexError("missing )");
            i_++;
            return n;
        }
        if (c == '[') { i_++; return char_class(); }
        if (c == '.') {
            i_++;
            NodeP n = mk(Node::SET);
            for (int x = 0; x < 128; x++) if (x != '\n') n->ascii.set(x);
            n->multibyte = true;
            return n;
        }
        if (c == '^') { i_++; return mk(Node::BOL); }
        if (c == '$') { i_++; return mk(Node::EOL); }
        if (c == '*' || c == '+' || c == '?') throw RegexError("missing argument to repetition operator");
        if (c == '\\') {
            i_++;
            NodeP n = mk(Node::SET);
            bool mb = false;
            int v = escape(n->ascii, mb);
            if (v < 0) n->multibyte = mb; else n->ascii.set(v);
            return n;
        }
        i_++;
        if (c < 0x80) { NodeP n = mk(Node::SET); n->ascii.set(c); return n; }
        // non-ASCII literal: the whole UTF-8 sequence is one atom
        int len = c >= 0xF0 ? 4 : c >= 0xE0 ? 3 : c >= 0xC2 ? 2 : 0;
        if (len == 0 || i_ + (len - 1) > s_.size()) throw RegexError("invalid UTF-8 in pattern");
        NodeP seq = mk(Node::SET);
        seq->ascii.set(c);
        for (int k = 1; k < len; k++) {
            uint8_t t = peek();
            i_++;
            if ((t & 0xC0) != 0x80) throw RegexError("invalid UTF-8 in pattern");
            NodeP b = mk(Node::SET);
            b->ascii.set(t);
            NodeP cat = mk(Node::CAT);
            cat->a = std::move(seq); cat->b = std::move(b);
            seq = std::move(cat);
        }
        return seq;
    }

    bool braces(int& lo, int& hi) { // {m} {m,} {m,n}; a '{' that is not a repeat is a literal
        size_t q = i_ + 1;
        auto num = [&](int& out) {
            int v = 0, nd = 0;
            while (q < s_.size() && s_[q] >= '0' && s_[q] <= '9') { v = v * 10 + (s_[q] - '0'); q++; if (++nd > 4) return false; }
            out = v;
            return nd > 0;
        };
        if (!num(lo)) return false;
        if (q < s_.size() && s_[q] == '}') { hi = lo; i_ = q + 1; return true; }
        if (q >= s_.size() || s_[q] != ',') return false;
        q++;
        if (q < s_.size() && s_[q] == '}') { hi = -1; i_ = q + 1; return true; }
        if (!num(hi) || q >= s_.size() || s_[q] != '}') return false;
        i_ = q + 1;
        return true;
    }

    NodeP repeat() {
        NodeP n = atom();
        while (more()) {
            uint8_t c = peek();
            NodeP r;
            if (c == '*') { i_++; r = mk(Node::STAR); }
            else if (c == '+') { i_++; r = mk(Node::PLUS); }
            else if (c == '?') { i_++; r = mk(Node::QUEST); }
            else if (c == '{') {
                int lo, hi;
                if (!braces(lo, hi)) break;
                if (lo > 1000 || hi > 1000) throw RegexError("bad repetition operator: repeat count above 1000");
                if (hi >= 0 && hi < lo) throw RegexError("bad repetition operator: min > max");
                r = mk(Node::REPEAT);
                r->lo = lo; r->hi = hi;
            } else break;
            if (n->kind == Node::BOL || n->kind == Node::EOL || n->kind == Node::EMPTY)
                throw RegexError("missing argument to repetition operator");
            r->a = std::move(n);
            n = std::move(r);
            if (more() && peek() == '?') i_++; // lazy form: same match set
            else if (more() && (peek() == '*' || peek() == '+')) throw RegexError("bad repetition operator");
        }
        return n;
    }

    NodeP cat() {
        NodeP n;
        while (more() && peek() != '|' && peek() != ')') {
            NodeP r = repeat();
            if (!n) n = std::move(r);
            else { NodeP c = mk(Node::CAT); c->a = std::move(n); c->b = std::move(r); n = std::move(c); }
        }
        return n ? std::move(n) : mk(Node::EMPTY);
    }

    NodeP alt() {
        NodeP n = cat();
        while (more() && peek() == '|') {
            i_++;
            NodeP r = cat();
            NodeP a = mk(Node::ALT);
            a->a = std::move(n); a->b = std::move(r);
            n = std::move(a);
        }
        return n;
    }
};

// ── NFA ──────────────────────────────────────────────────────────────────────────────────
struct NState {
    enum Kind { BYTES, EPS, SPLIT, BOL, EOL, MATCH } kind = EPS;
    ByteSet set;
    int out = -1, out2 = -1;
};

struct Nfa {
    std::vector<NState> st;
    int add(NState::Kind k) {
        if (st.size() > 60000) throw RegexError("pattern too large (NFA state limit)");
        st.emplace_back();
        st.back().kind = k;
        return static_cast<int>(st.size()) - 1;
    }
};

// A fragment: entry state + dangling exits.  An exit is (state << 1 | which): which = 0 is
// NState::out, 1 is NState::out2 (indices, not pointers: the state vector reallocates).
struct Frag { int start; std::vector<int> holes; };

class Builder {
public:
    Nfa nfa;
    Frag build(const Node* n) {
        switch (n->kind) {
            case Node::EMPTY: { int s = nfa.add(NState::EPS); return {s, {hole(s, 0)}}; }
            case Node::BOL: { int s = nfa.add(NState::BOL); return {s, {hole(s, 0)}}; }
            case Node::EOL: { int s = nfa.add(NState::EOL); return {s, {hole(s, 0)}}; }
            case Node::SET: return build_set(n);
            case Node::CAT: {
                Frag a = build(n->a.get());
                Frag b = build(n->b.get());
                patch(a, b.start);
                return {a.start, b.holes};
            }
            case Node::ALT: {
                Frag a = build(n->a.get());
                Frag b = build(n->b.get());
                int s = nfa.add(NState::SPLIT);
                nfa.st[s].out = a.start;
                nfa.st[s].out2 = b.start;
                a.holes.insert(a.holes.end(), b.holes.begin(), b.holes.end());
                return {s, a.holes};
            }
            case Node::QUEST: return quest(n->a.get());
            case Node::STAR: return star(n->a.get());
            case Node::PLUS: {
                Frag a = build(n->a.get());
                int s = nfa.add(NState::SPLIT);
                nfa.st[s].out = a.start;
                patch(a, s);
                return {a.start, {hole(s, 1)}};
            }
            case Node::REPEAT: {
                // x{m,n}: m copies, then a star (n unbounded) or n-m nested optionals
                int s0 = nfa.add(NState::EPS);
                Frag acc{s0, {hole(s0, 0)}};
                for (int i = 0; i < n->lo; i++) {
                    Frag c = build(n->a.get());
                    patch(acc, c.start);
                    acc.holes = c.holes;
                }
                if (n->hi < 0) {
                    Frag c = star(n->a.get());
                    patch(acc, c.start);
                    acc.holes = c.holes;
                } else {
                    std::vector<int> exits;
                    for (int i = n->lo; i < n->hi; i++) {
                        Frag c = build(n->a.get());
                        int s = nfa.add(NState::SPLIT);
                        nfa.st[s].out = c.start;
                        patch(acc, s);
                        exits.push_back(hole(s, 1));
                        acc.holes = c.holes;
                    }
                    acc.holes.insert(acc.holes.end(), exits.begin(), exits.end());
                }
                return acc;
            }
        }
        throw RegexError("internal: bad node");
    }
    void patch(Frag& f, int target) {
        for (int h : f.holes) {
            NState& st = nfa.st[static_cast<size_t>(h >> 1)];
            ((h & 1) ? st.out2 : st.out) = target;
        }
        f.holes.clear();
    }

private:
    static int hole(int s, int which) { return (s << 1) | which; }
    Frag quest(const Node* x) {
        Frag a = build(x);
        int s = nfa.add(NState::SPLIT);
        nfa.st[s].out = a.start;
        a.holes.push_back(hole(s, 1));
        return {s, a.holes};
    }
    Frag star(const Node* x) {
        Frag a = build(x);
        int s = nfa.add(NState::SPLIT);
        nfa.st[s].out = a.start;
        patch(a, s);
        return {s, {hole(s, 1)}};
    }
    int bytes_state(const ByteSet& set) { int s = nfa.add(NState::BYTES); nfa.st[s].set = set; return s; }
    static ByteSet range(int lo, int hi) { ByteSet b; for (int x = lo; x <= hi; x++) b.set(x); return b; }

    Frag build_set(const Node* n) {
        std::vector<int> firsts, lasts; // per alternative: first state, last state (its out is the exit)
        if (n->ascii.any()) { int s = bytes_state(n->ascii); firsts.push_back(s); lasts.push_back(s); }
        if (n->multibyte) {
            // well-formed UTF-8 multi-byte sequences (Unicode standard, table 3-7)
            struct Seq { int n; int lo[4], hi[4]; };
            static const Seq seqs[] = {
                {2, {0xC2, 0x80}, {0xDF, 0xBF}},
                {3, {0xE0, 0xA0, 0x80}, {0xE0, 0xBF, 0xBF}},
                {3, {0xE1, 0x80, 0x80}, {0xEC, 0xBF, 0xBF}},
                {3, {0xED, 0x80, 0x80}, {0xED, 0x9F, 0xBF}},
                {3, {0xEE, 0x80, 0x80}, {0xEF, 0xBF, 0xBF}},
                {4, {0xF0, 0x90, 0x80, 0x80}, {0xF0, 0xBF, 0xBF, 0xBF}},
                {4, {0xF1, 0x80, 0x80, 0x80}, {0xF3, 0xBF, 0xBF, 0xBF}},
                {4, {0xF4, 0x80, 0x80, 0x80}, {0xF4, 0x8F, 0xBF, 0xBF}},
            };
            for (const Seq& q : seqs) {
                int first = -1, prev = -1;
                for (int k = 0; k < q.n; k++) {
                    int s = bytes_state(range(q.lo[k], q.hi[k]));
                    if (prev >= 0) nfa.st[prev].out = s; else first = s;
                    prev = s;
                }
                firsts.push_back(first);
                lasts.push_back(prev);
            }
        }
        if (firsts.empty()) { // empty set: matches nothing
            int s = bytes_state(ByteSet());
            return {s, {hole(s, 0)}};
        }
        int start = firsts[0];
        for (size_t k = 1; k < firsts.size(); k++) {
            int sp = nfa.add(NState::SPLIT);
            nfa.st[sp].out = start;
            nfa.st[sp].out2 = firsts[k];
            start = sp;
        }
        std::vector<int> holes;
        for (int l : lasts) holes.push_back(hole(l, 0));
        return {start, holes};
    }
};

} // namespace

// ── DFA ──────────────────────────────────────────────────────────────────────────────────
struct DfaBuilder {
    const Nfa& nfa;
    int match_state;
    explicit DfaBuilder(const Nfa& n, int m) : nfa(n), match_state(m) {}

    // epsilon closure; bol/eol say whether those assertion edges may be followed
    void closure(std::vector<int>& set, bool bol, bool eol) const {
        std::vector<char> seen(nfa.st.size(), 0);
        std::vector<int> stack;
        for (int s : set) if (!seen[s]) { seen[s] = 1; stack.push_back(s); }
        set.clear();
        while (!stack.empty()) {
            int s = stack.back();
            stack.pop_back();
            set.push_back(s);
            const NState& st = nfa.st[s];
            auto push = [&](int t) { if (t >= 0 && !seen[t]) { seen[t] = 1; stack.push_back(t); } };
            switch (st.kind) {
                case NState::EPS: push(st.out); break;
                case NState::SPLIT: push(st.out); push(st.out2); break;
                case NState::BOL: if (bol) push(st.out); break;
                case NState::EOL: if (eol) push(st.out); break;
                default: break;
            }
        }
        std::sort(set.begin(), set.end());
    }
    bool has_match(const std::vector<int>& set) const { return std::binary_search(set.begin(), set.end(), match_state); }
};

CompiledDfa compile_regex(const std::string& pattern, uint32_t max_states) {
    Parser parser(pattern);
    NodeP ast = parser.parse();
    Builder b;
    Frag f = b.build(ast.get());
    int m = b.nfa.add(NState::MATCH);
    b.patch(f, m);
    const Nfa& nfa = b.nfa;
    // byte equivalence classes over every BYTES set
    std::vector<const ByteSet*> sets;
    for (const NState& s : nfa.st) if (s.kind == NState::BYTES) sets.push_back(&s.set);
    CompiledDfa out;
    {
        std::map<std::vector<bool>, int> sig;
        for (int c = 0; c < 256; c++) {
            std::vector<bool> v(sets.size());
            for (size_t k = 0; k < sets.size(); k++) v[k] = sets[k]->test(c);
            auto it = sig.find(v);
            if (it == sig.end()) it = sig.emplace(std::move(v), static_cast<int>(sig.size())).first;
            out.cls[c] = static_cast<uint8_t>(it->second);
        }
        out.n_classes = static_cast<uint32_t>(sig.size());
    }
    std::vector<int> rep(out.n_classes, -1); // a representative byte per class
    for (int c = 0; c < 256; c++) if (rep[out.cls[c]] < 0) rep[out.cls[c]] = c;

    DfaBuilder db(nfa, m);
    // DFA states: sorted NFA-state sets (closures without $-edges).  State 0 = absorbing
    // ACCEPT (a match that no longer depends on the rest of the input).
    std::map<std::vector<int>, int> ids;
    std::vector<std::vector<int>> dsets;
    std::vector<uint8_t> accept_end;
    std::vector<std::vector<int>> trans;
    auto add_state = [&](std::vector<int> set, bool initial) -> int {
        if (db.has_match(set)) return 0;
        auto it = ids.find(set);
        if (it != ids.end() && !initial) return it->second;
        if (dsets.size() >= max_states) throw RegexError("pattern too large: DFA exceeds " + std::to_string(max_states) + " states");
        int id = static_cast<int>(dsets.size());
        std::vector<int> endset = set;
        db.closure(endset, initial, true);
        accept_end.push_back(db.has_match(endset));
        if (!initial) ids.emplace(set, id);
        dsets.push_back(std::move(set));
        trans.emplace_back(out.n_classes, -1);
        return id;
    };
    dsets.push_back({}); accept_end.push_back(1); trans.emplace_back(out.n_classes, 0); // state 0: ACCEPT
    std::vector<int> seed{f.start};
    // the pattern is re-seeded at every later position (search semantics): closure of the
    // pattern start with ^-edges disabled
    std::vector<int> reseed{f.start};
    db.closure(reseed, false, false);
    std::vector<int> init = seed;
    db.closure(init, true, false);
    int start;
    if (db.has_match(init)) start = 0;
    else start = add_state(init, true);
    out.start = static_cast<uint32_t>(start);
    for (size_t cur = 1; cur < dsets.size(); cur++) {
        for (uint32_t c = 0; c < out.n_classes; c++) {
            std::vector<int> next;
            const std::vector<int> from = dsets[cur]; // copy: dsets may grow
            for (int s : from) {
                const NState& st = nfa.st[s];
                if (st.kind == NState::BYTES && st.set.test(rep[c]) && st.out >= 0) next.push_back(st.out);
            }
            next.insert(next.end(), reseed.begin(), reseed.end());
            db.closure(next, false, false);
            trans[cur][c] = add_state(std::move(next), false);
        }
    }
    // Moore minimisation
    size_t n = dsets.size();
    std::vector<int> part(n);
    for (size_t s = 0; s < n; s++) part[s] = accept_end[s] ? 1 : 0;
    part[0] = 2; // keep ACCEPT distinguishable only by behaviour: it loops to itself
    for (;;) {
        std::map<std::vector<int>, int> sigs;
        std::vector<int> np(n);
        for (size_t s = 0; s < n; s++) {
            std::vector<int> sg;
            sg.push_back(part[s]);
            for (uint32_t c = 0; c < out.n_classes; c++) sg.push_back(part[trans[s][c]]);
            auto it = sigs.find(sg);
            if (it == sigs.end()) it = sigs.emplace(std::move(sg), static_cast<int>(sigs.size())).first;
            np[s] = it->second;
        }
        bool same = true;
        {
            // partitions only ever split: equal count means stable
            int a = *std::max_element(part.begin(), part.end()), bmax = *std::max_element(np.begin(), np.end());
            std::map<int, int> distinct;
            for (int x : part) distinct[x] = 1;
            same = distinct.size() == static_cast<size_t>(bmax + 1);
            (void)a;
        }
        part = np;
        if (same) break;
    }
    // renumber: ACCEPT's block -> 0
    int nblocks = *std::max_element(part.begin(), part.end()) + 1;
    std::vector<int> remap(nblocks, -1);
    remap[part[0]] = 0;
    int nextid = 1;
    for (size_t s = 0; s < n; s++) if (remap[part[s]] < 0) remap[part[s]] = nextid++;
    out.n_states = static_cast<uint32_t>(nblocks);
    out.trans.assign(static_cast<size_t>(nblocks) * out.n_classes, 0);
    out.accept.assign(nblocks, 0);
    for (size_t s = 0; s < n; s++) {
        int id = remap[part[s]];
        out.accept[id] = accept_end[s];
        for (uint32_t c = 0; c < out.n_classes; c++) out.trans[static_cast<size_t>(id) * out.n_classes + c] = static_cast<uint16_t>(remap[part[trans[s][c]]]);
    }
    out.start = static_cast<uint32_t>(remap[part[start]]);
    // dead state: non-accepting and absorbing (anchored patterns): lets the scan stop early
    out.dead = UINT32_MAX;
    for (uint32_t s = 0; s < out.n_states; s++) {
        if (out.accept[s]) continue;
        bool absorbing = true;
        for (uint32_t c = 0; c < out.n_classes; c++) if (out.trans[static_cast<size_t>(s) * out.n_classes + c] != s) { absorbing = false; break; }
        if (absorbing) { out.dead = s; break; }
    }
    return out;
}

bool dfa_match(const CompiledDfa& d, const uint8_t* text, uint64_t len) {
    uint32_t s = d.start;
    for (uint64_t i = 0; i < len && s != 0; i++) s = d.trans[static_cast<size_t>(s) * d.n_classes + d.cls[text[i]]];
    return d.accept[s] != 0;
}

} // namespace pqg

struct pqg_dfa { pqg::CompiledDfa d; };

extern "C" {

int pqg_regex_compile(const char* pattern, pqg_dfa** out, char* err, size_t errlen) {
    if (out) *out = nullptr;
    if (!pattern || !out) {
        if (err && errlen) std::snprintf(err, errlen, "pqg_regex_compile: bad argument");
        return PQG_ERR_ARG;
    }
    try {
        auto* d = new pqg_dfa();
        try { d->d = pqg::compile_regex(pattern, 4096); } catch (...) { delete d; throw; }
        *out = d;
        return PQG_OK;
    } catch (const std::exception& e) {
        if (err && errlen) std::snprintf(err, errlen, "regex: %s", e.what());
        return PQG_ERR_REGEX;
    }
}

void pqg_dfa_free(pqg_dfa* dfa) { delete dfa; }
uint32_t pqg_dfa_num_states(const pqg_dfa* dfa) { return dfa ? dfa->d.n_states : 0; }
int pqg_dfa_match_host(const pqg_dfa* dfa, const uint8_t* text, uint64_t len) {
    if (!dfa) return -1;
    return pqg::dfa_match(dfa->d, text, len) ? 1 : 0;
}

} // extern "C"

namespace pqg {
const CompiledDfa& dfa_tables(const pqg_dfa* d) { return d->d; }
}
