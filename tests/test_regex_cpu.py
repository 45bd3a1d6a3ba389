"""CPU tests of the host regex -> DFA compiler (product, host/pq_regex.cpp) against two
checkers: RE2 itself (pyarrow.compute.match_substring_regex -- the library the reference's regex
mode links, README.md:7-29,54-64; its call sites are absent from the tree, so RE2's PartialMatch on
the column values IS the pin for a-19), the oracle's backtracking matcher (oracle/regex_oracle.c,
pinned to RE2 by the same tests) and Python's `re` on the common subset."""
import re

import numpy as np
import pytest

PATTERNS = [
    r"^[a-z0-9._]+@[a-z0-9.]+\.com$",  # BASELINE config 4
    r"@", r"^user", r"com$", r"^$", r"", r"a", r"abc", r"a|b", r"^a|b$", r"(ab)+", r"(?:ab|cd)*x",
    r"a.c", r"a.*c", r"^.*$", r"[^a]", r"[^a-z]+$", r"\d+", r"^\d{3}-\d{4}$", r"\w+\s\w+", r"\S+@\S+",
    r"[a-c]{2,3}z", r"x{0}y", r"x{2,}", r"(a|b|c){3}", r"a?b?c?d", r"\.", r"\\", r"[.]", r"[]a]", r"[a\]]",
    r"[\d\s]", r"[^\d]", r"\D\W", r"colou?r", r"^(?:[0-9]{1,3}\.){3}[0-9]{1,3}$", r"é", r"é+x", r".é.",
    r"[^x]é", r"日本", r"^.{3}$", r"a\x41b", r"\tq", r"(a*)*b", r"(a+)+$", r"(|a)b", r"a||b", r"^^a", r"a$$",
    r"$a", r"a^", r"(^a|b)c", r"a($|b)", r"x*", r"[a-z]+[0-9]+$", r"user[0-9]+@mail[0-9]+\.example\.com!!",
    r"a{,2}", r"a{", r"a{x}", r"}", r"]",
]
REJECTED = [r"(", r")", r"a)", r"[a", r"*a", r"a**", r"\1", r"\b", r"(?i)a", r"(?=a)", r"(?P<n>a)", r"[[:alpha:]]",
            r"[é]", r"a{1001}", r"a{3,2}", "\\", r"\pL", r"\Qa\E", r"[z-a]", r"\xff", r"+"]
TEXTS = [b"", b"a", b"b", b"ab", b"abc", b"abab", b"cdcdx", b"a\nc", b"axxc", b"user123@mail7.example.com",
         b"user123.mail7.example.com", b"user123@mail7.example.com!!", b"USER@x.com", b"x@y.com\n", b"\n", b"\n\n",
         b"123-4567", b"123-45678", b"hello world", b"  ", b"aaz", b"abcz", b"abcaz", b"y", b"xxy", b"xx", b"xxx",
         b"abcabc", b"d", b"abcd", b".", b"\\", b"]", b"a]", b"7", b" 7", b"q!", b"color", b"colour", b"colouur",
         b"192.168.1.1", b"192.168.1", b"1234.1.1.1", "é".encode(), "ééx".encode(), "aéb".encode(),
         "xé".encode(), "yé".encode(), "日本語".encode(), "日".encode(), b"\xff", b"\xc3", b"a\xffb", "aé".encode(),
         b"aAb", b"\tq", b"aaaaaaaaaaaaaaaaaaaaaaab", b"aaaaaaaaaaaaaaaaaaaaaaa!", b"b", b"ac", b"bc", b"ab9", b"zz99x",
         b"a{", b"a{x}", b"}", b"aa", b"{,2}"]


def py_translate(p):
    """the subset in Python-re spelling: $ -> \\Z, \\s without \\v, ASCII classes, bytes"""
    out, i, in_cls = [], 0, False
    while i < len(p):
        c = p[i]
        if c == "\\" and i + 1 < len(p):
            n = p[i + 1]
            rep = {"s": r" \t\n\f\r", "S": None, "d": "0-9", "D": None, "w": "0-9A-Za-z_", "W": None}.get(n, False)
            if rep is False:
                out.append(c + n)
            elif rep is None:
                inner = {"S": r" \t\n\f\r", "D": "0-9", "W": "0-9A-Za-z_"}[n]
                if in_cls:
                    return None  # negated perl class inside a class: skip the python cross-check
                out.append("(?:[^" + inner + r"\x80-\xff]|" + UTF8_MB + ")")
            else:
                out.append(rep if in_cls else "[" + rep + "]")
            i += 2
            continue
        if in_cls:
            if c == "]" and not cls_first:
                in_cls = False
            cls_first = False
            out.append(c)
        elif c == "[":
            in_cls, cls_first = True, True
            if i + 1 < len(p) and p[i + 1] == "^":
                return None  # negated classes match a code point: python-bytes differs, oracle covers it
            out.append(c)
        elif c == "$":
            out.append(r"\Z")
        elif c == ".":
            out.append("(?:[^\\n\\x80-\\xff]|" + UTF8_MB + ")")
        else:
            out.append(c)
        i += 1
    return "".join(out)


UTF8_MB = (r"[\xc2-\xdf][\x80-\xbf]|\xe0[\xa0-\xbf][\x80-\xbf]|[\xe1-\xec\xee\xef][\x80-\xbf]{2}|\xed[\x80-\x9f][\x80-\xbf]"
           r"|\xf0[\x90-\xbf][\x80-\xbf]{2}|[\xf1-\xf3][\x80-\xbf]{3}|\xf4[\x80-\x8f][\x80-\xbf]{2}")


@pytest.mark.parametrize("pattern", PATTERNS)
def test_dfa_matches_oracle_and_python(pq, oracle, pattern):
    dfa = pq.regex_compile(pattern)
    pyp = py_translate(pattern)
    pyre = None
    if pyp is not None and not any(ord(ch) > 127 for ch in pattern) and "{," not in pattern:
        try:
            pyre = re.compile(pyp.encode(), re.S if False else 0)
        except re.error:
            pyre = None
    for t in TEXTS:
        got = pq.dfa_match_host(dfa, t)
        exp = oracle.regex_search(pattern, t)
        assert got == int(exp), (pattern, t, got, exp)
        if pyre is not None:
            assert bool(pyre.search(t)) == exp, ("python re disagrees with the oracle", pattern, pyp, t)
    pq.lib().pqg_dfa_free(dfa)


@pytest.mark.parametrize("pattern", REJECTED)
def test_unsupported_patterns_are_rejected_explicitly(pq, oracle, pattern):
    with pytest.raises(ValueError) as e:
        pq.regex_compile(pattern)
    assert str(e.value).startswith("regex: ") and len(str(e.value)) > 10
    with pytest.raises(ValueError):
        oracle.regex_search(pattern, b"abc")


def test_dfa_random_corpus(pq, oracle):
    rng = np.random.default_rng(3)
    alphabet = b"ab.@c0\n"
    texts = [bytes(rng.choice(list(alphabet), size=rng.integers(0, 12))) for _ in range(300)]
    for pattern in [r"a+b", r"^[ab]*@", r"(a|b)*c$", r"\.\d", r"^a.b$", r"[^ab]{2}", r"(ab|ba)+0?$"]:
        dfa = pq.regex_compile(pattern)
        for t in texts:
            assert pq.dfa_match_host(dfa, t) == int(oracle.regex_search(pattern, t)), (pattern, t)
        pq.lib().pqg_dfa_free(dfa)


def test_dfa_size_limit(pq):
    with pytest.raises(ValueError, match="too large"):
        pq.regex_compile(r"(a|b)*a(a|b){14}")


# ---- a-19 pinned to RE2 -----------------------------------------------------------------------
# The reference's regex mode links RE2 (README.md:7-29); its source is absent, but pyarrow's
# `match_substring_regex` IS RE2 (UTF-8 mode on string arrays, unanchored = PartialMatch).  The host
# DFA compiler must agree with it on every supported pattern; whatever RE2 accepts and the DFA
# compiler does not support must be REJECTED explicitly, never matched differently.
RE2_ACCEPTS_WE_REJECT = [r"\b", r"(?i)a", r"(?P<n>a)", r"[[:alpha:]]", r"[é]", r"\pL", r"\Qa\E", r"\xff", r"\Ba", r"(?s).", r"(?m)^a$",
                         r"\A", r"\z", r"\C", r"[^é]", r"\x{10FFFF}"]


def _re2():
    pa = pytest.importorskip("pyarrow")
    pc = pytest.importorskip("pyarrow.compute")
    return pa, pc


def _valid_utf8(b):
    try:
        b.decode()
        return True
    except UnicodeDecodeError:
        return False


def test_dfa_agrees_with_re2_on_the_pattern_list(pq, oracle):
    pa, pc = _re2()
    texts = [t for t in TEXTS if _valid_utf8(t)]
    arr = pa.array([t.decode() for t in texts], type=pa.string())
    for pattern in PATTERNS:
        dfa = pq.regex_compile(pattern)
        exp = pc.match_substring_regex(arr, pattern).to_pylist()  # raises if RE2 rejects what we accept
        got = [bool(pq.dfa_match_host(dfa, t)) for t in texts]
        pq.lib().pqg_dfa_free(dfa)
        assert got == exp, (pattern, [t for t, g, e in zip(texts, got, exp) if g != e])
        assert [bool(oracle.regex_search(pattern, t)) for t in texts] == exp, ("oracle vs RE2", pattern)


def test_dfa_agrees_with_re2_on_a_random_corpus(pq):
    pa, pc = _re2()
    rng = np.random.default_rng(11)
    alphabet = ["a", "b", "c", ".", "@", "0", "7", "\n", " ", "_", "é", "日", "z", "-", "Z"]
    strs = ["".join(rng.choice(alphabet, size=rng.integers(0, 16))) for _ in range(3000)]
    strs += [f"user{rng.integers(0, 10**6)}@mail{rng.integers(0, 999)}.example.com" for _ in range(300)]
    strs += [f"user{rng.integers(0, 10**6)}.mail.example.co!" for _ in range(100)]
    arr = pa.array(strs, type=pa.string())
    enc = [s.encode() for s in strs]
    patterns = [r"^[a-z0-9._]+@[a-z0-9.]+\.com$", r"a+b", r"^[ab]*@", r"(a|b)*c$", r"\.\d", r"^a.b$", r"[^ab]{2}", r"(ab|ba)+0?$",
                r"^.{0,3}$", r"é.", r"[a-c]+[^a-c]", r"\w+@\w+", r"^\S+$", r"\s\S\s", r"(a|é|日)+z", r"^(?:[a-z]|\d){4,}$", r"[-_.]\D",
                r"^$", r"\W$", r"^[^@\n]*$", r"a{2,3}b{0,1}c", r"0|7|Z", r"[0-9][a-z]|[a-z][0-9]", r"\n.", r".\n", r"^\n",
                r"a*?b", r"^a+?$", r"ab??c"]  # lazy quantifiers: same language, so the same match / no-match answer
    for pattern in patterns:
        dfa = pq.regex_compile(pattern)
        exp = np.array(pc.match_substring_regex(arr, pattern).to_pylist())
        got = np.array([bool(pq.dfa_match_host(dfa, t)) for t in enc])
        pq.lib().pqg_dfa_free(dfa)
        bad = np.nonzero(got != exp)[0]
        assert len(bad) == 0, (pattern, [strs[i] for i in bad[:5]])


def test_divergences_from_re2_are_rejections(pq):
    """every construct RE2 takes and the DFA compiler does not: an explicit error, never a different answer"""
    pa, pc = _re2()
    arr = pa.array(["abc"], type=pa.string())
    for pattern in RE2_ACCEPTS_WE_REJECT:
        pc.match_substring_regex(arr, pattern)  # RE2 accepts it
        with pytest.raises(ValueError, match="^regex: "):
            pq.regex_compile(pattern)
    for pattern in REJECTED:
        re2_ok = True
        try:
            pc.match_substring_regex(arr, pattern)
        except Exception:
            re2_ok = False
        assert re2_ok == (pattern in RE2_ACCEPTS_WE_REJECT), pattern
