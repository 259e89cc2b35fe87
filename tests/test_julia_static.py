"""Static sanity of julia/OptimPackB200.jl and tools/make_reference_golden.jl.

There is no Julia runtime where this repository is built, so the module that IS the product
boundary for Julia callers can only be checked statically: (i) every block opener has its `end`,
brackets and strings balance; (ii) every `ccall((:sym, libopkb200), ret, (argtypes...), args...)`
names a symbol that include/opkb200.h declares, with the same number of parameters and
compatible parameter kinds (int / int64 / double / pointer), the same return kind, and as many
actual arguments as argument types."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODULE = os.path.join(ROOT, "julia", "OptimPackB200.jl")
GENERATOR = os.path.join(ROOT, "tools", "make_reference_golden.jl")

OPENERS = {"function", "if", "for", "while", "struct", "module", "let", "begin", "try", "do", "quote", "macro"}


def strip_comments_and_strings(src):
    """Julia source with comments and string / char literal CONTENTS blanked (same length)."""
    out = []
    i, n = 0, len(src)
    while i < n:
        c = src[i]
        if src.startswith("#=", i):
            j = src.find("=#", i + 2)
            j = n if j < 0 else j + 2
            out.append(re.sub(r"[^\n]", " ", src[i:j]))
            i = j
        elif c == "#":
            j = src.find("\n", i)
            j = n if j < 0 else j
            out.append(" " * (j - i))
            i = j
        elif src.startswith('"""', i):
            j = src.find('"""', i + 3)
            assert j >= 0, "unterminated triple-quoted string"
            out.append('"' + re.sub(r"[^\n]", " ", src[i + 1:j + 2]) + '"')
            i = j + 3
        elif c == '"':
            j = i + 1
            while j < n and src[j] != '"':
                j += 2 if src[j] == "\\" else 1
            assert j < n, "unterminated string"
            out.append('"' + " " * (j - i - 1) + '"')
            i = j + 1
        elif c == "'" and re.match(r"'(\\.|[^\\'])'", src[i:i + 4]):
            m = re.match(r"'(\\.|[^\\'])'", src[i:i + 4])
            out.append("'" + " " * (m.end() - 2) + "'")
            i += m.end()
        else:
            out.append(c)
            i += 1
    return "".join(out)


def check_blocks(path):
    """Blocks and brackets balance.  A `for` / `if` with an unclosed bracket opened since the innermost
    block started is a generator / comprehension clause (no `end`); an `end` with an unclosed `[`
    opened since then is an index (`a[end]`)."""
    src = strip_comments_and_strings(open(path).read())
    stack = []          # (opener, line, bracket depth when it opened)
    brackets = []       # (bracket, line)
    line = 1
    prev = ""
    for m in re.finditer(r"\n|[A-Za-z_][A-Za-z_0-9!]*|[()\[\]{}]|\S", src):
        tok = m.group(0)
        if tok == "\n":
            line += 1
            continue
        base = stack[-1][2] if stack else 0
        adjacent_colon = m.start() > 0 and src[m.start() - 1] == ":" and (m.start() < 2 or src[m.start() - 2] != ":")
        if tok in "([{":
            brackets.append((tok, line))
        elif tok in ")]}":
            assert len(brackets) > base, "%s:%d: unmatched %s" % (path, line, tok)
            o, _ = brackets.pop()
            assert "([{".index(o) == ")]}".index(tok), "%s:%d: %s closed by %s" % (path, line, o, tok)
        elif tok == "end" and not adjacent_colon:
            if any(b[0] == "[" for b in brackets[base:]):
                pass                                        # a[end], a[2:end]
            else:
                assert stack, "%s:%d: `end` without an opener" % (path, line)
                assert len(brackets) == base, "%s:%d: `end` of `%s` (line %d) inside an unclosed %s" % (
                    path, line, stack[-1][0], stack[-1][1], brackets[-1][0])
                stack.pop()
        elif tok in OPENERS and prev != "." and not adjacent_colon:
            if tok in ("for", "if") and len(brackets) > base:
                pass                                        # generator / comprehension clause
            elif tok == "struct" and prev == "mutable":
                stack.append(("mutable struct", line, len(brackets)))
            else:
                stack.append((tok, line, len(brackets)))
        prev = tok
    assert not brackets, "%s: unclosed %s opened at line %d" % (path, brackets[-1][0], brackets[-1][1])
    assert not stack, "%s: `%s` opened at line %d has no `end`" % (path, stack[-1][0], stack[-1][1])


def split_top(s):
    """Split on commas at bracket depth 0."""
    parts, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    last = "".join(cur).strip()
    if last:
        parts.append(last)
    return parts


def header_signatures():
    hdr = open(os.path.join(ROOT, "include", "opkb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    sigs = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z_0-9 \*]*?)\b(opkb_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", hdr):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "typedef" in ret:
            continue
        plist = [] if params in ("", "void") else split_top(params)
        sigs[name] = (ckind(ret), [ckind(p) for p in plist])
    return sigs


def ckind(decl):
    d = decl.strip()
    if "*" in d or "[" in d or "opkb_fg_callback" in d:
        return "ptr"
    if re.search(r"\buint64_t\b", d):
        return "u64"
    if re.search(r"\bint64_t\b", d):
        return "i64"
    if re.search(r"\bdouble\b", d):
        return "f64"
    if re.search(r"\bint\b", d):
        return "i32"
    if re.search(r"\bvoid\b", d):
        return "void"
    raise AssertionError("unknown C type: " + decl)


def jkind(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t in ("Cstring", "Ptr{Cvoid}"):
        return "ptr"
    return {"Cint": "i32", "Int32": "i32", "Int64": "i64", "Clonglong": "i64", "UInt64": "u64", "Cdouble": "f64",
            "Float64": "f64", "Cvoid": "void", "Nothing": "void"}.get(t, "?" + t)


def ccalls(path):
    src = strip_comments_and_strings(open(path).read())
    for m in re.finditer(r"ccall\(\(\s*:(opkb_[a-z0-9_]+)\s*,\s*libopkb200\s*\)\s*,", src):
        # the rest of the call up to its closing parenthesis
        depth, j = 1, m.end()
        while depth and j < len(src):
            depth += src[j] in "([{"
            depth -= src[j] in ")]}"
            j += 1
        args = split_top(src[m.end():j - 1])
        line = src.count("\n", 0, m.start()) + 1
        yield m.group(1), args, line


@pytest.mark.parametrize("path", [MODULE, GENERATOR])
def test_blocks_and_brackets_balance(path):
    check_blocks(path)


def test_every_ccall_matches_the_header():
    sigs = header_signatures()
    assert len(sigs) > 80
    n = 0
    used = set()
    for name, args, line in ccalls(MODULE):
        where = "julia/OptimPackB200.jl:%d %s" % (line, name)
        assert name in sigs, where + ": not declared in include/opkb200.h"
        ret, params = sigs[name]
        assert len(args) >= 2, where
        jret, jtypes = args[0], args[1]
        assert jtypes.startswith("(") and jtypes.endswith(")"), where + ": argument types must be a tuple"
        types = split_top(jtypes[1:-1])
        assert len(types) == len(params), "%s: %d argument types, the header declares %d" % (where, len(types), len(params))
        assert len(args) - 2 == len(params), "%s: %d arguments passed, %d declared" % (where, len(args) - 2, len(params))
        assert jkind(jret) == ret, "%s: return type %s vs C %s" % (where, jret, ret)
        for k, (jt, ck) in enumerate(zip(types, params)):
            assert jkind(jt) == ck, "%s: parameter %d is %s in Julia, %s in C" % (where, k + 1, jt, ck)
        used.add(name)
        n += 1
    assert n >= 60 and len(used) >= 50, (n, len(used))


def test_static_checker_catches_breakage(tmp_path):
    """The checker itself: a dropped `end`, a wrong arity."""
    good = "module M\nfunction f(x)\n    y = [i for i in 1:3 if i > 1]\n    return x[end] + y[1]\nend\nend\n"
    p = tmp_path / "a.jl"
    p.write_text(good)
    check_blocks(str(p))
    p.write_text(good.replace("    return x[end] + y[1]\nend\n", "    return x[end] + y[1]\n"))
    with pytest.raises(AssertionError):
        check_blocks(str(p))


# ---------------------------------------------------------------------------
# Drop-in signatures: every keyword of the reference's entry points is accepted by the module.
# The lists are the reference's (src/quasinewton.jl:226-242, src/spg.jl:168-178); when the reference
# tree is at hand (the build container, not the GPU box) they are re-read from it.
REF_VMLMB_KW = ["mem", "lower", "upper", "autodiff", "blmvm", "fmin", "maxiter", "maxeval", "xtol", "ftol", "gtol",
                "epsilon", "verb", "printer", "output", "lnsrch"]
REF_SPG_KW = ["autodiff", "ws", "maxit", "maxfc", "eps1", "eps2", "eta", "printer", "verb", "io"]


def _signature_keywords(src, head):
    """Keyword names of the Julia method whose definition starts with `head` (up to `) where` / `)`)."""
    i = src.index(head)
    j = src.index(";", i)
    depth, k = 1, j
    # the signature ends at the parenthesis that closes the one opened by `head`
    depth = 0
    for k in range(i + len("function"), len(src)):
        ch = src[k]
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                break
    sig = src[j + 1:k]
    names, depth, tok = [], 0, ""
    for ch in sig + ",":
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            name = re.match(r"\s*([A-Za-z_][A-Za-z_0-9!]*)", tok)
            if name:
                names.append(name.group(1))
            tok = ""
        else:
            tok += ch
    return names


def test_julia_entry_points_accept_the_reference_keywords():
    src = open(MODULE).read()
    vm = _signature_keywords(src, "function vmlmb!(fg!, x::DeviceVector{T};")
    vn = _signature_keywords(src, "function vmlmb_native!(fg!, x::DeviceVector{T};")
    sp = _signature_keywords(src, "function fused_spg!(fg!, box::BoundedSet, x::DeviceVector{T}, m::Int;")
    for kw in REF_VMLMB_KW:
        assert kw in vm, ("vmlmb!", kw, vm)
        assert kw in vn, ("vmlmb_native!", kw, vn)
    for kw in REF_SPG_KW:
        assert kw in sp, ("spg!", kw, sp)
    # the generic entry points forward every keyword
    assert "vmlmb(fg!, x0::Array; kwds...)" in src and "spg!(fg!, prj!, x::DeviceVector{T}, m::Integer; kwds...)" in src
    ref = "/root/reference/src"
    if os.path.isdir(ref):
        q = open(os.path.join(ref, "quasinewton.jl")).read()
        s = open(os.path.join(ref, "spg.jl")).read()
        assert _signature_keywords(q, "function vmlmb!(fg!, x::T;") == REF_VMLMB_KW
        assert _signature_keywords(s, "function spg!(fg!, prj!, x, m::Integer;") == REF_SPG_KW
