#!/usr/bin/env python
"""Mechanical check of julia/BridgeB200.jl against include/bridge_b200.h (Julia cannot run in this image).

  * every `ccall((:name, lib), Ret, (ArgTypes...), args...)` names a function the header declares, with the same number
    of arguments, each Julia argument type compatible with the C parameter (scalar width and signedness, pointer vs
    scalar, pointee type), the same return type, and as many call arguments as argument types;
  * the block structure of the file is sound: brackets close in order, every block opener has its `end` (jl_structure);
  * the structs the shim mirrors by hand (BBModel, BBAux, BBThetaSpec) have the byte layout gcc gives bb_model, bb_aux,
    bb_theta_spec (sizeof and every offsetof).

`python tools/check_shim.py` prints a report and exits non-zero on any mismatch; tests/test_cabi.py calls check()."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bridge_b200.h")
SHIM = os.path.join(ROOT, "julia", "BridgeB200.jl")

HANDLES = {"bb_ctx", "bb_ens", "bb_guide", "bb_comm", "bb_user_model", "void"}
C_SCALARS = {"int": "i32", "int32_t": "i32", "int64_t": "i64", "uint32_t": "u32", "uint64_t": "u64", "double": "f64",
             "uint8_t": "u8"}
JL_SCALARS = {"Cint": "i32", "Int32": "i32", "Int64": "i64", "UInt32": "u32", "UInt64": "u64", "Float64": "f64",
              "Cdouble": "f64", "UInt8": "u8"}
JL_STRUCTS = {"BBModel": "bb_model", "BBAux": "bb_aux", "BBThetaSpec": "bb_theta_spec"}


def c_class(param: str) -> str:
    p = re.sub(r"/\*.*?\*/", "", param).strip()
    p = re.sub(r"\bconst\b", "", p)
    stars = p.count("*")
    base = re.sub(r"[\*\s]+\w*$", "", p.replace("*", " * ")).split()[0] if stars else p.split()[0]
    if stars >= 2:
        return "pptr"
    if stars == 1:
        if base in HANDLES:
            return "ptr:handle"
        if base == "char":
            return "ptr:char"
        return "ptr:" + (C_SCALARS.get(base) or base)
    return C_SCALARS[base]


def header_prototypes():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(bb_\w+)\s*\(([^;{}]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "typedef" in ret or name in protos:
            continue
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        rcls = "ptr:char" if "char" in ret else ("ptr:handle" if "*" in ret else C_SCALARS[ret.replace("const", "").strip()])
        protos[name] = (rcls, [c_class(a) for a in params])
    return protos


def jl_class(t: str) -> str:
    t = t.strip()
    if t in JL_SCALARS:
        return JL_SCALARS[t]
    if t == "Cstring":
        return "ptr:char"
    if t in ("Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}", "Ptr{Cstring}"):
        return "pptr"
    if t == "Ptr{Cvoid}":
        return "ptr:handle"
    m = re.fullmatch(r"(?:Ptr|Ref)\{(\w+)\}", t)
    if m:
        inner = m.group(1)
        if inner in JL_STRUCTS:
            return "ptr:" + JL_STRUCTS[inner]
        if inner in JL_SCALARS:
            return "ptr:" + JL_SCALARS[inner]
    raise ValueError(f"unknown Julia ccall type {t!r}")


def split_top(s: str):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out]


def balanced(src: str, start: int) -> int:
    """index just past the parenthesis group that opens at src[start]"""
    depth = 0
    for i in range(start, len(src)):
        if src[i] == "(":
            depth += 1
        elif src[i] == ")":
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced parentheses")


def shim_ccalls():
    src = open(SHIM).read()
    src = re.sub(r"#[^\n]*", "", src)  # comments (no '#' occurs inside the strings of a ccall)
    calls = []
    for m in re.finditer(r"ccall\(", src):
        end = balanced(src, m.end() - 1)
        body = src[m.end():end - 1]
        parts = split_top(body)
        fm = re.fullmatch(r"\(:(\w+),\s*lib\)", parts[0])
        if not fm:
            raise ValueError(f"ccall target not understood: {parts[0]!r}")
        types = parts[2].strip()
        assert types.startswith("(") and types.endswith(")"), types
        inner = types[1:-1].strip()
        if inner.endswith(","):
            inner = inner[:-1]
        argtypes = split_top(inner) if inner else []
        line = src.count("\n", 0, m.start()) + 1
        calls.append((fm.group(1), parts[1].strip(), argtypes, len(parts) - 3, line))
    return calls


def compatible(jl: str, c: str) -> bool:
    if jl == c:
        return True
    if c == "pptr" and jl == "pptr":
        return True
    if c == "ptr:handle" and jl == "ptr:handle":
        return True
    return False


# ------------------------------------------------------------------------------------------- struct layouts
JL_SIZES = {"Int32": 4, "UInt32": 4, "Float64": 8, "Int64": 8, "UInt64": 8, "UInt8": 1}


def jl_struct_layout(name: str):
    src = open(SHIM).read()
    m = re.search(r"struct\s+" + name + r"\b(.*?)\nend", src, flags=re.S)
    assert m, f"struct {name} not found in the shim"
    body = re.sub(r"#[^\n]*", "", m.group(1))
    fields = []
    for f in re.split(r"[;\n]", body):
        f = f.strip()
        if not f:
            continue
        fname, ftype = [x.strip() for x in f.split("::")]
        nt = re.fullmatch(r"NTuple\{(\d+),\s*(\w+)\}", ftype)
        if nt:
            n, el = int(nt.group(1)), JL_SIZES[nt.group(2)]
            fields.append((fname, el, n * el))
        elif ftype.startswith("Ptr{"):
            fields.append((fname, 8, 8))
        else:
            fields.append((fname, JL_SIZES[ftype], JL_SIZES[ftype]))
    off, layout, maxal = 0, [], 1
    for fname, al, size in fields:
        off = (off + al - 1) // al * al
        layout.append((fname, off))
        off += size
        maxal = max(maxal, al)
    total = (off + maxal - 1) // maxal * maxal
    return layout, total


C_FIELDS = {
    "bb_model": ["id", "d", "dprime", "reserved", "par"],
    "bb_aux": ["d", "is_const", "B", "beta", "a", "a_left"],
    "bb_theta_spec": ["m", "aux_kind", "L", "Sigma", "eps", "v", "prior_kind", "prior_a", "prior_b", "start_sd", "start_dir"],
}


def c_struct_layout(cname: str):
    fields = C_FIELDS[cname]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(void){printf("%%zu", sizeof(%s));' % (HEADER, cname)
    for f in fields:
        prog += 'printf(" %%zu", offsetof(%s, %s));' % (cname, f)
    prog += "return 0;}\n"
    with tempfile.TemporaryDirectory() as td:
        cfile, exe = os.path.join(td, "l.c"), os.path.join(td, "l")
        open(cfile, "w").write(prog)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.run([cc, "-o", exe, cfile], check=True)
        nums = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    return list(zip(fields, nums[1:])), nums[0]


_JL_OPENERS = {"function", "macro", "module", "baremodule", "struct", "if", "for", "while", "let", "begin", "do", "try",
               "quote"}


def jl_strip(src: str) -> str:
    """Julia source with comments, strings, character literals and docstrings blanked (newlines kept)."""
    out, i, n = [], 0, len(src)
    while i < n:
        c = src[i]
        if src.startswith("#=", i):
            j = src.find("=#", i + 2)
            j = n if j < 0 else j + 2
            out.append(re.sub(r"[^\n]", " ", src[i:j])); i = j
        elif c == "#":
            j = src.find("\n", i)
            j = n if j < 0 else j
            out.append(" " * (j - i)); i = j
        elif src.startswith('"""', i):
            j = src.find('"""', i + 3)
            j = n if j < 0 else j + 3
            out.append(re.sub(r"[^\n]", " ", src[i:j])); i = j
        elif c == '"':
            j = i + 1
            while j < n and src[j] != '"':
                j += 2 if src[j] == "\\" else 1
            out.append(re.sub(r"[^\n]", " ", src[i:j + 1])); i = j + 1
        elif c == "'" and re.match(r"'(\\.|[^'\\])'", src[i:i + 4]):
            m = re.match(r"'(\\.|[^'\\])'", src[i:i + 4])
            out.append(" " * m.end()); i += m.end()
        else:
            out.append(c); i += 1
    return "".join(out)


def jl_structure(path: str = SHIM):
    """Block structure of the shim (no Julia here to parse it): every (, [, { closes in order, and outside brackets
    (where `for` / `if` belong to comprehensions and `end` is an index) every block opener has its `end`; the file ends
    at depth 0.  -> list of problems."""
    src = jl_strip(open(path).read())
    problems, brackets, blocks = [], [], []
    pairs = {")": "(", "]": "[", "}": "{"}
    line = 1
    for m in re.finditer(r"\n|[()\[\]{}]|:?[A-Za-z_\u0080-\uffff][\w!\u0080-\uffff]*", src):
        tok = m.group(0)
        if tok == "\n":
            line += 1
        elif tok in "([{":
            brackets.append((tok, line))
        elif tok in ")]}":
            if not brackets or brackets[-1][0] != pairs[tok]:
                problems.append(f"line {line}: unmatched {tok}")
                return problems
            brackets.pop()
        elif brackets:
            continue
        elif tok in _JL_OPENERS:
            before = src[:m.start()].rstrip()
            if tok == "struct" and before.endswith("mutable"):
                pass
            blocks.append((tok, line))
        elif tok == "type" and re.search(r"\b(abstract|primitive)\s*$", src[:m.start()]):
            blocks.append((tok, line))
        elif tok == "end":
            if not blocks:
                problems.append(f"line {line}: `end` without an open block")
                return problems
            blocks.pop()
    if brackets:
        problems.append(f"line {brackets[-1][1]}: {brackets[-1][0]} never closed")
    if blocks:
        problems.append(f"line {blocks[-1][1]}: `{blocks[-1][0]}` without `end`")
    return problems


def check(verbose: bool = False):
    """-> list of problems (empty = the shim agrees with the header)"""
    problems = [f"structure: {p}" for p in jl_structure()]
    protos = header_prototypes()
    calls = shim_ccalls()
    seen = set()
    for name, ret, argtypes, nargs, line in calls:
        seen.add(name)
        if name not in protos:
            problems.append(f"line {line}: ccall of {name}, which include/bridge_b200.h does not declare")
            continue
        rcls, pcls = protos[name]
        try:
            jr = jl_class(ret)
            ja = [jl_class(t) for t in argtypes]
        except ValueError as ex:
            problems.append(f"line {line}: {name}: {ex}")
            continue
        if not compatible(jr, rcls):
            problems.append(f"line {line}: {name}: return type {ret} vs C {rcls}")
        if len(ja) != len(pcls):
            problems.append(f"line {line}: {name}: {len(ja)} argument types, the header has {len(pcls)} parameters")
            continue
        if nargs != len(ja):
            problems.append(f"line {line}: {name}: {nargs} call arguments for {len(ja)} argument types")
        for k, (j, c) in enumerate(zip(ja, pcls)):
            if not compatible(j, c):
                problems.append(f"line {line}: {name}: argument {k + 1} is {argtypes[k]} ({j}), the header wants {c}")
    for jl, cn in JL_STRUCTS.items():
        jl_lay, jl_size = jl_struct_layout(jl)
        c_lay, c_size = c_struct_layout(cn)
        if jl_size != c_size:
            problems.append(f"struct {jl}: {jl_size} bytes, {cn} has {c_size}")
        if [o for _, o in jl_lay] != [o for _, o in c_lay]:
            problems.append(f"struct {jl}: field offsets {jl_lay} vs {cn} {c_lay}")
    if verbose:
        print(f"{len(calls)} ccalls of {len(seen)} distinct functions ({len(protos)} declared in the header); "
              f"structs checked: {', '.join(JL_STRUCTS)}")
        for p in problems:
            print("MISMATCH", p)
    return problems, len(calls), len(seen), len(protos)


if __name__ == "__main__":
    probs, *_ = check(verbose=True)
    sys.exit(1 if probs else 0)
