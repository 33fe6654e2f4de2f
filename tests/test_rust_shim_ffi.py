"""CPU: rust_shim/src/ffi.rs against include/kolibrie_b200.h. There is no Rust toolchain in this image, so the bindings cannot be
compiled; this test is the check they get: every `extern "C"` declaration names a function the header declares, with the same number
of parameters, the same scalar widths and the same pointer depth / constness per parameter and for the result; the #[repr(C)] structs
have the header's fields in the header's order; the constants carry the header's values."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def strip_c(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([<":
            depth += 1
        elif ch in ")]>":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


C_SCALAR = {"uint32_t": "u32", "uint64_t": "u64", "int32_t": "i32", "int": "i32", "uint8_t": "u8", "double": "f64", "char": "i8", "void": "void",
            "kb_status": "i32"}
RS_SCALAR = {"u32": "u32", "u64": "u64", "i32": "i32", "c_int": "i32", "u8": "u8", "c_double": "f64", "f64": "f64", "c_char": "i8", "c_void": "void",
             "kb_status": "i32"}
HANDLES = {"kb_ctx": "KbCtx", "kb_rel": "KbRel", "kb_groups": "KbGroups", "kb_strings": "KbStrings", "kb_plan": "KbPlan", "kb_pattern": "KbPattern",
           "kb_filter_op": "KbFilterOp", "kb_agg": "KbAgg", "kb_rule": "KbRule", "kb_fixpoint_stats": "KbFixpointStats", "kb_term": "KbTerm",
           "kb_rule_filter": "KbRuleFilter", "kb_stats": "KbStats"}


def c_type(decl, with_name=True):
    """'const uint32_t* const* cols' -> (base, [constness of each pointer level's pointee, outermost last])"""
    d = decl.strip()
    d = re.sub(r"\[[^\]]*\]", "*", d)  # arrays decay
    if with_name:
        d = re.sub(r"\b[A-Za-z_][A-Za-z_0-9]*\s*$", "", d) if not d.endswith("*") and len(d.split()) > 1 else d
    toks = re.findall(r"const|\*|[A-Za-z_][A-Za-z_0-9]*", d)
    toks = [t for t in toks if t not in ("struct", "unsigned") ]
    base = next(t for t in toks if t not in ("const", "*"))
    levels, const_pending = [], toks[0] == "const" or ("const" in toks[: toks.index(base) + 2] and toks[toks.index(base) + 1: toks.index(base) + 2] == ["const"])
    const_pending = "const" in toks[: toks.index(base)] or (toks.index(base) + 1 < len(toks) and toks[toks.index(base) + 1] == "const")
    i = toks.index(base) + 1
    if i < len(toks) and toks[i] == "const":
        i += 1
    while i < len(toks):
        if toks[i] == "*":
            levels.append(const_pending)
            const_pending = i + 1 < len(toks) and toks[i + 1] == "const"
            i += 2 if const_pending else 1
        else:
            i += 1
    return HANDLES.get(base, C_SCALAR.get(base, base)), levels


def rs_type(t):
    t = t.strip()
    levels = []
    while t.startswith("*"):
        m = re.match(r"\*(const|mut)\s+", t)
        levels.append(m.group(1) == "const")
        t = t[m.end():]
    levels.reverse()  # Rust writes the outermost pointer first, C the innermost
    return RS_SCALAR.get(t, t), levels


def header_functions():
    text = strip_c(open(os.path.join(ROOT, "include", "kolibrie_b200.h")).read())
    fns = {}
    for m in re.finditer(r"KB_API\s+([\w\s\*]+?)\b(kb_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        fns[name] = (c_type(ret, with_name=False), [] if args in ("void", "") else [c_type(a) for a in split_args(args)])
    return fns


def rust_functions():
    text = re.sub(r"//[^\n]*", " ", open(os.path.join(ROOT, "rust_shim", "src", "ffi.rs")).read())
    block = text[text.index('extern "C"'):]
    fns = {}
    for m in re.finditer(r"pub fn (kb_\w+)\s*\(([^;]*?)\)\s*(?:->\s*([^;]+?))?\s*;", block, flags=re.S):
        args = [a.split(":", 1)[1] for a in split_args(" ".join(m.group(2).split()))]
        fns[m.group(1)] = (rs_type(m.group(3)) if m.group(3) else ("void", []), [rs_type(a) for a in args])
    return fns, text


def test_every_rust_declaration_matches_the_header():
    hdr = header_functions()
    rs, _ = rust_functions()
    assert len(hdr) >= 60
    assert set(rs) == set(hdr), f"not bound: {sorted(set(hdr) - set(rs))}; not in the header: {sorted(set(rs) - set(hdr))}"
    for name, (ret, args) in sorted(rs.items()):
        assert name in hdr, f"{name}: bound in ffi.rs, not declared in the header"
        hret, hargs = hdr[name]
        assert len(args) == len(hargs), f"{name}: {len(args)} parameters in ffi.rs, {len(hargs)} in the header"
        assert ret == hret, f"{name}: result {ret} vs {hret}"
        for i, (a, h) in enumerate(zip(args, hargs)):
            assert a[0] == h[0] and len(a[1]) == len(h[1]), f"{name} parameter {i}: {a} in ffi.rs, {h} in the header"
            # a `*mut` where the header says const would let Rust write through it; `*const` where the header writes is a bug too
            assert a[1] == h[1] or (a[0] == "void"), f"{name} parameter {i}: constness {a[1]} in ffi.rs, {h[1]} in the header"


def test_structs_and_constants_match_the_header():
    hdr = strip_c(open(os.path.join(ROOT, "include", "kolibrie_b200.h")).read())
    _, rs = rust_functions()
    for c_name, rs_name in (("kb_term", "KbTerm"), ("kb_pattern", "KbPattern"), ("kb_filter_op", "KbFilterOp"), ("kb_agg", "KbAgg"),
                            ("kb_rule_filter", "KbRuleFilter"), ("kb_rule", "KbRule"), ("kb_fixpoint_stats", "KbFixpointStats"), ("kb_stats", "KbStats")):
        body = re.search(r"typedef struct " + c_name + r"\s*\{(.*?)\}\s*" + c_name + r"\s*;", hdr, flags=re.S).group(1)
        c_fields = []
        for f in [x.strip() for x in body.split(";") if x.strip()]:
            names = f.split(None, 1)[1] if not f.startswith("const") else f.split(None, 2)[2]
            c_fields += [re.sub(r"[\*\s]|\[.*\]", "", n) for n in names.split(",")]
        rbody = re.search(r"pub struct " + rs_name + r"\s*\{(.*?)\}", rs, flags=re.S).group(1)
        r_fields = re.findall(r"pub (\w+)\s*:", rbody)
        assert r_fields == c_fields, (c_name, r_fields, c_fields)
    consts = dict(re.findall(r"pub const (KB_\w+): \w+ = ([^;]+);", rs))
    enums = dict(re.findall(r"\b(KB_[A-Z_0-9]+)\s*=\s*(-?\d+)", hdr))
    checked = 0
    for k, v in consts.items():
        if k in enums:
            assert int(v.replace("_", ""), 0) == int(enums[k]), k
            checked += 1
    assert checked >= 25, checked
    assert int(consts["KB_TAG_INFERRED"].replace("_", ""), 16) == int(re.search(r"#define KB_TAG_INFERRED (0x[0-9A-Fa-f]+)", hdr).group(1), 16)
    assert int(consts["KB_MAX_COLS"]) == int(re.search(r"#define KB_MAX_COLS (\d+)", hdr).group(1))
