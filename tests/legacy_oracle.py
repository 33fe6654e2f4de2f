"""TEST INFRASTRUCTURE: a faithful, loop-for-loop Python restatement of the LEGACY executor's BGP + FILTER stage (small inputs only):

* perform_join_par_simd_with_strict_filter_1   kolibrie/src/sparql_database.rs:2056-2213 (rows of string bindings, four binding classes)
* apply_filters_simd                           kolibrie/src/sparql_database.rs:1381-1669 (i32 comparison, else byte-wise string = / !=)
* evaluate_filter_expression                   kolibrie/src/sparql_database.rs:1784-1836 (comparisons nested under AND / OR / NOT: f64, else strings)
* the pattern loop of execute_query            kolibrie/src/execute_query.rs:221-303

Operands containing + - * / (the reference's arithmetic-expression parser) are outside this restatement: NotImplementedError."""
from kolibrie_b200.engine import And, Comparison, FunctionCall, Not, Or, rust_parse_f64, rust_parse_i32


def join_strict_filter_1(subject_var, predicate, object_var, triples, final_results, literal_filter):
    """triples: decoded (subject, predicate, object) strings in store order"""
    if not final_results:
        return []
    both, subj_b, obj_b, neither = {}, {}, {}, []
    for r in final_results:
        sb, ob = r.get(subject_var), r.get(object_var)
        if sb is not None and ob is not None:
            both.setdefault((sb, ob), []).append(r)
        elif sb is not None:
            subj_b.setdefault(sb, []).append(r)
        elif ob is not None:
            obj_b.setdefault(ob, []).append(r)
        else:
            neither.append(r)
    out = []
    for s, p, o in triples:
        if p != predicate:
            continue
        if literal_filter is not None and o != literal_filter:
            continue
        for r in both.get((s, o), []):
            out.append(dict(r))
        for r in subj_b.get(s, []):
            e = dict(r)
            if object_var in e:
                if e[object_var] != o:
                    continue
            else:
                e[object_var] = o
            out.append(e)
        for r in obj_b.get(o, []):
            e = dict(r)
            if subject_var in e:
                if e[subject_var] != s:
                    continue
            else:
                e[subject_var] = s
            out.append(e)
        for r in neither:
            e = dict(r)
            if subject_var in e:
                if e[subject_var] != s:
                    continue
            else:
                e[subject_var] = s
            if object_var in e:
                if e[object_var] != o:
                    continue
            else:
                e[object_var] = o
            out.append(e)
    return out


def _arith(s):
    return any(ch in s for ch in "+-*/")


def _operand_f64(row, text):
    if text.startswith("?"):
        v = row.get(text)
        return None if v is None else rust_parse_f64(v)
    return rust_parse_f64(text)


def evaluate_filter_expression(row, e):
    """nested semantics (sparql_database.rs:1784-1836)"""
    if isinstance(e, Comparison):
        if _arith(e.var) or _arith(e.value):
            raise NotImplementedError("arithmetic operands")
        a, b = _operand_f64(row, e.var), _operand_f64(row, e.value)
        if a is not None and b is not None:
            return {"=": a == b, "!=": a != b, ">": a > b, ">=": a >= b, "<": a < b, "<=": a <= b}.get(e.op, False)
        ls = row.get(e.var, e.var) if e.var.startswith("?") else e.var
        rs = row.get(e.value, e.value) if e.value.startswith("?") else e.value
        return {"=": ls == rs, "!=": ls != rs}.get(e.op, False)
    if isinstance(e, And):
        return evaluate_filter_expression(row, e.left) and evaluate_filter_expression(row, e.right)
    if isinstance(e, Or):
        return evaluate_filter_expression(row, e.left) or evaluate_filter_expression(row, e.right)
    if isinstance(e, Not):
        return not evaluate_filter_expression(row, e.inner)
    if isinstance(e, FunctionCall):
        if e.name == "isTRIPLE" and e.args:
            a = e.args[0]
            v = row.get(a, "") if a.startswith("?") else a
            return v.startswith("<<") and v.endswith(">>")
        return False
    raise NotImplementedError(type(e))


def apply_filters_simd(rows, filters):
    out = []
    for row in rows:
        ok = True
        for f in filters:
            if isinstance(f, Comparison):
                if _arith(f.var) or _arith(f.value):
                    raise NotImplementedError("arithmetic operands")
                v = row.get(f.var)
                if v is None:
                    r = False
                else:
                    a, b = rust_parse_i32(v), rust_parse_i32(f.value)
                    if a is not None and b is not None:
                        r = {"=": a == b, "!=": a != b, ">": a > b, ">=": a >= b, "<": a < b, "<=": a <= b}.get(f.op, False)
                    else:
                        same = v.encode() == f.value.encode()
                        r = {"=": same, "!=": not same}.get(f.op, False)
            else:
                r = evaluate_filter_expression(row, f)
            if not r:
                ok = False
                break
        if ok:
            out.append(row)
    return out


def execute_bgp(triples, patterns, filters=()):
    """the pattern loop of execute_query (no VALUES clause: one empty row to start from), then the filters. Rows keep every binding,
    pseudo-bindings named after constants included."""
    rows = [{}]
    for s, p, o in patterns:
        rows = join_strict_filter_1(s, p, o, triples, rows, None if o.startswith("?") else o)
    return apply_filters_simd(rows, list(filters))
