"""Python mirror of the reference's host-side interfaces for the hot path, on top of the C ABI.

Same names, argument meaning and error behaviour as the reference (paths relative to /root/reference):

* `Dictionary`            — shared/src/dictionary.rs:17-51 (ids dense from 0 in first-seen order)
* `Term`/`TriplePattern`  — shared/src/terms.rs:13-23
* `PhysicalOperator`      — kolibrie/src/streamertail_optimizer/operators/physical.rs:16-76 (the operators on the hot path)
* `ExecutionEngine`       — kolibrie/src/streamertail_optimizer/execution/engine.rs:27,54 (`execute`, `execute_with_ids`)
* `Condition`             — kolibrie/src/streamertail_optimizer/types.rs:110-186 (FILTER expressions)
* `SparqlDatabase`        — kolibrie/src/sparql_database.rs:49-60,215-258,3364-3394 (triples + dictionary + build_all_indexes)
* `Rule`/`FilterCondition`/`Reasoner` — shared/src/rule.rs:14-25, datalog/src/reasoning.rs:31-38,60-100,
                            datalog/src/reasoning/materialisation/semi_naive.rs:89, my_naive.rs:74

The canonical C++ mirror (the reference is compiled code) is kolibrie_b200/host/kolibrie_host.hpp; this module exists because
pytest, bench.py and torch.distributed live in Python. All work is done by libkolibrie_b200.so — there is no CPU path here.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import capi as c

# ---------------------------------------------------------------------------------------------------------------------
_RUST_F64 = re.compile(r"[+-]?(?:(?:inf|infinity|nan)|(?:(?:[0-9]+\.?[0-9]*|\.[0-9]+)(?:[eE][+-]?[0-9]+)?))", re.IGNORECASE)
_RUST_F64_SPECIAL = re.compile(r"[+-]?(?:inf|infinity|nan)", re.IGNORECASE)


def rust_parse_f64(s: str) -> Optional[float]:
    """`str::parse::<f64>()` acceptance of Rust (no whitespace, no '_', no hex; inf/infinity/nan in any case; '5.', '.5', '1e5')."""
    if not isinstance(s, str) or not s.isascii() or not _RUST_F64.fullmatch(s):
        return None
    if _RUST_F64_SPECIAL.fullmatch(s):
        return float(s)  # Python accepts the same special spellings
    return float(s)


class Dictionary:
    def __init__(self):
        self.string_to_id: Dict[str, int] = {}
        self.id_to_string: List[str] = []

    def encode(self, value: str) -> int:
        i = self.string_to_id.get(value)
        if i is None:
            i = len(self.id_to_string)
            assert i < 0x8000_0000, "Dictionary ID space exhausted"  # dictionary.rs:36-40
            self.string_to_id[value] = i
            self.id_to_string.append(value)
        return i

    def encode_bulk(self, ctx, terms: Sequence[str]) -> np.ndarray:
        """the load-time encode of a whole batch (sparql_database.rs:1000-1013 calls encode once per term) on the device: kb_dict_encode
        hands out the ids the sequential loop would, and this dictionary learns the new strings from the positions that introduced them.
        The device dictionary must mirror this one (it does when every string entered through encode_bulk / dict_strings_load)."""
        n_dev, _ = ctx.dict_strings_info()
        if n_dev != len(self.id_to_string):
            ctx.dict_strings_load(self.id_to_string)
        ids, first = ctx.dict_encode(terms)
        for pos in first:
            t = terms[int(pos)]
            assert len(self.id_to_string) < 0x8000_0000, "Dictionary ID space exhausted"  # dictionary.rs:36-40
            self.string_to_id[t] = len(self.id_to_string)
            self.id_to_string.append(t)
        return ids

    def lookup(self, value: str) -> Optional[int]:
        return self.string_to_id.get(value)

    def decode(self, i: int) -> Optional[str]:
        return self.id_to_string[i] if 0 <= i < len(self.id_to_string) else None

    def numeric_table(self) -> Tuple[np.ndarray, np.ndarray]:
        n = len(self.id_to_string)
        num = np.zeros(n, dtype=np.float64)
        isn = np.zeros(n, dtype=np.uint8)
        for i, s in enumerate(self.id_to_string):
            v = rust_parse_f64(s)
            if v is not None:
                num[i] = v
                isn[i] = 1
        return num, isn


# ---------------------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Variable:
    name: str


@dataclass(frozen=True)
class Constant:
    id: int


@dataclass(frozen=True)
class QuotedTriple:
    """Term::QuotedTriple(Box<(Term, Term, Term)>) (shared/src/terms.rs:19): an RDF-star pattern `<< s p o >>` in subject or object position"""
    s: object
    p: object
    o: object


Term = Union[Variable, Constant, QuotedTriple]
TriplePattern = Tuple[Term, Term, Term]

QUOTED_TRIPLE_ID_BIT = 0x8000_0000  # shared/src/quoted_triple_store.rs:17


def is_quoted_triple_id(i: int) -> bool:
    return (i & QUOTED_TRIPLE_ID_BIT) != 0


class QuotedTripleStore:
    """shared/src/quoted_triple_store.rs:28-80: quoted triples as u32 ids with bit 31 set, bidirectional, deduplicating; components
    may themselves be quoted-triple ids (nesting)."""

    def __init__(self):
        self.id_to_components: Dict[int, Tuple[int, int, int]] = {}
        self.components_to_id: Dict[Tuple[int, int, int], int] = {}
        self.next_qt_id = QUOTED_TRIPLE_ID_BIT

    def encode(self, subject: int, predicate: int, obj: int) -> int:
        key = (int(subject), int(predicate), int(obj))
        i = self.components_to_id.get(key)
        if i is not None:
            return i
        i = self.next_qt_id
        self.next_qt_id += 1
        self.id_to_components[i] = key
        self.components_to_id[key] = i
        return i

    def decode(self, i: int) -> Optional[Tuple[int, int, int]]:
        return self.id_to_components.get(int(i))

    def __len__(self):
        return len(self.id_to_components)

    def is_empty(self) -> bool:
        return not self.id_to_components

    def merge(self, other: "QuotedTripleStore"):
        for i, comp in other.id_to_components.items():
            self.id_to_components.setdefault(i, comp)
            self.components_to_id.setdefault(comp, i)
        self.next_qt_id = max(self.next_qt_id, other.next_qt_id)


def _strip(v: str) -> str:
    return v[1:] if v.startswith("?") else v  # engine.rs strips '?' from variable names everywhere


class SlotMap:
    """variable name <-> integer slot (variables never reach the device as strings)"""

    def __init__(self):
        self.slot: Dict[str, int] = {}
        self.names: List[str] = []

    def of(self, name: str) -> int:
        name = _strip(name)
        if name not in self.slot:
            self.slot[name] = len(self.names)
            self.names.append(name)
        return self.slot[name]

    def term(self, t: Term) -> c.KbTerm:
        if isinstance(t, QuotedTriple):
            raise c.KolibrieError(c.KB_E_UNSUPPORTED, "a quoted-triple term must be resolved on the host first (ExecutionEngine._scan_quoted)")
        return c.V(self.of(t.name)) if isinstance(t, Variable) else c.K(t.id)

    def pattern(self, p: TriplePattern) -> c.KbPattern:
        return c.pattern(self.term(p[0]), self.term(p[1]), self.term(p[2]))


# ---------------------------------------------------------------------------------------------------------------------
# FILTER expressions (shared/src/query.rs:15-57) and their evaluation contract (types.rs:110-186)
@dataclass
class Comparison:
    var: str
    op: str
    value: str


@dataclass
class And:
    left: object
    right: object


@dataclass
class Or:
    left: object
    right: object


@dataclass
class Not:
    inner: object


@dataclass
class Arith:
    """ArithmeticExpression: ('op', l, r) trees with str operands; truthy when != 0"""
    expr: object


@dataclass
class FunctionCall:
    name: str
    args: List[str]


_CMP = {">": c.CMP_GT, ">=": c.CMP_GE, "<": c.CMP_LT, "<=": c.CMP_LE}


class Condition:
    def __init__(self, expression):
        self.expression = expression

    def compile(self, slots: SlotMap, dictionary: Dictionary) -> List[c.KbFilterOp]:
        ops: List[c.KbFilterOp] = []

        def arith(e):
            if isinstance(e, str):
                if e.startswith("?"):
                    ops.append(c.fop(c.F_PUSH_VAR, slot=slots.of(e)))
                else:
                    v = rust_parse_f64(e)
                    if v is None:
                        raise c.KolibrieError(c.KB_E_UNSUPPORTED, f"arithmetic operand {e!r} is not a number (the reference evaluates the filter to false)")
                    ops.append(c.fop(c.F_PUSH_CONST, value=v))
                return
            kind, l, r = e
            arith(l)
            arith(r)
            ops.append(c.fop({"+": c.F_ADD, "-": c.F_SUB, "*": c.F_MUL, "/": c.F_DIV}[kind]))

        def rec(e):
            if isinstance(e, Comparison):
                slot = slots.of(e.var)
                if e.op in ("=", "!="):
                    # decoded == literal  <=>  id == encode(literal) when the literal is in the dictionary (types.rs:131-132)
                    lit = dictionary.lookup(e.value)
                    ops.append(c.fop(c.F_EQ_ID if e.op == "=" else c.F_NE_ID, slot=slot, id=c.KB_ID_NONE if lit is None else lit))
                elif e.op in _CMP:
                    v = rust_parse_f64(e.value)  # types.rs:133-148: value.parse::<f64>().unwrap_or(0.0) — also for "?other" (quirk Q10)
                    ops.append(c.fop(c.F_CMP_NUM, slot=slot, cmp=_CMP[e.op], value=0.0 if v is None else v))
                else:
                    raise c.KolibrieError(c.KB_E_UNSUPPORTED, f"operator {e.op!r}: the reference evaluates it to false")
            elif isinstance(e, And):
                rec(e.left); rec(e.right); ops.append(c.fop(c.F_AND))
            elif isinstance(e, Or):
                rec(e.left); rec(e.right); ops.append(c.fop(c.F_OR))
            elif isinstance(e, Not):
                rec(e.inner); ops.append(c.fop(c.F_NOT))
            elif isinstance(e, Arith):
                arith(e.expr); ops.append(c.fop(c.F_TRUTHY))
            elif isinstance(e, FunctionCall):
                if e.name == "isTRIPLE" and e.args:
                    ops.append(c.fop(c.F_IS_TRIPLE, slot=slots.of(e.args[0])))
                else:
                    raise c.KolibrieError(c.KB_E_UNSUPPORTED, f"function {e.name!r}: the reference evaluates it to false")
            else:
                raise TypeError(e)

        rec(self.expression)
        return ops


# ---------------------------------------------------------------------------------------------------------------------
# PhysicalOperator (operators/physical.rs:16-76) — the variants on the hot path
@dataclass
class TableScan:
    pattern: TriplePattern


@dataclass
class IndexScan:
    pattern: TriplePattern


@dataclass
class Filter:
    input: object
    condition: Condition


@dataclass
class Projection:
    input: object
    variables: List[str]


@dataclass
class HashJoin:
    left: object
    right: object


@dataclass
class OptimizedHashJoin:
    left: object
    right: object


@dataclass
class NestedLoopJoin:
    left: object
    right: object


@dataclass
class ParallelJoin:
    left: object
    right: object


@dataclass
class StarJoin:
    join_var: str
    patterns: List[TriplePattern]


@dataclass
class InMemoryBuffer:
    content: List[Dict[str, int]]


class SparqlDatabase:
    """triples (set semantics, iteration in (s,p,o) order like the reference's BTreeSet) + dictionary + the device store."""

    def __init__(self, ctx: Optional[c.Context] = None, device: int = 0):
        self.dictionary = Dictionary()
        self.quoted_triple_store = QuotedTripleStore()  # sparql_database.rs:58
        self.triples: set = set()
        self.ctx = ctx or c.Context(device)
        self._uploaded_version = -1
        self._version = 0

    def encode_term(self, t) -> int:
        """a term given as a string, or as a nested (s, p, o) tuple for a quoted triple `<< s p o >>` (components encoded first, in
        s, p, o order, then the quoted triple itself: sparql_database.rs encode_quoted_triple)"""
        if isinstance(t, (tuple, list)):
            s, p, o = (self.encode_term(x) for x in t)
            return self.quoted_triple_store.encode(s, p, o)
        return self.dictionary.encode(t)

    def decode_term(self, i: int) -> Optional[str]:
        """Dictionary::decode_term (dictionary.rs:58-68): quoted-triple ids decode recursively to `<< s p o >>`"""
        if is_quoted_triple_id(i):
            comp = self.quoted_triple_store.decode(i)
            if comp is None:
                return None
            parts = [self.decode_term(x) for x in comp]
            if any(x is None for x in parts):
                return None
            return "<< " + " ".join(parts) + " >>"
        return self.dictionary.decode(i)

    def add_statement(self, s, p, o):
        """add one triple whose terms are strings or nested tuples (quoted triples)"""
        self.add_triple((self.encode_term(s), self.encode_term(p), self.encode_term(o)))

    def add_triple_parts(self, s: str, p: str, o: str):
        t = (self.dictionary.encode(s), self.dictionary.encode(p), self.dictionary.encode(o))
        self.add_triple(t)

    def add_triples_bulk(self, statements: Sequence[Tuple[str, str, str]]):
        """bulk load of parsed (s, p, o) STRING triples (what parse_ntriples / parse_rdf hand to the per-term encode loop,
        sparql_database.rs:1000-1013): one device encode of all 3n terms in s, p, o order per triple — the ids are those of the
        sequential loop (dictionary.rs:32-48) — then the triples join the set"""
        terms = [t for st in statements for t in st]
        ids = self.dictionary.encode_bulk(self.ctx, terms).reshape(-1, 3)
        for row in ids:
            self.add_triple((int(row[0]), int(row[1]), int(row[2])))

    def add_triple(self, t: Tuple[int, int, int]):  # sparql_database.rs:215-226
        if t not in self.triples:
            self.triples.add(t)
            self._version += 1

    def delete_triple(self, t: Tuple[int, int, int]) -> bool:  # sparql_database.rs:229-242
        if t in self.triples:
            self.triples.remove(t)
            self._version += 1
            return True
        return False

    def build_all_indexes(self):
        """The reference builds its six hash indexes here (sparql_database.rs:3364-3394); the device store is (re)uploaded and
        partitioned by predicate (kb_store_build_index)."""
        self._sync()
        self.ctx.build_index()

    def _sync(self):
        if self._uploaded_version == self._version:
            return
        arr = np.array(sorted(self.triples), dtype=np.uint32).reshape(-1, 3)  # (s,p,o) order: the reference's BTreeSet iteration
        self.ctx.store_load(arr[:, 0], arr[:, 1], arr[:, 2])
        num, isn = self.dictionary.numeric_table()
        self.ctx.dict_numeric_load(num, isn)
        self.ctx.dict_strings_load(self.dictionary.id_to_string)  # the final id -> string step runs on the device too
        self._uploaded_version = self._version


class ExecutionEngine:
    """engine.rs:24-143 — static functions on a unit struct."""

    @staticmethod
    def _run(op, db: SparqlDatabase, slots: SlotMap) -> c.Relation:
        ctx = db.ctx
        if isinstance(op, (TableScan, IndexScan)):
            if any(isinstance(t, QuotedTriple) for t in op.pattern):
                return ExecutionEngine._scan_quoted(op.pattern, db, slots)
            return ctx.scan([slots.pattern(op.pattern)])[0]
        if isinstance(op, Filter):
            # Selection directly over a star / scan: let the fused operator push conjuncts into its scans
            if isinstance(op.input, StarJoin):
                return ctx.star_join(slots.of(op.input.join_var), [slots.pattern(p) for p in op.input.patterns],
                                     op.condition.compile(slots, db.dictionary))
            rel = ExecutionEngine._run(op.input, db, slots)
            return ctx.filter(rel, op.condition.compile(slots, db.dictionary))
        if isinstance(op, Projection):
            rel = ExecutionEngine._run(op.input, db, slots)
            return ctx.project(rel, [slots.of(v) for v in op.variables])
        if isinstance(op, (HashJoin, OptimizedHashJoin, NestedLoopJoin)):
            return ctx.hash_join(ExecutionEngine._run(op.left, db, slots), ExecutionEngine._run(op.right, db, slots))
        if isinstance(op, ParallelJoin):
            # right side a scan -> bind join (engine.rs:935-937) == natural join of left with the pattern's matches: kb_bind_join looks the
            # pattern up in the index's persistent tables when it can, and scans + joins otherwise
            if isinstance(op.right, (TableScan, IndexScan)) and not any(isinstance(t, QuotedTriple) for t in op.right.pattern) and hasattr(ctx, "bind_join"):
                return ctx.bind_join(ExecutionEngine._run(op.left, db, slots), slots.pattern(op.right.pattern))
            return ctx.hash_join(ExecutionEngine._run(op.left, db, slots), ExecutionEngine._run(op.right, db, slots))
        if isinstance(op, StarJoin):
            return ctx.star_join(slots.of(op.join_var), [slots.pattern(p) for p in op.patterns])
        if isinstance(op, InMemoryBuffer):
            names = sorted({k for row in op.content for k in row})
            cols = [np.array([row[k] for row in op.content], dtype=np.uint32) for k in names]
            return ctx.rel_from_host([slots.of(k) for k in names], cols)
        raise c.KolibrieError(c.KB_E_UNSUPPORTED, f"operator {type(op).__name__} stays on the reference's CPU path")

    @staticmethod
    def _scan_quoted(pattern: TriplePattern, db: SparqlDatabase, slots: SlotMap):
        """resolve_quoted_triple_scan (engine.rs:1111-1188). The quoted-triple store is a host structure in the reference too: its
        entries are matched against the `<< s p o >>` term(s) on the host (match_term, engine.rs:1088-1107: constants by id, variables
        bound on first use and compared afterwards, a nested quoted term only requires a quoted-triple id), which yields one row
        (quoted id, inner bindings) per matching entry. The reference then runs one index scan per entry with the id substituted and
        merges the bindings, dropping conflicts; relationally that is the natural join of those rows with ONE scan of the pattern whose
        quoted position is a fresh variable — which is what runs on the device (kb_scan + kb_rel_from_host + kb_hash_join)."""
        ctx = db.ctx
        qt_s, qt_o = isinstance(pattern[0], QuotedTriple), isinstance(pattern[2], QuotedTriple)
        if isinstance(pattern[1], QuotedTriple):
            raise c.KolibrieError(c.KB_E_UNSUPPORTED, "quoted triple in predicate position")
        rows: List[Dict[str, int]] = []
        for qid, comp in db.quoted_triple_store.id_to_components.items():
            b: Dict[str, int] = {}

            def match(term, value) -> bool:
                if isinstance(term, Constant):
                    return term.id == value
                if isinstance(term, Variable):
                    name = _strip(term.name)
                    if name in b:
                        return b[name] == value
                    b[name] = value
                    return True
                return is_quoted_triple_id(value)  # nested quoted pattern: engine.rs:1100-1105

            ok = True
            for term in ([pattern[0]] if qt_s else []) + ([pattern[2]] if qt_o else []):  # the SAME entry serves both positions (engine.rs:1135-1150)
                ok = ok and match(term.s, comp[0]) and match(term.p, comp[1]) and match(term.o, comp[2])
            if ok:
                b["__qt"] = qid
                rows.append(b)
        qslot = slots.of("__qt_%d" % len(slots.names))
        names = sorted({k for r in rows for k in r if k != "__qt"})
        # the outer pattern with the quoted position(s) replaced by the fresh variable
        outer = (Variable(slots.names[qslot]) if qt_s else pattern[0], pattern[1], Variable(slots.names[qslot]) if qt_o else pattern[2])
        outer_rel = ctx.scan([slots.pattern(outer)])[0]
        cols = [np.array([r["__qt"] for r in rows], dtype=np.uint32)] + [np.array([r[k] for r in rows], dtype=np.uint32) for k in names]
        inner_rel = ctx.rel_from_host([qslot] + [slots.of(k) for k in names], cols)
        joined = ctx.hash_join(inner_rel, outer_rel)
        _, jslots = joined.info()
        return ctx.project(joined, [s_ for s_ in jslots if s_ != qslot])

    @staticmethod
    def execute_with_ids(op, db: SparqlDatabase) -> List[Dict[str, int]]:
        db._sync()
        slots = SlotMap()
        rel = ExecutionEngine._run(op, db, slots)
        n, rslots = rel.info()
        cols = [rel.column(i) for i in range(len(rslots))]
        names = [slots.names[s] for s in rslots]
        return [{names[j]: int(cols[j][i]) for j in range(len(names))} for i in range(n)]

    @staticmethod
    def execute(op, db: SparqlDatabase) -> List[Dict[str, str]]:
        # engine.rs:33-49: ids are turned into strings only at the very end — here by kb_rel_decode, column by column, on the device
        db._sync()
        slots = SlotMap()
        rel = ExecutionEngine._run(op, db, slots)
        n, rslots = rel.info()
        cols = []
        for i in range(len(rslots)):
            try:
                cols.append(rel.decode_strings(i))
            except c.KolibrieError as e:  # quoted-triple ids (bit 31) are decoded on the host, recursively (dictionary.rs:58-68)
                if e.status != c.KB_E_UNSUPPORTED:
                    raise
                cols.append([db.decode_term(int(x)) or "unknown" for x in rel.column(i)])
        names = [slots.names[s] for s in rslots]
        return [{names[j]: cols[j][i] for j in range(len(names))} for i in range(n)]


# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class FilterCondition:  # shared/src/rule.rs:14-18
    variable: str
    operator: str
    value: str


@dataclass
class Rule:  # shared/src/rule.rs:21-25
    premise: List[TriplePattern]
    conclusion: List[TriplePattern]
    filters: List[FilterCondition] = field(default_factory=list)


_RULE_CMP = {">": c.CMP_GT, ">=": c.CMP_GE, "<": c.CMP_LT, "<=": c.CMP_LE, "=": c.CMP_EQ, "!=": c.CMP_NE}


def compile_rule(rule: Rule) -> dict:
    slots = SlotMap()
    prem = [slots.pattern(p) for p in rule.premise]
    conc = [slots.pattern(p) for p in rule.conclusion]
    fl = []
    for f in rule.filters:
        cmp = _RULE_CMP.get(f.operator, 0)  # e.g. "OR:>" is a no-op in the reference (rules.rs:133-165)
        name = _strip(f.variable)
        if name not in slots.slot:
            continue  # unbound lhs: the reference skips the filter (rules.rs:139)
        rhs_name = _strip(f.value)
        if rhs_name in slots.slot:
            fl.append(c.KbRuleFilter(slots.slot[name], cmp, 1, slots.slot[rhs_name], 0.0))
        else:
            v = rust_parse_f64(f.value)
            fl.append(c.KbRuleFilter(slots.slot[name], cmp, 0, 0, 0.0 if v is None else v))
    return {"premise": prem, "conclusion": conc, "filters": fl}


class Reasoner:
    """datalog/src/reasoning.rs:31-100. Facts live in the device store; rules are compiled to slot form per call."""

    def __init__(self, ctx: Optional[c.Context] = None, device: int = 0):
        self.dictionary = Dictionary()
        self.rules: List[Rule] = []
        self.ctx = ctx or c.Context(device)
        self._facts: List[Tuple[int, int, int]] = []
        self._fact_set: set = set()
        self._dirty = True
        self.last_stats = None
        self._n_in_store: Optional[int] = None  # facts of self._facts the device store holds, once a fixpoint has closed it
        self._closed_rules = -1                 # len(self.rules) at that fixpoint

    def add_abox_triple(self, subject: str, predicate: str, obj: str):
        t = (self.dictionary.encode(subject), self.dictionary.encode(predicate), self.dictionary.encode(obj))
        if t not in self._fact_set:  # index_manager.insert dedups (index_manager.rs:41-57)
            self._fact_set.add(t)
            self._facts.append(t)
            self._dirty = True

    def add_rule(self, rule: Rule):
        self.rules.append(rule)

    def _sync(self):
        if self._dirty:
            arr = np.array(self._facts, dtype=np.uint32).reshape(-1, 3)
            self.ctx.store_load(arr[:, 0], arr[:, 1], arr[:, 2])
            self._dirty = False
            if self._n_in_store is not None and self._n_in_store != len(self._facts):
                self._n_in_store = None  # the reloaded store holds triples no fixpoint has seen: it is not a closed store plus a seed
        num, isn = self.dictionary.numeric_table()
        self.ctx.dict_numeric_load(num, isn)

    def _infer(self, strategy: int) -> List[Tuple[int, int, int]]:
        self._sync()
        rel, st = self.ctx.datalog_fixpoint([compile_rule(r) for r in self.rules], strategy)
        self.last_stats = st
        rows = rel.to_numpy([0, 1, 2])
        out = [tuple(int(x) for x in r) for r in rows]
        for t in out:  # the device already appended them to its store (infer_generic.rs:46)
            self._fact_set.add(t)
            self._facts.append(t)
        self._n_in_store, self._closed_rules = len(self._facts), len(self.rules)
        return out

    def infer_new_facts_incremental(self, strategy: int = c.SEMI_NAIVE):
        """`add_abox_triple` after a fixpoint, then infer again WITHOUT starting over (the reference starts over: semi_naive.rs:89 sets
        the delta to every fact): the triples added since the last fixpoint are the seed of kb_datalog_fixpoint_seed — they alone are
        the first delta, the closed store is OLD. Returns the facts inferred from them, the same set a fresh
        infer_new_facts_semi_naive() would add. Falls back to the full fixpoint when there is no closed store to extend."""
        if self._n_in_store is None or self._closed_rules != len(self.rules) or self._n_in_store > len(self._facts):
            return self._infer(strategy)
        rules = [compile_rule(r) for r in self.rules]
        rule_preds = {int(x.p.value) for r in rules for x in list(r["premise"]) + list(r["conclusion"]) if not x.p.is_var}
        added = np.array(self._facts[self._n_in_store:], dtype=np.uint32).reshape(-1, 3)
        num, isn = self.dictionary.numeric_table()
        self.ctx.dict_numeric_load(num, isn)
        in_rules = np.isin(added[:, 1], np.fromiter(rule_preds, np.uint32, len(rule_preds))) if len(added) else np.zeros(0, bool)
        other = added[~in_rules]
        if len(other):  # predicates no rule mentions: plain store rows
            self.ctx.store_append(other[:, 0], other[:, 1], other[:, 2], 0)
        sd = added[in_rules]
        seed = self.ctx.rel_from_host([0, 1, 2], [np.ascontiguousarray(sd[:, k]) for k in range(3)])
        rel, n_new, st = self.ctx.datalog_fixpoint_seed(rules, seed, strategy)
        self.last_stats = st
        rows = rel.to_numpy([0, 1, 2])[n_new:]
        out = [tuple(int(x) for x in r) for r in rows]
        for t in out:
            self._fact_set.add(t)
            self._facts.append(t)
        self._n_in_store, self._dirty = len(self._facts), False
        return out

    def infer_new_facts_semi_naive(self):
        return self._infer(c.SEMI_NAIVE)

    def infer_new_facts_naive(self):
        return self._infer(c.NAIVE)

    def infer_new_facts_semi_naive_parallel(self):
        return self._infer(c.SEMI_NAIVE_PARALLEL)  # semi_naive_parallel.rs:11: 1-2 premise rules, no filters, constants enforced

    def infer_new_facts(self):
        return self.infer_new_facts_naive()  # my_naive.rs:78-80 "for backward compatibility"

    def query_abox(self, subject: Optional[str], predicate: Optional[str], obj: Optional[str]):
        """reasoning.rs:79-93 — note: like the reference, querying ENCODES unknown strings (they simply match nothing)."""
        self._sync()
        ts = []
        for k, v in enumerate((subject, predicate, obj)):
            ts.append(c.V(k) if v is None else c.K(self.dictionary.encode(v)))
        rel = self.ctx.scan([c.pattern(*ts)])[0]
        n, slots = rel.info()
        cols = {s: rel.column(i) for i, s in enumerate(slots)}
        fixed = [None if v is None else self.dictionary.encode(v) for v in (subject, predicate, obj)]
        return [tuple(int(cols[k][i]) if fixed[k] is None else fixed[k] for k in range(3)) for i in range(n)]


# ---------------------------------------------------------------------------------------------------------------------
# The LEGACY executor's BGP + FILTER stage (kolibrie/src/execute_query.rs:151-341 — the path `execute_query` takes: CLI, criterion bench,
# README): patterns are joined one by one by perform_join_par_simd_with_strict_filter_1 (sparql_database.rs:2056-2213) over rows of
# STRING bindings, then apply_filters_simd (sparql_database.rs:1381-1669). Its quirks, kept here because they decide the answer:
#   * a pattern's predicate is compared as a string: a variable predicate ("?p") equals no predicate, the pattern matches nothing;
#   * a constant SUBJECT is not enforced: the constant's text is used as a binding NAME, so it behaves like a variable (that joins with
#     other occurrences of the same constant); a constant OBJECT is enforced (literal_filter) and bound under its own text as well;
#   * FILTER comparison at the top level: i32 comparison when the bound term and the constant both parse as i32, byte-wise string
#     = / != otherwise; the same comparison nested under AND / OR / NOT parses f64 instead (evaluate_filter_expression);
#   * an operand containing + - * / goes to the reference's arithmetic-expression parser: not evaluated on the device (UNSUPPORTED).
_I32 = re.compile(r"[+-]?[0-9]+")


def rust_parse_i32(s: str) -> Optional[int]:
    """`str::parse::<i32>()` acceptance: optional sign, ASCII digits, in range; nothing else"""
    if not isinstance(s, str) or not _I32.fullmatch(s):
        return None
    v = int(s)
    return v if -(1 << 31) <= v < (1 << 31) else None


def _has_arith(s: str) -> bool:
    return any(ch in s for ch in "+-*/")


class LegacyExecutor:
    def __init__(self, db: "SparqlDatabase"):
        self.db = db
        self._i32_loaded = -1

    def _sync(self):
        db = self.db
        db._sync()
        n = len(db.dictionary.id_to_string)
        if n != self._i32_loaded and hasattr(db.ctx, "dict_legacy_i32_load"):
            val = np.zeros(n, dtype=np.int32)
            isi = np.zeros(n, dtype=np.uint8)
            for i, st in enumerate(db.dictionary.id_to_string):
                v = rust_parse_i32(st)
                if v is not None:
                    val[i] = v
                    isi[i] = 1
            db.ctx.dict_legacy_i32_load(val, isi)
            self._i32_loaded = n

    def _compile_filter(self, e, slots: SlotMap, nested: bool, ops: List[c.KbFilterOp]):
        d = self.db.dictionary
        if isinstance(e, Comparison):
            if _has_arith(e.var) or _has_arith(e.value):
                raise c.KolibrieError(c.KB_E_UNSUPPORTED, "operands with + - * / go through the reference's arithmetic-expression parser")
            if not e.var.startswith("?") or e.value.startswith("?"):
                raise c.KolibrieError(c.KB_E_UNSUPPORTED, "legacy FILTER: variable <op> constant only")
            cmpc = {">": c.CMP_GT, ">=": c.CMP_GE, "<": c.CMP_LT, "<=": c.CMP_LE, "=": c.CMP_EQ, "!=": c.CMP_NE}.get(e.op)
            if cmpc is None:
                raise c.KolibrieError(c.KB_E_UNSUPPORTED, f"operator {e.op!r}: the reference evaluates it to false")
            name = _strip(e.var)
            if name not in slots.slot:
                raise c.KolibrieError(c.KB_E_UNSUPPORTED, "FILTER on an unbound variable: the reference evaluates it to false")
            lit = d.lookup(e.value)
            num = rust_parse_f64(e.value) if nested else rust_parse_i32(e.value)
            flags = (c.LEGACY_CONST_IS_I32 if num is not None else 0) | (c.LEGACY_NESTED_F64 if nested else 0)
            ops.append(c.fop(c.F_CMP_LEGACY, slot=slots.slot[name], cmp=cmpc | flags, id=c.KB_ID_NONE if lit is None else lit, value=float(num) if num is not None else 0.0))
        elif isinstance(e, And):
            self._compile_filter(e.left, slots, True, ops); self._compile_filter(e.right, slots, True, ops); ops.append(c.fop(c.F_AND))
        elif isinstance(e, Or):
            self._compile_filter(e.left, slots, True, ops); self._compile_filter(e.right, slots, True, ops); ops.append(c.fop(c.F_OR))
        elif isinstance(e, Not):
            self._compile_filter(e.inner, slots, True, ops); ops.append(c.fop(c.F_NOT))
        elif isinstance(e, FunctionCall) and e.name == "isTRIPLE" and e.args and e.args[0].startswith("?") and _strip(e.args[0]) in slots.slot:
            ops.append(c.fop(c.F_IS_TRIPLE, slot=slots.slot[_strip(e.args[0])]))
        else:
            raise c.KolibrieError(c.KB_E_UNSUPPORTED, f"legacy FILTER construct {type(e).__name__}")

    def execute_bgp(self, patterns: Sequence[Tuple[str, str, str]], filters: Sequence[object] = (), select: Optional[Sequence[str]] = None) -> List[Dict[str, str]]:
        """patterns: (subject, predicate, object) as RESOLVED strings, variables start with '?'. Returns rows of string bindings for the
        variables (pseudo-bindings named after constants are dropped unless selected)."""
        db = self.db
        self._sync()
        d = db.dictionary
        slots = SlotMap()
        pats, eqs = [], []
        for s_, p_, o_ in patterns:
            if p_.startswith("?"):
                return []  # the predicate string "?p" equals no predicate (sparql_database.rs:2136)
            pid = d.lookup(p_)
            if pid is None:
                return []
            # a constant subject / object is a binding NAME in the legacy join: one slot per distinct text
            st = c.V(slots.of(s_ if s_.startswith("?") else "?\x00" + s_))  # "\x00<text>": a name no variable can have
            ot = c.V(slots.of(o_ if o_.startswith("?") else "?\x00" + o_))
            if not o_.startswith("?"):  # literal_filter: the object string must equal the constant
                oid = d.lookup(o_)
                if oid is None:
                    return []
                eqs.append(c.fop(c.F_EQ_ID, slot=ot.value, id=oid))
            pats.append(c.pattern(st, c.K(pid), ot))
        ops: List[c.KbFilterOp] = []
        for i, e in enumerate(eqs):
            ops.append(e)
            if i:
                ops.append(c.fop(c.F_AND))
        for f in filters:  # filters.iter().all(...): a conjunction of top-level expressions
            had = bool(ops)
            self._compile_filter(f, slots, False, ops)
            if had:
                ops.append(c.fop(c.F_AND))
        rel = db.ctx.bgp_execute(pats, ops or None)
        n, rslots = rel.info()
        names = [slots.names[s_] for s_ in rslots]
        keep = [j for j, nm in enumerate(names) if (select is None and any(nm == _strip(t) for pt in patterns for t in (pt[0], pt[2]) if t.startswith("?")))
                or (select is not None and ("?" + nm) in select)]
        cols = {j: rel.column(j) for j in keep}
        return [{"?" + names[j]: db.decode_term(int(cols[j][i])) for j in keep} for i in range(n)]
