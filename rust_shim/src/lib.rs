//! kolibrie-gpu-b200 — the three seams through which Kolibrie hands its hot path to libkolibrie_b200.so:
//!
//! * [`try_execute`]            for `ExecutionEngine::execute_with_ids` (kolibrie/src/streamertail_optimizer/execution/engine.rs:54)
//! * [`GpuSemiNaiveStrategy`]   for `Reasoner::infer_with_strategy`    (datalog/src/reasoning/materialisation/infer_generic.rs:27-53)
//! * [`GpuR2R`]                 for the RSP engine's `R2ROperator`      (kolibrie/src/rsp/r2r.rs:17-30)
//!
//! plus [`planner`]: the device operators' cost constants for Streamertail's estimator (cost/estimator.rs:17-29).
//!
//! STATUS: written against the reference's types (commit 1d7c306c) and the C header of this repository; NOT compiled or tested in the
//! repository's build image, which has no Rust toolchain. The Python (`kolibrie_b200/engine.py`, `planner.py`) and C++
//! (`kolibrie_b200/host/kolibrie_host.hpp`) mirrors implement the same logic and ARE tested against the same C ABI.
pub mod ffi;
pub mod planner;
pub mod r2r;
pub mod reasoner;

use ffi::*;
use kolibrie::sparql_database::SparqlDatabase;
use kolibrie::streamertail_optimizer::{Condition, PhysicalOperator};
use shared::query::{ArithmeticExpression, FilterExpression};
use shared::terms::{Term, TriplePattern};
use std::collections::HashMap;
use std::ffi::CStr;
use std::ptr;

/// One device context + the store version it mirrors. Kolibrie holds `&mut SparqlDatabase` around every call, so no locking here.
pub struct DeviceStore {
    pub ctx: *mut KbCtx,
    synced_triples: usize,
    synced_dict: usize,
    indexed: bool,
}
unsafe impl Send for DeviceStore {}

impl DeviceStore {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut ctx: *mut KbCtx = ptr::null_mut();
        let rc = unsafe { kb_ctx_create(device, &mut ctx) };
        if rc != KB_OK {
            return Err(unsafe { CStr::from_ptr(kb_last_error(ptr::null())) }.to_string_lossy().into_owned());
        }
        Ok(Self { ctx, synced_triples: usize::MAX, synced_dict: usize::MAX, indexed: false })
    }

    pub fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(kb_last_error(self.ctx)) }.to_string_lossy().into_owned()
    }

    /// Mirror `database.triples` (BTreeSet iteration = (s,p,o) order) and the numeric side table on the device when they changed.
    /// The reference has no store version counter; `len()` of the set and of the dictionary is the cheap change signal used here —
    /// a caller that mutates without changing the sizes must call `invalidate()`.
    pub fn sync(&mut self, database: &SparqlDatabase) -> Result<(), String> {
        let dict = database.dictionary.read().unwrap();
        if dict.id_to_string.len() != self.synced_dict {
            let n = dict.id_to_string.len();
            let mut num = vec![0f64; n];
            let mut isn = vec![0u8; n];
            for (i, s) in dict.id_to_string.iter().enumerate() {
                if let Ok(v) = s.parse::<f64>() { num[i] = v; isn[i] = 1; }  // types.rs:133-148 / 163-167
            }
            if unsafe { kb_dict_numeric_load(self.ctx, num.as_ptr(), isn.as_ptr(), n as u32) } != KB_OK { return Err(self.last_error()); }
            let mut offs = Vec::with_capacity(n + 1);
            let mut bytes = Vec::new();
            offs.push(0u64);
            for s in dict.id_to_string.iter() { bytes.extend_from_slice(s.as_bytes()); offs.push(bytes.len() as u64); }
            if unsafe { kb_dict_strings_load(self.ctx, offs.as_ptr(), bytes.as_ptr(), n as u32) } != KB_OK { return Err(self.last_error()); }
            self.synced_dict = n;
        }
        if database.triples.len() != self.synced_triples {
            let n = database.triples.len();
            let (mut s, mut p, mut o) = (Vec::with_capacity(n), Vec::with_capacity(n), Vec::with_capacity(n));
            for t in database.triples.iter() { s.push(t.subject); p.push(t.predicate); o.push(t.object); }
            if unsafe { kb_store_load(self.ctx, s.as_ptr(), p.as_ptr(), o.as_ptr(), n as u64) } != KB_OK { return Err(self.last_error()); }
            self.synced_triples = n;
            self.indexed = false;
        }
        Ok(())
    }

    /// `SparqlDatabase::build_all_indexes` (sparql_database.rs:3364-3394) on the device; appends / evictions maintain it afterwards.
    pub fn build_all_indexes(&mut self) -> Result<(), String> {
        if unsafe { kb_store_build_index(self.ctx, ptr::null_mut(), ptr::null_mut()) } != KB_OK { return Err(self.last_error()); }
        self.indexed = true;
        Ok(())
    }

    pub fn invalidate(&mut self) { self.synced_triples = usize::MAX; self.synced_dict = usize::MAX; }

    /// The per-term `Dictionary::encode` loop of a bulk load (sparql_database.rs:1000-1013) as one device call. Returns the ids of
    /// `terms` (those the sequential loop would hand out, dictionary.rs:32-48) and, for every NEW string in id order, the index of the
    /// term that introduced it: the caller replays `dict.encode(terms[p])` for those positions only. The device dictionary must hold
    /// the host dictionary's strings (`sync` uploads them; afterwards the two grow in lock step through this call).
    pub fn encode_bulk(&mut self, terms: &[&str]) -> Result<(Vec<u32>, Vec<u64>), String> {
        let n = terms.len();
        let mut offs = Vec::with_capacity(n + 1);
        let mut bytes = Vec::new();
        offs.push(0u64);
        for t in terms { bytes.extend_from_slice(t.as_bytes()); offs.push(bytes.len() as u64); }
        let mut ids = vec![0u32; n];
        let mut first = vec![0u64; n];
        let mut n_new = 0u32;
        let rc = unsafe { kb_dict_encode(self.ctx, offs.as_ptr(), bytes.as_ptr(), n as u64, ids.as_mut_ptr(), &mut n_new, first.as_mut_ptr()) };
        if rc != KB_OK { return Err(self.last_error()); }
        first.truncate(n_new as usize);
        if self.synced_dict != usize::MAX { self.synced_dict += n_new as usize; }
        Ok((ids, first))
    }
}

impl Drop for DeviceStore {
    fn drop(&mut self) { unsafe { kb_ctx_destroy(self.ctx) } }
}

/// variable name (without '?', as engine.rs strips it everywhere) <-> slot
#[derive(Default)]
pub struct SlotMap { pub slot: HashMap<String, u32>, pub names: Vec<String> }
impl SlotMap {
    pub fn of(&mut self, name: &str) -> u32 {
        let name = name.strip_prefix('?').unwrap_or(name);
        if let Some(&s) = self.slot.get(name) { return s; }
        let s = self.names.len() as u32;
        self.slot.insert(name.to_string(), s);
        self.names.push(name.to_string());
        s
    }
    pub fn term(&mut self, t: &Term) -> Option<KbTerm> {
        match t {
            Term::Variable(v) => Some(KbTerm { is_var: 1, value: self.of(v) }),
            Term::Constant(c) => Some(KbTerm { is_var: 0, value: *c }),
            Term::QuotedTriple(_) => None,  // resolved on the host first (engine.rs:1111-1188); see kolibrie_b200/engine.py::_scan_quoted
        }
    }
    pub fn pattern(&mut self, p: &TriplePattern) -> Option<KbPattern> {
        Some(KbPattern { s: self.term(&p.0)?, p: self.term(&p.1)?, o: self.term(&p.2)? })
    }
}

/// `Condition` -> postfix filter program with the semantics of `evaluate_with_ids` (types.rs:110-186). None: a construct the device
/// does not evaluate (the caller falls back to the CPU operator).
pub fn compile_condition(cond: &Condition, slots: &mut SlotMap, database: &SparqlDatabase) -> Option<Vec<KbFilterOp>> {
    fn op(op: u32, slot: u32, cmp: u32, id: u32, value: f64) -> KbFilterOp { KbFilterOp { op, slot, cmp, id, value } }
    fn arith(e: &ArithmeticExpression, slots: &mut SlotMap, out: &mut Vec<KbFilterOp>) -> Option<()> {
        match e {
            ArithmeticExpression::Operand(s) => {
                if s.starts_with('?') { out.push(op(KB_F_PUSH_VAR, slots.of(s), 0, 0, 0.0)); }
                else { out.push(op(KB_F_PUSH_CONST, 0, 0, 0, s.parse::<f64>().ok()?)); }
            }
            ArithmeticExpression::Add(l, r) => { arith(l, slots, out)?; arith(r, slots, out)?; out.push(op(KB_F_ADD, 0, 0, 0, 0.0)); }
            ArithmeticExpression::Subtract(l, r) => { arith(l, slots, out)?; arith(r, slots, out)?; out.push(op(KB_F_SUB, 0, 0, 0, 0.0)); }
            ArithmeticExpression::Multiply(l, r) => { arith(l, slots, out)?; arith(r, slots, out)?; out.push(op(KB_F_MUL, 0, 0, 0, 0.0)); }
            ArithmeticExpression::Divide(l, r) => { arith(l, slots, out)?; arith(r, slots, out)?; out.push(op(KB_F_DIV, 0, 0, 0, 0.0)); }
        }
        Some(())
    }
    fn rec(e: &FilterExpression, slots: &mut SlotMap, db: &SparqlDatabase, out: &mut Vec<KbFilterOp>) -> Option<()> {
        match e {
            FilterExpression::Comparison(var, o, value) => {
                let slot = slots.of(var);
                match *o {
                    "=" | "!=" => {
                        // decoded == literal  <=>  id == lookup(literal); a literal the dictionary does not hold equals nothing
                        let id = db.dictionary.read().unwrap().string_to_id.get(*value).copied().unwrap_or(KB_ID_NONE);
                        out.push(op(if *o == "=" { KB_F_EQ_ID } else { KB_F_NE_ID }, slot, 0, id, 0.0));
                    }
                    ">" | ">=" | "<" | "<=" => {
                        let cmp = match *o { ">" => KB_CMP_GT, ">=" => KB_CMP_GE, "<" => KB_CMP_LT, _ => KB_CMP_LE };
                        out.push(op(KB_F_CMP_NUM, slot, cmp, 0, value.parse::<f64>().unwrap_or(0.0)));
                    }
                    _ => return None,
                }
            }
            FilterExpression::And(l, r) => { rec(l, slots, db, out)?; rec(r, slots, db, out)?; out.push(op(KB_F_AND, 0, 0, 0, 0.0)); }
            FilterExpression::Or(l, r) => { rec(l, slots, db, out)?; rec(r, slots, db, out)?; out.push(op(KB_F_OR, 0, 0, 0, 0.0)); }
            FilterExpression::Not(i) => { rec(i, slots, db, out)?; out.push(op(KB_F_NOT, 0, 0, 0, 0.0)); }
            FilterExpression::ArithmeticExpr(a) => { arith(a, slots, out)?; out.push(op(KB_F_TRUTHY, 0, 0, 0, 0.0)); }
            FilterExpression::FunctionCall(name, args) => {
                if *name == "isTRIPLE" && !args.is_empty() { out.push(op(KB_F_IS_TRIPLE, slots.of(args[0]), 0, 0, 0.0)); } else { return None; }
            }
        }
        Some(())
    }
    let mut out = Vec::new();
    rec(&cond.expression, slots, database, &mut out)?;
    Some(out)
}

struct Rel(*mut KbRel, *mut KbCtx);
impl Drop for Rel { fn drop(&mut self) { unsafe { kb_rel_free(self.1, self.0) } } }

fn run(op: &PhysicalOperator, dev: &DeviceStore, db: &SparqlDatabase, slots: &mut SlotMap) -> Option<Rel> {
    let ctx = dev.ctx;
    let mut out: *mut KbRel = ptr::null_mut();
    let ok = |rc: kb_status, out: *mut KbRel| if rc == KB_OK { Some(Rel(out, ctx)) } else { None };
    match op {
        PhysicalOperator::TableScan { pattern } | PhysicalOperator::IndexScan { pattern } => {
            let p = slots.pattern(pattern)?;
            ok(unsafe { kb_scan(ctx, &p, 1, ptr::null(), ptr::null(), &mut out) }, out)
        }
        PhysicalOperator::Filter { input, condition } => {
            let prog = compile_condition(condition, slots, db)?;
            if let PhysicalOperator::StarJoin { join_var, patterns } = input.as_ref() {
                // Selection over a star: the fused operator pushes the conjuncts into its scans / probe
                let pats: Option<Vec<KbPattern>> = patterns.iter().map(|p| slots.pattern(p)).collect();
                let pats = pats?;
                let js = slots.of(join_var);
                return ok(unsafe { kb_star_join(ctx, js, pats.as_ptr(), pats.len() as u32, prog.as_ptr(), prog.len() as u32, &mut out) }, out);
            }
            let r = run(input, dev, db, slots)?;
            ok(unsafe { kb_filter(ctx, r.0, prog.as_ptr(), prog.len() as u32, &mut out) }, out)
        }
        PhysicalOperator::Projection { input, variables } => {
            let r = run(input, dev, db, slots)?;
            let sl: Vec<u32> = variables.iter().map(|v| slots.of(v)).collect();
            ok(unsafe { kb_project(ctx, r.0, sl.as_ptr(), sl.len() as u32, &mut out) }, out)
        }
        PhysicalOperator::HashJoin { left, right } | PhysicalOperator::OptimizedHashJoin { left, right } | PhysicalOperator::NestedLoopJoin { left, right } => {
            let (l, r) = (run(left, dev, db, slots)?, run(right, dev, db, slots)?);
            ok(unsafe { kb_hash_join(ctx, l.0, r.0, &mut out) }, out)
        }
        PhysicalOperator::ParallelJoin { left, right } => {
            let l = run(left, dev, db, slots)?;
            match right.as_ref() {  // a scan on the right is a bind join (engine.rs:935-937): the index's persistent table answers it
                PhysicalOperator::TableScan { pattern } | PhysicalOperator::IndexScan { pattern } => {
                    let p = slots.pattern(pattern)?;
                    ok(unsafe { kb_bind_join(ctx, l.0, &p, &mut out) }, out)
                }
                _ => { let r = run(right, dev, db, slots)?; ok(unsafe { kb_hash_join(ctx, l.0, r.0, &mut out) }, out) }
            }
        }
        PhysicalOperator::StarJoin { join_var, patterns } => {
            let pats: Option<Vec<KbPattern>> = patterns.iter().map(|p| slots.pattern(p)).collect();
            let pats = pats?;
            let js = slots.of(join_var);
            ok(unsafe { kb_star_join(ctx, js, pats.as_ptr(), pats.len() as u32, ptr::null(), 0, &mut out) }, out)
        }
        PhysicalOperator::InMemoryBuffer { content, .. } => {
            let mut names: Vec<String> = content.iter().flat_map(|r| r.keys().cloned()).collect();
            names.sort(); names.dedup();
            let cols: Vec<Vec<u32>> = names.iter().map(|k| content.iter().map(|r| *r.get(k).unwrap_or(&KB_ID_NONE)).collect()).collect();
            let ptrs: Vec<*const u32> = cols.iter().map(|c| c.as_ptr()).collect();
            let sl: Vec<u32> = names.iter().map(|k| slots.of(k)).collect();
            ok(unsafe { kb_rel_from_host(ctx, sl.as_ptr(), sl.len() as u32, ptrs.as_ptr(), content.len() as u64, &mut out) }, out)
        }
        _ => None,  // Bind, Values, MLPredict, Subquery: the reference's CPU path
    }
}

/// The hook for `ExecutionEngine::execute_with_ids`:
/// ```ignore
/// #[cfg(feature = "cuda")]
/// if shared::GPU_MODE_ENABLED.load(Ordering::Relaxed) {
///     if let Some(rows) = kolibrie_gpu_b200::try_execute(operator, database, &mut DEVICE.lock().unwrap()) { return rows; }
/// }
/// ```
/// None means "shape not handled on the device" (KB_E_UNSUPPORTED or an operator outside the hot path): run the CPU path.
pub fn try_execute(op: &PhysicalOperator, database: &SparqlDatabase, dev: &mut DeviceStore) -> Option<Vec<HashMap<String, u32>>> {
    dev.sync(database).ok()?;
    let mut slots = SlotMap::default();
    let rel = run(op, dev, database, &mut slots)?;
    let (mut n, mut nc) = (0u64, 0u32);
    let mut rs = [0u32; KB_MAX_COLS];
    if unsafe { kb_rel_info(rel.0, &mut n, &mut nc, rs.as_mut_ptr()) } != KB_OK { return None; }
    let mut cols: Vec<Vec<u32>> = Vec::with_capacity(nc as usize);
    for c in 0..nc {
        let mut v = vec![0u32; n as usize];
        if unsafe { kb_rel_download(dev.ctx, rel.0, c, v.as_mut_ptr()) } != KB_OK { return None; }
        cols.push(v);
    }
    let names: Vec<&String> = (0..nc as usize).map(|c| &slots.names[rs[c] as usize]).collect();
    Some((0..n as usize).map(|i| names.iter().enumerate().map(|(c, k)| ((*k).clone(), cols[c][i])).collect()).collect())
}
