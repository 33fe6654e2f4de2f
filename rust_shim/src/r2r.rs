//! `R2ROperator` over the device store: a window firing is `remove(previous window)` + `add(current window)` + `materialize()` +
//! `execute_query(plan)` (rsp_engine.rs:94-109; simple_r2r.rs:95-142). The adds of one firing become ONE store segment tagged with the
//! firing number, the removes evict the segment of the firing they were added in; the store index is maintained across both (one chunk
//! per segment), so that the per-slide query keeps the one-kernel index path.
use crate::ffi::*;
use crate::{try_execute, DeviceStore};
use kolibrie::rsp::r2r::{AsAnyMut, R2ROperator};
use kolibrie::rsp::simple_r2r::SimpleR2R;
use kolibrie::streamertail_optimizer::PhysicalOperator;
use shared::triple::Triple;
use std::any::Any;
use std::collections::HashMap;

pub struct GpuR2R {
    pub inner: SimpleR2R,       // parsing, rules, dictionary: unchanged host logic
    pub dev: DeviceStore,
    pending: Vec<Triple>,       // adds of the current firing, not yet on the device
    firing: u64,
    tag_of: HashMap<Triple, u64>,
    live: HashMap<u64, usize>,  // tag -> triples of that segment still in the window
}

impl GpuR2R {
    pub fn new(inner: SimpleR2R, device: i32) -> Result<Self, String> {
        Ok(Self { inner, dev: DeviceStore::new(device)?, pending: Vec::new(), firing: 1, tag_of: HashMap::new(), live: HashMap::new() })
    }
    fn flush(&mut self) {
        if self.pending.is_empty() { return; }
        let (s, p, o): (Vec<u32>, Vec<u32>, Vec<u32>) = (self.pending.iter().map(|t| t.subject).collect(), self.pending.iter().map(|t| t.predicate).collect(),
                                                         self.pending.iter().map(|t| t.object).collect());
        unsafe { kb_store_append(self.dev.ctx, s.as_ptr(), p.as_ptr(), o.as_ptr(), s.len() as u64, self.firing) };
        self.live.insert(self.firing, self.pending.len());
        for t in self.pending.drain(..) { self.tag_of.insert(t, self.firing); }
        self.firing += 1;
    }
}

impl AsAnyMut for GpuR2R { fn as_any_mut(&mut self) -> &mut dyn Any { self } }

impl R2ROperator<Triple, Vec<(String, String)>, Vec<(String, String)>> for GpuR2R {
    fn load_triples(&mut self, data: &str, syntax: String) -> Result<(), String> { self.inner.load_triples(data, syntax) }
    fn load_rules(&mut self, data: &str) -> Result<(), &'static str> { self.inner.load_rules(data) }
    fn add(&mut self, data: Triple) { self.inner.add(data.clone()); self.pending.push(data); }
    fn remove(&mut self, data: &Triple) {
        self.inner.remove(data);
        if let Some(tag) = self.tag_of.remove(data) {
            let left = self.live.get_mut(&tag).map(|n| { *n -= 1; *n }).unwrap_or(0);
            if left == 0 { self.live.remove(&tag); unsafe { kb_store_evict(self.dev.ctx, tag) }; }  // the whole slide left the window
        }
    }
    fn materialize(&mut self) -> Vec<Triple> {
        self.flush();
        // rules on the device: kb_datalog_fixpoint over the live window; the inferred segment is dropped before the next firing
        // (SimpleR2R re-materialises from scratch, simple_r2r.rs:103-128). Rule compilation: reasoner.rs::compile_rule.
        unsafe { kb_store_evict(self.dev.ctx, KB_TAG_INFERRED) };
        self.inner.materialize()
    }
    fn execute_query(&mut self, op: &PhysicalOperator) -> Vec<Vec<(String, String)>> {
        self.flush();
        match try_execute(op, &self.inner.item, &mut self.dev) {
            Some(rows) => {
                let dict = self.inner.item.dictionary.read().unwrap();
                rows.into_iter().map(|r| r.into_iter().map(|(k, v)| (k, dict.decode(v).unwrap_or("unknown").to_string())).collect()).collect()
            }
            None => self.inner.execute_query(op),
        }
    }
    fn parse_data(&mut self, data: &str) -> Vec<Triple> { self.inner.parse_data(data) }
}
