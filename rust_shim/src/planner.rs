//! Device operators for Streamertail's plan search: the constants `CostEstimator::estimate_cost` (cost/estimator.rs:43-191) needs to
//! price them, in the estimator's own unit (COST_PER_ROW_INDEX_SCAN = 1 ~ 0.1 us of host time). Derivation and the tested Python
//! mirror: kolibrie_b200/planner.py. Integration: add the three variants below to `PhysicalOperator` (operators/physical.rs:16-76), the
//! arms of `device_cost` to `estimate_cost`, and push the device candidates next to the CPU ones in
//! `find_best_plan_recursive` (optimizer.rs:252-294) and `build_star_join_from_patterns` (optimizer.rs:400-480); `ExecutionEngine`
//! executes them through `kolibrie_gpu_b200::try_execute` (GpuStarJoin = StarJoin, GpuHashJoin = HashJoin, GpuBindJoin = ParallelJoin
//! with a scan on the right).
pub struct GpuCostConstants;
impl GpuCostConstants {
    /// one device operator: launch + stream synchronisation + row count back (~25 us)
    pub const LAUNCH: u64 = 250;
    /// probe / index-scan rows per cost unit (probe_index_kernel: 16.7 M rows in 0.094 ms)
    pub const ROWS_PER_UNIT_PROBE: u64 = 16_000;
    /// store-scanning rows per cost unit (scan_star_kernel: 100 M triples in ~0.35 ms)
    pub const ROWS_PER_UNIT_SCAN: u64 = 3_000;
    /// build + probe rows per cost unit of a materialised join (key-grouped join)
    pub const ROWS_PER_UNIT_JOIN: u64 = 4_000;
}

pub enum DeviceOp<'a> {
    /// every pattern `(?s P ?o)` over an indexed predicate: ONE kernel, the most selective slice probes the other patterns' tables
    StarJoin { cardinalities: &'a [u64], indexed: bool, total_triples: u64 },
    /// left relation joined with one store pattern through the index's persistent table
    BindJoin { left_cost: u64, left_cardinality: u64 },
    HashJoin { left_cost: u64, right_cost: u64, left_cardinality: u64, right_cardinality: u64 },
}

pub fn device_cost(op: &DeviceOp) -> u64 {
    use GpuCostConstants as G;
    match op {
        DeviceOp::StarJoin { cardinalities, indexed, total_triples } => {
            let mut c: Vec<u64> = cardinalities.to_vec();
            c.sort();
            let k = c.len() as u64;
            if *indexed { G::LAUNCH + c[0] * k / G::ROWS_PER_UNIT_PROBE }
            else { 2 * G::LAUNCH + total_triples / G::ROWS_PER_UNIT_SCAN + c[c.len() - 1] * k / G::ROWS_PER_UNIT_PROBE }
        }
        DeviceOp::BindJoin { left_cost, left_cardinality } => left_cost + G::LAUNCH + left_cardinality / G::ROWS_PER_UNIT_PROBE,
        DeviceOp::HashJoin { left_cost, right_cost, left_cardinality, right_cardinality } =>
            left_cost + right_cost + 2 * G::LAUNCH + (left_cardinality + right_cardinality) / G::ROWS_PER_UNIT_JOIN,
    }
}
