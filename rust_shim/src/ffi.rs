//! Raw bindings of include/kolibrie_b200.h (one declaration per exported entry point; layouts checked by tests/test_abi.py on the C side).
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_double, c_int, c_void};

pub type kb_status = i32;
pub const KB_OK: kb_status = 0;
pub const KB_E_INVALID: kb_status = -1;
pub const KB_E_CUDA: kb_status = -2;
pub const KB_E_OOM: kb_status = -3;
pub const KB_E_UNSUPPORTED: kb_status = -4;
pub const KB_E_NOT_FOUND: kb_status = -5;
pub const KB_E_LIMIT: kb_status = -6;
pub const KB_ID_NONE: u32 = 0xFFFF_FFFF;
pub const KB_MAX_COLS: usize = 16;
pub const KB_TAG_INFERRED: u64 = 0xFFFF_FFFF_FFFF_FFF0;

#[repr(C)] #[derive(Clone, Copy, Debug, Default)] pub struct KbTerm { pub is_var: u32, pub value: u32 }
#[repr(C)] #[derive(Clone, Copy, Debug, Default)] pub struct KbPattern { pub s: KbTerm, pub p: KbTerm, pub o: KbTerm }
#[repr(C)] #[derive(Clone, Copy, Debug, Default)] pub struct KbFilterOp { pub op: u32, pub slot: u32, pub cmp: u32, pub id: u32, pub value: c_double }
#[repr(C)] #[derive(Clone, Copy, Debug, Default)] pub struct KbAgg { pub kind: u32, pub slot: u32 }
#[repr(C)] #[derive(Clone, Copy, Debug, Default)] pub struct KbRuleFilter { pub lhs_slot: u32, pub cmp: u32, pub rhs_is_var: u32, pub rhs_slot: u32, pub rhs_value: c_double }
#[repr(C)] pub struct KbRule { pub premise: *const KbPattern, pub n_premise: u32, pub filters: *const KbRuleFilter, pub n_filters: u32,
                              pub conclusion: *const KbPattern, pub n_conclusion: u32 }
#[repr(C)] #[derive(Clone, Copy, Debug, Default)] pub struct KbStats { pub scan_ms: c_double, pub build_ms: c_double, pub probe_ms: c_double, pub filter_ms: c_double,
    pub group_ms: c_double, pub other_ms: c_double, pub total_ms: c_double, pub scan_launches: u64, pub build_launches: u64, pub probe_launches: u64,
    pub filter_launches: u64, pub group_launches: u64, pub other_launches: u64, pub rows_scanned: u64, pub rows_built: u64, pub rows_probed: u64,
    pub rows_out: u64, pub h2d_bytes: u64, pub d2h_bytes: u64, pub kernel_launches: u64, pub fused_scan_builds: u64, pub index_joins: u64 }
#[repr(C)] pub struct KbFixpointStats { pub rounds: u32, pub inferred: u64, pub derivations: u64, pub round_new: [u64; 64], pub device_ms: c_double }

// kb_filter_opcode / kb_cmp / kb_agg_kind / kb_strategy
pub const KB_F_CMP_NUM: u32 = 1; pub const KB_F_EQ_ID: u32 = 2; pub const KB_F_NE_ID: u32 = 3; pub const KB_F_AND: u32 = 4; pub const KB_F_OR: u32 = 5;
pub const KB_F_NOT: u32 = 6; pub const KB_F_PUSH_VAR: u32 = 7; pub const KB_F_PUSH_CONST: u32 = 8; pub const KB_F_ADD: u32 = 9; pub const KB_F_SUB: u32 = 10;
pub const KB_F_MUL: u32 = 11; pub const KB_F_DIV: u32 = 12; pub const KB_F_TRUTHY: u32 = 13; pub const KB_F_IS_TRIPLE: u32 = 14;
pub const KB_CMP_GT: u32 = 1; pub const KB_CMP_GE: u32 = 2; pub const KB_CMP_LT: u32 = 3; pub const KB_CMP_LE: u32 = 4; pub const KB_CMP_EQ: u32 = 5; pub const KB_CMP_NE: u32 = 6;
pub const KB_AGG_COUNT: u32 = 0; pub const KB_AGG_SUM: u32 = 1; pub const KB_AGG_MIN: u32 = 2; pub const KB_AGG_MAX: u32 = 3; pub const KB_AGG_AVG: u32 = 4;
pub const KB_SEMI_NAIVE: u32 = 0; pub const KB_NAIVE: u32 = 1; pub const KB_SEMI_NAIVE_PARALLEL: u32 = 2; pub const KB_SEMI_NAIVE_OLD_DELTA: u32 = 3;

pub enum KbCtx {} pub enum KbRel {} pub enum KbGroups {} pub enum KbStrings {} pub enum KbPlan {}

extern "C" {
    pub fn kb_version() -> *const c_char;
    pub fn kb_ctx_create(device: c_int, out: *mut *mut KbCtx) -> kb_status;
    pub fn kb_ctx_destroy(ctx: *mut KbCtx);
    pub fn kb_last_error(ctx: *const KbCtx) -> *const c_char;
    pub fn kb_synchronize(ctx: *mut KbCtx) -> kb_status;
    // store
    pub fn kb_store_load(ctx: *mut KbCtx, s: *const u32, p: *const u32, o: *const u32, n: u64) -> kb_status;
    pub fn kb_store_append(ctx: *mut KbCtx, s: *const u32, p: *const u32, o: *const u32, n: u64, tag: u64) -> kb_status;
    pub fn kb_store_evict(ctx: *mut KbCtx, tag: u64) -> kb_status;
    pub fn kb_store_delete(ctx: *mut KbCtx, s: *const u32, p: *const u32, o: *const u32, n: u64) -> kb_status;
    pub fn kb_store_clear(ctx: *mut KbCtx) -> kb_status;
    pub fn kb_store_build_index(ctx: *mut KbCtx, n_predicates: *mut u32, build_ms: *mut c_double) -> kb_status;
    pub fn kb_store_size(ctx: *mut KbCtx, n_triples: *mut u64, n_segments: *mut u32) -> kb_status;
    pub fn kb_dict_numeric_load(ctx: *mut KbCtx, num_or0: *const c_double, is_num: *const u8, n_ids: u32) -> kb_status;
    pub fn kb_dict_strings_load(ctx: *mut KbCtx, offsets: *const u64, bytes: *const u8, n_ids: u32) -> kb_status;
    pub fn kb_dict_encode(ctx: *mut KbCtx, offsets: *const u64, bytes: *const u8, n_terms: u64, out_ids: *mut u32, n_new: *mut u32, new_first_pos: *mut u64) -> kb_status;
    pub fn kb_dict_strings_info(ctx: *mut KbCtx, n_ids: *mut u32, n_bytes: *mut u64) -> kb_status;
    pub fn kb_store_append_device(ctx: *mut KbCtx, d_s: *const u32, d_p: *const u32, d_o: *const u32, n: u64, tag: u64) -> kb_status;
    // relations
    pub fn kb_rel_info(r: *const KbRel, n_rows: *mut u64, n_cols: *mut u32, slots: *mut u32) -> kb_status;
    pub fn kb_rel_download(ctx: *mut KbCtx, r: *const KbRel, col: u32, dst: *mut u32) -> kb_status;
    pub fn kb_rel_from_host(ctx: *mut KbCtx, slots: *const u32, n_cols: u32, cols: *const *const u32, n_rows: u64, out: *mut *mut KbRel) -> kb_status;
    pub fn kb_rel_decode(ctx: *mut KbCtx, r: *const KbRel, col: u32, out: *mut *mut KbStrings) -> kb_status;
    pub fn kb_strings_info(s: *const KbStrings, n_strings: *mut u64, total_bytes: *mut u64) -> kb_status;
    pub fn kb_strings_download(ctx: *mut KbCtx, s: *const KbStrings, offsets: *mut u64, bytes: *mut u8) -> kb_status;
    pub fn kb_strings_free(ctx: *mut KbCtx, s: *mut KbStrings);
    pub fn kb_rel_free(ctx: *mut KbCtx, r: *mut KbRel);
    // operators
    pub fn kb_scan(ctx: *mut KbCtx, pats: *const KbPattern, n: u32, pushdown: *const *const KbFilterOp, pushdown_len: *const u32, out: *mut *mut KbRel) -> kb_status;
    pub fn kb_filter(ctx: *mut KbCtx, r: *const KbRel, prog: *const KbFilterOp, n: u32, out: *mut *mut KbRel) -> kb_status;
    pub fn kb_project(ctx: *mut KbCtx, r: *const KbRel, slots: *const u32, n: u32, out: *mut *mut KbRel) -> kb_status;
    pub fn kb_hash_join(ctx: *mut KbCtx, l: *const KbRel, r: *const KbRel, out: *mut *mut KbRel) -> kb_status;
    pub fn kb_bind_join(ctx: *mut KbCtx, l: *const KbRel, pattern: *const KbPattern, out: *mut *mut KbRel) -> kb_status;
    pub fn kb_star_join(ctx: *mut KbCtx, join_slot: u32, pats: *const KbPattern, n: u32, f: *const KbFilterOp, nf: u32, out: *mut *mut KbRel) -> kb_status;
    pub fn kb_group_aggregate(ctx: *mut KbCtx, r: *const KbRel, group_slots: *const u32, n_group: u32, aggs: *const KbAgg, n_aggs: u32, out: *mut *mut KbGroups) -> kb_status;
    pub fn kb_star_join_aggregate(ctx: *mut KbCtx, join_slot: u32, pats: *const KbPattern, n: u32, f: *const KbFilterOp, nf: u32, group_slots: *const u32, n_group: u32,
                                  aggs: *const KbAgg, n_aggs: u32, out: *mut *mut KbGroups, n_rows: *mut u64) -> kb_status;
    pub fn kb_groups_info(g: *const KbGroups, n_groups: *mut u64, n_group_cols: *mut u32, n_aggs: *mut u32) -> kb_status;
    pub fn kb_groups_keys(g: *const KbGroups, col: u32, keys: *mut *const u32) -> kb_status;
    pub fn kb_groups_values(g: *const KbGroups, agg: u32, values: *mut *const c_double) -> kb_status;
    pub fn kb_groups_counts(g: *const KbGroups, counts: *mut *const u64) -> kb_status;
    pub fn kb_groups_free(g: *mut KbGroups);
    // prepared plans
    pub fn kb_star_join_prepare(ctx: *mut KbCtx, join_slot: u32, pats: *const KbPattern, n: u32, f: *const KbFilterOp, nf: u32, group_slots: *const u32, n_group: u32,
                                aggs: *const KbAgg, n_aggs: u32, ring: u32, out: *mut *mut KbPlan) -> kb_status;
    pub fn kb_plan_submit(ctx: *mut KbCtx, plan: *mut KbPlan, ticket: *mut u64) -> kb_status;
    pub fn kb_plan_collect(ctx: *mut KbCtx, plan: *mut KbPlan, ticket: u64, n_rows: *mut u64, rows: *mut *mut KbRel, groups: *mut *mut KbGroups) -> kb_status;
    pub fn kb_plan_free(ctx: *mut KbCtx, plan: *mut KbPlan);
    // Datalog
    pub fn kb_datalog_fixpoint(ctx: *mut KbCtx, rules: *const KbRule, n: u32, strategy: u32, out: *mut *mut KbRel, stats: *mut KbFixpointStats) -> kb_status;
    // incremental: the store is closed under the rules already, `seed` (s, p, o) are the triples added since
    pub fn kb_datalog_fixpoint_seed(ctx: *mut KbCtx, rules: *const KbRule, n: u32, strategy: u32, seed: *const KbRel, out: *mut *mut KbRel, n_seed_new: *mut u64,
                                    stats: *mut KbFixpointStats) -> kb_status;
    // multi-GPU helpers
    pub fn kb_shard_of(key: u32, n_shards: u32) -> u32;
    pub fn kb_set_sharding(ctx: *mut KbCtx, rank: u32, world: u32) -> kb_status;
    pub fn kb_shuffle_push(ctx: *mut KbCtx, r: *const KbRel, key_slot: u32, n_parts: u32, peer_cols: *const *mut u32, peer_cursors: *const *mut u32, capacity_rows: u64) -> kb_status;
    pub fn kb_rel_wrap_device(ctx: *mut KbCtx, slots: *const u32, n_cols: u32, d_cols: *const *mut u32, n_rows: u64, out: *mut *mut KbRel) -> kb_status;
    pub fn kb_groups_pack(g: *const KbGroups, dst: *mut c_void, capacity_bytes: u64, bytes: *mut u64) -> kb_status;
    // the rest of the header (timing, device-resident loads, the legacy executor's i32 view, BGP execution, plan introspection and the
    // cross-rank GROUP BY attachment, the count-matrix shuffle, store download, on-disk segments, the one-shot host-buffer joins)
    pub fn kb_set_timing(ctx: *mut KbCtx, enabled: c_int) -> kb_status;
    pub fn kb_get_stats(ctx: *mut KbCtx, out: *mut KbStats, reset: c_int) -> kb_status;
    pub fn kb_store_load_device(ctx: *mut KbCtx, d_s: *const u32, d_p: *const u32, d_o: *const u32, n: u64) -> kb_status;
    pub fn kb_set_use_index(ctx: *mut KbCtx, enabled: c_int) -> kb_status;
    pub fn kb_dict_legacy_i32_load(ctx: *mut KbCtx, val: *const i32, is_i32: *const u8, n_ids: u32) -> kb_status;
    pub fn kb_rel_device_col(r: *const KbRel, col: u32, d_ptr: *mut *const u32) -> kb_status;
    pub fn kb_bgp_execute(ctx: *mut KbCtx, pats: *const KbPattern, n_pats: u32, filter: *const KbFilterOp, n_filter_ops: u32, project_slots: *const u32, n_project: u32,
                          out: *mut *mut KbRel) -> kb_status;
    pub fn kb_plan_info(plan: *const KbPlan, ring: *mut u32, capacity_rows: *mut u64, n_cols: *mut u32, slots: *mut u32, grouped: *mut u32) -> kb_status;
    pub fn kb_plan_peer_scratch_bytes(plan: *const KbPlan) -> u64;
    pub fn kb_plan_attach_peers(ctx: *mut KbCtx, plan: *mut KbPlan, rank: u32, world: u32, peer_scratch: *const *mut c_void) -> kb_status;
    pub fn kb_partition(ctx: *mut KbCtx, r: *const KbRel, key_slot: u32, n_parts: u32, out: *mut *mut KbRel, part_offsets: *mut u64) -> kb_status;
    pub fn kb_partition_counts(ctx: *mut KbCtx, r: *const KbRel, key_slot: u32, n_parts: u32, counts: *mut u64) -> kb_status;
    pub fn kb_shuffle_scatter(ctx: *mut KbCtx, r: *const KbRel, key_slot: u32, n_parts: u32, peer_cols: *const *mut u32, base: *const u64, capacity_rows: u64) -> kb_status;
    pub fn kb_rel_from_device(ctx: *mut KbCtx, slots: *const u32, n_cols: u32, d_cols: *const *const u32, n_rows: u64, out: *mut *mut KbRel) -> kb_status;
    pub fn kb_store_download(ctx: *mut KbCtx, s: *mut u32, p: *mut u32, o: *mut u32, cap: u64, n: *mut u64) -> kb_status;
    pub fn kb_segment_write(path: *const c_char, s: *const u32, p: *const u32, o: *const u32, n: u64, tag: u64) -> kb_status;
    pub fn kb_segment_info(path: *const c_char, n_triples: *mut u64, tag: *mut u64, cmin: *mut u32, cmax: *mut u32) -> kb_status;
    pub fn kb_segment_save(ctx: *mut KbCtx, tag: u64, whole_store: c_int, path: *const c_char) -> kb_status;
    pub fn kb_store_append_file(ctx: *mut KbCtx, path: *const c_char, tag: u64, verify: c_int) -> kb_status;
    pub fn kb_star_join_host(ctx: *mut KbCtx, s: *const u32, p: *const u32, o: *const u32, n: u64, join_slot: u32, pats: *const KbPattern, n_pats: u32,
                             filter: *const KbFilterOp, n_filter_ops: u32, n_cols: *mut u32, slots: *mut u32, cols: *mut *mut u32, n_rows: *mut u64) -> kb_status;
    pub fn kb_star_join_host_into(ctx: *mut KbCtx, s: *const u32, p: *const u32, o: *const u32, n: u64, join_slot: u32, pats: *const KbPattern, n_pats: u32,
                                  filter: *const KbFilterOp, n_filter_ops: u32, n_cols: *mut u32, slots: *mut u32, cols: *const *mut u32, capacity_rows: u64,
                                  n_rows: *mut u64) -> kb_status;
    pub fn kb_groups_merge(ctx: *mut KbCtx, parts: *const *const c_void, part_bytes: *const u64, n_parts: u32, out: *mut *mut KbGroups) -> kb_status;
}
