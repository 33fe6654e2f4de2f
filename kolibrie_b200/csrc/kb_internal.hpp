// kb_internal.hpp — host-side state behind the C ABI (context, device buffers, relations, store segments).
#pragma once
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "kb_kernels.cuh"

namespace kb {

struct Life {  // shared by a context and every buffer it handed out: relations may outlive their context (bindings' GC order)
    bool alive = true;
    // Large buffers are recycled here instead of going back to the stream-ordered pool: a freed 267 MB result buffer that the pool has
    // to re-map for the next query costs ~0.5 ms (measured at N = 2: 0.69 ms per synchronous query instead of 0.12), a closure's
    // multi-GB join outputs several ms. Everything is allocated and released on the context's one stream, so a recycled buffer is
    // ordered behind its previous user exactly like a pool allocation would be.
    struct Cached { void* p; size_t bytes; };
    std::vector<Cached> cache;
    size_t cached_bytes = 0;
    bool cache_off = false;  // KOLIBRIE_BUF_CACHE=0
    // Sized for the working set of the largest workload (a config-4 closure releases ~30 GB of buffers when it ends and asks for the
    // same sizes in the same order the next time): with 24 GB / 48 entries the 8 GB known-fact set fell out of the cache, and one
    // closure in seven then waited 0.3-1.2 s for the pool to re-map it. When the cache is full the OLDEST entries go back to the pool.
    static constexpr size_t MIN_BYTES = 1u << 20, MAX_ENTRIES = 256, MAX_BYTES = 56ull << 30;
};
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaStream_t st = nullptr;
    std::shared_ptr<Life> life;
    ~DevBuf() {
        if (!p) return;
        if (life && life->alive) {
            if (!life->cache_off && bytes >= Life::MIN_BYTES && bytes <= Life::MAX_BYTES / 2) {
                while (!life->cache.empty() && (life->cache.size() >= Life::MAX_ENTRIES || life->cached_bytes + bytes > Life::MAX_BYTES)) {
                    cudaFreeAsync(life->cache.front().p, st);  // (entries are in release order: the front is the oldest)
                    life->cached_bytes -= life->cache.front().bytes;
                    life->cache.erase(life->cache.begin());
                }
                life->cache.push_back({p, bytes});
                life->cached_bytes += bytes;
                return;
            }
            cudaFreeAsync(p, st);  // stream-ordered: kernels still queued on st may be using it
        } else cudaFree(p);         // the context (and its stream) is gone
    }
};
using Buf = std::shared_ptr<DevBuf>;

struct Col {
    Buf buf;            // owner (may be shared by views / projections)
    u32* ptr = nullptr;  // 16-byte aligned, readable up to the next multiple of 256 B past the last row
};

// a timing event that is destroyed on every return path
struct ScopedEvent {
    cudaEvent_t e = nullptr;
    ScopedEvent() { cudaEventCreate(&e); }
    ~ScopedEvent() { if (e) cudaEventDestroy(e); }
    ScopedEvent(const ScopedEvent&) = delete;
    ScopedEvent& operator=(const ScopedEvent&) = delete;
    operator cudaEvent_t() const { return e; }
};

enum Family { F_SCAN = 0, F_BUILD, F_PROBE, F_FILTER, F_GROUP, F_OTHER, F_COUNT };

struct PendingTimer {
    cudaEvent_t a, b;
    int fam;
};

}  // namespace kb

struct kb_rel {
    std::vector<kb::u32> slots;
    std::vector<kb::Col> cols;
    kb::u64 n = 0;
    bool pair = false;  // internal only: slots = {subject var, object var}, cols[0] = interleaved uint2 (s,o) rows
    int col_of(kb::u32 slot) const {
        for (size_t i = 0; i < slots.size(); i++) if (slots[i] == slot) return (int)i;
        return -1;
    }
};

struct kb_strings {
    kb::Buf off;    // u32 [n + 1] (exclusive prefix of the lengths)
    kb::Buf bytes;
    kb::u64 n = 0, total = 0;
};

struct kb_groups {
    std::vector<std::vector<kb::u32>> keys;
    std::vector<std::vector<double>> vals;
    std::vector<uint64_t> counts;
    std::vector<kb::u32> kinds;            // kb_agg_kind of every aggregate
    std::vector<std::vector<double>> raw;  // the accumulators as the device left them (AVG: the SUM, before the division): what a
                                           // cross-rank merge combines (kb_groups_pack / kb_groups_merge)
};

namespace kb {
struct Segment {
    u64 tag = 0;
    Col s, p, o;
    u64 n = 0;
    cudaEvent_t ready = nullptr;  // set while a chunked upload is in flight: the scan waits on it
    u32 cmin[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};  // per-column id range (s,p,o), computed when the segment is loaded
    u32 cmax[3] = {0, 0, 0};
    bool has_stats = false;
    u32 stats_world = 0;   // sharding the statistics were checked against
    bool sharded_ok = true;  // every subject belongs to this context's shard (checked when kb_set_sharding is in effect)
    // distinct predicates of the segment with their row counts, from the same pass as the ranges (empty + preds_overflow when it
    // carries more than SEGP_SLOTS of them): what the index maintenance of a window slide allocates its chunks from
    std::vector<std::pair<u32, u32>> pred_rows;
    bool has_preds = false, preds_overflow = false;
};
}  // namespace kb

namespace kb {
struct SliceChunk {  // the rows of ONE store segment that carry the predicate: interleaved (subject, object)
    u64 tag = 0;           // the segment's tag (kb_store_append): evicting the segment drops the chunk
    Col pairs;
    u64 n = 0;
    Buf ynum;              // typed literal column: f64 value of every object (num_or0), kept when the chunk has numeric objects
    u64 ynum_version = 0;  // numeric side table version it was built from
    u32 xmin = 0xFFFFFFFFu, xmax = 0, ymin = 0xFFFFFFFFu, ymax = 0;  // id ranges of the chunk (an eviction re-derives the slice's from the survivors)
};
struct PredSlice {  // one predicate's (subject, object) rows of the store — the device analogue of pos[P] (index_manager.rs:18-26) — as one
                    // chunk per store segment, so that an RSP window slide (append one segment, evict another) maintains the index
                    // instead of dropping it (rsp_engine.rs:94-104; simple_r2r.rs:95-142)
    std::vector<SliceChunk> chunks;
    u64 n = 0;  // rows over all chunks
    u32 xmin = 0xFFFFFFFFu, xmax = 0, ymin = 0xFFFFFFFFu, ymax = 0;  // id range of subjects (x) and objects (y) over the live chunks
    // persistent direct tables kept while the column is unique and dense: xtab[subject - xtab_min] = object (what the reference's
    // spo[s][P] lookup answers, index_manager.rs:18-26) and ytab[object - ytab_min] = subject (pos[P][o]). *_range is the table's
    // CAPACITY in slots: it is allocated with headroom above the largest key (dictionary ids grow), appended chunks are inserted in
    // place, evicted chunks are cleared in place, and the table is rebuilt only when a key falls outside it.
    Buf xtab, ytab;
    u32 xtab_min = 0, xtab_range = 0, ytab_min = 0, ytab_range = 0, tab_cshift = 0;
    bool x_unique = false, y_unique = false;  // no subject (object) occurs twice -> builds keyed on it need no duplicate detection
                                              // (functional / inverse-functional predicate in this store)
    bool x_tried = false, y_tried = false;    // a table for the column was attempted (false: never dense enough)
    // key-grouped directories for the columns that are NOT unique (multi-valued predicates, objects shared by many subjects): the
    // slice sorted by that column — a counting sort, the ids being dense — as off[key - min .. ] + the other half in key order. What the
    // reference's spo[s][P] -> {o} and pos[P][o] -> {s} sets are (index_manager.rs:18-26, 253-340): a bound subject / object is a range
    // of val, not a scan. Built by kb_store_build_index for single-chunk slices; dropped when the predicate's slice is appended to.
    Buf xnum;                // typed values in TABLE order: xnum[compact(subject) - xtab_min] = num_or0[object] (built with xtab by kb_store_build_index
    u64 xnum_version = 0;    // when the slice has numeric objects; the table-mode probe filters on it sequentially)
    Buf xoff, xval, yoff, yval;
    u32 xcsr_min = 0, xcsr_range = 0, ycsr_min = 0, ycsr_range = 0;
    const SliceChunk* single() const { return chunks.size() == 1 ? &chunks[0] : nullptr; }
    bool typed(u64 num_version) const {
        for (auto& c : chunks) if (!c.ynum || c.ynum_version != num_version) return false;
        return !chunks.empty();
    }
};
}  // namespace kb

struct kb_ctx {
    int device = 0;
    int n_sms = 148;
    cudaStream_t st = nullptr;
    cudaStream_t st_copy = nullptr;
    cudaEvent_t ev_copy = nullptr;
    std::string err;
    std::shared_ptr<kb::Life> life = std::make_shared<kb::Life>();
    std::vector<kb::Segment> segs;
    kb::u64 n_triples = 0;
    kb::Buf num, isnum;
    kb::Buf i32val, isi32;  // kb_dict_legacy_i32_load: the legacy executor's integer view of the terms
    kb::u32 n_i32 = 0;
    kb::Buf dict_off, dict_bytes;  // kb_dict_strings_load / kb_dict_encode: u64 offsets [dict_ids + 1] + UTF-8 bytes
    kb::u32 dict_ids = 0;
    kb::Buf dict_index;            // kb_dict_encode: open-addressing index string -> id (hash tag << 32 | id), built lazily
    kb::u64 dict_index_slots = 0;
    kb::u32 dict_indexed = 0;      // ids [0, dict_indexed) are in the index
    kb::u32 n_ids = 0;
    kb::u64 num_version = 1;  // bumped by kb_dict_numeric_load: typed literal columns of the index are tied to it
    // tile-state buffer of the look-back prefix (never cleared: words carry the launch epoch)
    kb::Buf tile_state, block_state;
    size_t tile_state_tiles = 0;
    // 0 (default): compaction in tile-completion order (one atomic per tile; row order of results is unspecified, as in the
    // reference whose row order is hash-iteration order). 1 (KOLIBRIE_ORDERED=1, and always for the legacy FFI symbol): store order.
    kb::u32 ordered = 0;
    kb::u64 epoch = 1;
    // control arena: small device words (tickets, totals, flags) zeroed at the start of every API call, mirrored in pinned memory
    kb::u32* ctrl = nullptr;
    kb::u32* h_ctrl = nullptr;
    kb::u32 ctrl_used = 0;
    bool ctrl_dirty = true;  // words handed out since the arena was last zeroed (a call that hands out none skips the memset)
    // control block of the one-kernel index join: device words {ticket, total, 0, done} that the kernel itself leaves zeroed, and the
    // pinned host word the kernel publishes the row count to
    kb::u32* fast_cb = nullptr;
    kb::u32* h_fast = nullptr;
    kb::u32* d_fast = nullptr;  // device address of h_fast
    static constexpr kb::u32 CTRL_WORDS = 8192;
    // pinned staging for uploads / downloads
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    // timing
    bool timing = false;
    std::vector<kb::PendingTimer> timers;
    std::vector<cudaEvent_t> ev_pool;
    kb_stats stats{};
    // knowledge cached across calls: (predicate, key position) pairs whose direct build met duplicate keys
    std::map<kb::u32, kb::u64> fix_hint;  // head predicate -> facts it held when the last fixpoint ended (sizes the next run's known-fact sets)
    std::shared_ptr<void> fix_state;      // kb_datalog_fixpoint_seed: the per-predicate relations and known-fact sets of the last seeded
                                          // call (kb_datalog.cu: FixState), valid while store_version is the one it was left at
    std::set<std::pair<kb::u32, kb::u32>> multi_valued;
    std::set<std::pair<kb::u32, kb::u32>> single_valued;  // verified duplicate-free by an earlier direct build on this store version
    cudaStream_t st2 = nullptr;                            // second compute stream: independent direct builds run concurrently
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    kb::u64 store_version = 0;
    std::map<kb::u32, kb::PredSlice> index;  // kb_store_build_index: predicate -> slice; valid while index_version == store_version
    kb::u64 index_version = ~0ull;
    bool csr_join = true;                // KOLIBRIE_CSR_JOIN=0: 1:N joins always use the chained table (A/B switch)
    // Datalog candidate dedup: radix-partitioned above these sizes (KOLIBRIE_DERIVE_PART=0: never; =1 with KOLIBRIE_DERIVE_SLICE /
    // KOLIBRIE_DERIVE_MIN_ROWS: tests force it on small inputs)
    int derive_part = 1;
    kb::u64 derive_slice_bytes = 32ull << 20;
    kb::u64 derive_min_part_rows = 1ull << 20;
    kb::u64 derive_bucket_slack = 8192;  // KOLIBRIE_DERIVE_SLACK: rows a bucket holds beyond 9/8 of its fair share
    bool fast_index_kernel = true;       // KOLIBRIE_INDEX_KERNEL=0: index joins go through the generic probe kernel (A/B switch)
    bool use_index = true;               // KOLIBRIE_USE_INDEX=0 / kb_set_use_index: force the scanning path
    bool in_index_build = false;         // scan_impl must read the store itself (index build in progress / remaining patterns of a mixed scan)
    bool in_full_index_build = false;
    bool probe_table_mode = true;        // KOLIBRIE_PROBE_TABLE=0: the index probe always streams the slice (A/B switch)
    bool index_maintain = true;          // KOLIBRIE_INDEX_MAINTAIN=0: append / evict drop the index instead of maintaining it (A/B switch)
    kb::u32 shard_rank = 0, shard_world = 1;  // kb_set_sharding: the store is shard `rank` of `world`, sharded by subject
    int upload_stats_off = -1;  // chunked upload in flight: control words where the copy stream accumulates the column ranges
};

namespace kb {

kb_status fail(kb_ctx* ctx, kb_status code, const char* fmt, ...);
#define KB_CUDA(ctx, call)                                                                                             \
    do {                                                                                                               \
        cudaError_t _e = (call);                                                                                       \
        if (_e != cudaSuccess) return kb::fail((ctx), KB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define KB_TRY(expr)                     \
    do {                                 \
        kb_status _s = (expr);           \
        if (_s != KB_OK) return _s;      \
    } while (0)

kb_status alloc_buf(kb_ctx* ctx, size_t bytes, Buf* out);
kb_status alloc_col(kb_ctx* ctx, u64 rows, Col* out);
kb_status begin_call(kb_ctx* ctx);                      // zero the control arena
u32 ctrl_alloc(kb_ctx* ctx, u32 words);                 // returns word offset into ctx->ctrl
kb_status ctrl_read(kb_ctx* ctx);                       // D2H of the used part of the arena + stream sync
kb_status ensure_tile_state(kb_ctx* ctx, u64 tiles);
NumTab numtab(const kb_ctx* ctx);
void timer_begin(kb_ctx* ctx, int fam, int n_kernels = 1);
void timer_end(kb_ctx* ctx);
void timers_flush(kb_ctx* ctx);  // after a stream sync

// internal operators (device-resident, return relations whose row counts are known on the host)
struct FilterProg {
    std::vector<kb_filter_op> ops;
};
kb_status validate_filter(kb_ctx* ctx, const kb_filter_op* ops, u32 n);
// split a postfix program into its top-level AND conjuncts
bool split_conjuncts(const kb_filter_op* ops, u32 n, std::vector<FilterProg>* out);
std::set<u32> filter_slots(const FilterProg& f);

// pairs=true: 2-variable (?s P ?o)-shaped patterns are emitted as interleaved (s,o) pair relations (internal fast path)
struct ScanTable {  // scan+build fusion: pattern k inserts its matches into this direct table instead of emitting rows
    u32* tab = nullptr;
    u32 kmin = 0, range = 0;  // compacted key space when cshift != 0
    u32 cshift = 0;
    u32 key_is_o = 0, trusted = 0;
    u32* dup_flag = nullptr;
    bool clear = false;  // the table has not been set to 0xFF yet: scan_impl clears it (inside the scan kernel when it can)
};
kb_status scan_impl(kb_ctx* ctx, const kb_pattern* pats, u32 n_pats, const std::vector<FilterProg>& pushdown, bool want_index, bool pairs,
                    std::vector<std::unique_ptr<kb_rel>>* out, const std::vector<ScanTable>* tables = nullptr);
// fused GROUP BY of a star join (kb_star_join_aggregate): in = group / aggregate spec; out = groups when the one-kernel index path took it
struct AggSpec {
    u32 group_slot = 0;
    bool has_agg = false;
    u32 kind = 0, agg_slot = 0;
    bool applied = false;  // false on return: the caller joins and groups in two steps
    u64 n_rows = 0;        // joined rows (when applied)
    std::unique_ptr<kb_groups> groups;
};
// What the one-kernel index path resolves a star join into (kb_star_join_prepare keeps it; kb_plan_submit only launches): the kernel
// parameters minus everything that belongs to one launch (output buffers, control block, epoch).
struct IndexPlan {
    bool ok = false;
    ProbeIParams P{};
    std::vector<u32> out_slots;  // variable of kernel output column c
    std::vector<u32> all_slots;  // column order of the relation handed to the caller
    u32 n_out = 0;
    u64 probe_rows = 0;          // rows of the probe slice = capacity a result buffer needs
    bool agg = false;            // GROUP BY folded into the kernel (P.gsel / P.asel / P.akind are set)
    bool has_agg = false;
    u32 agg_kind = 0, agg_slot = 0, group_slot = 0;
};
// plan_only != nullptr: resolve the one-kernel index path and return without launching (KB_E_UNSUPPORTED when the query does not take it)
kb_status star_join_impl2(kb_ctx* ctx, u32 join_slot, const kb_pattern* pats, u32 n_pats, const kb_filter_op* filter, u32 n_ops, bool allow_fused_scan,
                          std::unique_ptr<kb_rel>* out, AggSpec* agg = nullptr, IndexPlan* plan_only = nullptr);
// GROUP BY hash table on the device (kb_group_aggregate and the fused path): one buffer [val | cnt | keys | state]
struct GroupTable {
    Buf buf;
    u64 slots = 0;
    size_t o_val = 0, o_cnt = 0, o_keys = 0, o_state = 0, bytes = 0;
};
kb_status group_table_create(kb_ctx* ctx, u64 slots, GroupParams* P, GroupTable* t);  // allocates, points P at it, runs the init kernel
// downloads the table and appends its groups to g (AVG divided, COUNT filled in)
kb_status group_table_collect(kb_ctx* ctx, const GroupTable& t, u32 n_group, const kb_agg* aggs, u32 n_aggs, kb_groups* g);
// the host half of it: `hb` = a host copy of the table's buffer
void groups_from_host_table(const char* hb, const GroupTable& t, u32 n_group, const kb_agg* aggs, u32 n_aggs, kb_groups* g);
// the same from a dense record list (launch_group_compact)
void groups_from_records(const GroupRecord* recs, u64 n, u32 n_group, const kb_agg* aggs, u32 n_aggs, kb_groups* g);
kb_status segment_stats(kb_ctx* ctx, Segment* sg);
kb_status index_add_segment(kb_ctx* ctx, size_t seg_idx, bool* indexable);
kb_status index_evict_tag(kb_ctx* ctx, u64 tag);
kb_status store_add_device_segment(kb_ctx* ctx, Segment& sg);
kb_status unpair_rel(kb_ctx* ctx, std::unique_ptr<kb_rel>* r);
kb_status filter_impl(kb_ctx* ctx, const kb_rel& in, const FilterProg& f, std::unique_ptr<kb_rel>* out);
kb_status hash_join_impl(kb_ctx* ctx, const kb_rel& L, const kb_rel& R, const FilterProg* post, std::unique_ptr<kb_rel>* out);
kb_status star_join_impl(kb_ctx* ctx, u32 join_slot, const kb_pattern* pats, u32 n_pats, const kb_filter_op* filter, u32 n_ops,
                         std::unique_ptr<kb_rel>* out);
void pattern_vars(const kb_pattern& pt, std::vector<u32>* slots, std::vector<u32>* src);
u32 pow2_at_least(u64 x);
kb_status check_pattern(kb_ctx* ctx, const kb_pattern& pt);

}  // namespace kb
