// kb_kernels.cuh — parameter blocks and launch wrappers of the sm_100a kernels (definitions in kb_kernels.cu).
#pragma once
#include "kb_device.cuh"

namespace kb {

// ---------------------------------------------------------------------------------------------------------------
// K_scan: fused multi-pattern triple scan + constant filter + pushed-down FILTER + ordered compaction.
// Replaces execute_table_scan_with_ids (engine.rs:510-584), the index scans (engine.rs:1192-1407), QueryBuilder's
// Exact filter scan (query_builder.rs:486-531) and the reference's hash_join_kernel (cuda_join.cu:26-45).
#ifndef KB_SCAN_THREADS
#define KB_SCAN_THREADS 256
#endif
constexpr int SCAN_THREADS = KB_SCAN_THREADS;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048 triples = 24 KB of shared memory per CTA

enum : u32 {
    SP_HAS_S = 1, SP_HAS_P = 2, SP_HAS_O = 4,      // constant in that position
    SP_EQ_SP = 8, SP_EQ_SO = 16, SP_EQ_PO = 32,    // same variable twice in one pattern (quirk Q4: enforced)
    SP_EMIT_S = 64, SP_EMIT_P = 128, SP_EMIT_O = 256,  // positions written to the output (columns come out in s,p,o order)
    SP_EMIT_IDX = 512,                              // global triple index (legacy FFI)
    SP_PAIR = 1024,  // emit (subject, object) interleaved as uint2 into outp[0]: one 8-byte store per match (2-variable patterns)
    // scan + build fused: the match goes straight into a direct join table, table[key - cs] = other half (cs = key base, co = key
    // range, outp[0] = table, outp[1] = duplicate flag); no intermediate relation is written or re-read
    SP_TABLE = 2048,
    SP_TKEY_O = 4096,    // the join key is the object (default: the subject)
    SP_TTRUSTED = 8192   // (kept for the A/B switch) atomicExch-with-return duplicate detection inside the scan when NOT set and KB_FUSED_ATOMIC
};
struct ScanPat {
    u32 cs, cp, co;
    u32 flags;
    u32* outp[4];        // output column of position s / p / o / index (null when not emitted)
    u32 f_begin, f_len;  // pushed-down FILTER program (slots = positions 0/1/2)
};
constexpr int SCAN_MAXSEG = 16;  // store segments one launch can walk (an RSP window of slides is scanned in ONE launch)
struct ScanSeg {
    const u32 *s, *p, *o;
    u32 n;           // triples
    u32 tile0;       // first tile of this segment in the launch's tile numbering
    u32 index_base;  // global index of the segment's first triple
};
struct ScanParams {
    ScanSeg seg[SCAN_MAXSEG];
    u32 n_seg;
    u32 n_tiles, K;  // n_tiles: over all segments of the launch
    u32 cshift;      // SP_TABLE patterns: log2(world) key compaction of a subject-sharded store (0 = none)
    ScanPat pat[MAXP];
    FilterOp ops[KB_MAX_FILTER_OPS];
    NumTab nt;
    // direct tables the FIRST launch of a scan clears itself (scan_star_kernel: every CTA fills a slice with 0xFF, then the grid meets
    // at clear_barrier[0] before any insert; clear_barrier[1] is raised when the meeting timed out) — two 67 MB cudaMemsetAsync calls
    // in front of a 0.32 ms kernel cost 45 us + two launch gaps, the same bytes written by the resident grid ~25 us
    u32* clear_tab[4];
    u32 clear_words[4];
    u32 n_clear;
    u32* clear_barrier;
    u64* tile_state;   // level 1: [n_tiles][MAXP]
    u64* block_state;  // level 2: [ceil(n_tiles/32)][MAXP]
    u32 ordered;       // 1: output in store order (two-level prefix)  0: tile completion order (atomic cursor)
    u32* ticket;
    u64 epoch;
    const u32* totals_in;  // [MAXP] rows written by earlier segments
    u32* totals_out;       // [MAXP] rows written so far (ordered: a different array than totals_in; unordered: the same, used as atomic cursor)
};
void launch_scan(const ScanParams& p, int n_sms, cudaStream_t st);
bool scan_clears_tables(const ScanParams& p);  // true when launch_scan(p) runs the kernel that honours clear_tab / n_clear

// ---------------------------------------------------------------------------------------------------------------
// K_build
// DIRECT: dense dictionary ids make the identity a perfect hash: table[key - kmin] = payload (u32, EMPTY32 = none).
void launch_build_direct(const u32* keys, const u32* vals /*null: row index*/, u32 n, u32* table, u32 kmin, u32 range, u32 cshift,
                         u32* dup_flag, int n_sms, cudaStream_t st);
// same from an interleaved (subject, object) pair relation; key_is_y selects which half is the key, the other is the payload
// trusted=1: the (predicate, key position) is known single-valued for this store version -> plain stores, no read-modify-write
void launch_build_direct_pairs(const uint2* kv, u32 key_is_y, u32 n, u32* table, u32 kmin, u32 range, u32 cshift, u32* dup_flag, u32 trusted,
                               int n_sms, cudaStream_t st);
// CHAINED multimap: open-addressing slots {key tag, head row} + next[] chains. Insert cost is O(1) whatever the key
// multiplicity (1:N joins and heavy hitters of the Datalog joins).
struct ChainTab {
    u64* slots;   // lo 32: key tag (EMPTY32 free), hi 32: head row index (EMPTY32 = end)
    u32 n_slots;  // power of two
    u32* next;    // [n_build]
    u32 n_keys;   // 1..4 key columns; tag = key itself when n_keys==1 else a 32-bit mix of all
    const u32* bkey[4];
};
void launch_build_chained(const ChainTab& t, u32 n, int n_sms, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------------
// K_probe (direct, fused multiway): one probe stream against T<=4 direct tables + FILTER + ordered compaction.
// Replaces execute_star_join_with_ids (engine.rs:587-691) / bind-join chains (engine.rs:840-885) on unique keys.
constexpr int PROBE_THREADS = 256;
constexpr int PROBE_ITEMS = 4;
constexpr int PROBE_TILE = PROBE_THREADS * PROBE_ITEMS;  // 1024 rows
constexpr int MAXT = 4;

struct DirectTab {
    const u32* tab;
    u32 kmin, range;  // in compacted key space when cshift != 0
    u32 cshift;
    u32 mode;  // 0: payload is a value column  1: payload is a build-row index (gather pay[])  2: existence only
    u32 n_pay;
    const u32* pay[2];
};
enum : u32 { OUT_PROBE = 0, OUT_TABVAL = 1, OUT_TABPAY = 2 };
struct OutCol {
    u32 kind, a, b;  // OUT_PROBE: a = probe column; OUT_TABVAL: a = table; OUT_TABPAY: a = table, b = payload column
};
struct ProbeDParams {
    const u32* pcol[KB_MAX_COLS];
    u32 n_pcols, key_col, n, n_tiles, T;
    DirectTab tab[MAXT];
    u32 n_out;
    OutCol oc[KB_MAX_COLS];
    u32* out[KB_MAX_COLS];
    u32 cap;
    FilterOp ops[KB_MAX_FILTER_OPS];  // slots = output column indices
    u32 n_ops;
    NumTab nt;
    u64* tile_state;
    u64* block_state;
    u32 ordered;
    u32* ticket;
    u64 epoch;
    u32* total;            // result row count; must be zero at launch
    const u32* zero_word;  // a word that stays zero (ordered mode reads its base from here)
    const u32* abort_flag;  // non-null: exit immediately if *abort_flag != 0 (a direct build met duplicate keys)
};
void launch_probe_direct(const ProbeDParams& p, int n_sms, cudaStream_t st);

// K_probe (direct, FAST): the probe side is a pair relation; each thread owns 4 consecutive rows held in registers, so the shared
// tile is refilled by TMA while the rows are probed (same register-staged double buffering as K_scan).
constexpr int PROBEF_THREADS = 256;
#ifndef KB_PROBEF_ITEMS
#define KB_PROBEF_ITEMS 4
#endif
constexpr int PROBEF_ITEMS = KB_PROBEF_ITEMS;  // rows per thread: 8 keeps twice the lookups in flight per thread (the kernel is latency-bound)
constexpr int PROBEF_TILE = PROBEF_THREADS * PROBEF_ITEMS;
struct ProbeFParams {
    const uint2* pairs;
    u32 key_is_y;  // which half of the pair is the join key
    u32 n, n_tiles, T;
    DirectTab tab[MAXT];  // mode 0 (value payload) or 2 (existence) only
    u32 n_out;
    OutCol oc[KB_MAX_COLS];  // OUT_PROBE: a = 0 (pair.x) / 1 (pair.y); OUT_TABVAL: a = table
    u32* out[KB_MAX_COLS];
    u32 cap;
    FilterOp ops[KB_MAX_FILTER_OPS];
    u32 n_ops;
    FilterOp pre_ops[8];  // conjuncts over the probe row alone (slot 0 = pair.x, 1 = pair.y): evaluated BEFORE any table lookup
    u32 n_pre;
    const double* pre_num;  // non-null: numeric value of pair.y per probe row (typed literal column of the index): a coalesced read
                            // replaces the random gather num_or0[pair.y] when the pre-filter is FILTER(?y <cmp> c)
    NumTab nt;
    u64* tile_state;
    u64* block_state;
    u32 ordered;
    u32* ticket;
    u64 epoch;
    u32* total;
    const u32* zero_word;
    const u32* abort_flag;
};
void launch_probe_fast(const ProbeFParams& p, int n_sms, cudaStream_t st);

// K_probe (INDEX): the star join over the store index when every non-probe pattern has a persistent per-predicate table. The probe
// rows are one predicate slice (x = subject, y = object) and, when the pattern carries FILTER(?y <cmp> c), the slice's typed literal
// column; both are TMA-staged into shared memory by the same mbarrier, so the FILTER is resolved BEFORE any table lookup is issued.
// Output layout is fixed: column 0 = x, 1 = y, 2+t = value found in table t. The launch owns a 4-word control block
// {ticket, total, zero, done} that the last CTA to finish resets to zero after publishing the row count to pinned host memory: a
// query is ONE device operation (no memset before, no copy after).
constexpr int PROBEI_MAXSEG = 16;  // chunks of the probe slice one launch can walk (one per store segment: an RSP window of slides)
struct ProbeISeg {
    const uint2* pairs;
    const double* ynum;  // typed literal column of the chunk (non-null iff pre_mode == 1)
    u32 n, tile0;        // rows; first tile of this chunk in the launch's tile numbering
};
struct ProbeIParams {
    ProbeISeg seg[PROBEI_MAXSEG];
    u32 n_seg;
    // table mode (ptab != nullptr): the probe stream is the probe pattern's own persistent table, slot by slot (keys ascending), and
    // pnum the numeric value of the slot's other half (pre_mode == 1); keys are rebuilt from the slot number (ptab_min + slot, then
    // the inverse of the shard compaction)
    const u32* ptab;
    const double* pnum;
    u32 ptab_min, ptab_range, ptab_cshift, shard_rank;
    u32 key_is_y, n, n_tiles, T;  // n = rows over all chunks
    DirectTab tab[MAXT];
    u32* out[2 + MAXT];
    u32 cap;
    u32 pre_mode;  // 0: none, 1: cmp_num(pre_cmp, ynum[row], pre_val), 2: pre_ops over (x, y)
    u32 pre_cmp;
    double pre_val;
    FilterOp pre_ops[8];
    u32 n_pre;
    FilterOp ops[KB_MAX_FILTER_OPS];  // conjuncts that need a looked-up value: evaluated over the output row
    u32 n_ops;
    NumTab nt;
    u64* tile_state;
    u64* block_state;
    u32 ordered;
    u64 epoch;
    u32 gsel, asel, akind;  // fused GROUP BY (launch_probe_index with `agg`): output column of the key / of the aggregated value, kb_agg kind
    u32* cb;          // device control block: [0] ticket, [1] total, [2] always 0, [3] CTAs done
    u32* host_total;  // mapped pinned host word
};
struct GroupParams;
// agg == nullptr: the joined rows are written to p.out; else they are folded into the group table *agg (GROUP BY column p.gsel)
void launch_probe_index(const ProbeIParams& p, const GroupParams* agg, int n_sms, cudaStream_t st);
void launch_unpair(const uint2* kv, u32 n, u32* x, u32* y, cudaStream_t st);
// direct build from a predicate slice of the store index, with the pattern's pushed-down FILTER evaluated on (s, P, o)
struct BuildPairsParams {
    const uint2* kv;
    u32 n, key_is_y, pred;
    u32* table;
    u32 kmin, range, cshift;
    u32* dup_flag;
    u32 trusted;
    FilterOp ops[KB_MAX_FILTER_OPS];  // slots = positions 0/1/2
    u32 n_ops;
    NumTab nt;
    u32* count;  // rows inserted (after the filter)
};
void launch_build_direct_pairs_filtered(const BuildPairsParams& p, int n_sms, cudaStream_t st);
// distinct values of a column into a small open-addressing set (EMPTY32 = free); *overflow set when it fills up
void launch_distinct(const u32* col, u32 n, u32* set, u32 set_slots, u32* overflow, int n_sms, cudaStream_t st);

// ---- window-slide maintenance of the store index: ONE pass profiles a new segment (column ranges, foreign subjects, its distinct
// predicates with their row counts), ONE pass splits it into the predicate slices' new chunks — pairs, id ranges, typed literal
// column, in-place inserts into the persistent tables — and ONE pass clears an evicted segment's keys from all of its slices' tables.
constexpr u32 SEGP_SLOTS = 32;  // open-addressing slots for a segment's distinct predicates (more than that: the batched scan path)
// control words of segment_profile: [0..2] min s/p/o, [3] unused, [4 .. 4+SLOTS) predicate slots (all 0xFF at launch);
// then (all 0 at launch) [P_MAX .. +3) max s/p/o, P_FOREIGN, P_OVERFLOW, [P_COUNT .. +SLOTS) rows per slot
constexpr u32 SEGP_MIN = 0, SEGP_SLOT = 4, SEGP_MAX = 4 + SEGP_SLOTS, SEGP_FOREIGN = SEGP_MAX + 3, SEGP_OVERFLOW = SEGP_MAX + 4, SEGP_COUNT = SEGP_MAX + 8,
              SEGP_WORDS = SEGP_COUNT + SEGP_SLOTS;
void launch_segment_profile(const u32* s, const u32* p, const u32* o, u32 n, u32 rank, u32 world, u32* ctrl, int n_sms, cudaStream_t st);

struct SplitEntry {
    u32 pred;
    u32 n;              // rows of the predicate in the segment (capacity of pairs / ynum)
    uint2* pairs;       // the new chunk
    double* ynum;       // its typed literal column (null: no numeric side table loaded)
    u32* xtab; u32 xtab_min, xtab_range, cshift;   // subject table to insert into (null: none)
    double* xnum;       // typed values in table order, maintained with xtab (null: none)
    u32* ytab; u32 ytab_min, ytab_range;           // object table (null: none)
};
// control words per entry (all 0 at launch): cursor, n_numeric, xdup, ydup, xout, yout, pad, pad, ~xmin, ~ymin, xmax, ymax
constexpr u32 SPLIT_CURSOR = 0, SPLIT_NNUM = 1, SPLIT_XDUP = 2, SPLIT_YDUP = 3, SPLIT_XOUT = 4, SPLIT_YOUT = 5, SPLIT_XMIN = 8, SPLIT_YMIN = 9, SPLIT_XMAX = 10,
              SPLIT_YMAX = 11, SPLIT_WORDS = 12;
struct SplitParams {
    const u32 *s, *p, *o;
    u32 n, k;
    SplitEntry e[MAXP];
    u32* ctrl;  // k * SPLIT_WORDS words
    NumTab nt;
};
void launch_segment_split(const SplitParams& p, int n_sms, cudaStream_t st);

constexpr u32 CLEAR_MAX = 16;
struct ClearEntry {
    const uint2* pairs; u32 n;
    u32* xtab; u32 xtab_min, xtab_range, cshift;
    u32* ytab; u32 ytab_min, ytab_range;
};
struct ClearParams { ClearEntry e[CLEAR_MAX]; u32 k; };
void launch_clear_chunks(const ClearParams& p, int n_sms, cudaStream_t st);
// min/max of both halves of a pair relation: out[0..3] = min x, min y, max x, max y
// typed literal column of a predicate slice: out[i] = num_or0[kv[i].y]; *n_numeric += rows whose object is numeric
void launch_pair_numcol(const uint2* kv, u32 n, NumTab nt, double* out, u32* n_numeric, int n_sms, cudaStream_t st);
// the same values laid out like the subject-keyed direct table: out[compact(subject) - kmin] = num_or0[object]
void launch_pair_numtab(const uint2* kv, u32 n, NumTab nt, double* out, u32 kmin, u32 range, u32 cshift, int n_sms, cudaStream_t st);
void launch_pair_minmax(const uint2* kv, u32 n, u32* out4, int n_sms, cudaStream_t st);
// number of occupied (non-EMPTY32) slots of a direct table: equals the number of inserted rows iff the keys were single-valued
void launch_count_nonempty(const u32* table, u32 n, u32* out_count, int n_sms, cudaStream_t st);
// eviction of a chunk: the table entries of its keys go back to EMPTY32 (keys are unique while a persistent table exists)
void launch_clear_direct_pairs(const uint2* kv, u32 key_is_y, u32 n, u32* table, u32 kmin, u32 range, u32 cshift, int n_sms, cudaStream_t st);
// number of keys that do NOT belong to shard `rank` of `world` (kb_shard_of): 0 for a correctly sharded column
void launch_count_foreign(const u32* col, u32 n, u32 rank, u32 world, u32* out, int n_sms, cudaStream_t st);
void launch_col_minmax(const u32* col, u32 n, u32* out_min, u32* out_max, int n_sms, cudaStream_t st);

// K_probe (chained, binary): general natural join with 1:N matches, multi-column keys, FILTER, ordered compaction.
// Replaces execute_optimized_hash_join_with_ids / execute_hash_join_with_ids / merge join (engine.rs:710-811, 970-1039) and
// perform_hash_join_for_rules (shared/src/join_algorithm.rs:499-677).
constexpr int PROBEC_THREADS = 256;
constexpr int PROBEC_ITEMS = 2;
constexpr int PROBEC_TILE = PROBEC_THREADS * PROBEC_ITEMS;
struct ProbeCParams {
    const u32* pcol[KB_MAX_COLS];
    u32 n_pcols, n, n_tiles;
    u32 pkey[4];  // probe columns matching tab.bkey[]
    ChainTab tab;
    u32 n_bpay;
    const u32* bpay[KB_MAX_COLS];  // build columns appended to the output
    u32 n_out;                      // = n_pcols + n_bpay
    u32* out[KB_MAX_COLS];
    u32 cap;
    FilterOp ops[KB_MAX_FILTER_OPS];
    u32 n_ops;
    NumTab nt;
    u64* tile_state;
    u64* block_state;
    u32 ordered;
    u32* ticket;
    u64 epoch;
    u32* total;
    const u32* zero_word;
    unsigned long long* total64;  // exact output size in 64 bits (the 32-bit positions wrap past 2^32 rows: the host checks this one)
};
void launch_probe_chained(const ProbeCParams& p, int n_sms, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------------
// K_build / K_probe (grouped): 1:N equi-join on ONE key column whose ids are dense. The build side is grouped by key into a CSR
// directory — off[key - kmin] .. off[key - kmin + 1] delimit the key's rows in payload columns permuted into key order — so a probe
// is two directory reads and a contiguous run of payload, and the expansion can hand consecutive OUTPUT rows to consecutive threads
// (coalesced stores). Replaces the (tag, head) + next[] chains wherever the key range allows; same bag of rows as
// perform_hash_join_for_rules (join_algorithm.rs:499-677) / execute_hash_join_with_ids (engine.rs:710-811).
struct CsrTab {
    const u32* off;  // [range + 1]
    u32 kmin, range;
    u32 n_pay;
    const u32* pay[KB_MAX_COLS];  // payload columns in key order
};
void launch_csr_count(const u32* keys, u32 n, u32 kmin, u32* counts, int n_sms, cudaStream_t st);
// in-place exclusive prefix sum of a[0..n) (n >= 1); scratch holds ceil(n / 2048) + 1 words
void launch_exclusive_scan_u32(u32* a, u32 n, u32* scratch, cudaStream_t st);
void launch_csr_fill(const u32* keys, u32 n, u32 kmin, u32* cursor, const u32* const* pay_in, u32* const* pay_out, u32 n_pay, int n_sms, cudaStream_t st);
void launch_csr_total(const u32* pkeys, u32 n, const CsrTab& tab, unsigned long long* total, int n_sms, cudaStream_t st);
// the same directory built from an interleaved (subject, object) slice of the store index: key = one half, payload = the other half
void launch_csr_count_pairs(const uint2* kv, u32 key_is_y, u32 n, u32 kmin, u32* counts, int n_sms, cudaStream_t st);
void launch_csr_fill_pairs(const uint2* kv, u32 key_is_y, u32 n, u32 kmin, u32* cursor, u32* val_out, int n_sms, cudaStream_t st);
constexpr int PROBEG_THREADS = 256;
constexpr int PROBEG_ITEMS = 4;
constexpr int PROBEG_TILE = PROBEG_THREADS * PROBEG_ITEMS;
struct ProbeGParams {
    const u32* pcol[KB_MAX_COLS];
    u32 n_pcols, n, n_tiles;
    u32 pkey;  // probe column holding the key
    CsrTab tab;
    u32* out[KB_MAX_COLS];  // n_pcols probe columns, then tab.n_pay payload columns
    u32 cap;
    u64* tile_state;
    u64* block_state;
    u32 ordered;
    u32* ticket;
    u64 epoch;
    u32* total;
    const u32* zero_word;
    // tiles whose expansion exceeds PROBEG_HEAVY output rows are not expanded by the CTA that scanned them: it appends them here
    // (heavy_count[0] entries) and a second launch spreads each one's output range over the whole grid in PROBEG_CHUNK-row pieces
    struct Heavy { u32 tile, gbase, total, pad; };
    Heavy* heavy;        // room for n_tiles entries (null: no splitting)
    u32* heavy_count;    // zeroed control word
};
constexpr u32 PROBEG_HEAVY = 32768, PROBEG_CHUNK = 4096;
void launch_probe_grouped(const ProbeGParams& p, int n_sms, cudaStream_t st);

// cartesian product (engine.rs:1054-1071) — small inputs only
void launch_cartesian(const u32* const* lcols, u32 nl, u32 n_lcols, const u32* const* rcols, u32 nr, u32 n_rcols, u32* const* out,
                      cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------------
// K_group: hash aggregate (execute_query.rs:1150-1227)
struct GroupParams {
    const u32* gcol[4];
    u32 n_gcols;
    const u32* acol[8];  // null for COUNT
    u32 akind[8];
    u32 n_aggs;
    u32 n;
    NumTab nt;
    // global group table
    u32 n_slots;   // power of two
    u32* gkeys;    // [n_slots][4] (EMPTY32 in [.,0] lane = free is tracked by gstate)
    u32* gstate;   // [n_slots] 0 free, 1 being written, 2 ready
    double* gval;  // [n_slots][8]
    unsigned long long* gcnt;  // [n_slots]
    u32* overflow;
};
void launch_group(const GroupParams& p, int n_sms, cudaStream_t st);
void launch_group_init(const GroupParams& p, cudaStream_t st);
// one partial group of a GROUP BY evaluated elsewhere (another rank): its key, its row count and the raw accumulators (AVG: the sum)
struct GroupRecord {
    u32 keys[4];
    unsigned long long count;
    double raw[8];
};
static_assert(sizeof(GroupRecord) == 88, "layout");
// folds n partial groups into the global table of p: counts add, SUM/AVG accumulators add, MIN/MAX fold (execute_query.rs:1150-1227
// applied to the union of the partials' rows)
void launch_group_merge(const GroupParams& p, const GroupRecord* recs, u32 n, int n_sms, cudaStream_t st);
// Cross-rank GROUP BY without a host round trip (one process per GPU, partial tables in peer-mapped memory):
//   launch_peer_barrier: rank `rank` stores `epoch` into slot [rank] of every peer's flag array (system-scope release) and waits until
//                        every slot of its OWN flag array has reached `epoch` (acquire): all ranks' earlier kernels are then visible
//   launch_group_merge_peers: folds the `world` partial group tables (layout of GroupTable: val | cnt | keys | state, then the overflow
//                        word) — read straight from the peers' memory over NVLink — into the local table of p
struct PeerTables {
    const char* table[64];   // device-visible address of rank r's partial table
    u32* flags[64];          // device-visible address of rank r's flag array [64]
    u32 world, rank;
    u32 n_slots;             // slots per partial table
    u32 o_val, o_cnt, o_keys, o_state, o_overflow;  // byte offsets inside a partial table
};
// the occupied slots of p's table as a dense GroupRecord list: out[0 .. *count), header[0] = count, header[1] = the table's overflow
// word. `out` / `header` may be mapped pinned host memory: a prepared query then needs no device-to-host copy of the sparse table
void launch_group_compact(const GroupParams& p, GroupRecord* out, u32* header, u32 cap, cudaStream_t st);
void launch_peer_barrier(const PeerTables& t, u32 epoch, cudaStream_t st);
void launch_group_merge_peers(const GroupParams& p, const PeerTables& t, int n_sms, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------------
// Datalog: instantiate rule heads from binding columns, apply rule filters (rules.rs:133-165), insert into the
// known-facts set (96-bit keys) and append the facts that were new (infer_generic.rs:42-48).
struct HeadTerm { u32 is_var, value; };  // value = binding column index or constant id
struct RuleFilterDev { u32 lhs_col, cmp, rhs_is_var, rhs_col; double rhs_value; };
struct DeriveParams {
    const u32* bcol[KB_MAX_COLS];
    u32 n;
    HeadTerm head_s, head_o;  // the head predicate is a constant: one known-fact set per predicate, keyed by (s << 32) | o
    RuleFilterDev filt[KB_MAX_RULE_FILTERS];
    u32 n_filt;
    NumTab nt;
    u64* set;       // open addressing, EMPTY64 = free; a fact is new iff its atomicCAS wins
    u32 set_slots;  // power of two
    u32 *out_s, *out_o;
    u32 out_cap;
    u32 budget;                   // stop inserting once this many facts were appended (keeps the set's load bounded): overflow = 2
    u32* out_count;               // appended so far (atomic cursor)
    unsigned long long* n_deriv;  // candidates that passed the filters
    u32* overflow;                // 1: set or output full (error); 2: budget reached, the host grows the set and runs the launch again
};
void launch_derive(const DeriveParams& p, int n_sms, cudaStream_t st);
// RADIX-PARTITIONED candidate dedup (the same work as launch_derive for large candidate sets). The known-fact set is an open-addressing
// table far larger than L2 (config 4: 8 GB), so a direct probe per candidate is one random DRAM line per candidate. Here the
// candidates are first partitioned by the HIGH bits of their home slot (tile sorted by partition in shared memory, one range
// reservation per tile and partition, coalesced runs), so that partition p only touches slice p of the table; the probe pass then walks
// the partitions in order, every CTA on (nearly) the same slice, which is sized to stay resident in L2.
struct DerivePartParams {
    DeriveParams d;            // candidates, head, filters, set, outputs, counters (as for launch_derive)
    u32 slice_bits;            // log2(slots per slice); partition = home slot >> slice_bits
    u32 n_parts;               // set_slots >> slice_bits, <= 1024
    u64* buckets;              // [n_parts][bucket_cap] candidate keys (s << 32 | o)
    u32 bucket_cap;
    u32* cursors;              // [n_parts] keys in each bucket (zeroed)
    u32* tile_start;           // [n_parts + 1] first probe tile of each partition (written by the launch)
    u32* tickets;              // [2] zeroed: tile counters of the two passes
};
void launch_derive_partitioned(const DerivePartParams& p, int n_sms, cudaStream_t st);
void launch_set64_insert(u64* set, u32 set_slots, const u32* s, const u32* o, u32 n, u32* overflow, int n_sms, cudaStream_t st);
void launch_set_insert(uint4* set, u32 set_slots, const u32* s, const u32* p /*null: p_const*/, u32 p_const, const u32* o, u32 n, u32* overflow,
                       int n_sms, cudaStream_t st);

// fused partition + transfer of the multi-GPU join-key shuffle: a tile of rows is sorted by destination rank d = shard_of(key, n_parts)
// in shared memory, one range per (tile, destination) is reserved with an atomicAdd on *cursor_ptrs[d] — the destination's own cursor
// in peer memory (push mode, base = 0) or a local cursor behind a precomputed base[d] — and the runs are streamed into d's receive
// buffer (peer memory over NVLink) as 128-byte-aligned warp stores
struct ShuffleParams {
    const u32* key;
    u32 n, n_parts, n_cols;
    const u32* in[KB_MAX_COLS];
    u32* const* peer_cols;     // device array [n_parts * n_cols]
    u32* const* cursor_ptrs;   // device array [n_parts]: word the range of destination d is reserved on
    const u32* base;           // device array [n_parts]
    u32 capacity;
    u32* overflow;
    u32* ticket;               // zeroed tile counter
    u32 tile, n_tiles, stride, pow2_mask;  // set by the launcher
};
void launch_shuffle_scatter(const ShuffleParams& p, int n_sms, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------------
// id -> string decode of a result column (engine.rs:27-51): lengths, exclusive scan (launch_exclusive_scan_u32), byte gather
// len[i] = length of the string of ids[i] ("unknown" = 7 bytes when the dictionary does not hold the id); *total += sum; *quoted = 1
// when an id carries the quoted-triple bit
void launch_decode_lengths(const u32* ids, u32 n, const unsigned long long* dict_off, u32 dict_ids, u32* len, unsigned long long* total, u32* quoted,
                           int n_sms, cudaStream_t st);
// out[off[i] .. off[i+1]) = bytes of the string of ids[i]; one warp per 32 rows, the lanes copy each string together
void launch_decode_gather(const u32* ids, u32 n, const unsigned long long* dict_off, const unsigned char* dict_bytes, u32 dict_ids, const u32* off,
                          unsigned char* out, int n_sms, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------------
// small utilities
void launch_fill_u32(u32* p, u32 v, u64 n, cudaStream_t st);
void launch_fill_const_col(u32* p, u32 v, u64 n, cudaStream_t st);
// partition rows by mix32(key) % n_parts (multi-GPU shuffle): count then scatter
void launch_part_count(const u32* key, u32 n, u32 n_parts, u32* counts, int n_sms, cudaStream_t st);
void launch_part_scatter(const u32* key, u32 n, u32 n_parts, u32* cursors, const u32* const* in_cols, u32* const* out_cols, u32 n_cols,
                         int n_sms, cudaStream_t st);
// mark rows of a segment that equal any triple of a (small) delete set held in a 96-bit set; then compaction is a scan with flags
void launch_delete_mark(const u32* s, const u32* p, const u32* o, u32 n, const uint4* set, u32 set_slots, u32* p_out /*p with EMPTY32 where deleted*/,
                        int n_sms, cudaStream_t st);

}  // namespace kb
