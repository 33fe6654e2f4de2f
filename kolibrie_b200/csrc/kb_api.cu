// kb_api.cu — C ABI of libkolibrie_b200.so: context, device triple store, relations and the operator entry points.
// Host code only orchestrates: every data-touching step is one of the sm_100a kernels in kb_kernels.cu.
#include <cstdarg>
#include <cstdlib>
#include <algorithm>
#include <chrono>
#include <mutex>

#include "kb_internal.hpp"

using namespace kb;

namespace kb {

static std::string g_create_err;

kb_status fail(kb_ctx* ctx, kb_status code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_err = buf;
    return code;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

static inline size_t round256(size_t b) { return (b + 255) & ~(size_t)255; }

kb_status alloc_buf(kb_ctx* ctx, size_t bytes, Buf* out) {
    auto b = std::make_shared<DevBuf>();
    b->bytes = round256(bytes) + 256;  // every column can be over-read to the next 16-byte boundary by the TMA tile loads
    b->st = ctx->st;
    b->life = ctx->life;
    static const bool cache_on = !(getenv("KOLIBRIE_BUF_CACHE") && getenv("KOLIBRIE_BUF_CACHE")[0] == '0');
    if (!cache_on) ctx->life->cache_off = true;
    if (b->bytes >= Life::MIN_BYTES && cache_on) {  // best fit among the recycled buffers (at most 25 % larger than asked)
        Life& L = *ctx->life;
        int best = -1;
        for (size_t i = 0; i < L.cache.size(); i++)
            if (L.cache[i].bytes >= b->bytes && L.cache[i].bytes <= b->bytes + b->bytes / 4 && (best < 0 || L.cache[i].bytes < L.cache[best].bytes)) best = (int)i;
        if (best >= 0) {
            b->p = L.cache[best].p;
            b->bytes = L.cache[best].bytes;
            L.cached_bytes -= b->bytes;
            L.cache.erase(L.cache.begin() + best);
            *out = b;
            return KB_OK;
        }
    }
    cudaError_t e = cudaMallocAsync(&b->p, b->bytes, ctx->st);
    if (e != cudaSuccess && !ctx->life->cache.empty()) {  // out of memory with buffers parked in the cache: give them back and try once more
        cudaGetLastError();
        for (auto& cb : ctx->life->cache) cudaFreeAsync(cb.p, ctx->st);
        ctx->life->cache.clear();
        ctx->life->cached_bytes = 0;
        cudaStreamSynchronize(ctx->st);
        e = cudaMallocAsync(&b->p, b->bytes, ctx->st);
    }
    if (e != cudaSuccess) {
        b->p = nullptr;
        cudaGetLastError();
        return fail(ctx, KB_E_OOM, "device allocation of %zu bytes failed: %s", b->bytes, cudaGetErrorString(e));
    }
    *out = b;
    return KB_OK;
}

kb_status alloc_col(kb_ctx* ctx, u64 rows, Col* out) {
    Buf b;
    KB_TRY(alloc_buf(ctx, (size_t)rows * sizeof(u32), &b));
    out->buf = b;
    out->ptr = static_cast<u32*>(b->p);
    return KB_OK;
}

kb_status begin_call(kb_ctx* ctx) {
    ctx->err.clear();
    ctx->ctrl_used = 0;
    if (ctx->ctrl_dirty) {
        KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl, 0, kb_ctx::CTRL_WORDS * sizeof(u32), ctx->st));
        ctx->ctrl_dirty = false;
    }
    return KB_OK;
}

u32 ctrl_alloc(kb_ctx* ctx, u32 words) {
    ctx->ctrl_dirty = true;
    u32 off = ctx->ctrl_used;
    ctx->ctrl_used += (words + 3u) & ~3u;
    if (ctx->ctrl_used > kb_ctx::CTRL_WORDS) {  // wrap: callers never keep more than a few hundred words live per call
        off = 0;
        ctx->ctrl_used = (words + 3u) & ~3u;
    }
    return off;
}

kb_status ctrl_read(kb_ctx* ctx) {
    const u32 words = kb_ctx::CTRL_WORDS;  // whole arena: allocations may have wrapped
    KB_CUDA(ctx, cudaMemcpyAsync(ctx->h_ctrl, ctx->ctrl, words * sizeof(u32), cudaMemcpyDeviceToHost, ctx->st));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
    ctx->stats.d2h_bytes += words * sizeof(u32);
    timers_flush(ctx);
    return KB_OK;
}

kb_status ensure_tile_state(kb_ctx* ctx, u64 tiles) {
    if (tiles <= ctx->tile_state_tiles) return KB_OK;
    u64 want = std::max<u64>(tiles, 1024) * 2;
    Buf b, b2;
    KB_TRY(alloc_buf(ctx, want * MAXP * sizeof(u64), &b));
    KB_TRY(alloc_buf(ctx, (want / 32 + 2) * MAXP * sizeof(u64), &b2));
    KB_CUDA(ctx, cudaMemsetAsync(b->p, 0, want * MAXP * sizeof(u64), ctx->st));  // epoch 0 is never used by a launch
    KB_CUDA(ctx, cudaMemsetAsync(b2->p, 0, (want / 32 + 2) * MAXP * sizeof(u64), ctx->st));
    ctx->tile_state = b;
    ctx->block_state = b2;
    ctx->tile_state_tiles = want;
    return KB_OK;
}

NumTab numtab(const kb_ctx* ctx) {
    NumTab nt;
    nt.num_or0 = ctx->num ? static_cast<const double*>(ctx->num->p) : nullptr;
    nt.is_num = ctx->isnum ? static_cast<const u8*>(ctx->isnum->p) : nullptr;
    nt.n_ids = ctx->n_ids;
    nt.i32_val = ctx->i32val ? static_cast<const int*>(ctx->i32val->p) : nullptr;
    nt.is_i32 = ctx->isi32 ? static_cast<const u8*>(ctx->isi32->p) : nullptr;
    nt.n_i32 = ctx->n_i32;
    return nt;
}

static cudaEvent_t get_event(kb_ctx* ctx) {
    if (!ctx->ev_pool.empty()) {
        cudaEvent_t e = ctx->ev_pool.back();
        ctx->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void timer_begin(kb_ctx* ctx, int fam, int n_kernels) {
    ctx->stats.kernel_launches += (u64)n_kernels;
    switch (fam) {
        case F_SCAN: ctx->stats.scan_launches++; break;
        case F_BUILD: ctx->stats.build_launches++; break;
        case F_PROBE: ctx->stats.probe_launches++; break;
        case F_FILTER: ctx->stats.filter_launches++; break;
        case F_GROUP: ctx->stats.group_launches++; break;
        default: ctx->stats.other_launches++; break;
    }
    if (!ctx->timing) return;
    PendingTimer t;
    t.a = get_event(ctx);
    t.b = get_event(ctx);
    t.fam = fam;
    cudaEventRecord(t.a, ctx->st);
    ctx->timers.push_back(t);
}
void timer_end(kb_ctx* ctx) {
    if (!ctx->timing || ctx->timers.empty()) return;
    cudaEventRecord(ctx->timers.back().b, ctx->st);
}
void timers_flush(kb_ctx* ctx) {
    size_t kept = 0;
    for (auto& t : ctx->timers) {
        float ms = 0.f;
        if (cudaEventQuery(t.b) == cudaErrorNotReady) {  // launched by an asynchronous submit that has not finished: next flush
            ctx->timers[kept++] = t;
            continue;
        }
        if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
            switch (t.fam) {
                case F_SCAN: ctx->stats.scan_ms += ms; break;
                case F_BUILD: ctx->stats.build_ms += ms; break;
                case F_PROBE: ctx->stats.probe_ms += ms; break;
                case F_FILTER: ctx->stats.filter_ms += ms; break;
                case F_GROUP: ctx->stats.group_ms += ms; break;
                default: ctx->stats.other_ms += ms; break;
            }
            ctx->stats.total_ms += ms;
        } else {
            cudaGetLastError();
        }
        ctx->ev_pool.push_back(t.a);
        ctx->ev_pool.push_back(t.b);
    }
    cudaGetLastError();  // cudaErrorNotReady is sticky in the per-thread last-error slot
    ctx->timers.resize(kept);
}

// ---------------------------------------------------------------------------------------------------------------
// FILTER program helpers
static int op_arity(u32 op) {
    switch (op) {
        case KB_F_CMP_NUM: case KB_F_EQ_ID: case KB_F_NE_ID: case KB_F_PUSH_VAR: case KB_F_PUSH_CONST: case KB_F_IS_TRIPLE: case KB_F_CMP_LEGACY: return 0;
        case KB_F_NOT: case KB_F_TRUTHY: return 1;
        case KB_F_AND: case KB_F_OR: case KB_F_ADD: case KB_F_SUB: case KB_F_MUL: case KB_F_DIV: return 2;
        default: return -1;
    }
}
static bool op_has_slot(u32 op) {
    return op == KB_F_CMP_NUM || op == KB_F_EQ_ID || op == KB_F_NE_ID || op == KB_F_PUSH_VAR || op == KB_F_IS_TRIPLE || op == KB_F_CMP_LEGACY;
}

kb_status validate_filter(kb_ctx* ctx, const kb_filter_op* ops, u32 n) {
    if (n == 0) return KB_OK;
    if (!ops) return fail(ctx, KB_E_INVALID, "filter program is NULL");
    if (n > KB_MAX_FILTER_OPS) return fail(ctx, KB_E_LIMIT, "filter program longer than %d ops", KB_MAX_FILTER_OPS);
    int depth = 0;
    for (u32 i = 0; i < n; i++) {
        int a = op_arity(ops[i].op);
        if (a < 0) return fail(ctx, KB_E_INVALID, "filter op %u: unknown opcode %u", i, ops[i].op);
        if (depth < a) return fail(ctx, KB_E_INVALID, "filter op %u: stack underflow", i);
        depth += 1 - a;
        if (depth > 10) return fail(ctx, KB_E_LIMIT, "filter expression too deep");
    }
    if (depth != 1) return fail(ctx, KB_E_INVALID, "filter program leaves %d values on the stack", depth);
    return KB_OK;
}

static void split_rec(const kb_filter_op* ops, const std::vector<u32>& begin, u32 b, u32 e, std::vector<FilterProg>* out) {
    if (ops[e - 1].op == KB_F_AND) {
        const u32 right_b = begin[e - 2];
        split_rec(ops, begin, b, right_b, out);
        split_rec(ops, begin, right_b, e - 1, out);
        return;
    }
    FilterProg f;
    f.ops.assign(ops + b, ops + e);
    out->push_back(std::move(f));
}

bool split_conjuncts(const kb_filter_op* ops, u32 n, std::vector<FilterProg>* out) {
    out->clear();
    if (n == 0) return true;
    std::vector<u32> begin(n), stack;
    for (u32 i = 0; i < n; i++) {
        int a = op_arity(ops[i].op);
        if (a < 0 || (int)stack.size() < a) return false;
        u32 b = i;
        for (int k = 0; k < a; k++) { b = stack.back(); stack.pop_back(); }
        begin[i] = b;
        stack.push_back(b);
    }
    if (stack.size() != 1) return false;
    split_rec(ops, begin, 0, n, out);
    return true;
}

std::set<u32> filter_slots(const FilterProg& f) {
    std::set<u32> s;
    for (auto& op : f.ops) if (op_has_slot(op.op)) s.insert(op.slot);
    return s;
}

// append `f` (slots remapped through `remap`) to the device op array, AND-ing it with what is already there
static bool append_prog(std::vector<FilterOp>* dst, const FilterProg& f, const std::map<u32, u32>& remap) {
    const bool had = !dst->empty();
    for (auto& op : f.ops) {
        FilterOp d;
        d.op = op.op; d.slot = op.slot; d.cmp = op.cmp; d.id = op.id; d.value = op.value;
        if (op_has_slot(op.op)) {
            auto it = remap.find(op.slot);
            if (it == remap.end()) return false;
            d.slot = it->second;
        }
        dst->push_back(d);
    }
    if (had) {
        FilterOp a{};
        a.op = KB_F_AND;
        dst->push_back(a);
    }
    return true;
}

void pattern_vars(const kb_pattern& pt, std::vector<u32>* slots, std::vector<u32>* src) {
    const kb_term* ts[3] = {&pt.s, &pt.p, &pt.o};
    for (u32 i = 0; i < 3; i++) {
        if (!ts[i]->is_var) continue;
        if (std::find(slots->begin(), slots->end(), ts[i]->value) == slots->end()) {
            slots->push_back(ts[i]->value);
            src->push_back(i);
        }
    }
}

kb_status check_pattern(kb_ctx* ctx, const kb_pattern& pt) {
    const kb_term* ts[3] = {&pt.s, &pt.p, &pt.o};
    for (auto* t : ts) {
        if (t->is_var > 1) return fail(ctx, KB_E_INVALID, "kb_term.is_var must be 0 or 1");
        if (!t->is_var && t->value == KB_ID_NONE) return fail(ctx, KB_E_INVALID, "constant id 0xFFFFFFFF is reserved (KB_ID_NONE)");
        if (t->is_var && t->value >= 4096) return fail(ctx, KB_E_LIMIT, "variable slot %u too large (max 4095)", t->value);
    }
    return KB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// scan
kb_status segment_stats(kb_ctx* ctx, Segment* sg) {
    if ((sg->has_stats && sg->stats_world == ctx->shard_world) || sg->n == 0) { sg->has_stats = true; sg->stats_world = ctx->shard_world; return KB_OK; }
    // one pass: column ranges, foreign subjects, distinct predicates + their row counts (segment_profile_kernel)
    const u32 off = ctrl_alloc(ctx, SEGP_WORDS);
    KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off, 0xFF, SEGP_MAX * sizeof(u32), ctx->st));
    KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off + SEGP_MAX, 0, (SEGP_WORDS - SEGP_MAX) * sizeof(u32), ctx->st));
    timer_begin(ctx, F_OTHER);
    launch_segment_profile(sg->s.ptr, sg->p.ptr, sg->o.ptr, (u32)sg->n, ctx->shard_rank, ctx->shard_world, ctx->ctrl + off, ctx->n_sms, ctx->st);
    timer_end(ctx);
    KB_CUDA(ctx, cudaGetLastError());
    KB_TRY(ctrl_read(ctx));
    const u32* h = ctx->h_ctrl + off;
    for (int c = 0; c < 3; c++) { sg->cmin[c] = h[SEGP_MIN + c]; sg->cmax[c] = h[SEGP_MAX + c]; }
    sg->sharded_ok = h[SEGP_FOREIGN] == 0;  // key compaction is only sound when every subject is ours
    sg->pred_rows.clear();
    sg->preds_overflow = h[SEGP_OVERFLOW] != 0;
    sg->has_preds = true;
    if (!sg->preds_overflow) {
        for (u32 i = 0; i < SEGP_SLOTS; i++) if (h[SEGP_SLOT + i] != EMPTY32) sg->pred_rows.push_back({h[SEGP_SLOT + i], h[SEGP_COUNT + i]});
        std::sort(sg->pred_rows.begin(), sg->pred_rows.end());
    }
    sg->stats_world = ctx->shard_world;
    sg->has_stats = true;
    return KB_OK;
}

kb_status unpair_rel(kb_ctx* ctx, std::unique_ptr<kb_rel>* r) {
    if (!(*r)->pair) return KB_OK;
    auto c = std::make_unique<kb_rel>();
    c->slots = (*r)->slots;
    c->n = (*r)->n;
    Col x, y;
    KB_TRY(alloc_col(ctx, c->n, &x));
    KB_TRY(alloc_col(ctx, c->n, &y));
    timer_begin(ctx, F_OTHER);
    launch_unpair(reinterpret_cast<const uint2*>((*r)->cols[0].ptr), (u32)c->n, x.ptr, y.ptr, ctx->st);
    timer_end(ctx);
    KB_CUDA(ctx, cudaGetLastError());
    c->cols = {x, y};
    *r = std::move(c);
    return KB_OK;
}

// IndexScan proper (engine.rs:1192-1407, the 8-case lookup of index_manager.rs:253-340 for a constant predicate): with a valid store
// index a pattern whose predicate is constant is answered from the predicate's slice — (c P ?o) from the subject table / directory,
// (?s P c) from the object table / directory, (?s P ?o) is the slice itself — without touching the 12-byte/triple store.
// *out stays null when the index cannot answer the pattern (the caller scans).
static kb_status index_lookup(kb_ctx* ctx, const kb_pattern& pt, std::unique_ptr<kb_rel>* out) {
    out->reset();
    if (!ctx->use_index || ctx->index_version != ctx->store_version || pt.p.is_var) return KB_OK;
    std::vector<u32> slots, src;
    pattern_vars(pt, &slots, &src);
    auto it = ctx->index.find(pt.p.value);
    auto rel = std::make_unique<kb_rel>();
    rel->slots = slots;
    if (it == ctx->index.end() || it->second.n == 0) {  // the predicate does not occur: empty answer
        for (size_t c = 0; c < slots.size(); c++) { Col col; KB_TRY(alloc_col(ctx, 0, &col)); rel->cols.push_back(col); }
        *out = std::move(rel);
        return KB_OK;
    }
    const PredSlice& ps = it->second;
    if (pt.s.is_var && pt.o.is_var) {
        if (pt.s.value == pt.o.value) return KB_OK;  // ?x P ?x: the scan enforces the equality
        Col x, y;  // the slice itself, de-interleaved (all chunks)
        KB_TRY(alloc_col(ctx, ps.n, &x));
        KB_TRY(alloc_col(ctx, ps.n, &y));
        u64 at = 0;
        timer_begin(ctx, F_SCAN, (int)ps.chunks.size());
        for (auto& ch : ps.chunks) {
            // (chunk boundaries keep 16-byte alignment only by luck: unpair writes element-wise, no TMA on the output side)
            launch_unpair(reinterpret_cast<const uint2*>(ch.pairs.ptr), (u32)ch.n, x.ptr + at, y.ptr + at, ctx->st);
            at += ch.n;
        }
        timer_end(ctx);
        KB_CUDA(ctx, cudaGetLastError());
        rel->cols = {x, y};
        rel->n = ps.n;
        ctx->stats.index_joins++;
        *out = std::move(rel);
        return KB_OK;
    }
    if (pt.s.is_var == pt.o.is_var) return KB_OK;  // both constant: left to the scan
    const bool by_y = !pt.o.is_var;                // the bound position is the object
    const u32 key = by_y ? pt.o.value : pt.s.value;
    const Buf& tab = by_y ? ps.ytab : ps.xtab;
    const Buf& off = by_y ? ps.yoff : ps.xoff;
    const Buf& val = by_y ? ps.yval : ps.xval;
    if (tab) {  // unique column: one table slot
        const u32 cs = by_y ? 0u : ps.tab_cshift;
        if (cs != 0u && shard_of(key, ctx->shard_world) != ctx->shard_rank) {  // a key of another shard: not in this store
            Col col; KB_TRY(alloc_col(ctx, 0, &col)); rel->cols.push_back(col);
            *out = std::move(rel);
            return KB_OK;
        }
        const u32 o = compact_key(key, cs) - (by_y ? ps.ytab_min : ps.xtab_min);
        u32 v = EMPTY32;
        if (o < (by_y ? ps.ytab_range : ps.xtab_range)) {
            KB_CUDA(ctx, cudaMemcpyAsync(&v, static_cast<const u32*>(tab->p) + o, sizeof(u32), cudaMemcpyDeviceToHost, ctx->st));
            KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
            ctx->stats.d2h_bytes += 4;
        }
        Col col;
        KB_TRY(alloc_col(ctx, v == EMPTY32 ? 0 : 1, &col));
        if (v != EMPTY32) KB_CUDA(ctx, cudaMemcpyAsync(col.ptr, &v, sizeof(u32), cudaMemcpyHostToDevice, ctx->st));
        KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
        rel->cols.push_back(col);
        rel->n = v == EMPTY32 ? 0 : 1;
        ctx->stats.index_joins++;
        *out = std::move(rel);
        return KB_OK;
    }
    if (off && val) {  // multi-valued column: the key's run in the directory
        const u32 kmin = by_y ? ps.ycsr_min : ps.xcsr_min, range = by_y ? ps.ycsr_range : ps.xcsr_range;
        u32 be[2] = {0, 0};
        if (key >= kmin && key - kmin < range) {
            KB_CUDA(ctx, cudaMemcpyAsync(be, static_cast<const u32*>(off->p) + (key - kmin), 2 * sizeof(u32), cudaMemcpyDeviceToHost, ctx->st));
            KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
            ctx->stats.d2h_bytes += 8;
        }
        const u64 n = be[1] - be[0];
        Col col;
        if ((be[0] & 3u) == 0 && n) {  // 16-byte aligned run: a view of the directory's values, no copy
            col.buf = val;
            col.ptr = static_cast<u32*>(val->p) + be[0];
        } else {
            KB_TRY(alloc_col(ctx, n, &col));
            if (n) KB_CUDA(ctx, cudaMemcpyAsync(col.ptr, static_cast<const u32*>(val->p) + be[0], n * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->st));
        }
        rel->cols.push_back(col);
        rel->n = n;
        ctx->stats.index_joins++;
        *out = std::move(rel);
        return KB_OK;
    }
    return KB_OK;
}

kb_status scan_impl(kb_ctx* ctx, const kb_pattern* pats, u32 K, const std::vector<FilterProg>& pushdown, bool want_index, bool pairs,
                    std::vector<std::unique_ptr<kb_rel>>* out, const std::vector<ScanTable>* tables) {
    if (K == 0 || K > (u32)MAXP) return fail(ctx, KB_E_LIMIT, "a fused scan takes 1..%d patterns (got %u)", MAXP, K);
    if (!want_index && !pairs && !tables && ctx->use_index && ctx->index_version == ctx->store_version && !ctx->in_index_build) {
        // patterns the index answers are not scanned; the rest (if any) go through one fused scan
        std::vector<std::unique_ptr<kb_rel>> got(K);
        std::vector<u32> rest;
        for (u32 k = 0; k < K; k++) {
            KB_TRY(check_pattern(ctx, pats[k]));
            KB_TRY(index_lookup(ctx, pats[k], &got[k]));
            if (!got[k]) { rest.push_back(k); continue; }
            if (k < pushdown.size() && !pushdown[k].ops.empty()) {  // the pattern's own FILTER, applied to the looked-up rows
                std::unique_ptr<kb_rel> f;
                KB_TRY(filter_impl(ctx, *got[k], pushdown[k], &f));
                got[k] = std::move(f);
            }
        }
        if (rest.size() < K) {
            if (!rest.empty()) {
                std::vector<kb_pattern> rp;
                std::vector<FilterProg> rf;
                for (u32 k : rest) { rp.push_back(pats[k]); rf.push_back(k < pushdown.size() ? pushdown[k] : FilterProg{}); }
                std::vector<std::unique_ptr<kb_rel>> sub;
                ctx->in_index_build = true;  // (re-entrancy guard: the remaining patterns are scanned, not looked up again)
                const kb_status rc = scan_impl(ctx, rp.data(), (u32)rp.size(), rf, false, false, &sub, nullptr);
                ctx->in_index_build = false;
                if (rc != KB_OK) return rc;
                for (size_t i = 0; i < rest.size(); i++) got[rest[i]] = std::move(sub[i]);
            }
            out->clear();
            for (u32 k = 0; k < K; k++) out->push_back(std::move(got[k]));
            return KB_OK;
        }
    }
    const u64 N = ctx->n_triples;
    if (N >= 0xFFFFFFF0ull) return fail(ctx, KB_E_LIMIT, "store holds %llu triples; row positions are 32-bit", (unsigned long long)N);
    ScanParams P{};
    P.K = K;
    P.nt = numtab(ctx);
    std::vector<FilterOp> ops;
    out->clear();
    for (u32 k = 0; k < K; k++) {
        KB_TRY(check_pattern(ctx, pats[k]));
        const kb_pattern& pt = pats[k];
        ScanPat& sp = P.pat[k];
        sp.flags = 0;
        if (!pt.s.is_var) { sp.flags |= SP_HAS_S; sp.cs = pt.s.value; }
        if (!pt.p.is_var) { sp.flags |= SP_HAS_P; sp.cp = pt.p.value; }
        if (!pt.o.is_var) { sp.flags |= SP_HAS_O; sp.co = pt.o.value; }
        if (pt.s.is_var && pt.p.is_var && pt.s.value == pt.p.value) sp.flags |= SP_EQ_SP;
        if (pt.s.is_var && pt.o.is_var && pt.s.value == pt.o.value) sp.flags |= SP_EQ_SO;
        if (pt.p.is_var && pt.o.is_var && pt.p.value == pt.o.value) sp.flags |= SP_EQ_PO;
        std::vector<u32> slots, src;
        pattern_vars(pt, &slots, &src);
        auto rel = std::make_unique<kb_rel>();
        if (want_index) {
            slots.assign(1, 0u);
            src.assign(1, 3u);
        }
        rel->slots = slots;
        for (int q = 0; q < 4; q++) sp.outp[q] = nullptr;
        if (tables && k < tables->size() && (*tables)[k].tab) {  // fused build: no output relation, the table is the output
            const ScanTable& tb = (*tables)[k];
            rel->pair = true;
            sp.outp[0] = tb.tab;
            sp.outp[1] = tb.dup_flag;
            sp.cs = tb.kmin;
            sp.co = tb.range;
            sp.flags |= SP_TABLE | (tb.key_is_o ? SP_TKEY_O : 0u) | (tb.trusted ? SP_TTRUSTED : 0u);
            P.cshift = tb.cshift;
        } else if (pairs && !want_index && slots.size() == 2 && src[0] == 0 && src[1] == 2) {
            Col col;  // interleaved (subject, object) rows: one 8-byte store per match
            KB_TRY(alloc_col(ctx, 2 * N, &col));
            rel->cols.push_back(col);
            rel->pair = true;
            sp.outp[0] = col.ptr;
            sp.flags |= SP_PAIR;
        } else {
            for (size_t c = 0; c < slots.size(); c++) {  // columns come out in s,p,o order = order of first appearance
                Col col;
                KB_TRY(alloc_col(ctx, N, &col));
                rel->cols.push_back(col);
                sp.outp[src[c]] = col.ptr;
                sp.flags |= src[c] == 0 ? SP_EMIT_S : src[c] == 1 ? SP_EMIT_P : src[c] == 2 ? SP_EMIT_O : SP_EMIT_IDX;
            }
        }
        sp.f_begin = (u32)ops.size();
        sp.f_len = 0;
        if (k < pushdown.size() && !pushdown[k].ops.empty()) {
            std::map<u32, u32> remap;
            std::vector<u32> vs, vsrc;
            pattern_vars(pt, &vs, &vsrc);
            for (size_t i = 0; i < vs.size(); i++) remap[vs[i]] = vsrc[i];
            std::vector<FilterOp> local;
            if (!append_prog(&local, pushdown[k], remap)) return fail(ctx, KB_E_INVALID, "pushed-down filter of pattern %u uses a variable the pattern does not bind", k);
            if (ops.size() + local.size() > KB_MAX_FILTER_OPS) return fail(ctx, KB_E_LIMIT, "pushed-down filters exceed %d ops", KB_MAX_FILTER_OPS);
            ops.insert(ops.end(), local.begin(), local.end());
            sp.f_len = (u32)local.size();
        }
        out->push_back(std::move(rel));
    }
    for (size_t i = 0; i < ops.size(); i++) P.ops[i] = ops[i];

    const u32 n_seg = (u32)ctx->segs.size();
    // the guard depends on the segment count alone: where the arena cursor stands is irrelevant, ctrl_alloc wraps (an index build
    // over hundreds of predicates hands out ~76 words per 8-predicate batch and every consumer clears its own words)
    if (std::max(n_seg, 1u) + 2u * MAXP > kb_ctx::CTRL_WORDS - 64u) return fail(ctx, KB_E_LIMIT, "too many store segments (%u)", n_seg);
    const u32 off_tot = ctrl_alloc(ctx, 2 * MAXP);
    const u32 off_ticket = ctrl_alloc(ctx, std::max(n_seg, 1u));
    KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off_tot, 0, 2 * MAXP * sizeof(u32), ctx->st));
    KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off_ticket, 0, std::max(n_seg, 1u) * sizeof(u32), ctx->st));
    u64 all_tiles = 0;
    for (auto& sg : ctx->segs) all_tiles += (sg.n + SCAN_TILE - 1) / SCAN_TILE;
    KB_TRY(ensure_tile_state(ctx, all_tiles * 4));  // the star-shape kernel may walk the store in smaller tiles than SCAN_TILE
    P.tile_state = static_cast<u64*>(ctx->tile_state->p);
    P.block_state = static_cast<u64*>(ctx->block_state->p);
    P.ordered = ctx->ordered;
    // One launch walks up to SCAN_MAXSEG consecutive segments (an RSP window is ~10 slides). A segment whose upload is still in
    // flight (chunked kb_star_join_host) closes the group before it, so the scan of the chunks that already landed can start.
    u64 index_base = 0;
    u32 n_launched = 0;
    u32 g = 0;
    // tables that still have to be set to 0xFF: by the first launch itself when it is the star-shape kernel, else by memsets here
    std::vector<std::pair<u32*, u32>> pending_clear;
    if (tables) for (auto& tb : *tables) if (tb.tab && tb.clear) pending_clear.push_back({tb.tab, tb.range});
    u32 off_clear = 0;
    bool clear_folded = false;
    while (g < n_seg) {
        P.n_seg = 0;
        P.n_tiles = 0;
        cudaEvent_t wait_ev = nullptr;
        while (g < n_seg && P.n_seg < (u32)SCAN_MAXSEG) {
            const Segment& sg = ctx->segs[g];
            if (sg.n == 0) { g++; continue; }
            if (sg.ready && P.n_seg > 0) break;  // starts its own group
            ScanSeg& d = P.seg[P.n_seg++];
            d.s = sg.s.ptr; d.p = sg.p.ptr; d.o = sg.o.ptr;
            d.n = (u32)sg.n;
            d.tile0 = P.n_tiles;
            d.index_base = (u32)index_base;
            P.n_tiles += (u32)((sg.n + SCAN_TILE - 1) / SCAN_TILE);
            index_base += sg.n;
            g++;
            if (sg.ready) { wait_ev = sg.ready; break; }
        }
        if (P.n_seg == 0) break;
        P.ticket = ctx->ctrl + off_ticket + n_launched;
        if (ctx->ordered) {  // ping-pong: a launch never writes the word its tiles read their base from
            P.totals_in = ctx->ctrl + off_tot + (n_launched & 1u) * MAXP;
            P.totals_out = ctx->ctrl + off_tot + ((n_launched + 1u) & 1u) * MAXP;
        } else {
            P.totals_in = P.totals_out = ctx->ctrl + off_tot;
        }
        n_launched++;
        P.epoch = ctx->epoch++;
        if (ctx->epoch >= (1ull << 30)) ctx->epoch = 1;
        P.n_clear = 0;
        if (!pending_clear.empty()) {
            if (pending_clear.size() <= 4 && scan_clears_tables(P)) {
                off_clear = ctrl_alloc(ctx, 4);
                KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off_clear, 0, 4 * sizeof(u32), ctx->st));
                for (auto& pc : pending_clear) { P.clear_tab[P.n_clear] = pc.first; P.clear_words[P.n_clear] = pc.second; P.n_clear++; }
                P.clear_barrier = ctx->ctrl + off_clear;
                clear_folded = true;
            } else {
                timer_begin(ctx, F_BUILD, 0);
                for (auto& pc : pending_clear) KB_CUDA(ctx, cudaMemsetAsync(pc.first, 0xFF, (size_t)pc.second * sizeof(u32), ctx->st));
                timer_end(ctx);
            }
            pending_clear.clear();
        }
        if (wait_ev) KB_CUDA(ctx, cudaStreamWaitEvent(ctx->st, wait_ev, 0));  // chunked upload: start as soon as this chunk landed
        timer_begin(ctx, F_SCAN);
        launch_scan(P, ctx->n_sms, ctx->st);
        timer_end(ctx);
    }
    for (auto& pc : pending_clear) KB_CUDA(ctx, cudaMemsetAsync(pc.first, 0xFF, (size_t)pc.second * sizeof(u32), ctx->st));  // (nothing was launched)
    KB_CUDA(ctx, cudaGetLastError());
    ctx->stats.rows_scanned += N;
    KB_TRY(ctrl_read(ctx));
    if (clear_folded && ctx->h_ctrl[off_clear + 1]) return fail(ctx, KB_E_CUDA, "scan: the grid did not meet at the table-clear barrier (another kernel holds SMs?)");
    for (u32 k = 0; k < K; k++) (*out)[k]->n = ctx->h_ctrl[off_tot + (ctx->ordered ? (n_launched & 1u) * MAXP : 0u) + k];
    return KB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// filter (= probe kernel with zero tables)
kb_status filter_impl(kb_ctx* ctx, const kb_rel& in, const FilterProg& f, std::unique_ptr<kb_rel>* out) {
    auto rel = std::make_unique<kb_rel>();
    rel->slots = in.slots;
    if (in.n == 0 || f.ops.empty()) {
        rel->cols = in.cols;
        rel->n = in.n;
        *out = std::move(rel);
        return KB_OK;
    }
    if (in.n >= 0xFFFFFFF0ull) return fail(ctx, KB_E_LIMIT, "relation too large");
    ProbeDParams P{};
    P.n_pcols = (u32)in.cols.size();
    if (P.n_pcols == 0) return fail(ctx, KB_E_UNSUPPORTED, "filter on a relation without columns");
    P.key_col = 0;
    P.n = (u32)in.n;
    P.n_tiles = (u32)((in.n + PROBE_TILE - 1) / PROBE_TILE);
    P.T = 0;
    P.n_out = P.n_pcols;
    std::map<u32, u32> remap;
    for (u32 c = 0; c < P.n_pcols; c++) {
        P.pcol[c] = in.cols[c].ptr;
        P.oc[c] = OutCol{OUT_PROBE, c, 0};
        Col col;
        KB_TRY(alloc_col(ctx, in.n, &col));
        rel->cols.push_back(col);
        P.out[c] = col.ptr;
        remap[in.slots[c]] = c;
    }
    std::vector<FilterOp> ops;
    if (!append_prog(&ops, f, remap)) return fail(ctx, KB_E_INVALID, "filter uses a variable that is not a column of the relation");
    P.n_ops = (u32)ops.size();
    for (size_t i = 0; i < ops.size(); i++) P.ops[i] = ops[i];
    P.cap = (u32)in.n;
    P.nt = numtab(ctx);
    KB_TRY(ensure_tile_state(ctx, P.n_tiles));
    P.tile_state = static_cast<u64*>(ctx->tile_state->p);
    P.block_state = static_cast<u64*>(ctx->block_state->p);
    P.ordered = ctx->ordered;
    const u32 off = ctrl_alloc(ctx, 4);
    KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off, 0, 4 * sizeof(u32), ctx->st));
    P.ticket = ctx->ctrl + off;
    P.total = ctx->ctrl + off + 1;
    P.zero_word = ctx->ctrl + off + 2;
    P.epoch = ctx->epoch++;
    P.abort_flag = nullptr;
    timer_begin(ctx, F_FILTER);
    launch_probe_direct(P, ctx->n_sms, ctx->st);
    timer_end(ctx);
    KB_CUDA(ctx, cudaGetLastError());
    KB_TRY(ctrl_read(ctx));
    rel->n = ctx->h_ctrl[off + 1];
    *out = std::move(rel);
    return KB_OK;
}

u32 pow2_at_least(u64 x) {
    u64 p = 1024;
    while (p < x) p <<= 1;
    return (u32)std::min<u64>(p, 1ull << 31);
}

// ---------------------------------------------------------------------------------------------------------------
// general natural join (chained multimap)
kb_status hash_join_impl(kb_ctx* ctx, const kb_rel& L, const kb_rel& R, const FilterProg* post, std::unique_ptr<kb_rel>* out) {
    auto rel = std::make_unique<kb_rel>();
    rel->slots = L.slots;
    std::vector<u32> common;
    for (u32 s : R.slots) {
        if (L.col_of(s) >= 0) common.push_back(s);
        else rel->slots.push_back(s);
    }
    if (rel->slots.size() > KB_MAX_COLS) return fail(ctx, KB_E_LIMIT, "join result has more than %d columns", KB_MAX_COLS);
    if (L.n == 0 || R.n == 0) {  // engine.rs:714-716
        rel->n = 0;
        for (size_t c = 0; c < rel->slots.size(); c++) { Col col; KB_TRY(alloc_col(ctx, 0, &col)); rel->cols.push_back(col); }
        *out = std::move(rel);
        return KB_OK;
    }
    if (L.n >= 0xFFFFFFF0ull || R.n >= 0xFFFFFFF0ull) return fail(ctx, KB_E_LIMIT, "relation too large");
    if (common.empty()) {  // cartesian product (engine.rs:1054-1071)
        const u64 total = L.n * R.n;
        if (total > (1ull << 28)) return fail(ctx, KB_E_LIMIT, "cartesian product of %llu x %llu rows exceeds 2^28", (unsigned long long)L.n, (unsigned long long)R.n);
        std::vector<const u32*> lc, rc;
        std::vector<u32*> oc;
        for (auto& c : L.cols) lc.push_back(c.ptr);
        for (auto& c : R.cols) rc.push_back(c.ptr);
        for (size_t c = 0; c < rel->slots.size(); c++) { Col col; KB_TRY(alloc_col(ctx, total, &col)); rel->cols.push_back(col); oc.push_back(col.ptr); }
        timer_begin(ctx, F_PROBE);
        launch_cartesian(lc.data(), (u32)L.n, (u32)lc.size(), rc.data(), (u32)R.n, (u32)rc.size(), oc.data(), ctx->st);
        timer_end(ctx);
        KB_CUDA(ctx, cudaGetLastError());
        rel->n = total;
        if (post && !post->ops.empty()) {
            std::unique_ptr<kb_rel> f;
            KB_TRY(filter_impl(ctx, *rel, *post, &f));
            rel = std::move(f);
        }
        *out = std::move(rel);
        return KB_OK;
    }
    if (common.size() > 4) return fail(ctx, KB_E_LIMIT, "join on more than 4 common variables");
    const bool build_left = L.n <= R.n;  // engine.rs:729-733: build on the smaller side
    const kb_rel& B = build_left ? L : R;
    const kb_rel& Pr = build_left ? R : L;

    // ---- grouped (CSR) join: one key column with a dense id range and no join filter
    if (ctx->csr_join && common.size() == 1 && (!post || post->ops.empty())) {
        const u32* bk = B.cols[B.col_of(common[0])].ptr;
        const u32 off = ctrl_alloc(ctx, 8);  // [0] min, [1] max, [2..3] u64 output rows, [4] ticket, [5] total, [6] zero, [7] heavy tiles
        KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off, 0xFF, sizeof(u32), ctx->st));
        KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off + 1, 0, 7 * sizeof(u32), ctx->st));
        timer_begin(ctx, F_BUILD);
        launch_col_minmax(bk, (u32)B.n, ctx->ctrl + off, ctx->ctrl + off + 1, ctx->n_sms, ctx->st);
        timer_end(ctx);
        KB_TRY(ctrl_read(ctx));
        const u32 kmin = ctx->h_ctrl[off], kmax = ctx->h_ctrl[off + 1];
        const u64 range = (u64)kmax - kmin + 1;
        if (kmax >= kmin && range <= std::max<u64>(16 * B.n, 1ull << 20) && range < (1ull << 31)) {
            Buf dir, cursor, scratch;
            KB_TRY(alloc_buf(ctx, (range + 1) * sizeof(u32), &dir));
            KB_TRY(alloc_buf(ctx, range * sizeof(u32), &cursor));
            KB_TRY(alloc_buf(ctx, ((range + 1) / 2048 + 4) * sizeof(u32), &scratch));
            CsrTab tab{};
            tab.off = static_cast<const u32*>(dir->p);
            tab.kmin = kmin;
            tab.range = (u32)range;
            std::vector<u32> kernel_slots = Pr.slots;  // kernel output order: probe columns, then build payload columns
            std::vector<const u32*> pay_in;
            std::vector<u32*> pay_out;
            std::vector<Col> pay_cols;
            for (size_t c = 0; c < B.slots.size(); c++) {
                if (B.slots[c] == common[0]) continue;
                Col pc;
                KB_TRY(alloc_col(ctx, B.n, &pc));
                pay_cols.push_back(pc);
                pay_in.push_back(B.cols[c].ptr);
                pay_out.push_back(pc.ptr);
                tab.pay[tab.n_pay++] = pc.ptr;
                kernel_slots.push_back(B.slots[c]);
            }
            timer_begin(ctx, F_BUILD, 5);
            KB_CUDA(ctx, cudaMemsetAsync(dir->p, 0, (range + 1) * sizeof(u32), ctx->st));
            launch_csr_count(bk, (u32)B.n, kmin, static_cast<u32*>(dir->p), ctx->n_sms, ctx->st);
            launch_exclusive_scan_u32(static_cast<u32*>(dir->p), (u32)range + 1, static_cast<u32*>(scratch->p), ctx->st);
            KB_CUDA(ctx, cudaMemcpyAsync(cursor->p, dir->p, range * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->st));
            launch_csr_fill(bk, (u32)B.n, kmin, static_cast<u32*>(cursor->p), pay_in.data(), pay_out.data(), tab.n_pay, ctx->n_sms, ctx->st);
            timer_end(ctx);
            ctx->stats.rows_built += B.n;
            const u32* pk = Pr.cols[Pr.col_of(common[0])].ptr;
            timer_begin(ctx, F_PROBE);
            launch_csr_total(pk, (u32)Pr.n, tab, reinterpret_cast<unsigned long long*>(ctx->ctrl + off + 2), ctx->n_sms, ctx->st);
            timer_end(ctx);
            KB_CUDA(ctx, cudaGetLastError());
            KB_TRY(ctrl_read(ctx));
            unsigned long long total = 0;
            memcpy(&total, ctx->h_ctrl + off + 2, sizeof total);
            if (total >= 0xFFFFFFF0ull) return fail(ctx, KB_E_LIMIT, "join result of %llu rows exceeds 2^32", total);
            ProbeGParams G{};
            G.n_pcols = (u32)Pr.cols.size();
            G.n = (u32)Pr.n;
            G.n_tiles = (u32)((Pr.n + PROBEG_TILE - 1) / PROBEG_TILE);
            for (u32 c = 0; c < G.n_pcols; c++) G.pcol[c] = Pr.cols[c].ptr;
            G.pkey = (u32)Pr.col_of(common[0]);
            G.tab = tab;
            rel->cols.clear();
            rel->cols.resize(rel->slots.size());
            for (size_t c = 0; c < kernel_slots.size(); c++) {
                Col col;
                KB_TRY(alloc_col(ctx, total, &col));
                rel->cols[rel->col_of(kernel_slots[c])] = col;
                G.out[c] = col.ptr;
            }
            G.cap = (u32)total;
            KB_TRY(ensure_tile_state(ctx, G.n_tiles));
            G.tile_state = static_cast<u64*>(ctx->tile_state->p);
            G.block_state = static_cast<u64*>(ctx->block_state->p);
            G.ordered = ctx->ordered;
            G.ticket = ctx->ctrl + off + 4;
            G.total = ctx->ctrl + off + 5;
            G.zero_word = ctx->ctrl + off + 6;
            G.epoch = ctx->epoch++;
            Buf heavy;  // tiles whose expansion is spread over the grid by the second launch (skewed keys)
            if (total > PROBEG_HEAVY) {
                KB_TRY(alloc_buf(ctx, (size_t)G.n_tiles * sizeof(ProbeGParams::Heavy), &heavy));
                G.heavy = static_cast<ProbeGParams::Heavy*>(heavy->p);
                G.heavy_count = ctx->ctrl + off + 7;
            }
            ctx->stats.rows_probed += Pr.n;
            if (total) {
                timer_begin(ctx, F_PROBE, G.heavy ? 2 : 1);
                launch_probe_grouped(G, ctx->n_sms, ctx->st);
                timer_end(ctx);
                KB_CUDA(ctx, cudaGetLastError());
            }
            rel->n = total;
            // dir / cursor / payload copies are released stream-ordered (cudaFreeAsync on ctx->st), i.e. after the launch above
            *out = std::move(rel);
            return KB_OK;
        }
    }

    ChainTab T{};
    T.n_slots = pow2_at_least(B.n * 2);
    T.n_keys = (u32)common.size();
    Buf slots_buf, next_buf;
    KB_TRY(alloc_buf(ctx, (size_t)T.n_slots * sizeof(u64), &slots_buf));
    KB_TRY(alloc_buf(ctx, (size_t)B.n * sizeof(u32), &next_buf));
    T.slots = static_cast<u64*>(slots_buf->p);
    T.next = static_cast<u32*>(next_buf->p);
    for (u32 q = 0; q < T.n_keys; q++) T.bkey[q] = B.cols[B.col_of(common[q])].ptr;
    timer_begin(ctx, F_BUILD);
    KB_CUDA(ctx, cudaMemsetAsync(T.slots, 0xFF, (size_t)T.n_slots * sizeof(u64), ctx->st));
    launch_build_chained(T, (u32)B.n, ctx->n_sms, ctx->st);
    timer_end(ctx);
    ctx->stats.rows_built += B.n;

    ProbeCParams P{};
    P.tab = T;
    P.n_pcols = (u32)Pr.cols.size();
    P.n = (u32)Pr.n;
    P.n_tiles = (u32)((Pr.n + PROBEC_TILE - 1) / PROBEC_TILE);
    for (u32 c = 0; c < P.n_pcols; c++) P.pcol[c] = Pr.cols[c].ptr;
    for (u32 q = 0; q < T.n_keys; q++) P.pkey[q] = (u32)Pr.col_of(common[q]);
    std::vector<u32> kernel_slots = Pr.slots;  // kernel output order: probe columns, then build payload columns
    P.n_bpay = 0;
    for (size_t c = 0; c < B.slots.size(); c++) {
        if (std::find(common.begin(), common.end(), B.slots[c]) != common.end()) continue;
        P.bpay[P.n_bpay++] = B.cols[c].ptr;
        kernel_slots.push_back(B.slots[c]);
    }
    P.n_out = (u32)kernel_slots.size();
    if (post && !post->ops.empty()) {
        std::map<u32, u32> remap;
        for (u32 c = 0; c < P.n_out; c++) remap[kernel_slots[c]] = c;
        std::vector<FilterOp> ops;
        if (!append_prog(&ops, *post, remap)) return fail(ctx, KB_E_INVALID, "join filter uses a variable that neither side binds");
        P.n_ops = (u32)ops.size();
        for (size_t i = 0; i < ops.size(); i++) P.ops[i] = ops[i];
    }
    P.nt = numtab(ctx);
    KB_TRY(ensure_tile_state(ctx, P.n_tiles));
    P.tile_state = static_cast<u64*>(ctx->tile_state->p);
    P.block_state = static_cast<u64*>(ctx->block_state->p);
    P.ordered = ctx->ordered;
    const u32 off = ctrl_alloc(ctx, 8);  // [0] ticket, [1] total (32-bit positions), [2] zero, [4..5] exact total in 64 bits
    P.total = ctx->ctrl + off + 1;
    P.zero_word = ctx->ctrl + off + 2;
    P.total64 = reinterpret_cast<unsigned long long*>(ctx->ctrl + off + 4);
    u64 cap = std::max(Pr.n, B.n);
    ctx->stats.rows_probed += Pr.n;
    for (int attempt = 0; attempt < 2; attempt++) {
        rel->cols.clear();
        rel->cols.resize(rel->slots.size());
        for (u32 c = 0; c < P.n_out; c++) {
            Col col;
            KB_TRY(alloc_col(ctx, cap, &col));
            rel->cols[rel->col_of(kernel_slots[c])] = col;
            P.out[c] = col.ptr;
        }
        P.cap = (u32)std::min<u64>(cap, 0xFFFFFFF0ull);
        KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off, 0, 8 * sizeof(u32), ctx->st));
        P.ticket = ctx->ctrl + off;
        P.epoch = ctx->epoch++;
        timer_begin(ctx, F_PROBE);
        launch_probe_chained(P, ctx->n_sms, ctx->st);
        timer_end(ctx);
        KB_CUDA(ctx, cudaGetLastError());
        KB_TRY(ctrl_read(ctx));
        unsigned long long total = 0;
        memcpy(&total, ctx->h_ctrl + off + 4, sizeof total);
        if (total >= 0xFFFFFFF0ull) return fail(ctx, KB_E_LIMIT, "join result of %llu rows exceeds 2^32", total);
        rel->n = total;
        if (total <= cap) break;
        if (attempt == 1) return fail(ctx, KB_E_LIMIT, "join output did not fit after resizing");
        cap = total;  // the first pass doubled as the exact count
    }
    *out = std::move(rel);
    return KB_OK;
}

// reorder / select columns (zero-copy)
static std::unique_ptr<kb_rel> select_cols(const kb_rel& in, const std::vector<u32>& slots) {
    auto r = std::make_unique<kb_rel>();
    r->n = in.n;
    for (u32 s : slots) {
        int c = in.col_of(s);
        if (c < 0) continue;
        r->slots.push_back(s);
        r->cols.push_back(in.cols[c]);
    }
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// star join: fused scan -> direct builds -> fused multiway probe; falls back to chained binary joins on 1:N keys
kb_status star_join_impl(kb_ctx* ctx, u32 join_slot, const kb_pattern* pats, u32 K, const kb_filter_op* filter, u32 n_ops,
                         std::unique_ptr<kb_rel>* out) {
    return star_join_impl2(ctx, join_slot, pats, K, filter, n_ops, true, out);
}

kb_status star_join_impl2(kb_ctx* ctx, u32 join_slot, const kb_pattern* pats, u32 K, const kb_filter_op* filter, u32 n_ops, bool allow_fused_scan,
                          std::unique_ptr<kb_rel>* out, AggSpec* agg, IndexPlan* plan_only) {
    if (K == 0 || K > (u32)MAXP) return fail(ctx, KB_E_LIMIT, "a star join takes 1..%d patterns (got %u)", MAXP, K);
    KB_TRY(validate_filter(ctx, filter, n_ops));
    std::vector<std::vector<u32>> pv(K), psrc(K);
    std::vector<u32> all_slots;
    for (u32 k = 0; k < K; k++) {
        KB_TRY(check_pattern(ctx, pats[k]));
        pattern_vars(pats[k], &pv[k], &psrc[k]);
        if (std::find(pv[k].begin(), pv[k].end(), join_slot) == pv[k].end())
            return fail(ctx, KB_E_INVALID, "star join: pattern %u does not contain the join variable (slot %u)", k, join_slot);
        for (u32 s : pv[k]) if (std::find(all_slots.begin(), all_slots.end(), s) == all_slots.end()) all_slots.push_back(s);
    }
    if (all_slots.size() > KB_MAX_COLS) return fail(ctx, KB_E_LIMIT, "more than %d variables", KB_MAX_COLS);
    // FILTER: conjuncts over one pattern's variables are evaluated while scanning that pattern
    std::vector<FilterProg> conj, pushdown(K);
    FilterProg post;
    if (!split_conjuncts(filter, n_ops, &conj)) return fail(ctx, KB_E_INVALID, "malformed filter program");
    for (auto& c : conj) {
        std::set<u32> fs = filter_slots(c);
        for (u32 s : fs) if (std::find(all_slots.begin(), all_slots.end(), s) == all_slots.end())
            return fail(ctx, KB_E_UNSUPPORTED, "filter references slot %u which no pattern binds (the reference evaluates it to false, types.rs:149-151)", s);
        int target = -1;
        for (u32 k = 0; k < K && target < 0; k++) {
            bool all = true;
            for (u32 s : fs) if (std::find(pv[k].begin(), pv[k].end(), s) == pv[k].end()) all = false;
            if (all) target = (int)k;
        }
        FilterProg* dst = target >= 0 ? &pushdown[target] : &post;
        const bool had = !dst->ops.empty();
        dst->ops.insert(dst->ops.end(), c.ops.begin(), c.ops.end());
        if (had) { kb_filter_op a{}; a.op = KB_F_AND; dst->ops.push_back(a); }
    }
    // key range of the join variable from the load-time column statistics of the store (its position is the same in every
    // pattern that can take the fused path)
    auto key_pos = [&](u32 k) { for (size_t i = 0; i < pv[k].size(); i++) if (pv[k][i] == join_slot) return psrc[k][i]; return 0u; };
    // fast path shape: every pattern binds exactly two variables in subject and object position (the join variable + one more)
    bool all_pairs = K >= 2;
    for (u32 k = 0; k < K; k++) if (!(pv[k].size() == 2 && psrc[k][0] == 0 && psrc[k][1] == 2)) all_pairs = false;

    // ---- INDEX path: every pattern is (?s P ?o) over a predicate the store index has a slice for: no store scan at all. Build sides
    // are inserted from their slices (pushed-down FILTER evaluated in the build kernel), the probe side IS its slice (zero copy).
    if (allow_fused_scan && ctx->use_index && ctx->index_version == ctx->store_version && all_pairs && K - 1 <= (u32)MAXT) {
        bool ok = true, empty = false;
        std::vector<const PredSlice*> sl(K, nullptr);
        for (u32 k = 0; k < K; k++) {
            if (pats[k].p.is_var) { ok = false; break; }
            auto it = ctx->index.find(pats[k].p.value);
            if (it == ctx->index.end() || it->second.n == 0) { empty = true; continue; }  // predicate absent from the store
            sl[k] = &it->second;
        }
        for (u32 a = 0; a < K && ok; a++)
            for (u32 b = a + 1; b < K && ok; b++)
                for (u32 s2 : pv[a]) if (s2 != join_slot && std::find(pv[b].begin(), pv[b].end(), s2) != pv[b].end()) ok = false;
        if (ok && empty && plan_only) return fail(ctx, KB_E_UNSUPPORTED, "prepared plan: a pattern's predicate does not occur in the store (the answer is empty)");
        if (ok && empty) {
            auto r = std::make_unique<kb_rel>();
            r->slots = all_slots;
            for (size_t c = 0; c < all_slots.size(); c++) { Col col; KB_TRY(alloc_col(ctx, 0, &col)); r->cols.push_back(col); }
            *out = std::move(r);
            return KB_OK;
        }
        u32 kmn = 0xFFFFFFFFu, kmx = 0;
        u32 cshift = 0;
        if (ok) {
            bool subj = true;
            for (u32 k = 0; k < K; k++) {
                const bool y = key_pos(k) == 2;
                kmn = std::min(kmn, y ? sl[k]->ymin : sl[k]->xmin);
                kmx = std::max(kmx, y ? sl[k]->ymax : sl[k]->xmax);
                if (y) subj = false;
            }
            for (auto& sg : ctx->segs) if (sg.n && (!sg.has_stats || sg.stats_world != ctx->shard_world || !sg.sharded_ok)) subj = false;
            const u32 w = ctx->shard_world;
            if (subj && w > 1 && (w & (w - 1)) == 0) while ((1u << cshift) < w) cshift++;
            kmn = compact_key(kmn, cshift);
            kmx = compact_key(kmx, cshift);
        }
        // a pattern whose key column has a persistent table in the index needs no build at all: it is looked up like spo[s][P]
        auto persistent = [&](u32 k) -> bool {
            if (!ok) return false;
            const bool y = key_pos(k) == 2;
            return y ? (bool)sl[k]->ytab : ((bool)sl[k]->xtab && sl[k]->tab_cshift == cshift);
        };
        // probe side: prefer a pattern WITHOUT a persistent table (it would need a build), then the smallest slice (fewest lookups);
        // when builds remain, the classic choice: the largest unfiltered slice probes, filtered sides build
        int probe_k = -1;
        bool all_persistent = ok;
        for (u32 k = 0; k < K && ok; k++) if (!persistent(k)) all_persistent = false;
        if (ok && all_persistent) {
            // no builds at all: the probe side should be the most selective stream — a pattern that carries its own FILTER (evaluated
            // on the probe row before any lookup), else the smallest slice
            probe_k = -1;
            for (u32 k = 0; k < K; k++) if (!pushdown[k].ops.empty() && pushdown[k].ops.size() <= 8 && (probe_k < 0 || sl[k]->n < sl[probe_k]->n)) probe_k = (int)k;
            if (probe_k < 0) { probe_k = 0; for (u32 k = 1; k < K; k++) if (sl[k]->n < sl[probe_k]->n) probe_k = (int)k; }
        } else {
            for (u32 k = 0; k < K && ok; k++) if (pushdown[k].ops.empty() && (probe_k < 0 || sl[k]->n > sl[probe_k]->n)) probe_k = (int)k;
            if (ok && probe_k < 0) { probe_k = 0; for (u32 k = 1; k < K; k++) if (sl[k]->n > sl[probe_k]->n) probe_k = (int)k; }
        }
        const u64 rng = (ok && kmx >= kmn) ? (u64)kmx - kmn + 1 : 0;
        for (u32 k = 0; k < K && ok; k++) {
            if ((int)k == probe_k || persistent(k)) continue;
            if (rng == 0 || rng > std::max<u64>(8 * sl[k]->n, 1ull << 16) || rng > (1ull << 31)) ok = false;
            if (ctx->multi_valued.count({pats[k].p.value, key_pos(k)})) ok = false;
        }
        if (ok && all_persistent && ctx->fast_index_kernel && sl[probe_k]->chunks.size() <= (size_t)PROBEI_MAXSEG) {
            // every build side is a persistent table: the whole join is one launch of probe_index_kernel
            const PredSlice& PS = *sl[probe_k];
            const u32 T = K - 1;
            ProbeIParams P{};
            std::vector<u32> out_slots = pv[probe_k];
            FilterProg postp = post;
            auto to_post = [&](u32 k) {
                if (pushdown[k].ops.empty()) return;
                const bool had = !postp.ops.empty();
                postp.ops.insert(postp.ops.end(), pushdown[k].ops.begin(), pushdown[k].ops.end());
                if (had) { kb_filter_op a{}; a.op = KB_F_AND; postp.ops.push_back(a); }
            };
            u32 t = 0;
            for (u32 k = 0; k < K; k++) {
                if ((int)k == probe_k) continue;
                const bool y = key_pos(k) == 2;
                DirectTab& D = P.tab[t++];
                D.mode = 0; D.n_pay = 0; D.pay[0] = D.pay[1] = nullptr;
                D.tab = static_cast<const u32*>(y ? sl[k]->ytab->p : sl[k]->xtab->p);
                D.kmin = y ? sl[k]->ytab_min : sl[k]->xtab_min;
                D.range = y ? sl[k]->ytab_range : sl[k]->xtab_range;
                D.cshift = y ? 0u : sl[k]->tab_cshift;
                to_post(k);
                out_slots.push_back(pv[k][y ? 0 : 1]);
            }
            const FilterProg& pf = pushdown[probe_k];
            if (!pf.ops.empty()) {
                if (pf.ops.size() > 8) to_post((u32)probe_k);
                else {
                    std::map<u32, u32> remap;  // slots of the probe pattern's own filter -> pair halves (x = subject, y = object)
                    for (size_t i = 0; i < pv[probe_k].size(); i++) remap[pv[probe_k][i]] = psrc[probe_k][i] == 0 ? 0u : 1u;
                    std::vector<FilterOp> fo;
                    if (!append_prog(&fo, pf, remap)) return fail(ctx, KB_E_INVALID, "filter uses an unbound variable");
                    const bool typed = PS.typed(ctx->num_version) && fo.size() == 1 && fo[0].op == KB_F_CMP_NUM && fo[0].slot == 1u;
                    if (typed) {
                        P.pre_mode = 1; P.pre_cmp = fo[0].cmp; P.pre_val = fo[0].value;
                    } else {
                        P.pre_mode = 2; P.n_pre = (u32)fo.size();
                        for (size_t i = 0; i < fo.size(); i++) P.pre_ops[i] = fo[i];
                    }
                }
            }
            const u32 n_out = 2 + T;
            if (!postp.ops.empty()) {
                std::map<u32, u32> remap;
                for (u32 c = 0; c < n_out; c++) remap[out_slots[c]] = c;
                std::vector<FilterOp> fops;
                if (!append_prog(&fops, postp, remap)) return fail(ctx, KB_E_INVALID, "filter uses an unbound variable");
                if (fops.size() > KB_MAX_FILTER_OPS) return fail(ctx, KB_E_LIMIT, "filter too long");
                P.n_ops = (u32)fops.size();
                for (size_t i = 0; i < fops.size(); i++) P.ops[i] = fops[i];
            }
            // table mode: the probe pattern's own subject table is the probe stream (keys ascending whatever the store / id order)
            const bool tab_mode = ctx->probe_table_mode && !ctx->ordered && persistent((u32)probe_k) && key_pos((u32)probe_k) == 0 &&
                                  (P.pre_mode != 1u || (PS.xnum && PS.xnum_version == ctx->num_version));
            P.n_seg = 0;
            P.n_tiles = 0;
            if (tab_mode) {
                // walk the slots that can hold a key (smallest to largest live subject), not the table's growth headroom nor the
                // slots an evicted past has left empty; the first slot stays 128-byte aligned for the bulk loads
                const u32 lo = ((compact_key(PS.xmin, PS.tab_cshift) - PS.xtab_min) / 32u) * 32u;
                const u32 hi = std::min<u32>(PS.xtab_range, compact_key(PS.xmax, PS.tab_cshift) - PS.xtab_min + 1u);
                P.ptab = static_cast<const u32*>(PS.xtab->p) + lo;
                P.pnum = P.pre_mode == 1u ? static_cast<const double*>(PS.xnum->p) + lo : nullptr;
                P.ptab_min = PS.xtab_min + lo;
                P.ptab_range = hi > lo ? hi - lo : 0u;
                P.ptab_cshift = PS.tab_cshift;
                P.shard_rank = ctx->shard_rank;
                P.n_tiles = (u32)(((u64)P.ptab_range + PROBEF_TILE - 1) / PROBEF_TILE);
            }
            for (auto& ch : PS.chunks) {  // one chunk per store segment: an RSP window of slides is walked in ONE launch
                if (tab_mode) break;
                if (ch.n == 0) continue;
                ProbeISeg& g = P.seg[P.n_seg++];
                g.pairs = reinterpret_cast<const uint2*>(ch.pairs.ptr);
                g.ynum = P.pre_mode == 1u ? static_cast<const double*>(ch.ynum->p) : nullptr;
                g.n = (u32)ch.n;
                g.tile0 = P.n_tiles;
                P.n_tiles += (u32)((ch.n + PROBEF_TILE - 1) / PROBEF_TILE);
            }
            P.key_is_y = key_pos((u32)probe_k) == 2 ? 1u : 0u;
            P.n = (u32)PS.n;
            P.T = T;
            P.cap = (u32)PS.n;
            P.nt = numtab(ctx);
            P.cb = ctx->fast_cb;
            P.host_total = ctx->d_fast;
            if (plan_only) {  // kb_star_join_prepare: keep the resolved launch, run nothing
                plan_only->P = P;
                plan_only->out_slots = out_slots;
                plan_only->all_slots = all_slots;
                plan_only->n_out = n_out;
                plan_only->probe_rows = PS.n;
                if (agg) {
                    int gsel = -1, asel = -1;
                    for (u32 c = 0; c < n_out; c++) {
                        if (out_slots[c] == agg->group_slot) gsel = (int)c;
                        if (agg->has_agg && out_slots[c] == agg->agg_slot) asel = (int)c;
                    }
                    const bool needs_value = agg->has_agg && agg->kind != KB_AGG_COUNT;
                    if (gsel < 0 || (needs_value && asel < 0))
                        return fail(ctx, KB_E_UNSUPPORTED, "prepared GROUP BY: the group or aggregate variable is not bound by the join");
                    plan_only->agg = true;
                    plan_only->has_agg = agg->has_agg;
                    plan_only->agg_kind = agg->has_agg ? agg->kind : (u32)KB_AGG_COUNT;
                    plan_only->agg_slot = agg->agg_slot;
                    plan_only->group_slot = agg->group_slot;
                    plan_only->P.gsel = (u32)gsel;
                    plan_only->P.asel = asel >= 0 ? (u32)asel : 0u;
                    plan_only->P.akind = plan_only->agg_kind;
                }
                plan_only->ok = true;
                return KB_OK;
            }
            if (agg) {
                // GROUP BY folded into the probe kernel: no joined row is written. One try with a 4096-slot table; more groups than
                // that (or a group / aggregate variable the join does not bind) leave agg->applied false: the caller groups separately
                int gsel = -1, asel = -1;
                for (u32 c = 0; c < n_out; c++) {
                    if (out_slots[c] == agg->group_slot) gsel = (int)c;
                    if (agg->has_agg && out_slots[c] == agg->agg_slot) asel = (int)c;
                }
                const bool needs_value = agg->has_agg && agg->kind != KB_AGG_COUNT;
                if (gsel >= 0 && (!needs_value || asel >= 0)) {
                    GroupParams G{};
                    G.n_gcols = 1;
                    G.n_aggs = agg->has_agg ? 1u : 0u;
                    G.akind[0] = agg->has_agg ? agg->kind : (u32)KB_AGG_COUNT;
                    G.nt = numtab(ctx);
                    GroupTable tab;
                    KB_TRY(group_table_create(ctx, 1u << 12, &G, &tab));
                    const u32 goff = ctrl_alloc(ctx, 4);
                    KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + goff, 0, 4 * sizeof(u32), ctx->st));
                    G.overflow = ctx->ctrl + goff;
                    P.gsel = (u32)gsel;
                    P.asel = asel >= 0 ? (u32)asel : 0u;
                    P.akind = G.akind[0];
                    P.ordered = 0;
                    P.epoch = ctx->epoch++;
                    timer_begin(ctx, F_PROBE);
                    launch_probe_index(P, &G, ctx->n_sms, ctx->st);
                    timer_end(ctx);
                    KB_CUDA(ctx, cudaGetLastError());
                    ctx->stats.rows_probed += PS.n;
                    ctx->stats.index_joins++;
                    KB_TRY(ctrl_read(ctx));
                    if (ctx->h_ctrl[goff] == 0) {
                        agg->n_rows = *reinterpret_cast<volatile u32*>(ctx->h_fast);
                        agg->groups = std::make_unique<kb_groups>();
                        agg->groups->keys.resize(1);
                        agg->groups->vals.resize(agg->has_agg ? 1 : 0);
                        kb_agg a1{agg->kind, agg->agg_slot};
                        KB_TRY(group_table_collect(ctx, tab, 1, &a1, agg->has_agg ? 1u : 0u, agg->groups.get()));
                        agg->applied = true;
                        return KB_OK;
                    }
                }
                // fall through: join without the fused GROUP BY (the control block was left clean by the kernel)
            }
            auto res = std::make_unique<kb_rel>();
            res->slots = out_slots;
            {   // one allocation holds all output columns
                const size_t stride = round256((size_t)PS.n * sizeof(u32)) + 256;
                Buf b;
                KB_TRY(alloc_buf(ctx, stride * n_out, &b));
                for (u32 c = 0; c < n_out; c++) {
                    Col col;
                    col.buf = b;
                    col.ptr = reinterpret_cast<u32*>(static_cast<char*>(b->p) + stride * c);
                    res->cols.push_back(col);
                    P.out[c] = col.ptr;
                }
            }
            P.ordered = ctx->ordered;
            if (P.ordered) {
                KB_TRY(ensure_tile_state(ctx, P.n_tiles));
                P.tile_state = static_cast<u64*>(ctx->tile_state->p);
                P.block_state = static_cast<u64*>(ctx->block_state->p);
            }
            P.epoch = ctx->epoch++;
            timer_begin(ctx, F_PROBE);
            launch_probe_index(P, nullptr, ctx->n_sms, ctx->st);
            timer_end(ctx);
            KB_CUDA(ctx, cudaGetLastError());
            ctx->stats.rows_probed += PS.n;
            ctx->stats.index_joins++;
            KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
            timers_flush(ctx);
            ctx->stats.d2h_bytes += sizeof(u32);
            res->n = *reinterpret_cast<volatile u32*>(ctx->h_fast);
            *out = select_cols(*res, all_slots);
            return KB_OK;
        }
        for (u32 k = 0; k < K && ok; k++) if (!sl[k]->single()) ok = false;  // the build-from-slice path reads one contiguous slice per pattern
        if (plan_only) return fail(ctx, KB_E_UNSUPPORTED, "prepared plans take the one-kernel index path only: every pattern (?s P ?o) over a predicate whose key column has a persistent table in the store index");
        if (ok) {
            const u32 range = (u32)rng;
            const u32 off = ctrl_alloc(ctx, 16 + 2 * MAXT);
            KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off, 0, (16 + 2 * MAXT) * sizeof(u32), ctx->st));
            std::vector<Buf> tabs;
            DirectTab dt[MAXT];
            std::vector<u32> tab_k;
            std::vector<u32> out_slots = pv[probe_k];
            std::vector<OutCol> ocs{OutCol{OUT_PROBE, 0, 0}, OutCol{OUT_PROBE, 1, 0}};
            const u32 T = K - 1;
            FilterProg postp = post;  // filters of patterns that are not built (probe side, persistent tables) run on the joined row
            auto to_post = [&](u32 k) {
                if (pushdown[k].ops.empty()) return;
                const bool had = !postp.ops.empty();
                postp.ops.insert(postp.ops.end(), pushdown[k].ops.begin(), pushdown[k].ops.end());
                if (had) { kb_filter_op a{}; a.op = KB_F_AND; postp.ops.push_back(a); }
            };
            const bool pre_filter = !pushdown[probe_k].ops.empty() && pushdown[probe_k].ops.size() <= 8;
            if (!pre_filter) to_post((u32)probe_k);
            u32 n_build = 0;
            for (u32 k = 0; k < K; k++) if ((int)k != probe_k && !persistent(k)) n_build++;
            if (n_build) timer_begin(ctx, F_BUILD, (int)n_build);
            for (u32 k = 0; k < K; k++) if ((int)k != probe_k && !persistent(k)) { Buf b; KB_TRY(alloc_buf(ctx, (size_t)range * sizeof(u32), &b)); tabs.push_back(b); }
            if (n_build > 1) {
                KB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->st));
                KB_CUDA(ctx, cudaStreamWaitEvent(ctx->st2, ctx->ev_fork, 0));
            }
            u32 t = 0, bi = 0;
            for (u32 k = 0; k < K; k++) {
                if ((int)k == probe_k) continue;
                DirectTab& D = dt[t];
                D.mode = 0; D.n_pay = 0; D.pay[0] = D.pay[1] = nullptr;
                if (persistent(k)) {
                    const bool y = key_pos(k) == 2;
                    D.tab = static_cast<const u32*>(y ? sl[k]->ytab->p : sl[k]->xtab->p);
                    D.kmin = y ? sl[k]->ytab_min : sl[k]->xtab_min;
                    D.range = y ? sl[k]->ytab_range : sl[k]->xtab_range;
                    D.cshift = y ? 0u : sl[k]->tab_cshift;
                    to_post(k);
                } else {
                    cudaStream_t bs = (n_build > 1 && (bi & 1)) ? ctx->st2 : ctx->st;
                    u32* tab = static_cast<u32*>(tabs[bi]->p);
                    KB_CUDA(ctx, cudaMemsetAsync(tab, 0xFF, (size_t)range * sizeof(u32), bs));
                    BuildPairsParams B{};
                    B.kv = reinterpret_cast<const uint2*>(sl[k]->single()->pairs.ptr);
                    B.n = (u32)sl[k]->n;
                    B.key_is_y = key_pos(k) == 2 ? 1u : 0u;
                    B.pred = pats[k].p.value;
                    B.table = tab; B.kmin = kmn; B.range = range; B.cshift = cshift;
                    B.dup_flag = ctx->ctrl + off + 16 + t;
                    B.count = ctx->ctrl + off + 16 + MAXT + t;
                    B.trusted = ((key_pos(k) == 2 ? sl[k]->y_unique : sl[k]->x_unique) || ctx->single_valued.count({pats[k].p.value, key_pos(k)})) ? 1u : 0u;
                    B.nt = numtab(ctx);
                    if (!pushdown[k].ops.empty()) {
                        std::map<u32, u32> remap;
                        for (size_t i = 0; i < pv[k].size(); i++) remap[pv[k][i]] = psrc[k][i];
                        std::vector<FilterOp> fo;
                        if (!append_prog(&fo, pushdown[k], remap)) return fail(ctx, KB_E_INVALID, "pushed-down filter uses a variable the pattern does not bind");
                        B.n_ops = (u32)fo.size();
                        for (size_t i = 0; i < fo.size(); i++) B.ops[i] = fo[i];
                    }
                    launch_build_direct_pairs_filtered(B, ctx->n_sms, bs);
                    D.tab = tab; D.kmin = kmn; D.range = range; D.cshift = cshift;
                    ctx->stats.rows_built += sl[k]->n;
                    bi++;
                }
                out_slots.push_back(pv[k][key_pos(k) == 0 ? 1 : 0]);
                ocs.push_back(OutCol{OUT_TABVAL, t, 0});
                tab_k.push_back(k);
                t++;
            }
            if (n_build > 1) {
                KB_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->st2));
                KB_CUDA(ctx, cudaStreamWaitEvent(ctx->st, ctx->ev_join, 0));
            }
            if (n_build) timer_end(ctx);
            const PredSlice& PS = *sl[probe_k];
            auto res = std::make_unique<kb_rel>();
            res->slots = out_slots;
            const u32 n_out = (u32)out_slots.size();
            std::vector<FilterOp> fops;
            if (!postp.ops.empty()) {
                std::map<u32, u32> remap;
                for (u32 c = 0; c < n_out; c++) remap[out_slots[c]] = c;
                if (!append_prog(&fops, postp, remap)) return fail(ctx, KB_E_INVALID, "filter uses an unbound variable");
                if (fops.size() > KB_MAX_FILTER_OPS) return fail(ctx, KB_E_LIMIT, "filter too long");
            }
            ProbeFParams P{};
            if (pre_filter) {  // slots of the probe pattern's own filter -> pair halves (x = subject, y = object)
                std::map<u32, u32> remap;
                for (size_t i = 0; i < pv[probe_k].size(); i++) remap[pv[probe_k][i]] = psrc[probe_k][i] == 0 ? 0u : 1u;
                std::vector<FilterOp> fo;
                if (!append_prog(&fo, pushdown[probe_k], remap)) return fail(ctx, KB_E_INVALID, "filter uses an unbound variable");
                P.n_pre = (u32)fo.size();
                for (size_t i = 0; i < fo.size(); i++) P.pre_ops[i] = fo[i];
                if (sl[probe_k]->typed(ctx->num_version)) P.pre_num = static_cast<const double*>(sl[probe_k]->single()->ynum->p);
            }
            P.pairs = reinterpret_cast<const uint2*>(PS.single()->pairs.ptr);
            P.key_is_y = key_pos((u32)probe_k) == 2 ? 1u : 0u;
            P.n = (u32)PS.n;
            P.n_tiles = (u32)((PS.n + PROBEF_TILE - 1) / PROBEF_TILE);
            P.T = T;
            for (u32 q = 0; q < T; q++) P.tab[q] = dt[q];
            P.n_out = n_out;
            for (u32 c = 0; c < n_out; c++) {
                Col col;
                KB_TRY(alloc_col(ctx, PS.n, &col));
                res->cols.push_back(col);
                P.oc[c] = ocs[c];
                P.out[c] = col.ptr;
            }
            P.cap = (u32)PS.n;
            P.n_ops = (u32)fops.size();
            for (size_t i = 0; i < fops.size(); i++) P.ops[i] = fops[i];
            P.nt = numtab(ctx);
            KB_TRY(ensure_tile_state(ctx, P.n_tiles));
            P.tile_state = static_cast<u64*>(ctx->tile_state->p);
            P.block_state = static_cast<u64*>(ctx->block_state->p);
            P.ordered = ctx->ordered;
            P.ticket = ctx->ctrl + off;
            P.total = ctx->ctrl + off + 1;
            P.zero_word = ctx->ctrl + off + 2;
            P.abort_flag = nullptr;
            P.epoch = ctx->epoch++;
            timer_begin(ctx, F_PROBE);
            launch_probe_fast(P, ctx->n_sms, ctx->st);
            timer_end(ctx);
            KB_CUDA(ctx, cudaGetLastError());
            ctx->stats.rows_probed += PS.n;
            ctx->stats.index_joins++;
            KB_TRY(ctrl_read(ctx));
            bool dup = false;
            for (u32 q = 0; q < T; q++) if (ctx->h_ctrl[off + 16 + q]) {
                dup = true;
                ctx->multi_valued.insert({pats[tab_k[q]].p.value, key_pos(tab_k[q])});
            }
            if (dup) return star_join_impl2(ctx, join_slot, pats, K, filter, n_ops, true, out, agg);  // multi-valued is now cached: takes the chained route
            for (u32 q = 0; q < T; q++)
                if (!persistent(tab_k[q]) && pushdown[tab_k[q]].ops.empty()) ctx->single_valued.insert({pats[tab_k[q]].p.value, key_pos(tab_k[q])});
            res->n = ctx->h_ctrl[off + 1];
            *out = select_cols(*res, all_slots);
            return KB_OK;
        }
    }

    if (plan_only) return fail(ctx, KB_E_UNSUPPORTED, "prepared plans need a valid store index (kb_store_build_index) and (?s P ?o) patterns");
    // ---- scan + build FUSED: the build-side patterns insert straight into their direct tables while the store is scanned; only the
    // probe-side pattern is materialised. Needs the key range before the scan (load-time statistics) and a probe side chosen without
    // knowing the counts: the last pattern that carries no pushed-down filter (a filtered side is the smaller build side).
    if (allow_fused_scan && all_pairs && K - 1 <= (u32)MAXT && ctx->upload_stats_off < 0 && ctx->n_triples > 0) {
        bool ok = true;
        u32 kmn = 0xFFFFFFFFu, kmx = 0;
        for (auto& sg : ctx->segs) {
            if (sg.n == 0) continue;
            if (!sg.has_stats) KB_TRY(segment_stats(ctx, &sg));
            for (u32 k = 0; k < K; k++) { const u32 c = key_pos(k); kmn = std::min(kmn, sg.cmin[c]); kmx = std::max(kmx, sg.cmax[c]); }
        }
        // a subject-sharded store (kb_set_sharding) owns every world-th block of ids: compact the key domain so the tables stay dense
        u32 cshift = 0;
        {
            bool subj = true;
            for (u32 k = 0; k < K; k++) if (key_pos(k) != 0) subj = false;
            for (auto& sg : ctx->segs) if (sg.n && !sg.sharded_ok) subj = false;
            const u32 w = ctx->shard_world;
            if (subj && w > 1 && (w & (w - 1)) == 0) while ((1u << cshift) < w) cshift++;
        }
        if (kmx >= kmn) { kmn = compact_key(kmn, cshift); kmx = compact_key(kmx, cshift); }
        const u64 rng = kmx >= kmn ? (u64)kmx - kmn + 1 : 0;
        if (rng == 0 || rng > std::max<u64>(4 * ctx->n_triples, 1ull << 16) || rng > (1ull << 31)) ok = false;
        for (u32 a = 0; a < K && ok; a++)
            for (u32 b = a + 1; b < K && ok; b++)
                for (u32 s2 : pv[a]) if (s2 != join_slot && std::find(pv[b].begin(), pv[b].end(), s2) != pv[b].end()) ok = false;
        int probe_k = -1;
        for (u32 k = 0; k < K; k++) if (pushdown[k].ops.empty()) probe_k = (int)k;
        if (probe_k < 0) probe_k = (int)K - 1;
        for (u32 k = 0; k < K && ok; k++)
            if ((int)k != probe_k && !pats[k].p.is_var && ctx->multi_valued.count({pats[k].p.value, key_pos(k)})) ok = false;
        if (ok) {
            const u32 range = (u32)rng;
            const u32 off = ctrl_alloc(ctx, 8 + MAXT);
            KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off, 0, (8 + MAXT) * sizeof(u32), ctx->st));
            std::vector<ScanTable> st(K);
            std::vector<Buf> tabs;
            std::vector<u32> tab_of(K, 0);
            DirectTab dt[MAXT];
            std::vector<u32> out_slots = pv[probe_k];
            std::vector<OutCol> ocs{OutCol{OUT_PROBE, 0, 0}, OutCol{OUT_PROBE, 1, 0}};
            timer_begin(ctx, F_BUILD, 0);
            u32 t = 0;
            for (u32 k = 0; k < K; k++) {
                if ((int)k == probe_k) continue;
                Buf b;
                KB_TRY(alloc_buf(ctx, (size_t)range * sizeof(u32), &b));
                tabs.push_back(b);
                st[k].tab = static_cast<u32*>(b->p);
                st[k].clear = true;  // cleared by the scan (scan_impl)
                st[k].kmin = kmn; st[k].range = range; st[k].cshift = cshift;
                st[k].key_is_o = key_pos(k) == 2 ? 1u : 0u;
                st[k].trusted = (!pats[k].p.is_var && ctx->single_valued.count({pats[k].p.value, key_pos(k)})) ? 1u : 0u;
                st[k].dup_flag = ctx->ctrl + off + 8 + t;
                DirectTab& D = dt[t];
                D.tab = st[k].tab; D.kmin = kmn; D.range = range; D.cshift = cshift; D.mode = 0; D.n_pay = 0; D.pay[0] = D.pay[1] = nullptr;
                out_slots.push_back(pv[k][key_pos(k) == 0 ? 1 : 0]);
                ocs.push_back(OutCol{OUT_TABVAL, t, 0});
                tab_of[k] = t;
                t++;
            }
            timer_end(ctx);
            const u32 T = t;
            std::vector<std::unique_ptr<kb_rel>> rels;
            KB_TRY(scan_impl(ctx, pats, K, pushdown, false, true, &rels, &st));
            ctx->stats.fused_scan_builds++;
            // duplicate keys: the scan stores without reading back, so a multi-valued key shows up as fewer occupied slots than
            // inserted rows. Tables of (predicate, position) pairs already verified for this store version are not re-counted.
            bool counted = false;
            const u32 offc = ctrl_alloc(ctx, MAXT);
            KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + offc, 0, MAXT * sizeof(u32), ctx->st));
            for (u32 k = 0; k < K; k++) {
                if ((int)k == probe_k || st[k].trusted) continue;
                if (!counted) timer_begin(ctx, F_BUILD, 0);
                counted = true;
                ctx->stats.kernel_launches++;
                launch_count_nonempty(st[k].tab, range, ctx->ctrl + offc + tab_of[k], ctx->n_sms, ctx->st);
            }
            if (counted) timer_end(ctx);  // the counts are read together with the probe's result: no extra synchronisation
            auto check_dups = [&]() -> bool {
                bool dup = false;
                for (u32 k = 0; k < K; k++) {
                    if ((int)k == probe_k) continue;
                    const bool fewer = !st[k].trusted && ctx->h_ctrl[offc + tab_of[k]] != rels[k]->n;
                    if (ctx->h_ctrl[off + 8 + tab_of[k]] || fewer) {
                        dup = true;
                        if (!pats[k].p.is_var) ctx->multi_valued.insert({pats[k].p.value, key_pos(k)});
                    }
                }
                return dup;
            };
            for (u32 k = 0; k < K; k++) if ((int)k != probe_k) ctx->stats.rows_built += rels[k]->n;
            auto empty_result2 = [&]() -> kb_status {
                auto r = std::make_unique<kb_rel>();
                r->slots = all_slots;
                for (size_t c = 0; c < all_slots.size(); c++) { Col col; KB_TRY(alloc_col(ctx, 0, &col)); r->cols.push_back(col); }
                *out = std::move(r);
                return KB_OK;
            };
            for (u32 k = 0; k < K; k++) if (rels[k]->n == 0) return empty_result2();
            for (u32 k = 0; k < K; k++) if ((int)k != probe_k && ctx->h_ctrl[off + 8 + tab_of[k]]) {  // key outside the table range (cannot happen with store statistics)
                KB_TRY(ctrl_read(ctx));
                check_dups();
                return star_join_impl2(ctx, join_slot, pats, K, filter, n_ops, false, out, agg);
            }
            const kb_rel& PR = *rels[probe_k];
            auto res = std::make_unique<kb_rel>();
            res->slots = out_slots;
            const u32 n_out = (u32)out_slots.size();
            std::vector<FilterOp> fops;
            if (!post.ops.empty()) {
                std::map<u32, u32> remap;
                for (u32 c = 0; c < n_out; c++) remap[out_slots[c]] = c;
                if (!append_prog(&fops, post, remap)) return fail(ctx, KB_E_INVALID, "filter uses an unbound variable");
            }
            ProbeFParams P{};
            P.pairs = reinterpret_cast<const uint2*>(PR.cols[0].ptr);
            P.key_is_y = key_pos((u32)probe_k) == 2 ? 1u : 0u;
            P.n = (u32)PR.n;
            P.n_tiles = (u32)((PR.n + PROBEF_TILE - 1) / PROBEF_TILE);
            P.T = T;
            for (u32 q = 0; q < T; q++) P.tab[q] = dt[q];
            P.n_out = n_out;
            for (u32 c = 0; c < n_out; c++) {
                Col col;
                KB_TRY(alloc_col(ctx, PR.n, &col));
                res->cols.push_back(col);
                P.oc[c] = ocs[c];
                P.out[c] = col.ptr;
            }
            P.cap = (u32)PR.n;
            P.n_ops = (u32)fops.size();
            for (size_t i = 0; i < fops.size(); i++) P.ops[i] = fops[i];
            P.nt = numtab(ctx);
            KB_TRY(ensure_tile_state(ctx, P.n_tiles));
            P.tile_state = static_cast<u64*>(ctx->tile_state->p);
            P.block_state = static_cast<u64*>(ctx->block_state->p);
            P.ordered = ctx->ordered;
            const u32 off2 = ctrl_alloc(ctx, 8);
            KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off2, 0, 8 * sizeof(u32), ctx->st));
            P.ticket = ctx->ctrl + off2;
            P.total = ctx->ctrl + off2 + 1;
            P.zero_word = ctx->ctrl + off2 + 2;
            P.abort_flag = nullptr;
            P.epoch = ctx->epoch++;
            timer_begin(ctx, F_PROBE);
            launch_probe_fast(P, ctx->n_sms, ctx->st);
            timer_end(ctx);
            KB_CUDA(ctx, cudaGetLastError());
            ctx->stats.rows_probed += PR.n;
            KB_TRY(ctrl_read(ctx));
            if (check_dups())  // a build side is multi-valued: redo without the fusion (the chained operator needs the build rows)
                return star_join_impl2(ctx, join_slot, pats, K, filter, n_ops, false, out, agg);
            for (u32 k = 0; k < K; k++)
                if ((int)k != probe_k && !pats[k].p.is_var && pushdown[k].ops.empty()) ctx->single_valued.insert({pats[k].p.value, key_pos(k)});
            res->n = ctx->h_ctrl[off2 + 1];
            *out = select_cols(*res, all_slots);
            return KB_OK;
        }
    }

    std::vector<std::unique_ptr<kb_rel>> rels;
    KB_TRY(scan_impl(ctx, pats, K, pushdown, false, all_pairs, &rels));
    if (ctx->upload_stats_off >= 0) {  // one-shot host call: the copy stream computed the column ranges chunk by chunk
        const u32 o = (u32)ctx->upload_stats_off;
        for (auto& sg : ctx->segs) {
            for (int c = 0; c < 3; c++) { sg.cmin[c] = ctx->h_ctrl[o + c]; sg.cmax[c] = ctx->h_ctrl[o + 4 + c]; }
            sg.has_stats = true;
        }
        ctx->upload_stats_off = -1;
    }
    u32 kmin = 0xFFFFFFFFu, kmax = 0;
    for (auto& sg : ctx->segs) {
        if (sg.n == 0) continue;
        if (!sg.has_stats) KB_TRY(segment_stats(ctx, &sg));
        for (u32 k = 0; k < K; k++) { const u32 c = key_pos(k); kmin = std::min(kmin, sg.cmin[c]); kmax = std::max(kmax, sg.cmax[c]); }
    }

    auto empty_result = [&]() -> kb_status {
        auto r = std::make_unique<kb_rel>();
        r->slots = all_slots;
        for (size_t c = 0; c < all_slots.size(); c++) { Col col; KB_TRY(alloc_col(ctx, 0, &col)); r->cols.push_back(col); }
        *out = std::move(r);
        return KB_OK;
    };
    for (u32 k = 0; k < K; k++) if (rels[k]->n == 0) return empty_result();

    if (K == 1) {
        std::unique_ptr<kb_rel> r = std::move(rels[0]);
        KB_TRY(unpair_rel(ctx, &r));
        if (!post.ops.empty()) { std::unique_ptr<kb_rel> f; KB_TRY(filter_impl(ctx, *r, post, &f)); r = std::move(f); }
        *out = std::move(r);
        return KB_OK;
    }
    // probe side = largest relation (ties: first)
    u32 probe = 0;
    for (u32 k = 1; k < K; k++) if (rels[k]->n > rels[probe]->n) probe = k;
    std::vector<u32> builds;
    for (u32 k = 0; k < K; k++) if (k != probe) builds.push_back(k);

    // the fused path needs: patterns share only the join variable (quirk Q3: the reference never checks the others — we must,
    // so such shapes take the natural-join chain), a dense key range, and single-valued keys
    bool fused_ok = kmax >= kmin;
    for (u32 a = 0; a < K && fused_ok; a++)
        for (u32 b = a + 1; b < K && fused_ok; b++)
            for (u32 s2 : pv[a]) if (s2 != join_slot && std::find(pv[b].begin(), pv[b].end(), s2) != pv[b].end()) fused_ok = false;
    const u64 range64 = fused_ok ? (u64)kmax - kmin + 1 : 0;
    for (u32 k : builds) {
        if (range64 > std::max<u64>(8 * rels[k]->n, 1ull << 16) || range64 > (1ull << 31)) fused_ok = false;
        if (!pats[k].p.is_var && ctx->multi_valued.count({pats[k].p.value, key_pos(k)})) fused_ok = false;
        if (pv[k].size() > 3) fused_ok = false;
    }
    const u32 range = (u32)range64;

    auto to_columnar = [&]() -> kb_status {
        for (u32 k = 0; k < K; k++) KB_TRY(unpair_rel(ctx, &rels[k]));
        return KB_OK;
    };

    if (fused_ok) {
        std::unique_ptr<kb_rel> cur;  // result of the batches done so far (columnar); null = still the scan output of `probe`
        bool dup_seen = false;
        size_t done = 0;
        while (done < builds.size() && !dup_seen) {
            const size_t nb = std::min<size_t>(MAXT, builds.size() - done);
            const bool last = done + nb == builds.size();
            const bool fast = all_pairs && !cur;  // first batch over pair relations -> register-staged fast probe
            const kb_rel& PR = cur ? *cur : *rels[probe];
            const u32 off = ctrl_alloc(ctx, 8 + MAXT);
            KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off, 0, (8 + MAXT) * sizeof(u32), ctx->st));
            std::vector<Buf> tables(nb);
            DirectTab dt[MAXT];
            std::vector<u32> out_slots = PR.slots;
            std::vector<OutCol> ocs;
            for (u32 c = 0; c < (u32)PR.slots.size(); c++) ocs.push_back(OutCol{OUT_PROBE, c, 0});
            timer_begin(ctx, F_BUILD, (int)nb);
            for (size_t t = 0; t < nb; t++) KB_TRY(alloc_buf(ctx, (size_t)range * sizeof(u32), &tables[t]));  // stream-ordered on st
            if (nb > 1) {  // independent tables: odd ones are memset + built on the second stream, concurrently with the even ones
                KB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->st));
                KB_CUDA(ctx, cudaStreamWaitEvent(ctx->st2, ctx->ev_fork, 0));
            }
            for (size_t t = 0; t < nb; t++) {
                const u32 k = builds[done + t];
                const kb_rel& B = *rels[k];
                cudaStream_t bs = (nb > 1 && (t & 1)) ? ctx->st2 : ctx->st;
                u32* tab = static_cast<u32*>(tables[t]->p);
                KB_CUDA(ctx, cudaMemsetAsync(tab, 0xFF, (size_t)range * sizeof(u32), bs));
                DirectTab& D = dt[t];
                D.tab = tab; D.kmin = kmin; D.range = range; D.cshift = 0; D.n_pay = 0; D.pay[0] = D.pay[1] = nullptr;
                const int kc = B.col_of(join_slot);
                if (B.pair) {
                    D.mode = 0;
                    out_slots.push_back(B.slots[kc == 0 ? 1 : 0]);
                    ocs.push_back(OutCol{OUT_TABVAL, (u32)t, 0});
                    const u32 trusted = (!pats[k].p.is_var && ctx->single_valued.count({pats[k].p.value, key_pos(k)})) ? 1u : 0u;
                    launch_build_direct_pairs(reinterpret_cast<const uint2*>(B.cols[0].ptr), kc == 1 ? 1u : 0u, (u32)B.n, tab, kmin, range, 0u,
                                              ctx->ctrl + off + 8 + (u32)t, trusted, ctx->n_sms, bs);
                } else {
                    const u32* vals = nullptr;
                    if (B.cols.size() == 1) { D.mode = 2; }
                    else if (B.cols.size() == 2) {
                        D.mode = 0;
                        const int vc = kc == 0 ? 1 : 0;
                        vals = B.cols[vc].ptr;
                        out_slots.push_back(B.slots[vc]);
                        ocs.push_back(OutCol{OUT_TABVAL, (u32)t, 0});
                    } else {
                        D.mode = 1;
                        for (size_t c = 0; c < B.cols.size(); c++) if ((int)c != kc) {
                            D.pay[D.n_pay] = B.cols[c].ptr;
                            out_slots.push_back(B.slots[c]);
                            ocs.push_back(OutCol{OUT_TABPAY, (u32)t, D.n_pay});
                            D.n_pay++;
                        }
                    }
                    launch_build_direct(B.cols[kc].ptr, vals, (u32)B.n, tab, kmin, range, 0u, ctx->ctrl + off + 8 + (u32)t, ctx->n_sms, bs);
                }
                ctx->stats.rows_built += B.n;
            }
            if (nb > 1) {
                KB_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->st2));
                KB_CUDA(ctx, cudaStreamWaitEvent(ctx->st, ctx->ev_join, 0));
            }
            timer_end(ctx);
            auto res = std::make_unique<kb_rel>();
            res->slots = out_slots;
            const u32 n_out = (u32)out_slots.size();
            std::vector<FilterOp> fops;
            if (last && !post.ops.empty()) {
                std::map<u32, u32> remap;
                for (u32 c = 0; c < n_out; c++) remap[out_slots[c]] = c;
                if (!append_prog(&fops, post, remap)) return fail(ctx, KB_E_INVALID, "filter uses an unbound variable");
            }
            std::vector<u32*> outp(n_out);
            for (u32 c = 0; c < n_out; c++) {
                Col col;
                KB_TRY(alloc_col(ctx, PR.n, &col));
                res->cols.push_back(col);
                outp[c] = col.ptr;
            }
            if (fast) {
                ProbeFParams P{};
                P.pairs = reinterpret_cast<const uint2*>(PR.cols[0].ptr);
                P.key_is_y = PR.col_of(join_slot) == 1 ? 1u : 0u;
                P.n = (u32)PR.n;
                P.n_tiles = (u32)((PR.n + PROBEF_TILE - 1) / PROBEF_TILE);
                P.T = (u32)nb;
                for (size_t t = 0; t < nb; t++) P.tab[t] = dt[t];
                P.n_out = n_out;
                for (u32 c = 0; c < n_out; c++) { P.oc[c] = ocs[c]; P.out[c] = outp[c]; }
                P.cap = (u32)PR.n;
                P.n_ops = (u32)fops.size();
                for (size_t i = 0; i < fops.size(); i++) P.ops[i] = fops[i];
                P.nt = numtab(ctx);
                KB_TRY(ensure_tile_state(ctx, P.n_tiles));
                P.tile_state = static_cast<u64*>(ctx->tile_state->p);
                P.block_state = static_cast<u64*>(ctx->block_state->p);
                P.ordered = ctx->ordered;
                P.ticket = ctx->ctrl + off;
                P.total = ctx->ctrl + off + 1;
                P.zero_word = ctx->ctrl + off + 2;
                P.abort_flag = ctx->ctrl + off + 8;
                P.epoch = ctx->epoch++;
                timer_begin(ctx, F_PROBE);
                launch_probe_fast(P, ctx->n_sms, ctx->st);
                timer_end(ctx);
            } else {
                std::unique_ptr<kb_rel> tmp;
                const kb_rel* src = &PR;
                if (PR.pair) {  // generic probe reads columnar input
                    tmp = std::make_unique<kb_rel>();
                    tmp->slots = PR.slots; tmp->cols = PR.cols; tmp->n = PR.n; tmp->pair = true;
                    KB_TRY(unpair_rel(ctx, &tmp));
                    src = tmp.get();
                }
                ProbeDParams P{};
                P.n_pcols = (u32)src->cols.size();
                for (u32 c = 0; c < P.n_pcols; c++) P.pcol[c] = src->cols[c].ptr;
                P.key_col = (u32)src->col_of(join_slot);
                P.n = (u32)src->n;
                P.n_tiles = (u32)((src->n + PROBE_TILE - 1) / PROBE_TILE);
                P.T = (u32)nb;
                for (size_t t = 0; t < nb; t++) P.tab[t] = dt[t];
                P.n_out = n_out;
                for (u32 c = 0; c < n_out; c++) { P.oc[c] = ocs[c]; P.out[c] = outp[c]; }
                P.cap = (u32)src->n;
                P.n_ops = (u32)fops.size();
                for (size_t i = 0; i < fops.size(); i++) P.ops[i] = fops[i];
                P.nt = numtab(ctx);
                KB_TRY(ensure_tile_state(ctx, P.n_tiles));
                P.tile_state = static_cast<u64*>(ctx->tile_state->p);
                P.block_state = static_cast<u64*>(ctx->block_state->p);
                P.ordered = ctx->ordered;
                P.ticket = ctx->ctrl + off;
                P.total = ctx->ctrl + off + 1;
                P.zero_word = ctx->ctrl + off + 2;
                P.abort_flag = ctx->ctrl + off + 8;
                P.epoch = ctx->epoch++;
                timer_begin(ctx, F_PROBE);
                launch_probe_direct(P, ctx->n_sms, ctx->st);
                timer_end(ctx);
            }
            KB_CUDA(ctx, cudaGetLastError());
            ctx->stats.rows_probed += PR.n;
            KB_TRY(ctrl_read(ctx));
            for (size_t t = 0; t < nb; t++) if (ctx->h_ctrl[off + 8 + t]) {
                dup_seen = true;
                const u32 k = builds[done + t];
                if (!pats[k].p.is_var) ctx->multi_valued.insert({pats[k].p.value, key_pos(k)});
            }
            if (dup_seen) break;
            for (size_t t = 0; t < nb; t++) {  // verified duplicate-free: later builds of the same (predicate, key position) use plain stores
                const u32 k = builds[done + t];
                // only when the build side was the WHOLE (?s P ?o) relation: a filtered subset proves nothing about the rest
                if (!pats[k].p.is_var && rels[k]->pair && pushdown[k].ops.empty()) ctx->single_valued.insert({pats[k].p.value, key_pos(k)});
            }
            res->n = ctx->h_ctrl[off + 1];
            cur = std::move(res);
            done += nb;
            if (cur->n == 0) return empty_result();
        }
        if (!dup_seen) {
            *out = select_cols(*cur, all_slots);
            return KB_OK;
        }
        // a build side has duplicate keys: join what is left with the chained (multimap) operator
        KB_TRY(to_columnar());
        if (!cur) cur = std::move(rels[probe]);
        std::vector<u32> rest(builds.begin() + done, builds.end());
        std::sort(rest.begin(), rest.end(), [&](u32 a, u32 b) { return rels[a]->n < rels[b]->n; });
        for (size_t i = 0; i < rest.size(); i++) {
            std::unique_ptr<kb_rel> j;
            KB_TRY(hash_join_impl(ctx, *cur, *rels[rest[i]], (i + 1 == rest.size() && !post.ops.empty()) ? &post : nullptr, &j));
            cur = std::move(j);
        }
        *out = select_cols(*cur, all_slots);
        return KB_OK;
    }
    // general chain: natural joins, smallest build sides first
    KB_TRY(to_columnar());
    std::unique_ptr<kb_rel> cur = std::move(rels[probe]);
    std::sort(builds.begin(), builds.end(), [&](u32 a, u32 b) { return rels[a]->n < rels[b]->n; });
    for (size_t i = 0; i < builds.size(); i++) {
        std::unique_ptr<kb_rel> j;
        KB_TRY(hash_join_impl(ctx, *cur, *rels[builds[i]], (i + 1 == builds.size() && !post.ops.empty()) ? &post : nullptr, &j));
        cur = std::move(j);
    }
    *out = select_cols(*cur, all_slots);
    return KB_OK;
}

}  // namespace kb

// =================================================================================================================
// C ABI
#define KB_ENTER(ctx)                                         \
    if (!(ctx)) return KB_E_INVALID;                          \
    kb::DeviceGuard _guard((ctx)->device);                    \
    KB_TRY(kb::begin_call(ctx))

namespace kb {
kb_status group_table_create(kb_ctx* ctx, u64 slots, GroupParams* P, GroupTable* t) {
    t->slots = slots;
    t->o_val = 0;
    t->o_cnt = t->o_val + slots * 8 * sizeof(double);
    t->o_keys = t->o_cnt + slots * sizeof(unsigned long long);
    t->o_state = t->o_keys + slots * 4 * sizeof(u32);
    t->bytes = t->o_state + slots * sizeof(u32);
    KB_TRY(alloc_buf(ctx, t->bytes, &t->buf));
    char* tb = static_cast<char*>(t->buf->p);
    P->n_slots = (u32)slots;
    P->gval = (double*)(tb + t->o_val);
    P->gcnt = (unsigned long long*)(tb + t->o_cnt);
    P->gkeys = (u32*)(tb + t->o_keys);
    P->gstate = (u32*)(tb + t->o_state);
    launch_group_init(*P, ctx->st);
    return KB_OK;
}
void groups_from_host_table(const char* hb, const GroupTable& t, u32 n_group, const kb_agg* aggs, u32 n_aggs, kb_groups* g) {
    const double* hval = (const double*)(hb + t.o_val);
    const unsigned long long* hcnt = (const unsigned long long*)(hb + t.o_cnt);
    const u32* hkeys = (const u32*)(hb + t.o_keys);
    const u32* hstate = (const u32*)(hb + t.o_state);
    if (g->kinds.empty()) for (u32 a = 0; a < n_aggs; a++) g->kinds.push_back(aggs[a].kind);
    g->raw.resize(n_aggs);
    for (u64 i = 0; i < t.slots; i++) {
        if (hstate[i] != 2u) continue;
        for (u32 c = 0; c < n_group; c++) g->keys[c].push_back(hkeys[i * 4 + c]);
        g->counts.push_back(hcnt[i]);
        for (u32 a = 0; a < n_aggs; a++) {
            double v = hval[i * 8 + a];
            g->raw[a].push_back(v);
            if (aggs[a].kind == KB_AGG_AVG) v = v / (double)hcnt[i];   // execute_query.rs:1216
            if (aggs[a].kind == KB_AGG_COUNT) v = (double)hcnt[i];
            g->vals[a].push_back(v);
        }
    }
}
void groups_from_records(const GroupRecord* recs, u64 n, u32 n_group, const kb_agg* aggs, u32 n_aggs, kb_groups* g) {
    if (g->kinds.empty()) for (u32 a = 0; a < n_aggs; a++) g->kinds.push_back(aggs[a].kind);
    g->raw.resize(n_aggs);
    for (u64 i = 0; i < n; i++) {
        const GroupRecord& r = recs[i];
        for (u32 c = 0; c < n_group; c++) g->keys[c].push_back(r.keys[c]);
        g->counts.push_back(r.count);
        for (u32 a = 0; a < n_aggs; a++) {
            double v = r.raw[a];
            g->raw[a].push_back(v);
            if (aggs[a].kind == KB_AGG_AVG) v = v / (double)r.count;   // execute_query.rs:1216
            if (aggs[a].kind == KB_AGG_COUNT) v = (double)r.count;
            g->vals[a].push_back(v);
        }
    }
}
kb_status group_table_collect(kb_ctx* ctx, const GroupTable& t, u32 n_group, const kb_agg* aggs, u32 n_aggs, kb_groups* g) {
    std::vector<char> big;  // tables past 32 MB (hundreds of thousands of groups) are not worth pinning
    void* dst = nullptr;
    if (t.bytes <= (32u << 20)) {
        if (ctx->pinned_bytes < t.bytes) {
            if (ctx->pinned) cudaFreeHost(ctx->pinned);
            ctx->pinned = nullptr; ctx->pinned_bytes = 0;
            KB_CUDA(ctx, cudaMallocHost(&ctx->pinned, t.bytes));
            ctx->pinned_bytes = t.bytes;
        }
        dst = ctx->pinned;
    } else {
        big.resize(t.bytes);
        dst = big.data();
    }
    KB_CUDA(ctx, cudaMemcpyAsync(dst, t.buf->p, t.bytes, cudaMemcpyDeviceToHost, ctx->st));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
    ctx->stats.d2h_bytes += t.bytes;
    groups_from_host_table(static_cast<const char*>(dst), t, n_group, aggs, n_aggs, g);
    return KB_OK;
}
}  // namespace kb

extern "C" {

const char* kb_version(void) { return "kolibrie_b200 0.1.0 (sm_100a)"; }

kb_status kb_ctx_create(int device, kb_ctx** out) {
    if (!out) return KB_E_INVALID;
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return kb::fail(nullptr, KB_E_CUDA, "no CUDA device available (%s): this library has no CPU fallback", cudaGetErrorString(e));
    }
    if (device < 0 || device >= n_dev) return kb::fail(nullptr, KB_E_INVALID, "device %d out of range (0..%d)", device, n_dev - 1);
    kb::DeviceGuard guard(device);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return kb::fail(nullptr, KB_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10) return kb::fail(nullptr, KB_E_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    kb_ctx* ctx = new kb_ctx;
    ctx->device = device;
    ctx->n_sms = prop.multiProcessorCount;
    auto bail = [&](const char* what, cudaError_t err) {
        kb::fail(nullptr, KB_E_CUDA, "%s: %s", what, cudaGetErrorString(err));
        delete ctx;
        return KB_E_CUDA;
    };
    if ((e = cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->st_copy, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->st2, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;  // keep freed blocks cached: steady-state queries never hit cudaMalloc
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    if (const char* ord = getenv("KOLIBRIE_ORDERED")) ctx->ordered = (ord[0] == '0') ? 0u : 1u;
    if (const char* ui = getenv("KOLIBRIE_USE_INDEX")) ctx->use_index = ui[0] != '0';
    if (const char* fk = getenv("KOLIBRIE_INDEX_KERNEL")) ctx->fast_index_kernel = fk[0] != '0';
    if (const char* cj = getenv("KOLIBRIE_CSR_JOIN")) ctx->csr_join = cj[0] != '0';
    if (const char* im = getenv("KOLIBRIE_INDEX_MAINTAIN")) ctx->index_maintain = im[0] != '0';
    if (const char* tm = getenv("KOLIBRIE_PROBE_TABLE")) ctx->probe_table_mode = tm[0] != '0';
    if (const char* dp = getenv("KOLIBRIE_DERIVE_PART")) ctx->derive_part = dp[0] != '0';
    if (const char* ds = getenv("KOLIBRIE_DERIVE_SLICE")) ctx->derive_slice_bytes = std::max<u64>(64, strtoull(ds, nullptr, 10));
    if (const char* dk = getenv("KOLIBRIE_DERIVE_SLACK")) ctx->derive_bucket_slack = strtoull(dk, nullptr, 10);
    if (const char* dr = getenv("KOLIBRIE_DERIVE_MIN_ROWS")) ctx->derive_min_part_rows = std::max<u64>(1, strtoull(dr, nullptr, 10));
    if ((e = cudaMalloc(&ctx->ctrl, kb_ctx::CTRL_WORDS * sizeof(u32))) != cudaSuccess) return bail("cudaMalloc(ctrl)", e);
    if ((e = cudaMallocHost(&ctx->h_ctrl, kb_ctx::CTRL_WORDS * sizeof(u32))) != cudaSuccess) return bail("cudaMallocHost(ctrl)", e);
    if ((e = cudaMalloc(&ctx->fast_cb, 64)) != cudaSuccess) return bail("cudaMalloc(fast_cb)", e);
    if ((e = cudaMemset(ctx->fast_cb, 0, 64)) != cudaSuccess) return bail("cudaMemset(fast_cb)", e);
    if ((e = cudaHostAlloc(&ctx->h_fast, 64, cudaHostAllocMapped)) != cudaSuccess) return bail("cudaHostAlloc(h_fast)", e);
    if ((e = cudaHostGetDevicePointer(&ctx->d_fast, ctx->h_fast, 0)) != cudaSuccess) return bail("cudaHostGetDevicePointer", e);
    *out = ctx;
    return KB_OK;
}

void kb_ctx_destroy(kb_ctx* ctx) {
    if (!ctx) return;
    kb::DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->st);
    cudaStreamSynchronize(ctx->st_copy);
    kb::timers_flush(ctx);
    ctx->segs.clear();
    ctx->index.clear();
    ctx->num.reset();
    ctx->isnum.reset();
    ctx->i32val.reset();
    ctx->isi32.reset();
    ctx->dict_off.reset();
    ctx->dict_bytes.reset();
    ctx->tile_state.reset();
    ctx->block_state.reset();
    for (auto& cb : ctx->life->cache) cudaFreeAsync(cb.p, ctx->st);
    ctx->life->cache.clear();
    ctx->life->cached_bytes = 0;
    cudaStreamSynchronize(ctx->st);
    ctx->life->alive = false;  // buffers still referenced by live relations fall back to cudaFree
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->ctrl) cudaFree(ctx->ctrl);
    if (ctx->h_ctrl) cudaFreeHost(ctx->h_ctrl);
    if (ctx->fast_cb) cudaFree(ctx->fast_cb);
    if (ctx->h_fast) cudaFreeHost(ctx->h_fast);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    cudaEventDestroy(ctx->ev_copy);
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_join);
    cudaStreamSynchronize(ctx->st2);
    cudaStreamDestroy(ctx->st2);
    cudaStreamDestroy(ctx->st);
    cudaStreamDestroy(ctx->st_copy);
    delete ctx;
}

const char* kb_last_error(const kb_ctx* ctx) { return ctx ? ctx->err.c_str() : kb::g_create_err.c_str(); }

kb_status kb_set_timing(kb_ctx* ctx, int enabled) {
    if (!ctx) return KB_E_INVALID;
    ctx->timing = enabled != 0;
    return KB_OK;
}
kb_status kb_get_stats(kb_ctx* ctx, kb_stats* out, int reset) {
    if (!ctx || !out) return KB_E_INVALID;
    kb::DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->st);
    kb::timers_flush(ctx);
    *out = ctx->stats;
    if (reset) ctx->stats = kb_stats{};
    return KB_OK;
}
kb_status kb_synchronize(kb_ctx* ctx) {
    if (!ctx) return KB_E_INVALID;
    kb::DeviceGuard guard(ctx->device);
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->st_copy));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
    kb::timers_flush(ctx);
    return KB_OK;
}

// ------------------------------------------------------------------ store
static kb_status store_add_segment(kb_ctx* ctx, const u32* s, const u32* p, const u32* o, u64 n, u64 tag, cudaMemcpyKind kind) {
    if (n && (!s || !p || !o)) return kb::fail(ctx, KB_E_INVALID, "NULL column pointer");
    if (ctx->n_triples + n >= 0xFFFFFFF0ull) return kb::fail(ctx, KB_E_LIMIT, "store would exceed 2^32-16 triples");
    kb::Segment sg;
    sg.tag = tag;
    sg.n = n;
    KB_TRY(kb::alloc_col(ctx, n, &sg.s));
    KB_TRY(kb::alloc_col(ctx, n, &sg.p));
    KB_TRY(kb::alloc_col(ctx, n, &sg.o));
    if (n) {
        static const bool trace = getenv("KOLIBRIE_TRACE") != nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        // one stream, back to back: three concurrent H2D copies on three streams measured 3x SLOWER at 16 MB per column
        // (2.8 ms vs 0.9 ms for 4 M triples) and no faster at 4 MB
        KB_CUDA(ctx, cudaMemcpyAsync(sg.s.ptr, s, n * sizeof(u32), kind, ctx->st));
        KB_CUDA(ctx, cudaMemcpyAsync(sg.p.ptr, p, n * sizeof(u32), kind, ctx->st));
        KB_CUDA(ctx, cudaMemcpyAsync(sg.o.ptr, o, n * sizeof(u32), kind, ctx->st));
        if (trace) KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
        const auto t1 = std::chrono::steady_clock::now();
        if (kind == cudaMemcpyHostToDevice) ctx->stats.h2d_bytes += 3 * n * sizeof(u32);
        // the statistics pass ends with a read-back on the same stream: when it returns the copies are done too (the caller's
        // buffers are borrowed for this call only)
        KB_TRY(kb::segment_stats(ctx, &sg));
        if (trace)
            fprintf(stderr, "[kb trace] segment of %llu triples: copies %.3f ms, statistics %.3f ms\n", (unsigned long long)n,
                    std::chrono::duration<double, std::milli>(t1 - t0).count(),
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
    }
    return kb::store_add_device_segment(ctx, sg);
}

}  // extern "C"
namespace kb {
// a segment whose columns are already in HBM joins the store: version bump, cached knowledge dropped, index maintained (or dropped)
kb_status store_add_device_segment(kb_ctx* ctx, Segment& sg) {
    if (sg.n && !sg.has_stats) KB_TRY(segment_stats(ctx, &sg));
    const u64 n = sg.n;
    const bool maintain = ctx->index_maintain && ctx->index_version == ctx->store_version && !ctx->index.empty();
    ctx->segs.push_back(sg);
    ctx->n_triples += n;
    ctx->store_version++;
    ctx->multi_valued.clear();
    ctx->single_valued.clear();
    if (maintain) {
        // the index stays alive across the mutation: the new segment becomes one more chunk of every predicate slice it touches and
        // its keys are inserted into the persistent tables in place (an RSP window slide keeps the one-kernel index path)
        bool indexable = true;
        const kb_status rc = kb::index_add_segment(ctx, ctx->segs.size() - 1, &indexable);
        if (rc == KB_OK && indexable) { ctx->index_version = ctx->store_version; return KB_OK; }
        ctx->index.clear();
        ctx->index_version = ~0ull;
        return rc;
    }
    ctx->index.clear();  // the predicate-partitioned index describes the previous store version
    ctx->index_version = ~0ull;
    return KB_OK;
}
}  // namespace kb
extern "C" {

kb_status kb_store_clear(kb_ctx* ctx) {
    KB_ENTER(ctx);
    ctx->segs.clear();
    ctx->n_triples = 0;
    ctx->store_version++;
    ctx->multi_valued.clear();
    ctx->single_valued.clear();
    ctx->index.clear();  // the predicate-partitioned index describes the previous store version
    return KB_OK;
}
kb_status kb_store_load(kb_ctx* ctx, const uint32_t* s, const uint32_t* p, const uint32_t* o, uint64_t n) {
    KB_ENTER(ctx);
    ctx->index.clear();
    ctx->index_version = ~0ull;  // a reload replaces the store: nothing to maintain
    ctx->segs.clear();
    ctx->n_triples = 0;
    return store_add_segment(ctx, s, p, o, n, 0, cudaMemcpyHostToDevice);
}
kb_status kb_store_load_device(kb_ctx* ctx, const uint32_t* s, const uint32_t* p, const uint32_t* o, uint64_t n) {
    KB_ENTER(ctx);
    ctx->index.clear();
    ctx->index_version = ~0ull;  // a reload replaces the store: nothing to maintain
    ctx->segs.clear();
    ctx->n_triples = 0;
    return store_add_segment(ctx, s, p, o, n, 0, cudaMemcpyDeviceToDevice);
}
kb_status kb_store_append(kb_ctx* ctx, const uint32_t* s, const uint32_t* p, const uint32_t* o, uint64_t n, uint64_t tag) {
    KB_ENTER(ctx);
    return store_add_segment(ctx, s, p, o, n, tag, cudaMemcpyHostToDevice);
}
kb_status kb_store_append_device(kb_ctx* ctx, const uint32_t* s, const uint32_t* p, const uint32_t* o, uint64_t n, uint64_t tag) {
    KB_ENTER(ctx);
    return store_add_segment(ctx, s, p, o, n, tag, cudaMemcpyDeviceToDevice);
}
kb_status kb_store_evict(kb_ctx* ctx, uint64_t tag) {
    KB_ENTER(ctx);
    bool found = false;
    for (size_t i = 0; i < ctx->segs.size();) {
        if (ctx->segs[i].tag == tag) {
            ctx->n_triples -= ctx->segs[i].n;
            ctx->segs.erase(ctx->segs.begin() + i);
            found = true;
        } else i++;
    }
    const bool maintain = ctx->index_maintain && ctx->index_version == ctx->store_version && !ctx->index.empty();
    ctx->store_version++;
    ctx->multi_valued.clear();
    ctx->single_valued.clear();
    if (maintain && found && kb::index_evict_tag(ctx, tag) == KB_OK) ctx->index_version = ctx->store_version;
    else { ctx->index.clear(); ctx->index_version = ~0ull; }
    return found ? KB_OK : kb::fail(ctx, KB_E_NOT_FOUND, "no segment with tag %llu", (unsigned long long)tag);
}
kb_status kb_set_use_index(kb_ctx* ctx, int enabled) {
    if (!ctx) return KB_E_INVALID;
    ctx->use_index = enabled != 0;
    return KB_OK;
}

}  // extern "C"

// ---- store index, one chunk per (predicate, store segment)
namespace kb {
// (re)build or extend the persistent table of one column of a slice after `added` (index into ps.chunks; -1: all chunks) arrived.
// flag_word: control word the insert kernels raise on a duplicate key. Returns through *touched whether a flag has to be read.
static kb_status slice_table_update(kb_ctx* ctx, PredSlice& ps, u32 y, int added, u32 cshift, u32 flag_off, bool* touched) {
    Buf& tab = y ? ps.ytab : ps.xtab;
    u32& tmin = y ? ps.ytab_min : ps.xtab_min;
    u32& tcap = y ? ps.ytab_range : ps.xtab_range;
    bool& tried = y ? ps.y_tried : ps.x_tried;
    bool& unique = y ? ps.y_unique : ps.x_unique;
    *touched = false;
    if (tried && !tab) return KB_OK;  // the column was found multi-valued or too sparse: no table, nothing to maintain
    const u32 cs = y ? 0u : cshift;    // only subjects are sharded
    const u32 lo = compact_key(y ? ps.ymin : ps.xmin, cs), hi = compact_key(y ? ps.ymax : ps.xmax, cs);
    const u64 range = (u64)hi - lo + 1;
    const bool fits = tab && lo >= tmin && (u64)hi < (u64)tmin + tcap && (y || ps.tab_cshift == cshift);
    if (!fits) {
        // (re)build over every chunk, with headroom above the largest key: dictionary ids grow, so appended segments bring larger ones
        tab.reset();
        if (!y) ps.xnum.reset();
        tried = true;
        unique = false;
        if (range > std::max<u64>(4 * ps.n + 65536, 1ull << 16) || range > (1ull << 28)) return KB_OK;  // not dense: keep no table
        // headroom: half the range for large tables, twice the range for small ones (a sliding window's ids move up by a slide per
        // slide: a 10-slide window then re-bases its tables every ~20 slides instead of every ~5)
        const u64 cap = std::min<u64>(range + (range <= (4ull << 20) ? 2 * range : range / 2) + 65536, 1ull << 29);
        KB_TRY(alloc_buf(ctx, cap * sizeof(u32), &tab));
        KB_CUDA(ctx, cudaMemsetAsync(tab->p, 0xFF, cap * sizeof(u32), ctx->st));
        tmin = lo;
        tcap = (u32)cap;
        if (!y) ps.tab_cshift = cshift;
        added = -1;
        unique = true;  // until a flag says otherwise
    }
    for (size_t c = 0; c < ps.chunks.size(); c++) {
        if (added >= 0 && (int)c != added) continue;
        const SliceChunk& ch = ps.chunks[c];
        if (!ch.n) continue;
        launch_build_direct_pairs(reinterpret_cast<const uint2*>(ch.pairs.ptr), y, (u32)ch.n, static_cast<u32*>(tab->p), tmin, tcap, cs, ctx->ctrl + flag_off, 0u,
                                  ctx->n_sms, ctx->st);
        ctx->stats.kernel_launches++;
        *touched = true;
    }
    // a subject table rebuilt outside the full build takes its typed column along (the table-mode probe filters on it)
    if (!y && !fits && !ctx->in_full_index_build && ctx->n_ids && ps.typed(ctx->num_version)) {
        KB_TRY(alloc_buf(ctx, (size_t)tcap * sizeof(double), &ps.xnum));
        KB_CUDA(ctx, cudaMemsetAsync(ps.xnum->p, 0, (size_t)tcap * sizeof(double), ctx->st));
        for (auto& ch : ps.chunks) {
            if (!ch.n) continue;
            launch_pair_numtab(reinterpret_cast<const uint2*>(ch.pairs.ptr), (u32)ch.n, numtab(ctx), static_cast<double*>(ps.xnum->p), tmin, tcap, cs, ctx->n_sms, ctx->st);
            ctx->stats.kernel_launches++;
        }
        ps.xnum_version = ctx->num_version;
    }
    return KB_OK;
}

// index the triples of ONE store segment: a chunk per predicate it carries, ranges, persistent tables, typed literal columns
// the slide-sized case of index_add_segment: the segment's predicates and their row counts are known from its profile pass, so the
// chunks are allocated exactly and ONE kernel files every triple (segment_split_kernel); one read-back
static kb_status index_add_segment_split(kb_ctx* ctx, const Segment& sg, u32 cshift) {
    const u32 k = (u32)sg.pred_rows.size();
    SplitParams P{};
    P.s = sg.s.ptr; P.p = sg.p.ptr; P.o = sg.o.ptr;
    P.n = (u32)sg.n;
    P.k = k;
    P.nt = numtab(ctx);
    const u32 off = ctrl_alloc(ctx, k * SPLIT_WORDS);
    KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off, 0, k * SPLIT_WORDS * sizeof(u32), ctx->st));
    P.ctrl = ctx->ctrl + off;
    std::vector<SliceChunk> chs(k);
    bool rebuild[MAXP][2];
    for (u32 i = 0; i < k; i++) {
        PredSlice& ps = ctx->index[sg.pred_rows[i].first];
        SliceChunk& ch = chs[i];
        ch.tag = sg.tag;
        ch.n = sg.pred_rows[i].second;
        KB_TRY(alloc_col(ctx, 2 * ch.n, &ch.pairs));
        if (ctx->n_ids) KB_TRY(alloc_buf(ctx, ch.n * sizeof(double), &ch.ynum));
        SplitEntry& e = P.e[i];
        e.pred = sg.pred_rows[i].first;
        e.n = (u32)ch.n;
        e.pairs = reinterpret_cast<uint2*>(ch.pairs.ptr);
        e.ynum = ch.ynum ? static_cast<double*>(ch.ynum->p) : nullptr;
        const bool xok = ps.xtab && ps.tab_cshift == cshift;
        rebuild[i][0] = ps.xtab ? !xok : !ps.x_tried;
        rebuild[i][1] = !ps.ytab && !ps.y_tried;
        if (xok) { e.xtab = static_cast<u32*>(ps.xtab->p); e.xtab_min = ps.xtab_min; e.xtab_range = ps.xtab_range; e.cshift = cshift; }
        if (xok && ps.xnum && ps.xnum_version == ctx->num_version && ctx->n_ids) e.xnum = static_cast<double*>(ps.xnum->p);
        else ps.xnum.reset();
        if (ps.ytab) { e.ytab = static_cast<u32*>(ps.ytab->p); e.ytab_min = ps.ytab_min; e.ytab_range = ps.ytab_range; }
    }
    timer_begin(ctx, F_OTHER);
    launch_segment_split(P, ctx->n_sms, ctx->st);
    timer_end(ctx);
    ctx->stats.kernel_launches++;
    KB_CUDA(ctx, cudaGetLastError());
    KB_TRY(ctrl_read(ctx));
    const u32 uoff = ctrl_alloc(ctx, 2 * MAXP);
    bool any_touched = false;
    std::vector<char> touched(2 * k, 0);
    std::vector<int> chunk_of(k, -1);
    for (u32 i = 0; i < k; i++) {
        const u32* h = ctx->h_ctrl + off + i * SPLIT_WORDS;
        PredSlice& ps = ctx->index[sg.pred_rows[i].first];
        SliceChunk& ch = chs[i];
        if (h[SPLIT_CURSOR] != ch.n) return fail(ctx, KB_E_CUDA, "index maintenance: predicate %u has %u rows, its profile said %llu", sg.pred_rows[i].first,
                                                  h[SPLIT_CURSOR], (unsigned long long)ch.n);
        ch.xmin = ~h[SPLIT_XMIN]; ch.ymin = ~h[SPLIT_YMIN]; ch.xmax = h[SPLIT_XMAX]; ch.ymax = h[SPLIT_YMAX];
        if (ch.ynum && h[SPLIT_NNUM] > 0) ch.ynum_version = ctx->num_version;
        else { ch.ynum.reset(); ps.xnum.reset(); }
        ps.xmin = std::min(ps.xmin, ch.xmin); ps.ymin = std::min(ps.ymin, ch.ymin);
        ps.xmax = std::max(ps.xmax, ch.xmax); ps.ymax = std::max(ps.ymax, ch.ymax);
        chunk_of[i] = (int)ps.chunks.size();
        ps.chunks.push_back(ch);
        ps.n += ch.n;
        if (ps.chunks.size() > 1) { ps.xoff.reset(); ps.xval.reset(); ps.yoff.reset(); ps.yval.reset(); }  // the directories describe one chunk
        if (h[SPLIT_XDUP]) { ps.x_unique = false; ps.xtab.reset(); ps.xnum.reset(); }  // a subject occurs twice
        if (h[SPLIT_YDUP]) { ps.y_unique = false; ps.ytab.reset(); }
        if (h[SPLIT_XOUT] && ps.xtab) rebuild[i][0] = true;  // a key outside the table: rebuild it over every chunk, with new headroom
        if (h[SPLIT_YOUT] && ps.ytab) rebuild[i][1] = true;
    }
    for (u32 i = 0; i < k; i++) {
        PredSlice& ps = ctx->index[sg.pred_rows[i].first];
        for (u32 y = 0; y < 2; y++) {
            if (!rebuild[i][y]) continue;
            if (!any_touched) { KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + uoff, 0, 2 * MAXP * sizeof(u32), ctx->st)); any_touched = true; }
            Buf& tab = y ? ps.ytab : ps.xtab;
            tab.reset();  // slice_table_update sees no table: a rebuild over all chunks (the new one included)
            if (y) ps.y_tried = false; else ps.x_tried = false;
            bool t = false;
            KB_TRY(slice_table_update(ctx, ps, y, chunk_of[i], cshift, uoff + 2 * i + y, &t));
            touched[2 * i + y] = t ? 1 : 0;
        }
    }
    if (any_touched) {
        KB_CUDA(ctx, cudaGetLastError());
        KB_TRY(ctrl_read(ctx));
        for (u32 i = 0; i < k; i++) {
            PredSlice& ps = ctx->index[sg.pred_rows[i].first];
            if (touched[2 * i] && ctx->h_ctrl[uoff + 2 * i]) { ps.x_unique = false; ps.xtab.reset(); }
            if (touched[2 * i + 1] && ctx->h_ctrl[uoff + 2 * i + 1]) { ps.y_unique = false; ps.ytab.reset(); }
        }
    }
    return KB_OK;
}

kb_status index_add_segment(kb_ctx* ctx, size_t seg_idx, bool* indexable) {
    *indexable = true;
    Segment sg = ctx->segs[seg_idx];
    if (sg.n == 0) return KB_OK;
    static const bool split_ok = !(getenv("KOLIBRIE_INDEX_SPLIT") && getenv("KOLIBRIE_INDEX_SPLIT")[0] == '0');
    if (split_ok && !ctx->in_full_index_build && sg.has_stats && sg.stats_world == ctx->shard_world && sg.has_preds && !sg.preds_overflow &&
        sg.pred_rows.size() <= MAXP) {
        size_t n_new = 0;
        for (auto& pr : sg.pred_rows) n_new += ctx->index.count(pr.first) ? 0 : 1;
        if (ctx->index.size() + n_new <= 4096) {
            u32 cshift = 0;  // subject-sharded store: compact the subject keys (see compact_key)
            bool sharded = ctx->shard_world > 1 && (ctx->shard_world & (ctx->shard_world - 1)) == 0;
            for (auto& g : ctx->segs) if (g.n && g.has_stats && !g.sharded_ok) sharded = false;
            if (sharded) while ((1u << cshift) < ctx->shard_world) cshift++;
            return index_add_segment_split(ctx, sg, cshift);
        }
    }
    // 1. distinct predicates of the segment
    const u32 set_slots = 8192;
    Buf set;
    KB_TRY(alloc_buf(ctx, set_slots * sizeof(u32), &set));
    KB_CUDA(ctx, cudaMemsetAsync(set->p, 0xFF, set_slots * sizeof(u32), ctx->st));
    const u32 off = ctrl_alloc(ctx, 4);
    KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off, 0, 4 * sizeof(u32), ctx->st));
    timer_begin(ctx, F_OTHER);
    launch_distinct(sg.p.ptr, (u32)sg.n, static_cast<u32*>(set->p), set_slots, ctx->ctrl + off, ctx->n_sms, ctx->st);
    timer_end(ctx);
    std::vector<u32> hset(set_slots);
    KB_CUDA(ctx, cudaMemcpyAsync(hset.data(), set->p, set_slots * sizeof(u32), cudaMemcpyDeviceToHost, ctx->st));
    KB_TRY(ctrl_read(ctx));
    std::vector<u32> preds;
    for (u32 v : hset) if (v != EMPTY32) preds.push_back(v);
    std::sort(preds.begin(), preds.end());
    {
        std::set<u32> all(preds.begin(), preds.end());
        for (auto& kv : ctx->index) all.insert(kv.first);
        if (ctx->h_ctrl[off] || all.size() > 4096) { *indexable = false; return KB_OK; }  // too many predicates: keep scanning
    }
    if (!sg.has_stats || sg.stats_world != ctx->shard_world) { KB_TRY(segment_stats(ctx, &ctx->segs[seg_idx])); sg = ctx->segs[seg_idx]; }
    u32 cshift = 0;  // subject-sharded store: compact the subject keys (see compact_key)
    {
        bool sharded = ctx->shard_world > 1 && (ctx->shard_world & (ctx->shard_world - 1)) == 0;
        for (auto& g : ctx->segs) if (g.n && g.has_stats && !g.sharded_ok) sharded = false;
        if (sharded) while ((1u << cshift) < ctx->shard_world) cshift++;
    }
    // 2. one fused scan of THIS segment per 8 predicates, pair output, shrunk to a chunk; id ranges of both halves
    std::vector<Segment> one{sg}, saved;
    const u64 saved_n = ctx->n_triples;
    for (size_t b = 0; b < preds.size(); b += MAXP) {
        const u32 k = (u32)std::min<size_t>(MAXP, preds.size() - b);
        kb_pattern pats[MAXP];
        for (u32 i = 0; i < k; i++) { pats[i].s = kb_term{1, 0}; pats[i].p = kb_term{0, preds[b + i]}; pats[i].o = kb_term{1, 1}; }
        std::vector<std::unique_ptr<kb_rel>> rels;
        std::vector<FilterProg> none;
        ctx->segs.swap(one);
        ctx->n_triples = sg.n;
        const kb_status rc = scan_impl(ctx, pats, k, none, false, true, &rels);
        ctx->segs.swap(one);
        ctx->n_triples = saved_n;
        if (rc != KB_OK) return rc;
        const u32 soff = ctrl_alloc(ctx, 4 * MAXP);
        KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + soff, 0, 4 * MAXP * sizeof(u32), ctx->st));
        for (u32 i = 0; i < k; i++) KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + soff + 4 * i, 0xFF, 2 * sizeof(u32), ctx->st));
        std::vector<int> chunk_of(k, -1);
        for (u32 i = 0; i < k; i++) {
            if (rels[i]->n == 0) continue;
            PredSlice& ps = ctx->index[preds[b + i]];
            SliceChunk ch;
            ch.tag = sg.tag;
            ch.n = rels[i]->n;
            KB_TRY(alloc_col(ctx, 2 * ch.n, &ch.pairs));
            KB_CUDA(ctx, cudaMemcpyAsync(ch.pairs.ptr, rels[i]->cols[0].ptr, ch.n * sizeof(uint2), cudaMemcpyDeviceToDevice, ctx->st));
            launch_pair_minmax(reinterpret_cast<const uint2*>(ch.pairs.ptr), (u32)ch.n, ctx->ctrl + soff + 4 * i, ctx->n_sms, ctx->st);
            ctx->stats.kernel_launches++;
            chunk_of[i] = (int)ps.chunks.size();
            ps.chunks.push_back(ch);
            ps.n += ch.n;
            if (ps.chunks.size() > 1) { ps.xoff.reset(); ps.xval.reset(); ps.yoff.reset(); ps.yval.reset(); }  // the directories describe one chunk
            if (!ctx->in_full_index_build) ps.xnum.reset();  // (kept in table order by the full build only)
        }
        KB_CUDA(ctx, cudaGetLastError());
        KB_TRY(ctrl_read(ctx));
        // 3. persistent tables (inserted in place; a duplicate key or a key outside the table ends / rebuilds it) and the typed literal
        //    column of the new chunk: the f64 value of every object, so that FILTER(?o <cmp> c) reads it sequentially instead of
        //    gathering num_or0[object] at random
        const u32 uoff = ctrl_alloc(ctx, 2 * MAXP);
        const u32 noff = ctrl_alloc(ctx, MAXP);
        KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + uoff, 0, 2 * MAXP * sizeof(u32), ctx->st));
        KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + noff, 0, MAXP * sizeof(u32), ctx->st));
        std::vector<char> touched(2 * k, 0);
        for (u32 i = 0; i < k; i++) {
            if (chunk_of[i] < 0) continue;
            PredSlice& ps = ctx->index[preds[b + i]];
            {
                SliceChunk& ch = ps.chunks[chunk_of[i]];
                ch.xmin = ctx->h_ctrl[soff + 4 * i]; ch.ymin = ctx->h_ctrl[soff + 4 * i + 1];
                ch.xmax = ctx->h_ctrl[soff + 4 * i + 2]; ch.ymax = ctx->h_ctrl[soff + 4 * i + 3];
                ps.xmin = std::min(ps.xmin, ch.xmin); ps.ymin = std::min(ps.ymin, ch.ymin);
                ps.xmax = std::max(ps.xmax, ch.xmax); ps.ymax = std::max(ps.ymax, ch.ymax);
            }
            for (u32 y = 0; y < 2; y++) {
                bool t = false;
                KB_TRY(slice_table_update(ctx, ps, y, chunk_of[i], cshift, uoff + 2 * i + y, &t));
                touched[2 * i + y] = t ? 1 : 0;
            }
            if (ctx->n_ids) {
                SliceChunk& ch = ps.chunks[chunk_of[i]];
                KB_TRY(alloc_buf(ctx, ch.n * sizeof(double), &ch.ynum));
                launch_pair_numcol(reinterpret_cast<const uint2*>(ch.pairs.ptr), (u32)ch.n, numtab(ctx), static_cast<double*>(ch.ynum->p), ctx->ctrl + noff + i,
                                   ctx->n_sms, ctx->st);
                ctx->stats.kernel_launches++;
            }
        }
        KB_CUDA(ctx, cudaGetLastError());
        KB_TRY(ctrl_read(ctx));
        for (u32 i = 0; i < k; i++) {
            if (chunk_of[i] < 0) continue;
            PredSlice& ps = ctx->index[preds[b + i]];
            SliceChunk& ch = ps.chunks[chunk_of[i]];
            if (ch.ynum && ctx->h_ctrl[noff + i] > 0) ch.ynum_version = ctx->num_version;
            else ch.ynum.reset();
            if (touched[2 * i] && ctx->h_ctrl[uoff + 2 * i]) { ps.x_unique = false; ps.xtab.reset(); }      // a subject occurs twice
            if (touched[2 * i + 1] && ctx->h_ctrl[uoff + 2 * i + 1]) { ps.y_unique = false; ps.ytab.reset(); }
        }
    }
    return KB_OK;
}

// eviction of the segment(s) tagged `tag`: their chunks leave the slices, their keys leave the persistent tables
kb_status index_evict_tag(kb_ctx* ctx, u64 tag) {
    ClearParams C{};
    std::vector<SliceChunk> dying;  // kept until the clears are queued: their buffers are released stream-ordered behind them
    auto flush = [&]() {
        if (C.k) { launch_clear_chunks(C, ctx->n_sms, ctx->st); ctx->stats.kernel_launches++; }
        C.k = 0;
    };
    for (auto it = ctx->index.begin(); it != ctx->index.end();) {
        PredSlice& ps = it->second;
        bool hit = false;
        for (size_t c = 0; c < ps.chunks.size();) {
            SliceChunk& ch = ps.chunks[c];
            if (ch.tag != tag) { c++; continue; }
            hit = true;
            if ((ps.xtab || ps.ytab) && ch.n) {
                ClearEntry& e = C.e[C.k++];
                e = ClearEntry{};
                e.pairs = reinterpret_cast<const uint2*>(ch.pairs.ptr);
                e.n = (u32)ch.n;
                if (ps.xtab) { e.xtab = static_cast<u32*>(ps.xtab->p); e.xtab_min = ps.xtab_min; e.xtab_range = ps.xtab_range; e.cshift = ps.tab_cshift; }
                if (ps.ytab) { e.ytab = static_cast<u32*>(ps.ytab->p); e.ytab_min = ps.ytab_min; e.ytab_range = ps.ytab_range; }
                if (C.k == CLEAR_MAX) flush();
            }
            ps.n -= ch.n;
            dying.push_back(ch);
            ps.chunks.erase(ps.chunks.begin() + c);
            ps.xoff.reset(); ps.xval.reset(); ps.yoff.reset(); ps.yval.reset();  // (xnum stays: its slots are only read where xtab holds a key)
        }
        if (hit) {  // the slice's id ranges shrink to the surviving chunks': a long-running window does not walk its whole past
            ps.xmin = ps.ymin = 0xFFFFFFFFu;
            ps.xmax = ps.ymax = 0u;
            for (auto& ch : ps.chunks) {
                ps.xmin = std::min(ps.xmin, ch.xmin); ps.ymin = std::min(ps.ymin, ch.ymin);
                ps.xmax = std::max(ps.xmax, ch.xmax); ps.ymax = std::max(ps.ymax, ch.ymax);
            }
        }
        if (ps.chunks.empty()) it = ctx->index.erase(it);
        else ++it;
    }
    flush();
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}
}  // namespace kb

extern "C" {

kb_status kb_store_build_index(kb_ctx* ctx, uint32_t* n_predicates, double* build_ms) {
    KB_ENTER(ctx);
    if (n_predicates) *n_predicates = 0;
    if (build_ms) *build_ms = 0.0;
    ctx->index.clear();
    ctx->index_version = ~0ull;
    if (ctx->n_triples == 0) return KB_OK;
    kb::ScopedEvent e0, e1;
    cudaEventRecord(e0, ctx->st);
    for (auto& sg : ctx->segs) if (sg.n && (!sg.has_stats || sg.stats_world != ctx->shard_world)) KB_TRY(kb::segment_stats(ctx, &sg));
    for (size_t g = 0; g < ctx->segs.size(); g++) {
        bool indexable = true;
        ctx->in_full_index_build = true;
        const kb_status rc = kb::index_add_segment(ctx, g, &indexable);
        ctx->in_full_index_build = false;
        if (rc != KB_OK || !indexable) {
            ctx->index.clear();
            if (rc != KB_OK) return rc;
            return KB_OK;  // too many predicates: keep scanning (n_predicates = 0)
        }
    }
    // typed values in table order for subject tables whose slice is numeric: the table-mode probe reads FILTER operands sequentially
    for (auto& kv : ctx->index) {
        kb::PredSlice& ps = kv.second;
        if (!ps.xtab || !ctx->n_ids || !ps.typed(ctx->num_version)) continue;
        KB_TRY(kb::alloc_buf(ctx, (size_t)ps.xtab_range * sizeof(double), &ps.xnum));
        KB_CUDA(ctx, cudaMemsetAsync(ps.xnum->p, 0, (size_t)ps.xtab_range * sizeof(double), ctx->st));
        for (auto& ch : ps.chunks) {
            kb::launch_pair_numtab(reinterpret_cast<const uint2*>(ch.pairs.ptr), (u32)ch.n, kb::numtab(ctx), static_cast<double*>(ps.xnum->p), ps.xtab_min, ps.xtab_range,
                                   ps.tab_cshift, ctx->n_sms, ctx->st);
            ctx->stats.kernel_launches++;
        }
        ps.xnum_version = ctx->num_version;
    }
    // key-grouped directories (counting sort by the dense key) for the columns that did not get a direct table
    for (auto& kv : ctx->index) {
        kb::PredSlice& ps = kv.second;
        const kb::SliceChunk* ch = ps.single();
        if (!ch || ch->n == 0) continue;
        for (u32 y = 0; y < 2; y++) {
            if (y ? (bool)ps.ytab : (bool)ps.xtab) continue;  // unique column: the direct table answers the lookup
            const u32 lo = y ? ps.ymin : ps.xmin, hi = y ? ps.ymax : ps.xmax;
            const u64 range = (u64)hi - lo + 1;
            if (range > std::max<u64>(8 * ps.n, 1ull << 20) || range > (1ull << 28)) continue;  // sparse ids: no directory
            kb::Buf off, val, cursor, scratch;
            KB_TRY(kb::alloc_buf(ctx, (range + 1) * sizeof(u32), &off));
            KB_TRY(kb::alloc_buf(ctx, ch->n * sizeof(u32), &val));
            KB_TRY(kb::alloc_buf(ctx, range * sizeof(u32), &cursor));
            KB_TRY(kb::alloc_buf(ctx, ((range + 1) / 2048 + 4) * sizeof(u32), &scratch));
            KB_CUDA(ctx, cudaMemsetAsync(off->p, 0, (range + 1) * sizeof(u32), ctx->st));
            const uint2* pairs = reinterpret_cast<const uint2*>(ch->pairs.ptr);
            kb::launch_csr_count_pairs(pairs, y, (u32)ch->n, lo, static_cast<u32*>(off->p), ctx->n_sms, ctx->st);
            kb::launch_exclusive_scan_u32(static_cast<u32*>(off->p), (u32)range + 1, static_cast<u32*>(scratch->p), ctx->st);
            KB_CUDA(ctx, cudaMemcpyAsync(cursor->p, off->p, range * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->st));
            kb::launch_csr_fill_pairs(pairs, y, (u32)ch->n, lo, static_cast<u32*>(cursor->p), static_cast<u32*>(val->p), ctx->n_sms, ctx->st);
            ctx->stats.kernel_launches += 5;
            if (y) { ps.yoff = off; ps.yval = val; ps.ycsr_min = lo; ps.ycsr_range = (u32)range; }
            else { ps.xoff = off; ps.xval = val; ps.xcsr_min = lo; ps.xcsr_range = (u32)range; }
        }
    }
    KB_CUDA(ctx, cudaGetLastError());
    cudaEventRecord(e1, ctx->st);
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    ctx->index_version = ctx->store_version;
    if (n_predicates) *n_predicates = (uint32_t)ctx->index.size();
    if (build_ms) *build_ms = ms;
    return KB_OK;
}

kb_status kb_store_size(kb_ctx* ctx, uint64_t* n, uint32_t* n_seg) {
    if (!ctx) return KB_E_INVALID;
    if (n) *n = ctx->n_triples;
    if (n_seg) *n_seg = (uint32_t)ctx->segs.size();
    return KB_OK;
}

kb_status kb_store_delete(kb_ctx* ctx, const uint32_t* s, const uint32_t* p, const uint32_t* o, uint64_t n) {
    KB_ENTER(ctx);
    if (n == 0 || ctx->n_triples == 0) return KB_OK;
    if (!s || !p || !o) return kb::fail(ctx, KB_E_INVALID, "NULL column pointer");
    // 1. the delete set (96-bit keys), 2. mark matching triples by overwriting a copy of the predicate column with KB_ID_NONE,
    // 3. rescan every segment with (?s ?p ?o) — a row whose predicate is KB_ID_NONE can never equal a real constant, and the
    //    pattern's pushed-down filter NE_ID drops it.
    const u32 set_slots = kb::pow2_at_least(n * 2);
    kb::Buf set, ds, dp, dob;
    KB_TRY(kb::alloc_buf(ctx, (size_t)set_slots * sizeof(uint4), &set));
    KB_TRY(kb::alloc_buf(ctx, n * sizeof(u32), &ds));
    KB_TRY(kb::alloc_buf(ctx, n * sizeof(u32), &dp));
    KB_TRY(kb::alloc_buf(ctx, n * sizeof(u32), &dob));
    KB_CUDA(ctx, cudaMemsetAsync(set->p, 0, (size_t)set_slots * sizeof(uint4), ctx->st));
    KB_CUDA(ctx, cudaMemcpyAsync(ds->p, s, n * sizeof(u32), cudaMemcpyHostToDevice, ctx->st));
    KB_CUDA(ctx, cudaMemcpyAsync(dp->p, p, n * sizeof(u32), cudaMemcpyHostToDevice, ctx->st));
    KB_CUDA(ctx, cudaMemcpyAsync(dob->p, o, n * sizeof(u32), cudaMemcpyHostToDevice, ctx->st));
    const u32 off = kb::ctrl_alloc(ctx, 4);
    kb::launch_set_insert(static_cast<uint4*>(set->p), set_slots, (const u32*)ds->p, (const u32*)dp->p, 0u, (const u32*)dob->p, (u32)n,
                          ctx->ctrl + off, ctx->n_sms, ctx->st);
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
    std::vector<kb::Segment> old;
    old.swap(ctx->segs);
    ctx->n_triples = 0;
    kb_status rc = KB_OK;
    for (auto& sg : old) {
        kb::Segment marked = sg;
        if (sg.n) {
            rc = kb::alloc_col(ctx, sg.n, &marked.p);
            if (rc != KB_OK) break;  // the restore path below puts the untouched segments back
            kb::launch_delete_mark(sg.s.ptr, sg.p.ptr, sg.o.ptr, (u32)sg.n, static_cast<const uint4*>(set->p), set_slots, marked.p.ptr,
                                   ctx->n_sms, ctx->st);
        }
        // rescan this one segment
        std::vector<kb::Segment> one{marked};
        ctx->segs.swap(one);
        const kb::u64 saved = ctx->n_triples;
        ctx->n_triples = sg.n;
        kb_pattern pt{{1, 0}, {1, 1}, {1, 2}};
        kb::FilterProg keep;
        kb_filter_op tr{};
        tr.op = KB_F_IS_TRIPLE; tr.slot = 1;  // marked predicates (0xFFFFFFFF) have bit 31 set; real predicate ids never do
        kb_filter_op nt{};
        nt.op = KB_F_NOT;
        keep.ops.push_back(tr);
        keep.ops.push_back(nt);
        std::vector<kb::FilterProg> pd{keep};
        std::vector<std::unique_ptr<kb_rel>> rels;
        rc = kb::scan_impl(ctx, &pt, 1, pd, false, false, &rels);
        ctx->segs.swap(one);
        ctx->n_triples = saved;
        if (rc != KB_OK) break;
        kb::Segment ns;
        ns.tag = sg.tag;
        ns.n = rels[0]->n;
        ns.s = rels[0]->cols[0]; ns.p = rels[0]->cols[1]; ns.o = rels[0]->cols[2];
        ctx->segs.push_back(ns);
        ctx->n_triples += ns.n;
    }
    if (rc != KB_OK) { ctx->segs.swap(old); ctx->n_triples = 0; for (auto& g : ctx->segs) ctx->n_triples += g.n; return rc; }
    ctx->store_version++;
    ctx->multi_valued.clear();
    ctx->single_valued.clear();
    ctx->index.clear();  // the predicate-partitioned index describes the previous store version
    return KB_OK;
}

kb_status kb_store_download(kb_ctx* ctx, uint32_t* s, uint32_t* p, uint32_t* o, uint64_t cap, uint64_t* n) {
    KB_ENTER(ctx);
    if (n) *n = ctx->n_triples;
    if (cap < ctx->n_triples) return kb::fail(ctx, KB_E_LIMIT, "buffer holds %llu triples, store has %llu", (unsigned long long)cap, (unsigned long long)ctx->n_triples);
    u64 off = 0;
    for (auto& sg : ctx->segs) {
        if (!sg.n) continue;
        KB_CUDA(ctx, cudaMemcpyAsync(s + off, sg.s.ptr, sg.n * sizeof(u32), cudaMemcpyDeviceToHost, ctx->st));
        KB_CUDA(ctx, cudaMemcpyAsync(p + off, sg.p.ptr, sg.n * sizeof(u32), cudaMemcpyDeviceToHost, ctx->st));
        KB_CUDA(ctx, cudaMemcpyAsync(o + off, sg.o.ptr, sg.n * sizeof(u32), cudaMemcpyDeviceToHost, ctx->st));
        off += sg.n;
    }
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
    return KB_OK;
}

kb_status kb_dict_numeric_load(kb_ctx* ctx, const double* num_or0, const uint8_t* is_num, uint32_t n_ids) {
    KB_ENTER(ctx);
    ctx->num.reset();
    ctx->isnum.reset();
    ctx->n_ids = 0;
    ctx->num_version++;
    if (n_ids == 0) return KB_OK;
    if (!num_or0 || !is_num) return kb::fail(ctx, KB_E_INVALID, "NULL numeric table");
    KB_TRY(kb::alloc_buf(ctx, (size_t)n_ids * sizeof(double), &ctx->num));
    KB_TRY(kb::alloc_buf(ctx, (size_t)n_ids, &ctx->isnum));
    KB_CUDA(ctx, cudaMemcpyAsync(ctx->num->p, num_or0, (size_t)n_ids * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
    KB_CUDA(ctx, cudaMemcpyAsync(ctx->isnum->p, is_num, (size_t)n_ids, cudaMemcpyHostToDevice, ctx->st));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
    ctx->n_ids = n_ids;
    ctx->num_version++;
    return KB_OK;
}

kb_status kb_dict_legacy_i32_load(kb_ctx* ctx, const int32_t* val, const uint8_t* is_i32, uint32_t n_ids) {
    KB_ENTER(ctx);
    ctx->i32val.reset();
    ctx->isi32.reset();
    ctx->n_i32 = 0;
    if (n_ids == 0) return KB_OK;
    if (!val || !is_i32) return kb::fail(ctx, KB_E_INVALID, "NULL table");
    KB_TRY(kb::alloc_buf(ctx, (size_t)n_ids * sizeof(int32_t), &ctx->i32val));
    KB_TRY(kb::alloc_buf(ctx, (size_t)n_ids, &ctx->isi32));
    KB_CUDA(ctx, cudaMemcpyAsync(ctx->i32val->p, val, (size_t)n_ids * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->st));
    KB_CUDA(ctx, cudaMemcpyAsync(ctx->isi32->p, is_i32, (size_t)n_ids, cudaMemcpyHostToDevice, ctx->st));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
    ctx->stats.h2d_bytes += (u64)n_ids * 5;
    ctx->n_i32 = n_ids;
    return KB_OK;
}

kb_status kb_dict_strings_load(kb_ctx* ctx, const uint64_t* offsets, const uint8_t* bytes, uint32_t n_ids) {
    KB_ENTER(ctx);
    ctx->dict_off.reset();
    ctx->dict_bytes.reset();
    ctx->dict_ids = 0;
    ctx->dict_index.reset();  // (kb_dict_encode rebuilds its string -> id index lazily)
    ctx->dict_index_slots = 0;
    ctx->dict_indexed = 0;
    if (n_ids == 0) return KB_OK;
    if (!offsets) return kb::fail(ctx, KB_E_INVALID, "NULL offsets");
    const uint64_t total = offsets[n_ids];
    if (offsets[0] != 0 || (total && !bytes)) return kb::fail(ctx, KB_E_INVALID, "offsets must start at 0 and bytes must not be NULL");
    for (uint32_t i = 0; i < n_ids; i++) if (offsets[i + 1] < offsets[i]) return kb::fail(ctx, KB_E_INVALID, "offsets decrease at id %u", i);
    KB_TRY(kb::alloc_buf(ctx, ((size_t)n_ids + 1) * sizeof(uint64_t), &ctx->dict_off));
    KB_TRY(kb::alloc_buf(ctx, (size_t)total + 16, &ctx->dict_bytes));
    KB_CUDA(ctx, cudaMemcpyAsync(ctx->dict_off->p, offsets, ((size_t)n_ids + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->st));
    if (total) KB_CUDA(ctx, cudaMemcpyAsync(ctx->dict_bytes->p, bytes, (size_t)total, cudaMemcpyHostToDevice, ctx->st));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
    ctx->stats.h2d_bytes += ((size_t)n_ids + 1) * sizeof(uint64_t) + total;
    ctx->dict_ids = n_ids;
    return KB_OK;
}

kb_status kb_rel_decode(kb_ctx* ctx, const kb_rel* r, uint32_t col, kb_strings** out) {
    KB_ENTER(ctx);
    if (!r || !out) return kb::fail(ctx, KB_E_INVALID, "NULL argument");
    if (col >= r->cols.size()) return kb::fail(ctx, KB_E_INVALID, "column %u out of range", col);
    if (r->n >= 0xFFFFFFF0ull) return kb::fail(ctx, KB_E_LIMIT, "relation too large");
    auto res = std::make_unique<kb_strings>();
    res->n = r->n;
    const u32 n = (u32)r->n;
    KB_TRY(kb::alloc_buf(ctx, ((size_t)n + 1) * sizeof(u32), &res->off));
    u32* off = static_cast<u32*>(res->off->p);
    const u32 c = kb::ctrl_alloc(ctx, 4);  // [0..1] u64 total bytes, [2] quoted-triple id seen
    KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + c, 0, 4 * sizeof(u32), ctx->st));
    KB_CUDA(ctx, cudaMemsetAsync(off + n, 0, sizeof(u32), ctx->st));  // the scan runs over n + 1 entries: off[n] becomes the total
    const auto* doff = ctx->dict_ids ? static_cast<const unsigned long long*>(ctx->dict_off->p) : nullptr;
    const auto* dbytes = ctx->dict_ids ? static_cast<const unsigned char*>(ctx->dict_bytes->p) : nullptr;
    kb::Buf scratch;
    KB_TRY(kb::alloc_buf(ctx, (((size_t)n + 1) / 2048 + 4) * sizeof(u32), &scratch));
    kb::timer_begin(ctx, kb::F_OTHER, 4);
    kb::launch_decode_lengths(r->cols[col].ptr, n, doff, ctx->dict_ids, off, reinterpret_cast<unsigned long long*>(ctx->ctrl + c), ctx->ctrl + c + 2,
                              ctx->n_sms, ctx->st);
    kb::launch_exclusive_scan_u32(off, n + 1, static_cast<u32*>(scratch->p), ctx->st);
    kb::timer_end(ctx);
    KB_CUDA(ctx, cudaGetLastError());
    KB_TRY(kb::ctrl_read(ctx));
    unsigned long long total = 0;
    memcpy(&total, ctx->h_ctrl + c, sizeof total);
    if (ctx->h_ctrl[c + 2]) return kb::fail(ctx, KB_E_UNSUPPORTED, "the column holds quoted-triple ids (bit 31): decode them on the host (dictionary.rs:58-68)");
    if (total > 0xFFFFFFFFull) return kb::fail(ctx, KB_E_LIMIT, "decoded column would take %llu bytes (limit 2^32-1 per call)", total);
    res->total = total;
    KB_TRY(kb::alloc_buf(ctx, (size_t)total + 16, &res->bytes));
    kb::timer_begin(ctx, kb::F_OTHER);
    kb::launch_decode_gather(r->cols[col].ptr, n, doff, dbytes, ctx->dict_ids, off, static_cast<unsigned char*>(res->bytes->p), ctx->n_sms, ctx->st);
    kb::timer_end(ctx);
    KB_CUDA(ctx, cudaGetLastError());
    *out = res.release();
    return KB_OK;
}
kb_status kb_strings_info(const kb_strings* s, uint64_t* n_strings, uint64_t* total_bytes) {
    if (!s) return KB_E_INVALID;
    if (n_strings) *n_strings = s->n;
    if (total_bytes) *total_bytes = s->total;
    return KB_OK;
}
kb_status kb_strings_download(kb_ctx* ctx, const kb_strings* s, uint64_t* offsets, uint8_t* bytes) {
    if (!ctx || !s) return KB_E_INVALID;
    kb::DeviceGuard guard(ctx->device);
    if (offsets) {
        std::vector<u32> h((size_t)s->n + 1);
        KB_CUDA(ctx, cudaMemcpyAsync(h.data(), s->off->p, h.size() * sizeof(u32), cudaMemcpyDeviceToHost, ctx->st));
        KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
        for (size_t i = 0; i < h.size(); i++) offsets[i] = h[i];
    }
    if (bytes && s->total) {
        KB_CUDA(ctx, cudaMemcpyAsync(bytes, s->bytes->p, (size_t)s->total, cudaMemcpyDeviceToHost, ctx->st));
        KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
    }
    ctx->stats.d2h_bytes += ((size_t)s->n + 1) * sizeof(u32) + s->total;
    return KB_OK;
}
void kb_strings_free(kb_ctx* ctx, kb_strings* s) {
    if (!s) return;
    if (ctx) { kb::DeviceGuard guard(ctx->device); delete s; }
    else delete s;
}

// ------------------------------------------------------------------ relations
kb_status kb_rel_info(const kb_rel* r, uint64_t* n_rows, uint32_t* n_cols, uint32_t* slots) {
    if (!r) return KB_E_INVALID;
    if (n_rows) *n_rows = r->n;
    if (n_cols) *n_cols = (uint32_t)r->slots.size();
    if (slots) for (size_t i = 0; i < r->slots.size() && i < KB_MAX_COLS; i++) slots[i] = r->slots[i];
    return KB_OK;
}
kb_status kb_rel_download(kb_ctx* ctx, const kb_rel* r, uint32_t col, uint32_t* dst) {
    if (!ctx || !r) return KB_E_INVALID;
    kb::DeviceGuard guard(ctx->device);
    if (col >= r->cols.size()) return kb::fail(ctx, KB_E_INVALID, "column %u out of range", col);
    if (r->n == 0) return KB_OK;
    if (!dst) return kb::fail(ctx, KB_E_INVALID, "NULL destination");
    KB_CUDA(ctx, cudaMemcpyAsync(dst, r->cols[col].ptr, r->n * sizeof(u32), cudaMemcpyDeviceToHost, ctx->st));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
    ctx->stats.d2h_bytes += r->n * sizeof(u32);
    return KB_OK;
}
kb_status kb_rel_device_col(const kb_rel* r, uint32_t col, const uint32_t** d_ptr) {
    if (!r || !d_ptr || col >= r->cols.size()) return KB_E_INVALID;
    *d_ptr = r->cols[col].ptr;
    return KB_OK;
}
static kb_status rel_from(kb_ctx* ctx, const uint32_t* slots, uint32_t n_cols, const uint32_t* const* cols, uint64_t n_rows, cudaMemcpyKind kind, kb_rel** out) {
    if (!out || (n_cols && (!slots || !cols))) return kb::fail(ctx, KB_E_INVALID, "NULL argument");
    if (n_cols > KB_MAX_COLS) return kb::fail(ctx, KB_E_LIMIT, "more than %d columns", KB_MAX_COLS);
    if (n_rows >= 0xFFFFFFF0ull) return kb::fail(ctx, KB_E_LIMIT, "relation too large");
    auto r = std::make_unique<kb_rel>();
    r->n = n_rows;
    for (u32 c = 0; c < n_cols; c++) {
        for (u32 d = 0; d < c; d++) if (slots[d] == slots[c]) return kb::fail(ctx, KB_E_INVALID, "duplicate slot %u", slots[c]);
        kb::Col col;
        KB_TRY(kb::alloc_col(ctx, n_rows, &col));
        if (n_rows) KB_CUDA(ctx, cudaMemcpyAsync(col.ptr, cols[c], n_rows * sizeof(u32), kind, ctx->st));
        r->slots.push_back(slots[c]);
        r->cols.push_back(col);
    }
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
    if (kind == cudaMemcpyHostToDevice) ctx->stats.h2d_bytes += (u64)n_cols * n_rows * sizeof(u32);
    *out = r.release();
    return KB_OK;
}
kb_status kb_rel_from_host(kb_ctx* ctx, const uint32_t* slots, uint32_t n_cols, const uint32_t* const* cols, uint64_t n_rows, kb_rel** out) {
    KB_ENTER(ctx);
    return rel_from(ctx, slots, n_cols, cols, n_rows, cudaMemcpyHostToDevice, out);
}
kb_status kb_rel_from_device(kb_ctx* ctx, const uint32_t* slots, uint32_t n_cols, const uint32_t* const* cols, uint64_t n_rows, kb_rel** out) {
    KB_ENTER(ctx);
    return rel_from(ctx, slots, n_cols, cols, n_rows, cudaMemcpyDeviceToDevice, out);
}
void kb_rel_free(kb_ctx* ctx, kb_rel* r) {
    if (!r) return;
    if (ctx) {
        kb::DeviceGuard guard(ctx->device);
        delete r;
    } else delete r;
}

// ------------------------------------------------------------------ operators
kb_status kb_scan(kb_ctx* ctx, const kb_pattern* pats, uint32_t n_pats, const kb_filter_op* const* pushdown, const uint32_t* pushdown_len, kb_rel** out) {
    KB_ENTER(ctx);
    if (!pats || !out) return kb::fail(ctx, KB_E_INVALID, "NULL argument");
    std::vector<kb::FilterProg> pd(n_pats);
    if (pushdown && pushdown_len)
        for (u32 k = 0; k < n_pats; k++) if (pushdown[k] && pushdown_len[k]) {
            KB_TRY(kb::validate_filter(ctx, pushdown[k], pushdown_len[k]));
            pd[k].ops.assign(pushdown[k], pushdown[k] + pushdown_len[k]);
        }
    std::vector<std::unique_ptr<kb_rel>> rels;
    KB_TRY(kb::scan_impl(ctx, pats, n_pats, pd, false, false, &rels));
    for (u32 k = 0; k < n_pats; k++) out[k] = rels[k].release();
    if (n_pats) ctx->stats.rows_out = out[n_pats - 1]->n;
    return KB_OK;
}

kb_status kb_filter(kb_ctx* ctx, const kb_rel* in, const kb_filter_op* prog, uint32_t n_ops, kb_rel** out) {
    KB_ENTER(ctx);
    if (!in || !out) return kb::fail(ctx, KB_E_INVALID, "NULL argument");
    KB_TRY(kb::validate_filter(ctx, prog, n_ops));
    kb::FilterProg f;
    if (n_ops) f.ops.assign(prog, prog + n_ops);
    for (u32 s : kb::filter_slots(f)) if (in->col_of(s) < 0)
        return kb::fail(ctx, KB_E_UNSUPPORTED, "filter references slot %u which is not bound (the reference evaluates it to false, types.rs:149-151)", s);
    std::unique_ptr<kb_rel> r;
    KB_TRY(kb::filter_impl(ctx, *in, f, &r));
    ctx->stats.rows_out = r->n;
    *out = r.release();
    return KB_OK;
}

kb_status kb_project(kb_ctx* ctx, const kb_rel* in, const uint32_t* slots, uint32_t n_slots, kb_rel** out) {
    if (!ctx || !in || !out || (n_slots && !slots)) return KB_E_INVALID;
    std::vector<u32> v(slots, slots + n_slots);
    *out = kb::select_cols(*in, v).release();
    return KB_OK;
}

kb_status kb_hash_join(kb_ctx* ctx, const kb_rel* left, const kb_rel* right, kb_rel** out) {
    KB_ENTER(ctx);
    if (!left || !right || !out) return kb::fail(ctx, KB_E_INVALID, "NULL argument");
    std::unique_ptr<kb_rel> r;
    KB_TRY(kb::hash_join_impl(ctx, *left, *right, nullptr, &r));
    ctx->stats.rows_out = r->n;
    *out = r.release();
    return KB_OK;
}

// BindJoin (engine.rs:840-885): every row of `left` is extended by the matches of ONE store pattern. With the store index valid, the
// pattern (?x P ?y) and `left` binding exactly one of its variables through a unique, dense column, this is one probe kernel against the
// persistent table of the index (spo[x][P] / pos[P][y] lookups, no scan, no build); every other shape = scan of the pattern + natural join.
kb_status kb_bind_join(kb_ctx* ctx, const kb_rel* left, const kb_pattern* pat, kb_rel** out) {
    KB_ENTER(ctx);
    if (!left || !pat || !out) return kb::fail(ctx, KB_E_INVALID, "NULL argument");
    KB_TRY(kb::check_pattern(ctx, *pat));
    std::vector<u32> pv, psrc;
    kb::pattern_vars(*pat, &pv, &psrc);
    const bool pair_shape = pv.size() == 2 && psrc[0] == 0 && psrc[1] == 2 && !pat->p.is_var;
    if (pair_shape && ctx->use_index && ctx->index_version == ctx->store_version && left->n > 0 && left->n < 0xFFFFFFF0ull && !left->pair &&
        left->cols.size() < KB_MAX_COLS) {
        const int cx = left->col_of(pv[0]), cy = left->col_of(pv[1]);
        auto it = ctx->index.find(pat->p.value);
        if ((cx >= 0) != (cy >= 0) && it != ctx->index.end() && it->second.n > 0) {
            const kb::PredSlice& ps = it->second;
            const bool by_y = cy >= 0;  // the bound variable is the pattern's object: look up pos[P][y]
            const kb::Buf& tab = by_y ? ps.ytab : ps.xtab;
            // the persistent tables of a subject-sharded store are keyed by the COMPACTED subject: a relation that went through a
            // shuffle holds exactly the keys this shard owns, so the compaction applies to its rows too
            bool usable = (bool)tab;
            if (usable && !by_y && ps.tab_cshift != 0u) {
                // compacted keys alias across shards: the lookup is only sound when every key of `left` belongs to this shard
                const u32 foff = kb::ctrl_alloc(ctx, 4);
                KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + foff, 0, 4 * sizeof(u32), ctx->st));
                kb::timer_begin(ctx, kb::F_OTHER);
                kb::launch_count_foreign(left->cols[cx].ptr, (u32)left->n, ctx->shard_rank, ctx->shard_world, ctx->ctrl + foff, ctx->n_sms, ctx->st);
                kb::timer_end(ctx);
                KB_CUDA(ctx, cudaGetLastError());
                KB_TRY(kb::ctrl_read(ctx));
                usable = ctx->h_ctrl[foff] == 0;
            }
            if (usable) {
                kb::ProbeDParams P{};
                P.n_pcols = (u32)left->cols.size();
                P.key_col = (u32)(by_y ? cy : cx);
                P.n = (u32)left->n;
                P.n_tiles = (u32)((left->n + kb::PROBE_TILE - 1) / kb::PROBE_TILE);
                P.T = 1;
                kb::DirectTab& D = P.tab[0];
                D.tab = static_cast<const u32*>(tab->p);
                D.kmin = by_y ? ps.ytab_min : ps.xtab_min;
                D.range = by_y ? ps.ytab_range : ps.xtab_range;
                D.cshift = by_y ? 0u : ps.tab_cshift;
                D.mode = 0; D.n_pay = 0; D.pay[0] = D.pay[1] = nullptr;
                auto rel = std::make_unique<kb_rel>();
                rel->slots = left->slots;
                rel->slots.push_back(by_y ? pv[0] : pv[1]);
                P.n_out = (u32)rel->slots.size();
                const size_t stride = round256((size_t)left->n * sizeof(u32)) + 256;
                kb::Buf b;
                KB_TRY(kb::alloc_buf(ctx, stride * P.n_out, &b));
                for (u32 c2 = 0; c2 < P.n_out; c2++) {
                    kb::Col col;
                    col.buf = b;
                    col.ptr = reinterpret_cast<u32*>(static_cast<char*>(b->p) + stride * c2);
                    rel->cols.push_back(col);
                    P.out[c2] = col.ptr;
                    P.oc[c2] = c2 + 1 < P.n_out ? kb::OutCol{kb::OUT_PROBE, c2, 0} : kb::OutCol{kb::OUT_TABVAL, 0, 0};
                    if (c2 + 1 < P.n_out) P.pcol[c2] = left->cols[c2].ptr;
                }
                P.cap = (u32)left->n;
                P.nt = kb::numtab(ctx);
                KB_TRY(kb::ensure_tile_state(ctx, P.n_tiles));
                P.tile_state = static_cast<u64*>(ctx->tile_state->p);
                P.block_state = static_cast<u64*>(ctx->block_state->p);
                P.ordered = ctx->ordered;
                const u32 off = kb::ctrl_alloc(ctx, 4);
                KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off, 0, 4 * sizeof(u32), ctx->st));
                P.ticket = ctx->ctrl + off;
                P.total = ctx->ctrl + off + 1;
                P.zero_word = ctx->ctrl + off + 2;
                P.epoch = ctx->epoch++;
                P.abort_flag = nullptr;
                kb::timer_begin(ctx, kb::F_PROBE);
                kb::launch_probe_direct(P, ctx->n_sms, ctx->st);
                kb::timer_end(ctx);
                KB_CUDA(ctx, cudaGetLastError());
                ctx->stats.rows_probed += left->n;
                ctx->stats.index_joins++;
                KB_TRY(kb::ctrl_read(ctx));
                rel->n = ctx->h_ctrl[off + 1];
                ctx->stats.rows_out = rel->n;
                *out = rel.release();
                return KB_OK;
            }
        }
    }
    std::vector<std::unique_ptr<kb_rel>> rels;
    std::vector<kb::FilterProg> none;
    KB_TRY(kb::scan_impl(ctx, pat, 1, none, false, false, &rels));
    std::unique_ptr<kb_rel> r;
    KB_TRY(kb::hash_join_impl(ctx, *left, *rels[0], nullptr, &r));
    ctx->stats.rows_out = r->n;
    *out = r.release();
    return KB_OK;
}

kb_status kb_star_join(kb_ctx* ctx, uint32_t join_slot, const kb_pattern* pats, uint32_t n_pats, const kb_filter_op* filter, uint32_t n_ops, kb_rel** out) {
    KB_ENTER(ctx);
    if (!pats || !out) return kb::fail(ctx, KB_E_INVALID, "NULL argument");
    std::unique_ptr<kb_rel> r;
    KB_TRY(kb::star_join_impl(ctx, join_slot, pats, n_pats, filter, n_ops, &r));
    ctx->stats.rows_out = r->n;
    *out = r.release();
    return KB_OK;
}

kb_status kb_bgp_execute(kb_ctx* ctx, const kb_pattern* pats, uint32_t n_pats, const kb_filter_op* filter, uint32_t n_ops,
                         const uint32_t* project, uint32_t n_project, kb_rel** out) {
    KB_ENTER(ctx);
    if (!pats || !out || n_pats == 0) return kb::fail(ctx, KB_E_INVALID, "NULL or empty BGP");
    if (n_pats > (u32)kb::MAXP) return kb::fail(ctx, KB_E_LIMIT, "more than %d patterns", kb::MAXP);
    KB_TRY(kb::validate_filter(ctx, filter, n_ops));
    std::vector<std::vector<u32>> pv(n_pats), ps(n_pats);
    std::vector<u32> all_slots;
    for (u32 k = 0; k < n_pats; k++) {
        kb::pattern_vars(pats[k], &pv[k], &ps[k]);
        for (u32 s : pv[k]) if (std::find(all_slots.begin(), all_slots.end(), s) == all_slots.end()) all_slots.push_back(s);
    }
    // a variable shared by every pattern -> star join (optimizer.rs:84-152); prefer the subject position
    int star = -1;
    for (size_t i = 0; i < pv[0].size() && star < 0; i++) {
        bool all = true;
        for (u32 k = 1; k < n_pats; k++) if (std::find(pv[k].begin(), pv[k].end(), pv[0][i]) == pv[k].end()) all = false;
        if (all) star = (int)pv[0][i];
    }
    std::unique_ptr<kb_rel> cur;
    if (star >= 0) {
        KB_TRY(kb::star_join_impl(ctx, (u32)star, pats, n_pats, filter, n_ops, &cur));
    } else {
        // left-deep natural joins in textual order (build_logical_plan, utils.rs:101-191)
        std::vector<kb::FilterProg> conj, pushdown(n_pats);
        kb::FilterProg post;
        if (!kb::split_conjuncts(filter, n_ops, &conj)) return kb::fail(ctx, KB_E_INVALID, "malformed filter program");
        for (auto& c : conj) {
            std::set<u32> fs = kb::filter_slots(c);
            for (u32 s : fs) if (std::find(all_slots.begin(), all_slots.end(), s) == all_slots.end())
                return kb::fail(ctx, KB_E_UNSUPPORTED, "filter references slot %u which no pattern binds", s);
            int target = -1;
            for (u32 k = 0; k < n_pats && target < 0; k++) {
                bool all = true;
                for (u32 s : fs) if (std::find(pv[k].begin(), pv[k].end(), s) == pv[k].end()) all = false;
                if (all) target = (int)k;
            }
            kb::FilterProg* dst = target >= 0 ? &pushdown[target] : &post;
            const bool had = !dst->ops.empty();
            dst->ops.insert(dst->ops.end(), c.ops.begin(), c.ops.end());
            if (had) { kb_filter_op a{}; a.op = KB_F_AND; dst->ops.push_back(a); }
        }
        std::vector<std::unique_ptr<kb_rel>> rels;
        KB_TRY(kb::scan_impl(ctx, pats, n_pats, pushdown, false, false, &rels));
        cur = std::move(rels[0]);
        for (u32 k = 1; k < n_pats; k++) {
            std::unique_ptr<kb_rel> j;
            KB_TRY(kb::hash_join_impl(ctx, *cur, *rels[k], (k + 1 == n_pats && !post.ops.empty()) ? &post : nullptr, &j));
            cur = std::move(j);
        }
        if (n_pats == 1 && !post.ops.empty()) { std::unique_ptr<kb_rel> f; KB_TRY(kb::filter_impl(ctx, *cur, post, &f)); cur = std::move(f); }
        cur = kb::select_cols(*cur, all_slots);
    }
    if (project) {
        std::vector<u32> v(project, project + n_project);
        cur = kb::select_cols(*cur, v);
    }
    ctx->stats.rows_out = cur->n;
    *out = cur.release();
    return KB_OK;
}

// ------------------------------------------------------------------ GROUP BY
kb_status kb_group_aggregate(kb_ctx* ctx, const kb_rel* in, const uint32_t* group_slots, uint32_t n_group, const kb_agg* aggs, uint32_t n_aggs, kb_groups** out) {
    KB_ENTER(ctx);
    if (!in || !out || (n_group && !group_slots) || (n_aggs && !aggs)) return kb::fail(ctx, KB_E_INVALID, "NULL argument");
    if (n_group == 0 || n_group > 4) return kb::fail(ctx, KB_E_LIMIT, "GROUP BY takes 1..4 variables");
    if (n_aggs > 8) return kb::fail(ctx, KB_E_LIMIT, "at most 8 aggregates");
    if (in->n >= 0xFFFFFFF0ull) return kb::fail(ctx, KB_E_LIMIT, "relation too large");
    kb::GroupParams P{};
    P.n_gcols = n_group;
    for (u32 c = 0; c < n_group; c++) {
        int ci = in->col_of(group_slots[c]);
        if (ci < 0) return kb::fail(ctx, KB_E_INVALID, "GROUP BY slot %u is not a column", group_slots[c]);
        P.gcol[c] = in->cols[ci].ptr;
    }
    P.n_aggs = n_aggs;
    for (u32 a = 0; a < n_aggs; a++) {
        P.akind[a] = aggs[a].kind;
        P.acol[a] = nullptr;
        if (aggs[a].kind > KB_AGG_AVG) return kb::fail(ctx, KB_E_INVALID, "unknown aggregate kind %u", aggs[a].kind);
        if (aggs[a].kind != KB_AGG_COUNT) {
            int ci = in->col_of(aggs[a].slot);
            if (ci < 0) return kb::fail(ctx, KB_E_INVALID, "aggregate slot %u is not a column", aggs[a].slot);
            P.acol[a] = in->cols[ci].ptr;
        }
    }
    P.n = (u32)in->n;
    P.nt = kb::numtab(ctx);
    auto g = std::make_unique<kb_groups>();
    g->keys.resize(n_group);
    g->vals.resize(n_aggs);
    if (in->n == 0) { *out = g.release(); return KB_OK; }
    u64 slots = 1u << 12;  // small first try: the table is downloaded whole; overflow -> 16x larger and rerun
    for (;;) {
        kb::GroupTable tab;
        KB_TRY(kb::group_table_create(ctx, slots, &P, &tab));
        const u32 off = kb::ctrl_alloc(ctx, 4);
        KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off, 0, 4 * sizeof(u32), ctx->st));
        P.overflow = ctx->ctrl + off;
        kb::timer_begin(ctx, kb::F_GROUP, 2);
        kb::launch_group(P, ctx->n_sms, ctx->st);
        kb::timer_end(ctx);
        KB_CUDA(ctx, cudaGetLastError());
        KB_TRY(kb::ctrl_read(ctx));
        if (ctx->h_ctrl[off] == 0) {
            KB_TRY(kb::group_table_collect(ctx, tab, n_group, aggs, n_aggs, g.get()));
            break;
        }
        if (slots >= 2 * in->n && slots >= (1u << 20)) return kb::fail(ctx, KB_E_LIMIT, "group table overflow");
        slots *= 16;
        if (slots > (1ull << 30)) slots = 1ull << 30;
    }
    *out = g.release();
    return KB_OK;
}
kb_status kb_groups_info(const kb_groups* g, uint64_t* n_groups, uint32_t* n_group_cols, uint32_t* n_aggs) {
    if (!g) return KB_E_INVALID;
    if (n_groups) *n_groups = g->counts.size();
    if (n_group_cols) *n_group_cols = (uint32_t)g->keys.size();
    if (n_aggs) *n_aggs = (uint32_t)g->vals.size();
    return KB_OK;
}
kb_status kb_groups_keys(const kb_groups* g, uint32_t col, const uint32_t** keys) {
    if (!g || !keys || col >= g->keys.size()) return KB_E_INVALID;
    *keys = g->keys[col].data();
    return KB_OK;
}
kb_status kb_groups_values(const kb_groups* g, uint32_t agg, const double** values) {
    if (!g || !values || agg >= g->vals.size()) return KB_E_INVALID;
    *values = g->vals[agg].data();
    return KB_OK;
}
kb_status kb_groups_counts(const kb_groups* g, const uint64_t** counts) {
    if (!g || !counts) return KB_E_INVALID;
    *counts = g->counts.data();
    return KB_OK;
}
void kb_groups_free(kb_groups* g) { delete g; }

// ---- partial GROUP BY results across ranks: pack on every rank, gather, merge on the device
namespace {
struct GroupsHeader {
    uint64_t magic, n_groups;
    uint32_t n_group_cols, n_aggs;
    uint32_t kinds[8];
};
static_assert(sizeof(GroupsHeader) == 56, "layout");
constexpr uint64_t GROUPS_MAGIC = 0x4B4247524F555031ull;  // "KBGROUP1"
}  // namespace

kb_status kb_groups_pack(const kb_groups* g, void* dst, uint64_t capacity_bytes, uint64_t* bytes) {
    if (!g || !bytes) return KB_E_INVALID;
    const uint64_t n = g->counts.size();
    const uint64_t need = sizeof(GroupsHeader) + n * sizeof(kb::GroupRecord);
    *bytes = need;
    if (!dst) return KB_OK;  // size query
    if (capacity_bytes < need) return KB_E_LIMIT;
    if (g->keys.size() > 4 || g->vals.size() > 8 || g->raw.size() != g->vals.size() || g->kinds.size() != g->vals.size()) return KB_E_INVALID;
    GroupsHeader h{};
    h.magic = GROUPS_MAGIC;
    h.n_groups = n;
    h.n_group_cols = (uint32_t)g->keys.size();
    h.n_aggs = (uint32_t)g->vals.size();
    for (size_t a = 0; a < g->kinds.size(); a++) h.kinds[a] = g->kinds[a];
    memcpy(dst, &h, sizeof h);
    auto* rec = reinterpret_cast<kb::GroupRecord*>(static_cast<char*>(dst) + sizeof h);
    for (uint64_t i = 0; i < n; i++) {
        kb::GroupRecord r{};
        for (size_t c = 0; c < g->keys.size(); c++) r.keys[c] = g->keys[c][i];
        r.count = g->counts[i];
        for (size_t a = 0; a < g->raw.size(); a++) r.raw[a] = g->raw[a][i];
        memcpy(rec + i, &r, sizeof r);
    }
    return KB_OK;
}

kb_status kb_groups_merge(kb_ctx* ctx, const void* const* parts, const uint64_t* part_bytes, uint32_t n_parts, kb_groups** out) {
    KB_ENTER(ctx);
    if (!parts || !part_bytes || !out || n_parts == 0) return kb::fail(ctx, KB_E_INVALID, "NULL or empty argument");
    GroupsHeader h0{};
    uint64_t total = 0;
    for (u32 i = 0; i < n_parts; i++) {
        if (!parts[i] || part_bytes[i] < sizeof(GroupsHeader)) return kb::fail(ctx, KB_E_INVALID, "partial %u is truncated", i);
        GroupsHeader h;
        memcpy(&h, parts[i], sizeof h);
        if (h.magic != GROUPS_MAGIC) return kb::fail(ctx, KB_E_INVALID, "partial %u is not a kb_groups_pack buffer", i);
        if (part_bytes[i] < sizeof h + h.n_groups * sizeof(kb::GroupRecord)) return kb::fail(ctx, KB_E_INVALID, "partial %u is truncated", i);
        if (i == 0) h0 = h;
        else if (h.n_group_cols != h0.n_group_cols || h.n_aggs != h0.n_aggs || memcmp(h.kinds, h0.kinds, sizeof h.kinds) != 0)
            return kb::fail(ctx, KB_E_INVALID, "partial %u was produced by a different GROUP BY (columns / aggregates differ)", i);
        total += h.n_groups;
    }
    if (h0.n_group_cols == 0 || h0.n_group_cols > 4 || h0.n_aggs > 8) return kb::fail(ctx, KB_E_INVALID, "bad header");
    if (total >= (1ull << 30)) return kb::fail(ctx, KB_E_LIMIT, "too many partial groups");
    auto g = std::make_unique<kb_groups>();
    g->keys.resize(h0.n_group_cols);
    g->vals.resize(h0.n_aggs);
    std::vector<kb_agg> aggs(h0.n_aggs);
    for (u32 a = 0; a < h0.n_aggs; a++) { aggs[a].kind = h0.kinds[a]; aggs[a].slot = 0; }
    if (total == 0) {
        for (u32 a = 0; a < h0.n_aggs; a++) g->kinds.push_back(h0.kinds[a]);
        g->raw.resize(h0.n_aggs);
        *out = g.release();
        return KB_OK;
    }
    // partial records -> device, merged by find-or-insert into a group table sized for the worst case (every partial group distinct)
    kb::Buf recs;
    KB_TRY(kb::alloc_buf(ctx, total * sizeof(kb::GroupRecord), &recs));
    uint64_t at = 0;
    for (u32 i = 0; i < n_parts; i++) {
        GroupsHeader h;
        memcpy(&h, parts[i], sizeof h);
        if (h.n_groups == 0) continue;
        KB_CUDA(ctx, cudaMemcpyAsync(static_cast<kb::GroupRecord*>(recs->p) + at, static_cast<const char*>(parts[i]) + sizeof h, h.n_groups * sizeof(kb::GroupRecord),
                                     cudaMemcpyHostToDevice, ctx->st));
        at += h.n_groups;
    }
    ctx->stats.h2d_bytes += total * sizeof(kb::GroupRecord);
    kb::GroupParams P{};
    P.n_gcols = h0.n_group_cols;
    P.n_aggs = h0.n_aggs;
    for (u32 a = 0; a < h0.n_aggs; a++) P.akind[a] = h0.kinds[a];
    P.nt = kb::numtab(ctx);
    kb::GroupTable tab;
    KB_TRY(kb::group_table_create(ctx, kb::pow2_at_least(std::max<u64>(2 * total, 1024)), &P, &tab));
    const u32 off = kb::ctrl_alloc(ctx, 4);
    KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off, 0, 4 * sizeof(u32), ctx->st));
    P.overflow = ctx->ctrl + off;
    kb::timer_begin(ctx, kb::F_GROUP);
    kb::launch_group_merge(P, static_cast<const kb::GroupRecord*>(recs->p), (u32)total, ctx->n_sms, ctx->st);
    kb::timer_end(ctx);
    KB_CUDA(ctx, cudaGetLastError());
    KB_TRY(kb::ctrl_read(ctx));
    if (ctx->h_ctrl[off]) return kb::fail(ctx, KB_E_LIMIT, "group table overflow while merging partials");
    KB_TRY(kb::group_table_collect(ctx, tab, h0.n_group_cols, aggs.data(), h0.n_aggs, g.get()));
    *out = g.release();
    return KB_OK;
}

kb_status kb_star_join_aggregate(kb_ctx* ctx, uint32_t join_slot, const kb_pattern* pats, uint32_t n_pats, const kb_filter_op* filter, uint32_t n_ops,
                                 const uint32_t* group_slots, uint32_t n_group, const kb_agg* aggs, uint32_t n_aggs, kb_groups** out, uint64_t* n_rows) {
    KB_ENTER(ctx);
    if (!pats || !out || (n_group && !group_slots) || (n_aggs && !aggs)) return kb::fail(ctx, KB_E_INVALID, "NULL argument");
    if (n_group == 0 || n_group > 4) return kb::fail(ctx, KB_E_LIMIT, "GROUP BY takes 1..4 variables");
    if (n_aggs > 8) return kb::fail(ctx, KB_E_LIMIT, "at most 8 aggregates");
    for (u32 a = 0; a < n_aggs; a++) if (aggs[a].kind > KB_AGG_AVG) return kb::fail(ctx, KB_E_INVALID, "unknown aggregate kind %u", aggs[a].kind);
    kb::AggSpec spec;
    const bool fusable = n_group == 1 && n_aggs <= 1;
    if (fusable) {
        spec.group_slot = group_slots[0];
        spec.has_agg = n_aggs == 1;
        if (n_aggs) { spec.kind = aggs[0].kind; spec.agg_slot = aggs[0].slot; }
    }
    std::unique_ptr<kb_rel> rel;
    KB_TRY(kb::star_join_impl2(ctx, join_slot, pats, n_pats, filter, n_ops, true, &rel, fusable ? &spec : nullptr));
    if (spec.applied) {
        if (n_rows) *n_rows = spec.n_rows;
        *out = spec.groups.release();
        return KB_OK;
    }
    if (n_rows) *n_rows = rel->n;
    return kb_group_aggregate(ctx, rel.get(), group_slots, n_group, aggs, n_aggs, out);
}

// ------------------------------------------------------------------ multi-GPU helpers
uint32_t kb_shard_of(uint32_t key, uint32_t n_shards) { return n_shards ? kb::shard_of(key, n_shards) : 0; }

kb_status kb_set_sharding(kb_ctx* ctx, uint32_t rank, uint32_t world) {
    if (!ctx) return KB_E_INVALID;
    if (world == 0 || rank >= world) return kb::fail(ctx, KB_E_INVALID, "rank %u / world %u", rank, world);
    ctx->shard_rank = rank;
    ctx->shard_world = world;
    return KB_OK;
}

kb_status kb_partition_counts(kb_ctx* ctx, const kb_rel* in, uint32_t key_slot, uint32_t n_parts, uint64_t* counts) {
    KB_ENTER(ctx);
    if (!in || !counts) return kb::fail(ctx, KB_E_INVALID, "NULL argument");
    if (n_parts == 0 || n_parts > 64) return kb::fail(ctx, KB_E_LIMIT, "1..64 partitions");
    const int kc = in->col_of(key_slot);
    if (kc < 0) return kb::fail(ctx, KB_E_INVALID, "partition key slot %u is not a column", key_slot);
    if (in->n >= 0xFFFFFFF0ull) return kb::fail(ctx, KB_E_LIMIT, "relation too large");
    const u32 off = kb::ctrl_alloc(ctx, 64);
    KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off, 0, 64 * sizeof(u32), ctx->st));
    kb::timer_begin(ctx, kb::F_OTHER);
    kb::launch_part_count(in->cols[kc].ptr, (u32)in->n, n_parts, ctx->ctrl + off, ctx->n_sms, ctx->st);
    kb::timer_end(ctx);
    KB_CUDA(ctx, cudaGetLastError());
    KB_TRY(kb::ctrl_read(ctx));
    for (u32 i = 0; i < n_parts; i++) counts[i] = ctx->h_ctrl[off + i];
    return KB_OK;
}

static kb_status shuffle_common(kb_ctx* ctx, const kb_rel* in, uint32_t key_slot, uint32_t n_parts, uint32_t* const* peer_cols, const uint64_t* base,
                                uint32_t* const* peer_cursors, uint64_t capacity_rows) {
    if (!in || !peer_cols || (!base && !peer_cursors)) return kb::fail(ctx, KB_E_INVALID, "NULL argument");
    if (n_parts == 0 || n_parts > 64) return kb::fail(ctx, KB_E_LIMIT, "1..64 partitions");
    const int kc = in->col_of(key_slot);
    if (kc < 0) return kb::fail(ctx, KB_E_INVALID, "partition key slot %u is not a column", key_slot);
    if (in->n >= 0xFFFFFFF0ull || capacity_rows >= 0xFFFFFFF0ull) return kb::fail(ctx, KB_E_LIMIT, "relation too large");
    const u32 n_cols = (u32)in->cols.size();
    for (u32 i = 0; i < n_parts * n_cols; i++) if (!peer_cols[i]) return kb::fail(ctx, KB_E_INVALID, "peer column %u is NULL", i);
    if (peer_cursors) for (u32 i = 0; i < n_parts; i++) if (!peer_cursors[i]) return kb::fail(ctx, KB_E_INVALID, "peer cursor %u is NULL", i);
    // device-side tables: peer column addresses | cursor addresses | bases | local cursors | overflow flag, ticket
    kb::Buf tab;
    const size_t col_bytes = (size_t)n_parts * n_cols * sizeof(u32*);
    const size_t cur_bytes = 64 * sizeof(u32*);
    KB_TRY(kb::alloc_buf(ctx, col_bytes + cur_bytes + 3 * 64 * sizeof(u32), &tab));
    char* tb = static_cast<char*>(tab->p);
    u32* d_base = reinterpret_cast<u32*>(tb + col_bytes + cur_bytes);
    u32* d_cursors = d_base + 64;
    u32* d_flags = d_cursors + 64;  // [0] overflow, [1] tile ticket
    std::vector<u32> hbase(64, 0);
    std::vector<u32*> hcur(64, nullptr);
    for (u32 i = 0; i < n_parts; i++) {
        if (base) {
            if (base[i] > capacity_rows) return kb::fail(ctx, KB_E_INVALID, "base[%u] beyond the receive capacity", i);
            hbase[i] = (u32)base[i];
        }
        hcur[i] = peer_cursors ? peer_cursors[i] : d_cursors + i;
    }
    KB_CUDA(ctx, cudaMemcpyAsync(tb, peer_cols, col_bytes, cudaMemcpyHostToDevice, ctx->st));
    KB_CUDA(ctx, cudaMemcpyAsync(tb + col_bytes, hcur.data(), cur_bytes, cudaMemcpyHostToDevice, ctx->st));
    KB_CUDA(ctx, cudaMemcpyAsync(d_base, hbase.data(), 64 * sizeof(u32), cudaMemcpyHostToDevice, ctx->st));
    KB_CUDA(ctx, cudaMemsetAsync(d_cursors, 0, 2 * 64 * sizeof(u32), ctx->st));
    kb::ShuffleParams P{};
    P.key = in->cols[kc].ptr;
    P.n = (u32)in->n;
    P.n_parts = n_parts;
    P.n_cols = n_cols;
    for (u32 c = 0; c < n_cols; c++) P.in[c] = in->cols[c].ptr;
    P.peer_cols = reinterpret_cast<u32* const*>(tb);
    P.cursor_ptrs = reinterpret_cast<u32* const*>(tb + col_bytes);
    P.base = d_base;
    P.overflow = d_flags;
    P.ticket = d_flags + 1;
    P.capacity = (u32)capacity_rows;
    kb::timer_begin(ctx, kb::F_OTHER);
    kb::launch_shuffle_scatter(P, ctx->n_sms, ctx->st);
    kb::timer_end(ctx);
    KB_CUDA(ctx, cudaGetLastError());
    u32 ovf = 0;
    KB_CUDA(ctx, cudaMemcpyAsync(&ovf, P.overflow, sizeof ovf, cudaMemcpyDeviceToHost, ctx->st));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
    kb::timers_flush(ctx);
    if (ovf) return kb::fail(ctx, KB_E_LIMIT, "a receive buffer is smaller than the rows sent to it (capacity %llu rows)", (unsigned long long)capacity_rows);
    return KB_OK;
}

kb_status kb_shuffle_scatter(kb_ctx* ctx, const kb_rel* in, uint32_t key_slot, uint32_t n_parts, uint32_t* const* peer_cols, const uint64_t* base,
                             uint64_t capacity_rows) {
    KB_ENTER(ctx);
    if (!base) return kb::fail(ctx, KB_E_INVALID, "NULL argument");
    return shuffle_common(ctx, in, key_slot, n_parts, peer_cols, base, nullptr, capacity_rows);
}

kb_status kb_shuffle_push(kb_ctx* ctx, const kb_rel* in, uint32_t key_slot, uint32_t n_parts, uint32_t* const* peer_cols, uint32_t* const* peer_cursors,
                          uint64_t capacity_rows) {
    KB_ENTER(ctx);
    if (!peer_cursors) return kb::fail(ctx, KB_E_INVALID, "NULL argument");
    return shuffle_common(ctx, in, key_slot, n_parts, peer_cols, nullptr, peer_cursors, capacity_rows);
}

kb_status kb_rel_wrap_device(kb_ctx* ctx, const uint32_t* slots, uint32_t n_cols, uint32_t* const* d_cols, uint64_t n_rows, kb_rel** out) {
    if (!ctx) return KB_E_INVALID;
    if (!out || (n_cols && (!slots || !d_cols))) return kb::fail(ctx, KB_E_INVALID, "NULL argument");
    if (n_cols > KB_MAX_COLS) return kb::fail(ctx, KB_E_LIMIT, "more than %d columns", KB_MAX_COLS);
    if (n_rows >= 0xFFFFFFF0ull) return kb::fail(ctx, KB_E_LIMIT, "relation too large");
    auto r = std::make_unique<kb_rel>();
    r->n = n_rows;
    for (u32 c = 0; c < n_cols; c++) {
        for (u32 d = 0; d < c; d++) if (slots[d] == slots[c]) return kb::fail(ctx, KB_E_INVALID, "duplicate slot %u", slots[c]);
        if (!d_cols[c] || (reinterpret_cast<uintptr_t>(d_cols[c]) & 15u)) return kb::fail(ctx, KB_E_INVALID, "column %u must be a 16-byte aligned device pointer", c);
        kb::Col col;  // no owner: the caller keeps the memory alive (and readable to the next 256-byte boundary) while the relation is used
        col.ptr = d_cols[c];
        r->slots.push_back(slots[c]);
        r->cols.push_back(col);
    }
    *out = r.release();
    return KB_OK;
}

kb_status kb_partition(kb_ctx* ctx, const kb_rel* in, uint32_t key_slot, uint32_t n_parts, kb_rel** out, uint64_t* part_offsets) {
    KB_ENTER(ctx);
    if (!in || !out || !part_offsets) return kb::fail(ctx, KB_E_INVALID, "NULL argument");
    if (n_parts == 0 || n_parts > 64) return kb::fail(ctx, KB_E_LIMIT, "1..64 partitions");
    const int kc = in->col_of(key_slot);
    if (kc < 0) return kb::fail(ctx, KB_E_INVALID, "partition key slot %u is not a column", key_slot);
    if (in->n >= 0xFFFFFFF0ull) return kb::fail(ctx, KB_E_LIMIT, "relation too large");
    auto r = std::make_unique<kb_rel>();
    r->slots = in->slots;
    r->n = in->n;
    const u32 off = kb::ctrl_alloc(ctx, 128);
    KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off, 0, 128 * sizeof(u32), ctx->st));
    kb::timer_begin(ctx, kb::F_OTHER);
    kb::launch_part_count(in->cols[kc].ptr, (u32)in->n, n_parts, ctx->ctrl + off, ctx->n_sms, ctx->st);
    kb::timer_end(ctx);
    KB_TRY(kb::ctrl_read(ctx));
    u64 run = 0;
    std::vector<u32> cursors(n_parts);
    for (u32 i = 0; i < n_parts; i++) { part_offsets[i] = run; cursors[i] = (u32)run; run += ctx->h_ctrl[off + i]; }
    part_offsets[n_parts] = run;
    KB_CUDA(ctx, cudaMemcpyAsync(ctx->ctrl + off + 64, cursors.data(), n_parts * sizeof(u32), cudaMemcpyHostToDevice, ctx->st));
    std::vector<const u32*> ic;
    std::vector<u32*> oc;
    for (auto& c : in->cols) {
        kb::Col col;
        KB_TRY(kb::alloc_col(ctx, in->n, &col));
        r->cols.push_back(col);
        ic.push_back(c.ptr);
        oc.push_back(col.ptr);
    }
    kb::timer_begin(ctx, kb::F_OTHER);
    kb::launch_part_scatter(in->cols[kc].ptr, (u32)in->n, n_parts, ctx->ctrl + off + 64, ic.data(), oc.data(), (u32)ic.size(), ctx->n_sms, ctx->st);
    kb::timer_end(ctx);
    KB_CUDA(ctx, cudaGetLastError());
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
    kb::timers_flush(ctx);
    *out = r.release();
    return KB_OK;
}

// ------------------------------------------------------------------ one-shot host-buffer star join
static kb_status star_join_host_common(kb_ctx* ctx, const uint32_t* s, const uint32_t* p, const uint32_t* o, uint64_t n, uint32_t join_slot,
                                       const kb_pattern* pats, uint32_t n_pats, const kb_filter_op* filter, uint32_t n_ops, uint32_t* n_cols,
                                       uint32_t* slots, uint32_t** cols, uint64_t* n_rows, bool caller_bufs, uint64_t caller_cap) {
    if (!pats || !n_cols || !slots || !cols || !n_rows) return kb::fail(ctx, KB_E_INVALID, "NULL argument");
    if (n && (!s || !p || !o)) return kb::fail(ctx, KB_E_INVALID, "NULL column pointer");
    if (n >= 0xFFFFFFF0ull) return kb::fail(ctx, KB_E_LIMIT, "too many triples");
    // everything that can be rejected without looking at the data is rejected BEFORE the upload starts: an early return must not
    // leave copies in flight that still read the caller's (borrowed) buffers
    if (n_pats == 0 || n_pats > (u32)kb::MAXP) return kb::fail(ctx, KB_E_LIMIT, "a star join takes 1..%d patterns (got %u)", kb::MAXP, n_pats);
    KB_TRY(kb::validate_filter(ctx, filter, n_ops));
    for (u32 k = 0; k < n_pats; k++) {
        KB_TRY(kb::check_pattern(ctx, pats[k]));
        std::vector<u32> vs, src;
        kb::pattern_vars(pats[k], &vs, &src);
        if (std::find(vs.begin(), vs.end(), join_slot) == vs.end())
            return kb::fail(ctx, KB_E_INVALID, "star join: pattern %u does not contain the join variable (slot %u)", k, join_slot);
    }
    // upload in chunks on the copy stream; each chunk is a store segment the scan starts on as soon as its copy has landed
    ctx->segs.clear();
    ctx->n_triples = 0;
    ctx->store_version++;
    ctx->multi_valued.clear();
    ctx->single_valued.clear();
    ctx->index.clear();  // the predicate-partitioned index describes the previous store version
    kb::Col cs, cp, co;
    KB_TRY(kb::alloc_col(ctx, n, &cs));
    KB_TRY(kb::alloc_col(ctx, n, &cp));
    KB_TRY(kb::alloc_col(ctx, n, &co));
    const u32 soff = kb::ctrl_alloc(ctx, 8);
    KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + soff, 0xFF, 4 * sizeof(u32), ctx->st));
    KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + soff + 4, 0, 4 * sizeof(u32), ctx->st));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));  // allocations are stream-ordered on st; the copy stream must see them
    const u64 chunk = (u64)kb::SCAN_TILE * 4096;   // 8 Mi triples = 32 MiB per column
    std::vector<cudaEvent_t> evs;
    // whatever happens from here on, the caller's buffers are not read after this function returns and a failed call leaves an
    // EMPTY store (never segments whose copies did not complete)
    auto finish = [&](kb_status rc) -> kb_status {
        cudaStreamSynchronize(ctx->st_copy);
        for (auto& sg : ctx->segs) sg.ready = nullptr;
        ctx->upload_stats_off = -1;
        for (auto ev : evs) cudaEventDestroy(ev);
        if (rc != KB_OK) {
            cudaStreamSynchronize(ctx->st);
            ctx->segs.clear();
            ctx->n_triples = 0;
        }
        return rc;
    };
#define KB_HOST_CUDA(call)                                                                                                    \
    do {                                                                                                                      \
        cudaError_t _e = (call);                                                                                              \
        if (_e != cudaSuccess) return finish(kb::fail(ctx, KB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__)); \
    } while (0)
    for (u64 b = 0; b < n; b += chunk) {
        const u64 m = std::min(chunk, n - b);
        KB_HOST_CUDA(cudaMemcpyAsync(cs.ptr + b, s + b, m * sizeof(u32), cudaMemcpyHostToDevice, ctx->st_copy));
        KB_HOST_CUDA(cudaMemcpyAsync(cp.ptr + b, p + b, m * sizeof(u32), cudaMemcpyHostToDevice, ctx->st_copy));
        KB_HOST_CUDA(cudaMemcpyAsync(co.ptr + b, o + b, m * sizeof(u32), cudaMemcpyHostToDevice, ctx->st_copy));
        {  // column ranges of this chunk, on the copy stream right behind its copies (the direct tables need the key range)
            const u32* cc[3] = {cs.ptr + b, cp.ptr + b, co.ptr + b};
            for (int c = 0; c < 3; c++) kb::launch_col_minmax(cc[c], (u32)m, ctx->ctrl + soff + c, ctx->ctrl + soff + 4 + c, ctx->n_sms, ctx->st_copy);
            ctx->stats.kernel_launches += 3;
        }
        cudaEvent_t ev;
        KB_HOST_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        evs.push_back(ev);
        KB_HOST_CUDA(cudaEventRecord(ev, ctx->st_copy));
        kb::Segment sg;
        sg.tag = 0; sg.n = m;
        sg.ready = ev;
        sg.s.buf = cs.buf; sg.s.ptr = cs.ptr + b;
        sg.p.buf = cp.buf; sg.p.ptr = cp.ptr + b;
        sg.o.buf = co.buf; sg.o.ptr = co.ptr + b;
        ctx->segs.push_back(sg);
        ctx->n_triples += m;
    }
    ctx->stats.h2d_bytes += 3 * n * sizeof(u32);
    ctx->upload_stats_off = (int)soff;
    // scan_impl makes st wait on each segment's `ready` event right before that segment's scan kernel: copy i+1 overlaps scan i
    std::unique_ptr<kb_rel> r;
    kb_status rc = kb::star_join_impl(ctx, join_slot, pats, n_pats, filter, n_ops, &r);
    if (rc != KB_OK) return finish(rc);
    *n_cols = (u32)r->slots.size();
    *n_rows = r->n;
    for (size_t c = 0; c < r->slots.size(); c++) slots[c] = r->slots[c];
    if (caller_bufs && caller_cap < r->n)
        return finish(kb::fail(ctx, KB_E_LIMIT, "caller buffers hold %llu rows, result has %llu", (unsigned long long)caller_cap, (unsigned long long)r->n));
    if (!caller_bufs) for (size_t c = 0; c < KB_MAX_COLS; c++) cols[c] = nullptr;
    for (size_t c = 0; c < r->slots.size(); c++) {
        if (!caller_bufs) {
            cols[c] = (uint32_t*)malloc(std::max<size_t>(r->n * sizeof(u32), 4));
            if (!cols[c]) {
                for (size_t d = 0; d < c; d++) { free(cols[d]); cols[d] = nullptr; }
                return finish(kb::fail(ctx, KB_E_OOM, "malloc failed"));
            }
        } else if (!cols[c]) return finish(kb::fail(ctx, KB_E_INVALID, "caller-provided column buffer %zu is NULL", c));
        if (r->n) KB_HOST_CUDA(cudaMemcpyAsync(cols[c], r->cols[c].ptr, r->n * sizeof(u32), cudaMemcpyDeviceToHost, ctx->st));
    }
    KB_HOST_CUDA(cudaStreamSynchronize(ctx->st));
#undef KB_HOST_CUDA
    ctx->stats.d2h_bytes += r->slots.size() * r->n * sizeof(u32);
    ctx->stats.rows_out = r->n;
    return finish(KB_OK);
}

kb_status kb_star_join_host(kb_ctx* ctx, const uint32_t* s, const uint32_t* p, const uint32_t* o, uint64_t n, uint32_t join_slot,
                            const kb_pattern* pats, uint32_t n_pats, const kb_filter_op* filter, uint32_t n_ops, uint32_t* n_cols,
                            uint32_t* slots, uint32_t** cols, uint64_t* n_rows) {
    KB_ENTER(ctx);
    return star_join_host_common(ctx, s, p, o, n, join_slot, pats, n_pats, filter, n_ops, n_cols, slots, cols, n_rows, false, 0);
}

kb_status kb_star_join_host_into(kb_ctx* ctx, const uint32_t* s, const uint32_t* p, const uint32_t* o, uint64_t n, uint32_t join_slot,
                                 const kb_pattern* pats, uint32_t n_pats, const kb_filter_op* filter, uint32_t n_ops, uint32_t* n_cols,
                                 uint32_t* slots, uint32_t* const* cols, uint64_t capacity_rows, uint64_t* n_rows) {
    KB_ENTER(ctx);
    return star_join_host_common(ctx, s, p, o, n, join_slot, pats, n_pats, filter, n_ops, n_cols, slots, const_cast<uint32_t**>(cols), n_rows, true,
                                 capacity_rows);
}

}  // extern "C"
