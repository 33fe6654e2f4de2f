// kb_datalog.cu — semi-naive / naive Datalog materialisation on the device.
// Replaces Reasoner::infer_with_strategy (datalog/src/reasoning/materialisation/infer_generic.rs:27-53) driving
// SemiNaiveStrategy (semi_naive.rs:17-85) or NaiveStrategy (my_naive.rs:18-69), whose join is
// perform_hash_join_for_rules (shared/src/join_algorithm.rs:499-677).
//
// Layout: one (s,o) column pair per predicate that occurs in a rule, append-only, exactly like the reference's
// `all_facts: Vec<Triple>` with its delta suffix (semi_naive.rs:57-59) but partitioned by predicate once instead of
// re-filtered by predicate for every premise of every rule of every round (join_algorithm.rs:528-534).
// known_facts (infer_generic.rs:30) is a device hash set of 96-bit (s,p,o) keys; a derived fact is new iff its insert wins.
// Round structure is the reference's: every rule of a round sees the same snapshot; new facts become visible (and become the
// next delta) only after the round (infer_generic.rs:42-48) — so round counts and per-round fact counts match the oracle.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "kb_internal.hpp"

using namespace kb;

namespace {

struct PredRel {
    u32 pred = 0;
    Col s, o;
    u64 n = 0, cap = 0;
    u64 base_n = 0;       // facts present before inference (with a seed: after the seed's new facts were accepted)
    u64 seed_from = 0;    // kb_datalog_fixpoint_seed: rows [seed_from, base_n) are the seed facts that were new
    u64 delta_start = 0;  // facts [delta_start, snapshot) were added by the previous round
    u64 snapshot = 0;
    bool is_head = false;
    Buf set;              // head predicates only: known (s,o) pairs of this predicate, 64-bit keys
    u32 set_slots = 0;
    u64 set_count = 0;
};

struct Premise {
    int s_var, o_var;  // variable slots (real or synthetic, quirk Q6)
    bool pred_const;
    u32 pred;
    bool s_is_const = false, o_is_const = false;  // enforced only by KB_SEMI_NAIVE_PARALLEL (matches_rule_pattern, rules.rs:9-72)
    u32 s_const = 0, o_const = 0;
};
struct RulePlan {
    std::vector<Premise> prem;
    u32 n_vars = 0;
};

struct Pending {  // facts derived in the current round, not yet visible to the joins
    u32 pred;
    Col s, o;
    u64 count;
};

// KOLIBRIE_TRACE=1: synchronise after every phase of the fixpoint and print where the wall-clock time went (stderr)
struct Trace {
    bool on = getenv("KOLIBRIE_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void mark(kb_ctx* ctx, const char* what, unsigned long long rows = 0) {
        if (!on) return;
        cudaStreamSynchronize(ctx->st);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[kb trace] %-28s %9.3f ms  rows=%llu\n", what, std::chrono::duration<double, std::milli>(now - t).count(), rows);
        t = std::chrono::steady_clock::now();
    }
};

struct Fix {
    kb_ctx* ctx;
    std::map<u32, PredRel> rels;
    std::vector<Pending> pend;
    Buf buckets;  // candidate buckets of the radix-partitioned dedup, kept across the derive calls of one fixpoint (a 3 GB allocation
                  // from the stream-ordered pool can cost milliseconds when the pool has to remap)
};

// What a seeded call leaves in the context for the next one: the store split by predicate and a known-fact set per rule predicate. While
// the store is the one that call left (store_version) and the rules are the same, the next seed costs what its delta costs — no scan
// of the store, no set rebuilt. KOLIBRIE_FIX_STATE=0: every seeded call starts from the store (A/B switch).
struct FixState {
    std::string sig;  // the rules, byte for byte
    u64 version = 0;  // ctx->store_version the state belongs to
    std::map<u32, PredRel> rels;
};
std::string rules_signature(const kb_rule* rules, u32 n_rules) {
    std::string s;
    auto put = [&](const void* p, size_t n) { s.append(static_cast<const char*>(p), n); };
    auto pat = [&](const kb_pattern& x) {
        const u32 w[6] = {x.s.is_var, x.s.value, x.p.is_var, x.p.value, x.o.is_var, x.o.value};
        put(w, sizeof w);
    };
    for (u32 r = 0; r < n_rules; r++) {
        const kb_rule& k = rules[r];
        const u32 h[3] = {k.n_premise, k.n_conclusion, k.n_filters};
        put(h, sizeof h);
        for (u32 i = 0; i < k.n_premise; i++) pat(k.premise[i]);
        for (u32 i = 0; i < k.n_conclusion; i++) pat(k.conclusion[i]);
        for (u32 i = 0; i < k.n_filters; i++) {
            const kb_rule_filter& f = k.filters[i];
            const u32 w[4] = {f.lhs_slot, f.cmp, f.rhs_is_var, f.rhs_slot};
            put(w, sizeof w);
            put(&f.rhs_value, sizeof f.rhs_value);
        }
    }
    return s;
}

kb_status grow_rel(kb_ctx* ctx, PredRel& r, u64 need) {
    if (need <= r.cap) return KB_OK;
    u64 cap = std::max<u64>(need, r.cap * 2);
    cap = std::max<u64>(cap, 1024);
    Col ns, no;
    KB_TRY(alloc_col(ctx, cap, &ns));
    KB_TRY(alloc_col(ctx, cap, &no));
    if (r.n) {
        KB_CUDA(ctx, cudaMemcpyAsync(ns.ptr, r.s.ptr, r.n * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->st));
        KB_CUDA(ctx, cudaMemcpyAsync(no.ptr, r.o.ptr, r.n * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->st));
    }
    r.s = ns; r.o = no; r.cap = cap;
    return KB_OK;
}

// (re)build the known-fact set of predicate `r` so that it can take the facts the next launch is EXPECTED to add at load <= 0.5.
// `extra` counts candidates, most of which are usually known already (cfg4: 1.25e9 candidates for 2.9e8 new facts), so the
// expectation is capped at half the set's size (and at least 2^20 facts); derive_kernel enforces the real bound (budget) and the launch is repeated on a
// larger table when the cap was too optimistic. A rebuild re-inserts every known fact, so every (re)build allocates twice the need.
kb_status ensure_set(Fix& fx, PredRel& r, u64 extra, bool must_grow = false, bool exact_expectation = false) {
    kb_ctx* ctx = fx.ctx;
    const u64 expect = exact_expectation ? extra : std::min<u64>(extra, std::max<u64>(r.set_count / 2, 1u << 20));
    const u64 need = exact_expectation ? r.set_count + expect : (r.set_count + expect) * 2;  // (a known final size: load <= 1/2 is enough)
    if (r.set && need <= r.set_slots && !must_grow) return KB_OK;
    static const u64 slack = getenv("KOLIBRIE_SET_SLACK") ? std::max(1, atoi(getenv("KOLIBRIE_SET_SLACK"))) : 2;
    static const u64 tight = getenv("KOLIBRIE_SET_TIGHT") ? (u64)atoi(getenv("KOLIBRIE_SET_TIGHT")) : 0;  // experiment: known final size at load <= 2/3
    u64 slots = 1024;
    if (exact_expectation && tight) { while (slots * 2 < need * 3) slots <<= 1; }
    else while (slots < need * slack) slots <<= 1;
    if (must_grow) slots = std::max<u64>(slots, (u64)r.set_slots * 2);
    if (slots > (1ull << 31)) slots = 1ull << 31;
    if (slots < need || (must_grow && r.set && slots <= r.set_slots))
        return fail(ctx, KB_E_LIMIT, "known-fact set of predicate %u would need more than 2^31 slots", r.pred);
    Buf nb;
    KB_TRY(alloc_buf(ctx, slots * sizeof(u64), &nb));
    KB_CUDA(ctx, cudaMemsetAsync(nb->p, 0xFF, slots * sizeof(u64), ctx->st));
    const u32 off = ctrl_alloc(ctx, 4);
    KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + off, 0, 4 * sizeof(u32), ctx->st));
    u64 count = 0;
    timer_begin(ctx, F_BUILD);
    launch_set64_insert(static_cast<u64*>(nb->p), (u32)slots, r.s.ptr, r.o.ptr, (u32)r.n, ctx->ctrl + off, ctx->n_sms, ctx->st);
    count += r.n;
    for (auto& pd : fx.pend) {  // a rebuild in the middle of a round must keep this round's facts
        if (pd.pred != r.pred) continue;
        launch_set64_insert(static_cast<u64*>(nb->p), (u32)slots, pd.s.ptr, pd.o.ptr, (u32)pd.count, ctx->ctrl + off, ctx->n_sms, ctx->st);
        count += pd.count;
    }
    timer_end(ctx);
    KB_CUDA(ctx, cudaGetLastError());
    r.set = nb;
    r.set_slots = (u32)slots;
    r.set_count = count;
    return KB_OK;
}

// a relation view over rows [start, start+n) of a predicate's columns; the probe kernels TMA-load their inputs, which needs
// 16-byte aligned column starts, so a misaligned suffix (the delta) is copied
kb_status make_view(kb_ctx* ctx, const PredRel& r, u64 start, u64 n, u32 s_slot, u32 o_slot, std::unique_ptr<kb_rel>* out) {
    auto v = std::make_unique<kb_rel>();
    v->slots = {s_slot, o_slot};
    v->n = n;
    const Col* src[2] = {&r.s, &r.o};
    for (int c = 0; c < 2; c++) {
        Col col;
        if ((start & 3ull) == 0) {
            col.buf = src[c]->buf;
            col.ptr = src[c]->ptr + start;
        } else {
            KB_TRY(alloc_col(ctx, n, &col));
            if (n) KB_CUDA(ctx, cudaMemcpyAsync(col.ptr, src[c]->ptr + start, n * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->st));
        }
        v->cols.push_back(col);
    }
    *out = std::move(v);
    return KB_OK;
}

// KB_SEMI_NAIVE_PARALLEL: keep only the rows whose subject / object equal the premise's constants (the synthetic columns then
// carry a single value and simply ride along)
kb_status enforce_constants(kb_ctx* ctx, const Premise& pr, std::unique_ptr<kb_rel>* v) {
    if (!pr.s_is_const && !pr.o_is_const) return KB_OK;
    FilterProg f;
    auto eq = [&](u32 slot, u32 id) { kb_filter_op op{}; op.op = KB_F_EQ_ID; op.slot = slot; op.id = id; f.ops.push_back(op); };
    if (pr.s_is_const) eq((u32)pr.s_var, pr.s_const);
    if (pr.o_is_const) eq((u32)pr.o_var, pr.o_const);
    if (pr.s_is_const && pr.o_is_const) { kb_filter_op a{}; a.op = KB_F_AND; f.ops.push_back(a); }
    std::unique_ptr<kb_rel> out;
    KB_TRY(filter_impl(ctx, **v, f, &out));
    *v = std::move(out);
    return KB_OK;
}

}  // namespace

// `seed` == nullptr: kb_datalog_fixpoint (every fact of the store is the first delta). With a seed (kb_datalog_fixpoint_seed): the store
// is taken as closed under the rules already, the seed facts the store does not hold yet are accepted into it and form the first delta.
static kb_status fixpoint_impl(kb_ctx* ctx, const kb_rule* rules, uint32_t n_rules, uint32_t strategy, bool seeded, const kb_rel* seed,
                               kb_rel** inferred, uint64_t* n_seed_new, kb_fixpoint_stats* stats) {
    if (!ctx) return KB_E_INVALID;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    cudaSetDevice(ctx->device);
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev};
    KB_TRY(begin_call(ctx));
    if ((n_rules && !rules) || !inferred) return fail(ctx, KB_E_INVALID, "NULL argument");
    if (strategy != KB_SEMI_NAIVE && strategy != KB_NAIVE && strategy != KB_SEMI_NAIVE_PARALLEL && strategy != KB_SEMI_NAIVE_OLD_DELTA)
        return fail(ctx, KB_E_INVALID, "unknown strategy %u", strategy);
    if (seeded && !seed) return fail(ctx, KB_E_INVALID, "NULL seed");
    if (seed && (seed->pair || seed->cols.size() != 3 || seed->col_of(0) != 0 || seed->col_of(1) != 1 || seed->col_of(2) != 2))
        return fail(ctx, KB_E_INVALID, "the seed must be a 3-column relation with slots 0,1,2 = s,p,o");
    if (seed && seed->n >= 0xFFFFFFF0ull) return fail(ctx, KB_E_LIMIT, "seed too large");
    const bool old_delta = strategy == KB_SEMI_NAIVE_OLD_DELTA;
    const bool strict = strategy == KB_SEMI_NAIVE_PARALLEL;
    kb_fixpoint_stats st{};
    ScopedEvent ev0, ev1;
    cudaEventRecord(ev0, ctx->st);

    // ---- compile the rules (host): variable slots, synthetic variables for constants in s/o (quirk Q6), safety checks
    std::vector<RulePlan> plans(n_rules);
    Fix fx;
    Trace tr;
    fx.ctx = ctx;
    for (u32 r = 0; r < n_rules; r++) {
        const kb_rule& rule = rules[r];
        if (rule.n_premise > KB_MAX_PREMISES || rule.n_conclusion > KB_MAX_CONCLUSIONS || rule.n_filters > KB_MAX_RULE_FILTERS)
            return fail(ctx, KB_E_LIMIT, "rule %u exceeds the premise/conclusion/filter limits", r);
        u32 mx = 0;
        auto upd = [&](const kb_term& t) { if (t.is_var) mx = std::max(mx, t.value + 1); };
        for (u32 i = 0; i < rule.n_premise; i++) { upd(rule.premise[i].s); upd(rule.premise[i].p); upd(rule.premise[i].o); }
        for (u32 i = 0; i < rule.n_conclusion; i++) { upd(rule.conclusion[i].s); upd(rule.conclusion[i].p); upd(rule.conclusion[i].o); }
        u32 next = mx;
        std::map<std::pair<u32, u32>, u32> synth;  // (position, constant) -> synthetic slot: "__const_subj_{c}" / "__const_obj_{c}"
        auto syn = [&](u32 pos, u32 c) {
            auto it = synth.find({pos, c});
            if (it != synth.end()) return it->second;
            synth[{pos, c}] = next;
            return next++;
        };
        std::set<u32> bound;
        for (u32 i = 0; i < rule.n_premise; i++) {
            const kb_pattern& p = rule.premise[i];
            Premise pr;
            pr.s_var = (int)(p.s.is_var ? p.s.value : syn(0, p.s.value));
            pr.o_var = (int)(p.o.is_var ? p.o.value : syn(2, p.o.value));
            pr.pred_const = !p.p.is_var;
            pr.pred = p.p.value;
            pr.s_is_const = !p.s.is_var; pr.s_const = p.s.value;
            pr.o_is_const = !p.o.is_var; pr.o_const = p.o.value;
            if (pr.s_var == pr.o_var)
                return fail(ctx, KB_E_UNSUPPORTED, "rule %u premise %u repeats one variable in subject and object", r, i);
            if (pr.pred_const) { fx.rels[pr.pred].pred = pr.pred; bound.insert((u32)pr.s_var); bound.insert((u32)pr.o_var); }
            plans[r].prem.push_back(pr);
        }
        plans[r].n_vars = next;
        if (next > KB_MAX_COLS) return fail(ctx, KB_E_LIMIT, "rule %u uses more than %d variables", r, KB_MAX_COLS);
        for (u32 c = 0; c < rule.n_conclusion; c++) {
            const kb_pattern& h = rule.conclusion[c];
            if (h.p.is_var) return fail(ctx, KB_E_UNSUPPORTED, "rule %u: variable predicate in a conclusion", r);
            const kb_term* ts[2] = {&h.s, &h.o};
            for (auto* t : ts) if (t->is_var && !bound.count(t->value))
                return fail(ctx, KB_E_UNSUPPORTED, "rule %u: head variable (slot %u) is not bound by the premises (quirk Q8: the reference invents ml_output_placeholder terms / id 0)", r, t->value);
            fx.rels[h.p.value].pred = h.p.value;
            fx.rels[h.p.value].is_head = true;
        }
        for (u32 f = 0; f < rule.n_filters; f++) {
            if (rule.filters[f].cmp > KB_CMP_NE) return fail(ctx, KB_E_INVALID, "rule %u filter %u: bad comparison", r, f);
        }
    }
    // a seed fact may belong to any predicate of the rules: each of them needs its known-fact set (to tell the new seed facts from the old)
    if (seed) for (auto& kv : fx.rels) kv.second.is_head = true;

    // ---- a seeded call may find the state its predecessor left (FixState): then nothing below reads the store
    const bool state_off = getenv("KOLIBRIE_FIX_STATE") && getenv("KOLIBRIE_FIX_STATE")[0] == '0';  // (read per call: a test flips it)
    const std::string sig = seed && !state_off ? rules_signature(rules, n_rules) : std::string();
    bool reused = false;
    if (seed && !state_off && ctx->fix_state) {
        FixState* fs = static_cast<FixState*>(ctx->fix_state.get());
        if (fs->version == ctx->store_version && fs->sig == sig) {
            fx.rels = std::move(fs->rels);
            reused = true;
        }
    }
    ctx->fix_state.reset();  // taken, stale, or about to be (an unseeded run appends to the store)

    // ---- split the store by predicate: one fused scan per 8 predicates
    if (!reused) {
        std::vector<u32> preds;
        for (auto& kv : fx.rels) preds.push_back(kv.first);
        for (size_t b = 0; b < preds.size(); b += MAXP) {
            const u32 k = (u32)std::min<size_t>(MAXP, preds.size() - b);
            kb_pattern pats[MAXP];
            for (u32 i = 0; i < k; i++) { pats[i].s = kb_term{1, 0}; pats[i].p = kb_term{0, preds[b + i]}; pats[i].o = kb_term{1, 1}; }
            std::vector<std::unique_ptr<kb_rel>> out;
            std::vector<FilterProg> none;
            // the predicate relations are append-only buffers with the capacity of a whole-store scan output: take the scan, not the
            // (exactly sized) index slices
            ctx->in_index_build = true;
            const kb_status src = scan_impl(ctx, pats, k, none, false, false, &out);
            ctx->in_index_build = false;
            KB_TRY(src);
            for (u32 i = 0; i < k; i++) {
                PredRel& r = fx.rels[preds[b + i]];
                r.s = out[i]->cols[0];
                r.o = out[i]->cols[1];
                r.n = r.base_n = out[i]->n;
                r.cap = ctx->n_triples;
                if (r.n * 4 < r.cap) {  // shrink: the scan allocates for the worst case
                    r.cap = 0;
                    Col os = r.s, oo = r.o;
                    u64 n = r.n;
                    r.n = 0;
                    KB_TRY(grow_rel(ctx, r, std::max<u64>(n * 2, 1024)));
                    if (n) {
                        KB_CUDA(ctx, cudaMemcpyAsync(r.s.ptr, os.ptr, n * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->st));
                        KB_CUDA(ctx, cudaMemcpyAsync(r.o.ptr, oo.ptr, n * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->st));
                    }
                    r.n = n;
                }
            }
        }
    }
    tr.mark(ctx, "split store by predicate");
    if (!reused) for (auto& kv : fx.rels) if (kv.second.is_head) {
        kv.second.set_count = kv.second.n;
        // A window that re-materialises the same rules every firing (simple_r2r.rs:103-128) ends with about as many facts as last time:
        // the set is built for that size at once (config 4: the 2^28 -> 2^30 rebuild in the middle of round 0 cost 6.7 ms of 64)
        u64 expect = 1024;
        auto hint = ctx->fix_hint.find(kv.first);
        if (hint != ctx->fix_hint.end() && hint->second > kv.second.n) expect = std::min<u64>(hint->second - kv.second.n, 0x7FFFFFFFull);
        KB_TRY(ensure_set(fx, kv.second, expect, false, /*exact_expectation=*/expect > 1024));
    }
    tr.mark(ctx, "initial known-fact sets");

    // ---- one head of one rule over the joined bindings `cur`: filters + instantiate + dedup against the known facts + append to the
    // round's pending facts (`rule_p` == nullptr: no rule filters — the seed facts of kb_datalog_fixpoint_seed take the same way in)
    auto derive_head = [&](const kb_rel* cur, const kb_pattern& h, const kb_rule* rule_p) -> kb_status {
        DeriveParams D{};
        for (size_t k = 0; k < cur->cols.size(); k++) D.bcol[k] = cur->cols[k].ptr;
        D.n = (u32)cur->n;
        auto term = [&](const kb_term& t) {
            HeadTerm ht;
            ht.is_var = t.is_var;
            ht.value = t.is_var ? (u32)cur->col_of(t.value) : t.value;
            return ht;
        };
        D.head_s = term(h.s);
        D.head_o = term(h.o);
        D.n_filt = 0;
        for (u32 f = 0; f < (strict || !rule_p ? 0u : rule_p->n_filters); f++) {  // the parallel variant never evaluates the rule filters
            const kb_rule_filter& rf = rule_p->filters[f];
            const int lc = cur->col_of(rf.lhs_slot);
            if (lc < 0 || rf.cmp == 0) continue;  // unbound lhs: the reference skips the filter (rules.rs:139)
            RuleFilterDev d;
            d.lhs_col = (u32)lc;
            d.cmp = rf.cmp;
            d.rhs_value = rf.rhs_value;
            d.rhs_is_var = 0; d.rhs_col = 0;
            if (rf.rhs_is_var) {
                const int rc = cur->col_of(rf.rhs_slot);
                if (rc >= 0) { d.rhs_is_var = 1; d.rhs_col = (u32)rc; }
                // rhs names an unbound variable: the reference falls to the numeric branch with value.parse() (rules.rs:148-151)
            }
            D.filt[D.n_filt++] = d;
        }
        D.nt = numtab(ctx);
        PredRel& hr = fx.rels[h.p.value];
        KB_TRY(ensure_set(fx, hr, cur->n));
        tr.mark(ctx, "  ensure_set", hr.set_slots);
        Pending pd;
        pd.pred = h.p.value;
        pd.count = 0;
        KB_TRY(alloc_col(ctx, cur->n, &pd.s));
        KB_TRY(alloc_col(ctx, cur->n, &pd.o));
        const u32 coff = ctrl_alloc(ctx, 8);
        KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + coff, 0, 8 * sizeof(u32), ctx->st));
        D.out_s = pd.s.ptr; D.out_o = pd.o.ptr;
        D.out_cap = (u32)cur->n;
        D.out_count = ctx->ctrl + coff;
        D.overflow = ctx->ctrl + coff + 1;
        D.n_deriv = reinterpret_cast<unsigned long long*>(ctx->ctrl + coff + 4);
        const u64 count_before = hr.set_count;
        for (;;) {
            D.set = static_cast<u64*>(hr.set->p);
            D.set_slots = hr.set_slots;
            // the set may fill up to 3/4 before it has to grow (expected: 1/2)
            D.budget = (u32)std::min<u64>((u64)hr.set_slots / 4 * 3 - count_before, 0xFFFFFFFFull);
            // Large candidate sets against a table far beyond L2: radix-partition the candidates by the high bits of their
            // home slot and probe partition by partition (each slice stays L2-resident) instead of one random DRAM line per
            // candidate. Partitions: ~32 MB slices, at least ~1 M candidates each (so that the CTAs of the probe pass sit on
            // one or two slices at a time), at most 1024.
            u64 n_parts = 0;
            {
                const u64 set_bytes = (u64)hr.set_slots * sizeof(u64);
                u64 want = std::min<u64>(set_bytes / ctx->derive_slice_bytes, std::min<u64>(cur->n / ctx->derive_min_part_rows, 1024));
                while (want >= 2 && n_parts * 2 <= want) n_parts = n_parts ? n_parts * 2 : 2;  // power of two
                if (ctx->derive_part == 0) n_parts = 0;
            }
            if (n_parts >= 2) {
                DerivePartParams Q{};
                Q.d = D;
                Q.n_parts = (u32)n_parts;
                u32 set_bits = 0;
                while ((1ull << set_bits) < hr.set_slots) set_bits++;
                u32 part_bits = 0;
                while ((1ull << part_bits) < n_parts) part_bits++;
                Q.slice_bits = set_bits - part_bits;
                Q.bucket_cap = (u32)std::min<u64>(cur->n / n_parts + cur->n / n_parts / 8 + ctx->derive_bucket_slack, 0xFFFFFFF0ull);
                Buf ctl;
                const size_t bucket_bytes = n_parts * (u64)Q.bucket_cap * sizeof(u64);
                if (!fx.buckets || fx.buckets->bytes < bucket_bytes) {
                    fx.buckets.reset();
                    KB_TRY(alloc_buf(ctx, bucket_bytes + bucket_bytes / 4, &fx.buckets));
                }
                Buf buckets = fx.buckets;
                KB_TRY(alloc_buf(ctx, (2 * n_parts + 8) * sizeof(u32), &ctl));
                KB_CUDA(ctx, cudaMemsetAsync(ctl->p, 0, (2 * n_parts + 8) * sizeof(u32), ctx->st));
                Q.buckets = static_cast<u64*>(buckets->p);
                Q.cursors = static_cast<u32*>(ctl->p);
                Q.tile_start = Q.cursors + n_parts;
                Q.tickets = Q.tile_start + n_parts + 1 + ((n_parts + 1) & 1u);
                tr.mark(ctx, "    derive: buckets allocated", n_parts);
                timer_begin(ctx, F_OTHER, 3);
                launch_derive_partitioned(Q, ctx->n_sms, ctx->st);
                timer_end(ctx);
                ctx->stats.rows_built += cur->n;  // (counted as table traffic: partitioned candidates)
            } else {
                timer_begin(ctx, F_OTHER);
                launch_derive(D, ctx->n_sms, ctx->st);
                timer_end(ctx);
            }
            KB_CUDA(ctx, cudaGetLastError());
            // counts are read immediately: the control arena may be recycled by later joins of this round
            KB_TRY(ctrl_read(ctx));
            pd.count = ctx->h_ctrl[coff];
            if (ctx->h_ctrl[coff + 1] == 1u) return fail(ctx, KB_E_LIMIT, "known-fact set overflow");
            if (ctx->h_ctrl[coff + 1] == 0u) break;
            // budget reached: rebuild the set larger WITH the facts appended so far, then repeat the launch (facts already
            // inserted are found present and are not appended twice; the derivation count is the last, complete pass's)
            fx.pend.push_back(pd);
            const kb_status gs = ensure_set(fx, hr, cur->n, true);
            fx.pend.pop_back();
            if (gs != KB_OK) return gs;
            KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + coff + 1, 0, sizeof(u32), ctx->st));
            KB_CUDA(ctx, cudaMemsetAsync(ctx->ctrl + coff + 4, 0, 2 * sizeof(u32), ctx->st));
        }
        unsigned long long nd;
        memcpy(&nd, ctx->h_ctrl + coff + 4, sizeof nd);
        st.derivations += nd;
        hr.set_count = count_before;
        hr.set_count += pd.count;
        if (pd.count) fx.pend.push_back(pd);
        tr.mark(ctx, "  derive", pd.count);
        return KB_OK;
    };

    // ---- seed: its facts take the way of derived heads (dedup against the known facts); the winners become visible now and are the
    // first delta, everything the store held is OLD
    u64 seed_new = 0;
    if (seed) {
        for (auto& kv : fx.rels) kv.second.snapshot = kv.second.seed_from = kv.second.n;
        for (auto& kv : fx.rels) {
            if (seed->n == 0) break;
            FilterProg f;
            kb_filter_op op{};
            op.op = KB_F_EQ_ID; op.slot = 1; op.id = kv.first;
            f.ops.push_back(op);
            std::unique_ptr<kb_rel> sub;
            KB_TRY(filter_impl(ctx, *seed, f, &sub));
            if (sub->n == 0) continue;
            kb_pattern h{};
            h.s = kb_term{1, 0}; h.p = kb_term{0, kv.first}; h.o = kb_term{1, 2};
            KB_TRY(derive_head(sub.get(), h, nullptr));
        }
        for (auto& kv : fx.rels) kv.second.delta_start = kv.second.snapshot;
        for (auto& pd : fx.pend) {
            PredRel& r = fx.rels[pd.pred];
            const u64 cnt = pd.count;
            KB_TRY(grow_rel(ctx, r, r.n + cnt));
            KB_CUDA(ctx, cudaMemcpyAsync(r.s.ptr + r.n, pd.s.ptr, cnt * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->st));
            KB_CUDA(ctx, cudaMemcpyAsync(r.o.ptr + r.n, pd.o.ptr, cnt * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->st));
            r.n += cnt;
            seed_new += cnt;
        }
        fx.pend.clear();
        for (auto& kv : fx.rels) kv.second.base_n = kv.second.n;
        st.derivations = 0;  // (the seed's facts are not derivations)
        tr.mark(ctx, "seed accepted", seed_new);
    }

    // ---- rounds
    for (u32 round = 0;; round++) {
        for (auto& kv : fx.rels) kv.second.snapshot = kv.second.n;
        fx.pend.clear();
        for (u32 r = 0; r < n_rules; r++) {
            const kb_rule& rule = rules[r];
            const RulePlan& pl = plans[r];
            const u32 np = (u32)pl.prem.size();
            if (np == 0) continue;
            bool dead = false;
            for (auto& pr : pl.prem) if (!pr.pred_const) dead = true;  // a variable predicate never matches (join_algorithm.rs:515-521)
            if (dead) continue;
            if (strict && np != 1 && np != 2) continue;  // semi_naive_parallel.rs:149: other arities are skipped
            const u32 n_start = strategy == KB_NAIVE ? 1 : np;
            for (u32 i = 0; i < n_start; i++) {
                // premise i over the delta (semi-naive) or over all facts (naive), then the others over all facts
                std::unique_ptr<kb_rel> cur;
                {
                    const PredRel& pr = fx.rels[pl.prem[i].pred];
                    const u64 a = strategy == KB_NAIVE ? 0 : pr.delta_start;
                    const u64 n = pr.snapshot - a;
                    if (n == 0) continue;
                    KB_TRY(make_view(ctx, pr, a, n, (u32)pl.prem[i].s_var, (u32)pl.prem[i].o_var, &cur));
                    if (strict) KB_TRY(enforce_constants(ctx, pl.prem[i], &cur));
                }
                for (u32 j = 0; j < np && cur->n; j++) {
                    if (j == i) continue;
                    const PredRel& pr = fx.rels[pl.prem[j].pred];
                    // KB_SEMI_NAIVE_OLD_DELTA: a premise before the delta premise sees only the facts older than its delta
                    const u64 visible = old_delta && j < i ? pr.delta_start : pr.snapshot;
                    if (visible == 0) { cur->n = 0; break; }
                    std::unique_ptr<kb_rel> all, joined;
                    KB_TRY(make_view(ctx, pr, 0, visible, (u32)pl.prem[j].s_var, (u32)pl.prem[j].o_var, &all));
                    if (strict) KB_TRY(enforce_constants(ctx, pl.prem[j], &all));
                    tr.mark(ctx, "  views");
                    KB_TRY(hash_join_impl(ctx, *cur, *all, nullptr, &joined));
                    cur = std::move(joined);
                    tr.mark(ctx, "  join", cur->n);
                }
                if (cur->n == 0) continue;
                // heads: filters + instantiate + dedup against known facts + append
                for (u32 c = 0; c < rule.n_conclusion; c++) {
                    KB_TRY(derive_head(cur.get(), rule.conclusion[c], &rule));
                }
            }
        }
        // ---- end of round: the facts that were new become visible (infer_generic.rs:42-48)
        u64 round_new = 0;
        for (auto& kv : fx.rels) kv.second.delta_start = kv.second.snapshot;
        for (auto& pd : fx.pend) {
            PredRel& r = fx.rels[pd.pred];
            const u64 cnt = pd.count;
            KB_TRY(grow_rel(ctx, r, r.n + cnt));
            KB_CUDA(ctx, cudaMemcpyAsync(r.s.ptr + r.n, pd.s.ptr, cnt * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->st));
            KB_CUDA(ctx, cudaMemcpyAsync(r.o.ptr + r.n, pd.o.ptr, cnt * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->st));
            r.n += cnt;
            round_new += cnt;
        }
        fx.pend.clear();
        tr.mark(ctx, "round end: append", round_new);
        if (round_new == 0) break;
        if (st.rounds < 64) st.round_new[st.rounds] = round_new;
        st.rounds++;
        st.inferred += round_new;
        if (round > 100000) return fail(ctx, KB_E_LIMIT, "fixpoint did not converge");
    }

    // ---- result: inferred facts as (s,p,o) columns; also appended to the store (infer_generic.rs:46 index_manager.insert)
    auto res = std::make_unique<kb_rel>();
    res->slots = {0, 1, 2};
    // (with a seed: the accepted seed facts first, then the inferred ones)
    const u64 total = seed_new + st.inferred;
    if (total >= 0xFFFFFFF0ull || ctx->n_triples + total >= 0xFFFFFFF0ull) return fail(ctx, KB_E_LIMIT, "the closure exceeds 2^32 - 16 triples");
    res->n = total;
    Col cs, cp, co;
    KB_TRY(alloc_col(ctx, total, &cs));
    KB_TRY(alloc_col(ctx, total, &cp));
    KB_TRY(alloc_col(ctx, total, &co));
    u64 off = 0;
    for (int pass = seed ? 0 : 1; pass < 2; pass++) {
        for (auto& kv : fx.rels) {
            PredRel& r = kv.second;
            const u64 from = pass == 0 ? r.seed_from : r.base_n, to = pass == 0 ? r.base_n : r.n;
            const u64 cnt = to - from;
            if (!cnt) continue;
            KB_CUDA(ctx, cudaMemcpyAsync(cs.ptr + off, r.s.ptr + from, cnt * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->st));
            KB_CUDA(ctx, cudaMemcpyAsync(co.ptr + off, r.o.ptr + from, cnt * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->st));
            launch_fill_u32(cp.ptr + off, r.pred, cnt, ctx->st);
            off += cnt;
        }
    }
    res->cols = {cs, cp, co};
    for (auto& kv : fx.rels) if (kv.second.is_head) ctx->fix_hint[kv.first] = kv.second.n;
    if (total) {
        Segment sg;
        sg.tag = KB_TAG_INFERRED;
        sg.n = total;
        sg.s = cs; sg.p = cp; sg.o = co;
        ctx->segs.push_back(sg);
        ctx->n_triples += total;
        ctx->store_version++;
        ctx->multi_valued.clear();
        ctx->single_valued.clear();
        ctx->index.clear();
    }
    tr.mark(ctx, "result assembly", total);
    cudaEventRecord(ev1, ctx->st);
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->st));
    timers_flush(ctx);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    st.device_ms = ms;
    if (seed && !state_off) {  // the next seeded call continues from here
        auto fs = std::make_shared<FixState>();
        fs->sig = sig;
        fs->version = ctx->store_version;
        fs->rels = std::move(fx.rels);
        ctx->fix_state = fs;
    }
    if (stats) *stats = st;
    if (n_seed_new) *n_seed_new = seed_new;
    ctx->stats.rows_out = st.inferred;
    *inferred = res.release();
    return KB_OK;
}

extern "C" kb_status kb_datalog_fixpoint(kb_ctx* ctx, const kb_rule* rules, uint32_t n_rules, uint32_t strategy, kb_rel** inferred,
                                         kb_fixpoint_stats* stats) {
    return fixpoint_impl(ctx, rules, n_rules, strategy, false, nullptr, inferred, nullptr, stats);
}

// Incremental materialisation: what Reasoner::add_abox_triple + infer_new_facts_semi_naive (reasoning.rs, semi_naive.rs:89) reach by
// starting over, reached from the delta alone. Used per window slide (simple_r2r.rs:95-128 adds the slide's triples and materialises
// again) and by the sharded fixpoint (kolibrie_b200/dist.py: facts derived on another rank arrive as seeds).
extern "C" kb_status kb_datalog_fixpoint_seed(kb_ctx* ctx, const kb_rule* rules, uint32_t n_rules, uint32_t strategy, const kb_rel* seed,
                                              kb_rel** out, uint64_t* n_seed_new, kb_fixpoint_stats* stats) {
    return fixpoint_impl(ctx, rules, n_rules, strategy, true, seed, out, n_seed_new, stats);
}
