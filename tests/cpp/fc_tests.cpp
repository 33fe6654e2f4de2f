// fc_tests.cpp — the reference's forward-chaining tests (datalog/tests/reasoning_tests.rs:28-404) and two executor queries
// (kolibrie/tests/integration_test.rs fixture; simple_select_synth_data.rs), written against the C++ host mirror so they read like
// the reference's own tests. Built and run by tests/test_gpu_cpp_host.py on the GPU box. Exit code = number of failed checks.
#include <array>
#include <cstdio>
#include <string>

#include "../../kolibrie_b200/host/kolibrie_host.hpp"

using namespace kolibrie;
static int g_failed = 0, g_checks = 0;
#define CHECK(cond, msg) do { g_checks++; if (!(cond)) { g_failed++; std::fprintf(stderr, "FAILED %s:%d %s\n", __func__, __LINE__, msg); } } while (0)

static std::shared_ptr<Device> g_dev;
static uint32_t enc(Reasoner& r, const char* s) { return r.dictionary.encode(s); }
static Rule rule(std::vector<TriplePattern> premises, std::vector<TriplePattern> conclusions) { return Rule{std::move(premises), {}, std::move(conclusions)}; }
static bool inferred(Reasoner& r, const char* s, const char* p, const char* o) { return !r.query_abox(std::string(s), std::string(p), std::string(o)).empty(); }
static Term V(const char* n) { return Term::Variable(n); }
static Term C(uint32_t c) { return Term::Constant(c); }

static void fc_1hop_base() {
    Reasoner r(g_dev);
    r.add_abox_triple("A", "parent", "B");
    auto parent = enc(r, "parent"), ancestor = enc(r, "ancestor");
    r.add_rule(rule({{V("X"), C(parent), V("Y")}}, {{V("X"), C(ancestor), V("Y")}}));
    r.infer_new_facts_semi_naive();
    CHECK(inferred(r, "A", "ancestor", "B"), "A ancestor B");
}
static void fc_3hop_transitive() {
    Reasoner r(g_dev);
    r.add_abox_triple("A", "parent", "B");
    r.add_abox_triple("B", "parent", "C");
    r.add_abox_triple("C", "parent", "D");
    auto parent = enc(r, "parent"), ancestor = enc(r, "ancestor");
    r.add_rule(rule({{V("X"), C(parent), V("Y")}}, {{V("X"), C(ancestor), V("Y")}}));
    r.add_rule(rule({{V("X"), C(ancestor), V("Y")}, {V("Y"), C(ancestor), V("Z")}}, {{V("X"), C(ancestor), V("Z")}}));
    r.infer_new_facts_semi_naive();
    CHECK(inferred(r, "A", "ancestor", "B"), "A anc B");
    CHECK(inferred(r, "A", "ancestor", "C"), "A anc C");
    CHECK(inferred(r, "A", "ancestor", "D"), "A anc D");
    CHECK(inferred(r, "B", "ancestor", "D"), "B anc D");
}
static void fc_sibling_three_children() {
    Reasoner r(g_dev);
    r.add_abox_triple("A", "parent", "P");
    r.add_abox_triple("B", "parent", "P");
    r.add_abox_triple("C", "parent", "P");
    auto parent = enc(r, "parent"), sibling = enc(r, "sibling");
    r.add_rule(Rule{{{V("X"), C(parent), V("Z")}, {V("Y"), C(parent), V("Z")}}, {FilterCondition{"X", "!=", "Y"}}, {{V("X"), C(sibling), V("Y")}}});
    r.infer_new_facts_semi_naive();
    const char* names[3] = {"A", "B", "C"};
    for (auto s : names) for (auto o : names) {
        if (std::string(s) != o) CHECK(inferred(r, s, "sibling", o), "sibling pair");
        else CHECK(!inferred(r, s, "sibling", o), "no self sibling");
    }
}
static void fc_three_premise_rule() {
    Reasoner r(g_dev);
    r.add_abox_triple("A", "R", "B");
    r.add_abox_triple("B", "S", "C");
    r.add_abox_triple("C", "T", "D");
    auto rp = enc(r, "R"), sp = enc(r, "S"), tp = enc(r, "T"), connected = enc(r, "connected");
    r.add_rule(rule({{V("X"), C(rp), V("Y")}, {V("Y"), C(sp), V("Z")}, {V("Z"), C(tp), V("W")}}, {{V("X"), C(connected), V("W")}}));
    r.infer_new_facts_semi_naive();
    CHECK(inferred(r, "A", "connected", "D"), "A connected D");
}
static void fc_multi_conclusion_and_cascade() {
    Reasoner r(g_dev);
    r.add_abox_triple("A", "marriedTo", "B");
    r.add_abox_triple("A", "worksFor", "Corp");
    auto married = enc(r, "marriedTo"), spouse = enc(r, "spouse"), partner = enc(r, "partner");
    auto works = enc(r, "worksFor"), employed = enc(r, "employed"), affiliated = enc(r, "affiliated");
    r.add_rule(rule({{V("X"), C(married), V("Y")}}, {{V("X"), C(spouse), V("Y")}, {V("X"), C(partner), V("Y")}}));
    r.add_rule(rule({{V("X"), C(works), V("Y")}}, {{V("X"), C(employed), V("Y")}}));
    r.add_rule(rule({{V("X"), C(employed), V("Y")}}, {{V("X"), C(affiliated), V("Y")}}));
    r.infer_new_facts_semi_naive();
    CHECK(inferred(r, "A", "spouse", "B") && inferred(r, "A", "partner", "B"), "multi conclusion");
    CHECK(inferred(r, "A", "employed", "Corp") && inferred(r, "A", "affiliated", "Corp"), "cascade");
}
static void fc_diamond_and_disconnected() {
    Reasoner r(g_dev);
    r.add_abox_triple("A", "parent", "B"); r.add_abox_triple("A", "parent", "C");
    r.add_abox_triple("B", "parent", "D"); r.add_abox_triple("C", "parent", "D");
    r.add_abox_triple("X", "parent", "Y");
    auto parent = enc(r, "parent"), ancestor = enc(r, "ancestor");
    r.add_rule(rule({{V("X"), C(parent), V("Y")}}, {{V("X"), C(ancestor), V("Y")}}));
    r.add_rule(rule({{V("X"), C(ancestor), V("Y")}, {V("Y"), C(ancestor), V("Z")}}, {{V("X"), C(ancestor), V("Z")}}));
    r.infer_new_facts_semi_naive();
    CHECK(inferred(r, "A", "ancestor", "D") && inferred(r, "B", "ancestor", "D") && inferred(r, "C", "ancestor", "D"), "diamond");
    CHECK(!inferred(r, "A", "ancestor", "A") && !inferred(r, "D", "ancestor", "A"), "no cycles invented");
    CHECK(inferred(r, "X", "ancestor", "Y") && !inferred(r, "A", "ancestor", "Y") && !inferred(r, "X", "ancestor", "B"), "disconnected graphs");
}
static void fc_no_matching_and_idempotent() {
    Reasoner r(g_dev);
    r.add_abox_triple("A", "likes", "B");
    auto parent = enc(r, "parent"), ancestor = enc(r, "ancestor");
    r.add_rule(rule({{V("X"), C(parent), V("Y")}}, {{V("X"), C(ancestor), V("Y")}}));
    CHECK(r.infer_new_facts_semi_naive().empty(), "no facts when no premise matches");
    r.add_abox_triple("A", "parent", "B");
    r.infer_new_facts_semi_naive();
    CHECK(r.infer_new_facts_semi_naive().empty(), "second pass derives nothing");
    CHECK(r.query_abox(std::string("A"), std::string("ancestor"), std::string("B")).size() == 1, "exactly one ancestor triple");
}
static void fc_uncle_derived() {
    Reasoner r(g_dev);
    r.add_abox_triple("A", "parent", "P"); r.add_abox_triple("B", "parent", "P"); r.add_abox_triple("C", "parent", "A");
    auto parent = enc(r, "parent"), sibling = enc(r, "sibling"), uncle = enc(r, "uncle");
    r.add_rule(Rule{{{V("X"), C(parent), V("Z")}, {V("Y"), C(parent), V("Z")}}, {FilterCondition{"X", "!=", "Y"}}, {{V("X"), C(sibling), V("Y")}}});
    r.add_rule(rule({{V("U"), C(sibling), V("Par")}, {V("N"), C(parent), V("Par")}}, {{V("U"), C(uncle), V("N")}}));
    r.infer_new_facts_semi_naive();
    CHECK(inferred(r, "A", "sibling", "B") && inferred(r, "B", "sibling", "A"), "siblings");
    CHECK(inferred(r, "B", "uncle", "C"), "B uncle C");
    CHECK(!inferred(r, "A", "uncle", "C"), "A is the parent, not the uncle");
}

// executor: the 4-employee dataset (simple_select_synth_data.rs:16-52) through PhysicalOperator trees
static void executor_employee4() {
    SparqlDatabase db(g_dev);
    const char* emp[4][3] = {{"http://example.org/employee1", "Developer", "73681"}, {"http://example.org/employee2", "Developer", "83504"},
                             {"http://example.org/employee3", "Developer", "90065"}, {"http://example.org/employee4", "Manager", "67751"}};
    for (auto& e : emp) {
        db.add_triple_parts(e[0], "foaf:name", e[0]);
        db.add_triple_parts(e[0], "foaf:title", e[1]);
        db.add_triple_parts(e[0], "foaf:workplaceHomepage", "Company Name");
        db.add_triple_parts(e[0], "ds:annual_salary", e[2]);
    }
    db.build_all_indexes();
    auto K = [&](const char* s) { return C(db.dictionary.encode(s)); };
    auto scan_sal = PhysicalOperator::Scan(PhysicalOperator::IndexScan, {V("?employee"), K("ds:annual_salary"), V("?salary")});
    auto rows = ExecutionEngine::execute(scan_sal, db);
    CHECK(rows.size() == 4, "4 salaries");
    // ?employee foaf:workplaceHomepage ?w . ?employee ds:annual_salary ?salary (benches/my_benchmark.rs:29-41) as a bind join
    auto join = PhysicalOperator::Join(PhysicalOperator::ParallelJoin,
                                       PhysicalOperator::Scan(PhysicalOperator::IndexScan, {V("?employee"), K("foaf:workplaceHomepage"), V("?w")}), scan_sal);
    rows = ExecutionEngine::execute(join, db);
    CHECK(rows.size() == 4, "4 joined rows");
    for (auto& r : rows) CHECK(r.at("w") == "Company Name" && r.count("salary") && r.count("employee"), "joined row shape");
    // 3-pattern star + FILTER(?salary > 80000) + projection
    auto star = PhysicalOperator::Star("?employee", {{V("?employee"), K("foaf:title"), V("?t")}, {V("?employee"), K("ds:annual_salary"), V("?salary")},
                                                     {V("?employee"), K("foaf:name"), V("?n")}});
    auto plan = PhysicalOperator::ProjectionOf(PhysicalOperator::FilterOf(star, Condition{FilterExpression::Cmp("?salary", ">", "80000")}), {"?employee", "?salary"});
    rows = ExecutionEngine::execute(plan, db);
    CHECK(rows.size() == 2, "two salaries above 80000");
    for (auto& r : rows) CHECK(r.size() == 2 && std::stod(r.at("salary")) > 80000, "projection + filter");
}

// round 2: prepared plan, bulk encode, window slides through the C++ mirror
static void prepared_plan_bulk_load_and_slides() {
    SparqlDatabase db(g_dev);
    // bulk load = one device encode of all terms; must equal the per-term loop
    SparqlDatabase seq(g_dev);
    std::vector<std::string> terms;
    for (int e = 0; e < 500; e++) {
        const std::string s = "http://example.org/employee" + std::to_string(e);
        const std::string sal = std::to_string(60000 + (e * 7919) % 50000);
        const char* title = e % 3 ? "Developer" : "Manager";
        for (auto& t : {std::array<std::string, 3>{s, "foaf:name", s}, {s, "foaf:title", title}, {s, "ds:annual_salary", sal}}) {
            terms.insert(terms.end(), t.begin(), t.end());
            seq.add_triple_parts(t[0], t[1], t[2]);
        }
    }
    db.add_triples_bulk(terms);
    CHECK(db.dictionary.id_to_string == seq.dictionary.id_to_string, "bulk encode == sequential encode (dictionary)");
    CHECK(db.triples == seq.triples, "bulk encode == sequential encode (triples)");
    // the same star + FILTER through a prepared plan (asynchronous submit / collect) and through the synchronous operator
    auto K = [&](const char* s) { return C(db.dictionary.encode(s)); };
    std::vector<TriplePattern> pats = {{V("?employee"), K("foaf:title"), V("?t")}, {V("?employee"), K("ds:annual_salary"), V("?salary")},
                                       {V("?employee"), K("foaf:name"), V("?n")}};
    Condition cond{FilterExpression::Cmp("?salary", ">", "90000")};
    auto rows = ExecutionEngine::execute_with_ids(PhysicalOperator::FilterOf(PhysicalOperator::Star("?employee", pats), cond), db);
    size_t want = 0;
    for (int e = 0; e < 500; e++) want += (60000 + (e * 7919) % 50000) > 90000;
    CHECK(rows.size() == want, "synchronous star join rows");
    PreparedStarJoin plan(db, "?employee", pats, &cond, 4);
    uint64_t t[3] = {plan.submit(), plan.submit(), plan.submit()};
    for (uint64_t tk : t) CHECK(plan.collect(tk) == want, "prepared plan rows");
    // window slides: index built once, maintained by append / evict
    WindowStore w(g_dev);
    g_dev->check(kb_store_clear(g_dev->get()));
    const uint32_t P1 = 7, P2 = 8;
    auto slide = [&](uint32_t first, uint32_t n) {
        std::vector<Triple> v;
        for (uint32_t i = 0; i < n; i++) { v.push_back(Triple{1000 + first + i, P1, 50}); v.push_back(Triple{1000 + first + i, P2, 60 + i % 5}); }
        return v;
    };
    w.append_slide(1, slide(0, 300));
    w.build_index();
    std::vector<kb_pattern> ps = {kb_pattern{{1, 0}, {0, P1}, {1, 1}}, kb_pattern{{1, 0}, {0, P2}, {1, 2}}};
    CHECK(w.star_join_rows(0, ps) == 300, "first slide");
    w.append_slide(2, slide(300, 200));
    CHECK(w.star_join_rows(0, ps) == 500, "two slides");
    w.evict_slide(1);
    CHECK(w.star_join_rows(0, ps) == 200, "after evicting the first slide");
    kb_stats st{};
    g_dev->check(kb_get_stats(g_dev->get(), &st, 0));
    CHECK(st.index_joins >= 3, "the slides stayed on the index path");
}

int main() {
    try {
        g_dev = std::make_shared<Device>(0);
        fc_1hop_base(); fc_3hop_transitive(); fc_sibling_three_children(); fc_three_premise_rule(); fc_multi_conclusion_and_cascade();
        fc_diamond_and_disconnected(); fc_no_matching_and_idempotent(); fc_uncle_derived(); executor_employee4();
        prepared_plan_bulk_load_and_slides();
    } catch (const GpuError& e) {
        std::fprintf(stderr, "GpuError %d: %s\n", e.status, e.what());
        return 100;
    }
    std::printf("%d checks, %d failed\n", g_checks, g_failed);
    return g_failed;
}
