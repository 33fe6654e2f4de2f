"""Writes the JSON fixtures in this directory: hand transcriptions of the INPUTS and ASSERTIONS of the reference's own tests
(each file's "source" cites them). The reference is a Rust workspace and cannot run in this image, so no reference OUTPUT is
captured here — the fixtures pin the oracle (tests/test_oracle_golden.py) and then the device (tests/test_gpu_*.py).
Run: python tests/golden/make_fixtures.py"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def R_base(a, b):
    return {"premise": [["?X", a, "?Y"]], "conclusion": [["?X", b, "?Y"]], "filters": []}


trans = {"premise": [["?X", "ancestor", "?Y"], ["?Y", "ancestor", "?Z"]], "conclusion": [["?X", "ancestor", "?Z"]], "filters": []}


def sib(z):
    return {"premise": [["?X", "parent", "?" + z], ["?Y", "parent", "?" + z]], "conclusion": [["?X", "sibling", "?Y"]], "filters": [["X", "!=", "Y"]]}


cases = [
    {"name": "fc_1hop_base", "line": 29, "facts": [["A", "parent", "B"]], "encode_after": ["parent", "ancestor"], "rules": [R_base("parent", "ancestor")],
     "present": [["A", "ancestor", "B"]], "absent": []},
    {"name": "fc_2hop_transitive", "line": 47, "facts": [["A", "parent", "B"], ["B", "parent", "C"]], "encode_after": ["parent", "ancestor"],
     "rules": [R_base("parent", "ancestor"), trans], "present": [["A", "ancestor", "B"], ["B", "ancestor", "C"], ["A", "ancestor", "C"]], "absent": []},
    {"name": "fc_3hop_transitive", "line": 77, "facts": [["A", "parent", "B"], ["B", "parent", "C"], ["C", "parent", "D"]], "encode_after": ["parent", "ancestor"],
     "rules": [R_base("parent", "ancestor"), trans],
     "present": [["A", "ancestor", "B"], ["A", "ancestor", "C"], ["A", "ancestor", "D"], ["B", "ancestor", "D"]], "absent": []},
    {"name": "fc_join_sibling", "line": 107, "facts": [["A", "parent", "P"], ["B", "parent", "P"]], "encode_after": ["parent", "sibling"], "rules": [sib("P2")],
     "present": [["A", "sibling", "B"], ["B", "sibling", "A"]], "absent": [["A", "sibling", "A"]]},
    {"name": "fc_multi_rule_cascade", "line": 137, "facts": [["A", "worksFor", "Corp"]], "encode_after": ["worksFor", "employed", "affiliated"],
     "rules": [R_base("worksFor", "employed"), R_base("employed", "affiliated")], "present": [["A", "employed", "Corp"], ["A", "affiliated", "Corp"]], "absent": []},
    {"name": "fc_three_premise_rule", "line": 161, "facts": [["A", "R", "B"], ["B", "S", "C"], ["C", "T", "D"]], "encode_after": ["R", "S", "T", "connected"],
     "rules": [{"premise": [["?X", "R", "?Y"], ["?Y", "S", "?Z"], ["?Z", "T", "?W"]], "conclusion": [["?X", "connected", "?W"]], "filters": []}],
     "present": [["A", "connected", "D"]], "absent": []},
    {"name": "fc_no_spurious", "line": 187, "facts": [["A", "parent", "B"], ["C", "unrelated", "D"]], "encode_after": ["parent", "ancestor"],
     "rules": [R_base("parent", "ancestor")], "present": [["A", "ancestor", "B"]], "absent": [["C", "ancestor", "D"]]},
    {"name": "fc_sibling_three_children", "line": 207, "facts": [["A", "parent", "P"], ["B", "parent", "P"], ["C", "parent", "P"]], "encode_after": ["parent", "sibling"],
     "rules": [sib("Z")], "present": [[a, "sibling", b] for a in "ABC" for b in "ABC" if a != b], "absent": [[x, "sibling", x] for x in "ABC"]},
    {"name": "fc_multi_conclusion", "line": 241, "facts": [["A", "marriedTo", "B"]], "encode_after": ["marriedTo", "spouse", "partner"],
     "rules": [{"premise": [["?X", "marriedTo", "?Y"]], "conclusion": [["?X", "spouse", "?Y"], ["?X", "partner", "?Y"]], "filters": []}],
     "present": [["A", "spouse", "B"], ["A", "partner", "B"]], "absent": []},
    {"name": "fc_diamond_ancestor", "line": 265, "facts": [["A", "parent", "B"], ["A", "parent", "C"], ["B", "parent", "D"], ["C", "parent", "D"]],
     "encode_after": ["parent", "ancestor"], "rules": [R_base("parent", "ancestor"), trans],
     "present": [["A", "ancestor", "D"], ["B", "ancestor", "D"], ["C", "ancestor", "D"]], "absent": [["A", "ancestor", "A"], ["D", "ancestor", "A"]]},
    {"name": "fc_disconnected_graphs", "line": 298, "facts": [["A", "parent", "B"], ["X", "parent", "Y"]], "encode_after": ["parent", "ancestor"],
     "rules": [{"premise": [["?P", "parent", "?Q"]], "conclusion": [["?P", "ancestor", "?Q"]], "filters": []}],
     "present": [["A", "ancestor", "B"], ["X", "ancestor", "Y"]], "absent": [["A", "ancestor", "Y"], ["X", "ancestor", "B"]]},
    {"name": "fc_no_matching_facts", "line": 321, "facts": [["A", "likes", "B"]], "encode_after": ["parent", "ancestor"], "rules": [R_base("parent", "ancestor")],
     "present": [], "absent": [], "expect_empty": True},
    {"name": "fc_idempotent", "line": 340, "facts": [["A", "parent", "B"]], "encode_after": ["parent", "ancestor"], "rules": [R_base("parent", "ancestor")],
     "present": [["A", "ancestor", "B"]], "absent": [], "idempotent": True, "exactly_once": [["A", "ancestor", "B"]]},
    {"name": "fc_uncle_derived", "line": 362, "facts": [["A", "parent", "P"], ["B", "parent", "P"], ["C", "parent", "A"]], "encode_after": ["parent", "sibling", "uncle"],
     "rules": [sib("Z"), {"premise": [["?U", "sibling", "?Par"], ["?N", "parent", "?Par"]], "conclusion": [["?U", "uncle", "?N"]], "filters": []}],
     "present": [["A", "sibling", "B"], ["B", "sibling", "A"], ["B", "uncle", "C"]], "absent": [["A", "uncle", "C"]]},
]


def dump(name, doc):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(doc, f, indent=1)


dump("datalog_fc.json", {
    "source": "/root/reference/datalog/tests/reasoning_tests.rs:28-404 (14 fc_* known-answer tests: facts, rules, asserted presence/absence)",
    "cases": cases})

terms = ["http://example.org/person1", "http://example.org/person2", "http://example.org/company1", "ex:name", "ex:age", "ex:email", "ex:worksFor",
         "ex:founded", "ex:industry", "John Smith", "Jane Doe", "ACME Corp", "30", "25", "john@example.com", "jane@example.com", "2000", "Technology"]
triples = [(0, 3, 9), (0, 4, 12), (0, 5, 14), (0, 6, 2), (1, 3, 10), (1, 4, 13), (1, 5, 15), (1, 6, 2), (2, 3, 11), (2, 7, 16), (2, 8, 17)]
dump("integration_fixture.json", {
    "source": "/root/reference/kolibrie/tests/integration_test.rs:19-76 (fixture), :131-299 (asserted counts), :286-299 and :302-342 (two joins whose answers the test asserts)",
    "terms": terms, "triples": triples,
    "expect": {"subject==person1": 4, "predicate==ex:name": 3, "object==Jane Doe": 1, "numeric_objects": 3, "count(ex:name)": 3,
               "worksFor_company1_subjects": [0, 1], "person1_emails_after_add": 2,
               # :286-299  ?c ex:industry "Technology" . ?e ex:worksFor ?c      -> 2 employees, person1 among them
               "tech_employees": [0, 1],
               # :302-342  ?c ex:name "ACME Corp" . ?e ex:worksFor ?c . ?e ex:age ?a FILTER(?a < 30)  -> person2 only
               "young_acme_employees": [1]}})

dump("employee4.json", {
    "source": "/root/reference/kolibrie/examples/sparql_syntax/simple_select/simple_select_synth_data.rs:16-52",
    "employees": [["http://example.org/employee1", "Developer", "73681"], ["http://example.org/employee2", "Developer", "83504"],
                  ["http://example.org/employee3", "Developer", "90065"], ["http://example.org/employee4", "Manager", "67751"]],
    "workplace": "Company Name", "expect_rows": 4})

dump("rust_parse_f64.json", {
    "source": "Rust core::num::dec2flt grammar (SURVEY.md section 7 'FILTER parity'); reached through str::parse::<f64> at kolibrie/src/streamertail_optimizer/types.rs:133-148",
    "accept": {"1": 1.0, "+1.0": 1.0, "-2.5": -2.5, ".5": 0.5, "5.": 5.0, "1e5": 100000.0, "1E-2": 0.01, "1.5e+3": 1500.0, "inf": "inf", "-inf": "-inf",
               "Infinity": "inf", "INFINITY": "inf", "nan": "nan", "NaN": "nan", "+nan": "nan", "100000": 100000.0, "0": 0.0, "007": 7.0},
    "reject": ["", " 1", "1 ", "1_000", "0x10", "1e", "e5", ".", "+", "-", "1.2.3", "--1", "infinit", "nane", "1f", "Developer",
               "http://example.org/employee1", "١", "5\n", " 5", "5 ", "5\t", "\n5", "1e5\n"]})

EX = "http://example.org/"
dump("rdf_star.json", {
    "source": "/root/reference/kolibrie/tests/rdf_star_test.rs:107-145 (SPARQL-star scans), :281-329 (quoted triple as a bound value), :384-405 (DELETE WHERE); "
              "/root/reference/shared/src/quoted_triple_store.rs:82-157 (QuotedTripleStore unit tests, replayed by tests/test_rdf_star_rsp_golden.py)",
    "note": "terms are strings; a nested [s, p, o] list is a quoted triple << s p o >>; '?x' is a variable",
    "cases": [
        {"name": "constant_quoted_triple", "line": 107,
         "data": [[[EX + "emp38", EX + "jobTitle", EX + "AssistantDesigner"], EX + "statedBy", EX + "emp22"]],
         "pattern": [[EX + "emp38", EX + "jobTitle", EX + "AssistantDesigner"], EX + "statedBy", "?who"],
         "select": ["who"], "rows": [[EX + "emp22"]]},
        {"name": "variable_in_quoted_triple", "line": 126,
         "data": [[[EX + "emp38", EX + "jobTitle", EX + "AssistantDesigner"], EX + "statedBy", EX + "emp22"],
                  [[EX + "emp39", EX + "jobTitle", EX + "Designer"], EX + "statedBy", EX + "emp23"]],
         "pattern": [["?emp", EX + "jobTitle", "?title"], EX + "statedBy", "?who"],
         "select": ["emp", "title", "who"], "n_rows": 2,
         "rows": [[EX + "emp38", EX + "AssistantDesigner", EX + "emp22"], [EX + "emp39", EX + "Designer", EX + "emp23"]]},
        {"name": "quoted_triple_bound_to_a_variable", "line": 281,
         "data": [[[EX + "alice", EX + "knows", EX + "bob"], EX + "source", EX + "doc1"]],
         "pattern": ["?t", EX + "source", EX + "doc1"],
         "select": ["t"], "n_rows": 1, "rows": [["<< " + EX + "alice " + EX + "knows " + EX + "bob >>"]], "subject_of_t": EX + "alice"},
    ],
    "delete_where": {"line": 384,
                     "data": [[EX + "alice", EX + "knows", EX + "bob"], [EX + "alice", EX + "knows", EX + "carol"], [EX + "alice", EX + "name", "Alice"]],
                     "delete_pattern": ["?s", EX + "knows", "?o"], "triples_before": 3, "triples_after": 1}})

T = "http://test/"
dump("rsp_windows.json", {
    "source": "/root/reference/kolibrie/tests/rsp_engine_test.rs:24-112 (rsp_ql_istream_semantics), :935-1027 (rsp_ql_istream_range3_step1), "
              ":1029-1084 (test_window_evicts_old_data), :1103-1200 (rsp_ql_istream_same_sp_diff_object)",
    "note": "The window operator (S2R) and the stream operator (R2S) are out of the hot path; what the hot path sees per firing is the window's "
            "content (rsp_engine.rs:94-104: evict the previous window, add the current one) and the query over it. `firings` lists each "
            "firing's window content exactly as the reference test's comments state it, and the rows its R2S operator emits.",
    "cases": [
        {"name": "istream_semantics", "line": 24, "stream": "ISTREAM", "pattern": ["?s", "a", T + "IType"], "select": ["s"],
         "firings": [{"window": [[T + "subjectA", "a", T + "IType"]], "emit": [[T + "subjectA"]]},
                     {"window": [[T + "subjectA", "a", T + "IType"], [T + "subjectB", "a", T + "IType"]], "emit": [[T + "subjectB"]]},
                     {"window": [[T + "subjectA", "a", T + "IType"], [T + "subjectB", "a", T + "IType"], [T + "subjectC", "a", T + "IType"]], "emit": [[T + "subjectC"]]}]},
        {"name": "istream_range3_step1", "line": 935, "stream": "ISTREAM", "pattern": ["?s", "a", T + "RType"], "select": ["s"],
         "firings": [{"window": [[T + "subjectA", "a", T + "RType"]], "emit": [[T + "subjectA"]]},
                     {"window": [[T + "subjectA", "a", T + "RType"], [T + "subjectB", "a", T + "RType"]], "emit": [[T + "subjectB"]]},
                     {"window": [[T + "subjectA", "a", T + "RType"], [T + "subjectB", "a", T + "RType"], [T + "subjectC", "a", T + "RType"]], "emit": [[T + "subjectC"]]}]},
        {"name": "window_evicts_old_data", "line": 1029, "stream": "RSTREAM", "pattern": ["?s", "a", EX + "Type"], "select": ["s"],
         "firings": [{"window": [[EX + "subject1", "a", EX + "Type"]], "emit": [[EX + "subject1"]]},
                     {"window": [[EX + "subject2", "a", EX + "Type"]], "emit": [[EX + "subject2"]]},
                     {"window": [[EX + "subject3", "a", EX + "Type"]], "emit": [[EX + "subject3"]]}]},
        {"name": "istream_same_sp_diff_object", "line": 1103, "stream": "ISTREAM", "pattern": ["?reading", T + "hasTemp", "?temp"], "select": ["reading", "temp"],
         "firings": [{"window": [[T + "reading1", T + "hasTemp", "\"1\""]], "emit": [[T + "reading1", "\"1\""]]},
                     {"window": [[T + "reading1", T + "hasTemp", "\"1\""], [T + "reading1", T + "hasTemp", "\"2\""]], "emit": [[T + "reading1", "\"2\""]]},
                     {"window": [[T + "reading1", T + "hasTemp", "\"1\""], [T + "reading1", T + "hasTemp", "\"2\""], [T + "reading1", T + "hasTemp", "\"3\""]],
                      "emit": [[T + "reading1", "\"3\""]]}]},
    ]})
