// CPU-only checks of the C++ host mirror (kolibrie_b200/host/kolibrie_host.hpp): no device is touched.
// stdin: one hex-encoded string per line. stdout per line: "1 <value as %a>" when rust_parse_f64 accepts it, "0" when it does not,
// then "D <id>" = the id Dictionary::encode hands the string (first-seen order, dictionary.rs:32-48). Driven by tests/test_cpp_host_cpu.py.
#include <cstdio>
#include <iostream>
#include <string>

#include "../../kolibrie_b200/host/kolibrie_host.hpp"

static int hexval(char c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : -1; }

int main() {
    kolibrie::Dictionary dict;
    std::string line;
    while (std::getline(std::cin, line)) {
        std::string s;
        for (size_t i = 0; i + 1 < line.size(); i += 2) s.push_back((char)(hexval(line[i]) * 16 + hexval(line[i + 1])));
        auto v = kolibrie::rust_parse_f64(s);
        if (v) std::printf("1 %a\n", *v);
        else std::printf("0\n");
        std::printf("D %u\n", dict.encode(s));
    }
    std::vector<double> num;
    std::vector<uint8_t> isn;
    dict.numeric_table(&num, &isn);
    size_t n_num = 0;
    for (auto b : isn) n_num += b;
    std::printf("N %zu %zu\n", dict.id_to_string.size(), n_num);
    // FILTER compilation (types.rs:110-186 contract): the same expressions tests/test_cpp_host_cpu.py compiles with the Python mirror
    using FE = kolibrie::FilterExpression;
    const std::string lit = dict.id_to_string.empty() ? std::string("x") : dict.id_to_string[0];
    std::vector<FE> exprs = {
        FE::Cmp("?s", ">", "100000"),
        FE::AndOf(FE::Cmp("?s", ">=", "5."), FE::Cmp("?t", "=", lit)),
        FE::OrOf(FE::NotOf(FE::Cmp("?n", "!=", "a literal no triple mentions")), FE::Cmp("?s", "<", "abc")),
        FE::Cmp("?s", "<=", "?t"),
        FE::AndOf(FE::OrOf(FE::Cmp("t", "=", lit), FE::Cmp("?t", "!=", lit)), FE::NotOf(FE::Cmp("?s", ">", "-1e3"))),
    };
    for (size_t k = 0; k < exprs.size(); k++) {
        kolibrie::SlotMap sm;
        std::vector<kb_filter_op> ops;
        kolibrie::Condition{exprs[k]}.compile(exprs[k], sm, dict, &ops);
        for (auto& op : ops) std::printf("F %zu %u %u %u %u %a\n", k, op.op, op.slot, op.cmp, op.id, op.value);
    }
    return 0;
}
