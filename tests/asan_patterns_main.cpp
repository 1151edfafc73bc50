// Test harness (tests/test_host_cpu.py::test_expand_patterns_under_asan): the product's pattern table producer
// (teloscope_b200/csrc/patterns.cpp) compiled with -fsanitize=address and run on one seed set.
//   argv: edit_distance canonical_fwd seed [seed ...]   ->  one "PATTERN is_forward is_canonical" line per pattern
#include <cstdio>
#include <cstdlib>

#include "patterns.hpp"

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    std::vector<std::string> raw;
    for (int i = 3; i < argc; ++i) raw.push_back(argv[i]);
    std::vector<tls::Pattern> out;
    std::string err;
    if (!tls::expandPatterns(raw, std::atoi(argv[1]), argv[2], out, err)) {
        std::printf("ERR %s\n", err.c_str());
        return 1;
    }
    for (const tls::Pattern &p : out) std::printf("%s %d %d\n", p.text.c_str(), int(p.isForward), int(p.isCanonical));
    return 0;
}
