// Exercises include/bvgraph_b200.hpp (the C++ mirror of ImmutableGraph/NodeIterator/LazyIntIterator) on cnr-2000.
// Built and run by tests/test_gpu_parity.py::test_cpp_mirror on the GPU box; prints arcs and the XOR checksum.
#include <algorithm>
#include <cstdio>
#include "bvgraph_b200.hpp"
int main(int argc, char** argv) {
    using namespace webgraph;
    BVGraph g = BVGraph::load(argv[1]);
    uint64_t cs = 0;
    int64_t arcs = 0;
    NodeIterator it = g.nodeIterator();
    while (it.hasNext()) {
        const int32_t x = it.nextInt();
        const int32_t d = it.outdegree();
        const int32_t* s = it.successorArray();
        for (int32_t j = 0; j < d; j++) cs ^= (uint64_t)(uint32_t)x * 0x9E3779B97F4A7C15ull + (uint64_t)(uint32_t)s[j];
        arcs += d;
    }
    LazyIntIterator li = g.successors(0);
    int32_t first = li.nextInt();
    for (int i = 0; i < 20; i++) li.nextInt();
    bool caught = false;
    try { g.outdegree(g.numNodes()); } catch (const std::invalid_argument&) { caught = true; }
    auto sc = g.scanRange(0, g.numNodes());
    std::printf("%lld %llx %d %d %d %lld %llx\n", (long long)arcs, (unsigned long long)cs, first, li.nextInt(), (int)caught,
                (long long)sc.first, (unsigned long long)sc.second);
    if (argc > 3) {  // argv[2]: the same graph stored as an EFGraph; argv[3]: gamma labels over argv[1] (label of arc j = j % 1000)
        EFGraph e = EFGraph::load(argv[2]);
        auto es = e.scanRange(0, e.numNodes());
        const bool same0 = e.successorArray(0) == g.successorArray(0) && e.outdegree(7) == g.outdegree(7);
        ArcLabels labels(g, argv[3]);
        auto lab = labels.decodeRange(0, g.numNodes());
        bool ok = (int64_t)lab.second.size() == g.numArcs() && ArcLabels::underlyingBasename(argv[3]) == std::string(argv[1]);
        for (size_t j = 0; j < lab.second.size() && ok; j += 997) ok = lab.second[j] == (int32_t)(j % 1000);
        std::printf("%lld %llx %d %d\n", (long long)es.first, (unsigned long long)es.second, (int)same0, (int)ok);
    }
    return 0;
}
