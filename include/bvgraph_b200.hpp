// bvgraph_b200.hpp -- header-only C++ mirror of the reference's graph API over the C ABI (bvgraph_b200.h).
//
// Same surface as the reference's Java classes (reference src/it/unimi/dsi/webgraph/): ImmutableGraph
// (ImmutableGraph.java:254-420), BVGraph loaders (BVGraph.java:1380-1500), NodeIterator (BVGraph.java:1136-1281),
// LazyIntIterator (LazyIntIterator.java:28-44).  bvg_status codes become the exceptions the reference throws:
// IllegalArgumentException -> std::invalid_argument, IllegalStateException -> std::logic_error,
// UnsupportedOperationException -> bvg::unsupported_operation, NoSuchElementException -> std::out_of_range,
// IOException -> std::runtime_error.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "bvgraph_b200.h"

namespace webgraph {

struct unsupported_operation : std::runtime_error { using std::runtime_error::runtime_error; };

inline void check(int rc) {
    if (rc == BVG_OK) return;
    const std::string msg = bvg_strerror(rc);
    switch (rc) {
        case BVG_EINVAL: throw std::invalid_argument(msg);
        case BVG_ESTATE: throw std::logic_error(msg);
        case BVG_EUNSUPPORTED: throw unsupported_operation(msg);
        case BVG_EEND: throw std::out_of_range(msg);
        case BVG_ENOMEM: throw std::bad_alloc();
        default: throw std::runtime_error(msg);
    }
}

// LazyIntIterator over a decoded list: nextInt() returns the next successor, then -1 forever; skip(n) as in the reference.
class LazyIntIterator {
    std::vector<int32_t> a_;
    size_t i_ = 0;
public:
    explicit LazyIntIterator(std::vector<int32_t> a) : a_(std::move(a)) {}
    int32_t nextInt() { return i_ < a_.size() ? a_[i_++] : -1; }
    int32_t skip(int32_t n) { const size_t k = std::min<size_t>(n < 0 ? 0 : (size_t)n, a_.size() - i_); i_ += k; return (int32_t)k; }
};

class BVGraph;

// BVGraphNodeIterator: hasNext / nextInt / outdegree / successorArray / copy(upperBound).
class NodeIterator {
    friend class BVGraph;
    bvg_cursor* c_ = nullptr;
    int32_t from_, curr_, upper_, d_ = 0;
    const int32_t* succ_ = nullptr;
    NodeIterator(bvg_cursor* c, int32_t from, int32_t upper) : c_(c), from_(from), curr_(from - 1), upper_(upper) {}
public:
    NodeIterator(NodeIterator&& o) noexcept : c_(o.c_), from_(o.from_), curr_(o.curr_), upper_(o.upper_), d_(o.d_), succ_(o.succ_) { o.c_ = nullptr; }
    NodeIterator(const NodeIterator&) = delete;
    ~NodeIterator() { if (c_) bvg_cursor_close(c_); }
    bool hasNext() const { return curr_ < upper_ - 1; }
    int32_t nextInt() { int32_t node; check(bvg_cursor_next(c_, &node, &d_, &succ_)); return curr_ = node; }
    int32_t outdegree() const { if (curr_ == from_ - 1) throw std::logic_error("nextInt() never called"); return d_; }
    // valid until the next nextInt() on this iterator (BVGraph.java:1228-1233)
    const int32_t* successorArray() const { if (curr_ == from_ - 1) throw std::logic_error("nextInt() never called"); return succ_; }
    LazyIntIterator successors() const { return LazyIntIterator(std::vector<int32_t>(successorArray(), successorArray() + d_)); }
    NodeIterator copy(int32_t upperBound) const { bvg_cursor* n; check(bvg_cursor_copy(c_, upperBound, &n)); return NodeIterator(n, curr_ + 1, std::min(upperBound, upper_)); }
};

class ImmutableGraph {
public:
    virtual ~ImmutableGraph() = default;
    virtual int32_t numNodes() const = 0;
    virtual int64_t numArcs() const = 0;
    virtual bool randomAccess() const = 0;
    virtual int32_t outdegree(int32_t x) const = 0;
    virtual std::vector<int32_t> successorArray(int32_t x) const = 0;
    virtual LazyIntIterator successors(int32_t x) const { return LazyIntIterator(successorArray(x)); }
    virtual NodeIterator nodeIterator(int32_t from = 0) const = 0;
};

class BVGraph : public ImmutableGraph {
    std::shared_ptr<bvg_graph> g_;
    int32_t n_ = 0;
    int64_t m_ = 0;
    explicit BVGraph(bvg_graph* g) : g_(g, bvg_close) { check(bvg_info(g, &n_, &m_, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr)); }
    static BVGraph open(const std::string& basename, int offsetType) { bvg_graph* g; check(bvg_open(basename.c_str(), offsetType, nullptr, 0, &g)); return BVGraph(g); }
public:
    static BVGraph load(const std::string& basename) { return open(basename, 1); }
    static BVGraph loadMapped(const std::string& basename) { return open(basename, 2); }
    static BVGraph loadSequential(const std::string& basename) { return open(basename, 0); }
    static BVGraph loadOffline(const std::string& basename) { return open(basename, -1); }
    BVGraph copy() const { return *this; }  // flyweight: the native graph is immutable (ImmutableGraph.java:157-165)
    int32_t numNodes() const override { return n_; }
    int64_t numArcs() const override { return m_; }
    bool randomAccess() const override { return bvg_random_access(g_.get()) != 0; }
    int32_t outdegree(int32_t x) const override { int32_t d; check(bvg_outdegree(g_.get(), x, &d)); return d; }
    std::vector<int32_t> successorArray(int32_t x) const override {
        if (x < 0 || x >= n_) throw std::invalid_argument("Node index out of range");
        if (!randomAccess()) throw unsupported_operation("Random access to successor lists is not possible with sequential or offline graphs");
        std::vector<int32_t> out((size_t)outdegree(x));
        int32_t d = 0;
        check(bvg_successors(g_.get(), x, out.data(), (int32_t)out.size(), &d));
        return out;
    }
    NodeIterator nodeIterator(int32_t from = 0) const override { bvg_cursor* c; check(bvg_cursor_open(g_.get(), from, INT32_MAX, &c)); return NodeIterator(c, from, n_); }
    // bulk entry points
    std::pair<std::vector<int64_t>, std::vector<int32_t>> decodeRange(int32_t from, int32_t to) const {
        int64_t arcs; check(bvg_range_arcs(g_.get(), from, to, &arcs));
        std::vector<int64_t> off((size_t)(to - from) + 1);
        std::vector<int32_t> succ((size_t)arcs);
        check(bvg_decode_range(g_.get(), from, to, off.data(), succ.data(), arcs, 0));
        return { std::move(off), std::move(succ) };
    }
    std::pair<int64_t, uint64_t> scanRange(int32_t from, int32_t to) const { int64_t a; uint64_t c; check(bvg_scan_range(g_.get(), from, to, &a, &c)); return { a, c }; }
    bvg_graph* handle() const { return g_.get(); }
};

// EFGraph (reference EFGraph.java): the same surface over the bvg_ef_* entry points.  nodeIterator() is not offered through
// a cursor here (the reference inherits ImmutableGraph's generic iterator): use decodeRange for sequential access.
class EFGraph {
    std::shared_ptr<bvg_efgraph> g_;
    int32_t n_ = 0;
    int64_t m_ = 0;
    explicit EFGraph(bvg_efgraph* g) : g_(g, bvg_ef_close) { check(bvg_ef_info(g, &n_, &m_, nullptr, nullptr, nullptr)); }
public:
    static EFGraph load(const std::string& basename, int device = -1) { bvg_efgraph* g; check(bvg_ef_open(basename.c_str(), device, &g)); return EFGraph(g); }
    int32_t numNodes() const { return n_; }
    int64_t numArcs() const { return m_; }
    bool randomAccess() const { return true; }  // EFGraph.java:1045-1047
    int32_t outdegree(int32_t x) const { int32_t d; check(bvg_ef_outdegree(g_.get(), x, &d)); return d; }
    std::vector<int32_t> successorArray(int32_t x) const {
        std::vector<int32_t> out((size_t)outdegree(x));
        int32_t d = 0;
        check(bvg_ef_successors(g_.get(), x, out.data(), (int32_t)out.size(), &d));
        return out;
    }
    LazyIntIterator successors(int32_t x) const { return LazyIntIterator(successorArray(x)); }
    std::pair<std::vector<int64_t>, std::vector<int32_t>> decodeRange(int32_t from, int32_t to) const {
        int64_t arcs; check(bvg_ef_range_arcs(g_.get(), from, to, &arcs));
        std::vector<int64_t> off((size_t)(to - from) + 1);
        std::vector<int32_t> succ((size_t)arcs);
        check(bvg_ef_decode_range(g_.get(), from, to, off.data(), succ.data(), arcs, 0));
        return { std::move(off), std::move(succ) };
    }
    std::pair<int64_t, uint64_t> scanRange(int32_t from, int32_t to) const { int64_t a; uint64_t c; check(bvg_ef_scan_range(g_.get(), from, to, &a, &c)); return { a, c }; }
};

// The label stream of a BitStreamArcLabelledImmutableGraph over an open BVGraph (labelling/BitStreamArcLabelledImmutableGraph.java).
class ArcLabels {
    std::shared_ptr<bvg_labels> l_;
    BVGraph g_;   // keeps the underlying graph alive
public:
    ArcLabels(const BVGraph& g, const std::string& labelledBasename) : g_(g) {
        bvg_labels* l;
        check(bvg_labels_open(g.handle(), labelledBasename.c_str(), &l));
        l_.reset(l, bvg_labels_close);
    }
    static std::string underlyingBasename(const std::string& labelledBasename) {
        char buf[4096];
        check(bvg_labels_underlying(labelledBasename.c_str(), buf, (int)sizeof buf));
        return buf;
    }
    // labels of the arcs of [from, to) in successor order: (list offsets per arc, values)
    std::pair<std::vector<int64_t>, std::vector<int32_t>> decodeRange(int32_t from, int32_t to) const {
        int64_t arcs, nv = 0;
        check(bvg_range_arcs(g_.handle(), from, to, &arcs));
        check(bvg_labels_decode_range(l_.get(), from, to, nullptr, nullptr, 0, 0, &nv));
        std::vector<int64_t> lo((size_t)arcs + 1);
        std::vector<int32_t> vals((size_t)nv);
        check(bvg_labels_decode_range(l_.get(), from, to, lo.data(), vals.data(), nv, 0, &nv));
        return { std::move(lo), std::move(vals) };
    }
    std::vector<int32_t> labelArray(int32_t x) const { return decodeRange(x, x + 1).second; }
};

}  // namespace webgraph
