/*
 * bvgraph_tools.h -- host-side (CPU) tools that PRODUCE BVGraph inputs: a compressor and a seeded
 * synthetic power-law generator.  C ABI of libbvgraph_tools.so.
 *
 * These are the cold side of the format: the reference's BVGraph.store /
 * CompressionThread.diffComp / intervalize (reference src/it/unimi/dsi/webgraph/BVGraph.java:
 * 1631-1654, 2049-2219, 2221-2386, 2436-2650) re-done natively because no JVM exists in the image,
 * so that tests and bench.py can make .graph/.offsets/.properties files.  They never decode: the
 * decode path is CUDA-only (include/bvgraph_b200.h).
 */
#ifndef BVGRAPH_TOOLS_H
#define BVGRAPH_TOOLS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Coding ids, reference CompressionFlags.java:26-44; flag word layout BVGraph.java:474-523, 1317-1325. */
enum { BVGT_DELTA = 1, BVGT_GAMMA = 2, BVGT_GOLOMB = 3, BVGT_SKEWED_GOLOMB = 4, BVGT_UNARY = 5,
       BVGT_ZETA = 6, BVGT_NIBBLE = 7 };

typedef struct bvgt_store_stats {
    int64_t nodes, arcs;
    int64_t graph_bits, offsets_bits;
    int64_t bits_outdegrees, bits_references, bits_blocks, bits_intervals, bits_residuals;
    int64_t copied_arcs, intervalised_arcs, residual_arcs;
    int64_t tot_ref, tot_dist;      /* sums behind avgref / avgdist (BVGraph.java:2336-2338) */
    int32_t max_outdegree, max_ref_chain;
    uint64_t xor_checksum;          /* XOR over arcs of (x*0x9E3779B97F4A7C15 + y) */
    uint64_t sum_successors;
} bvgt_store_stats;

/* BVGraph.store(graph, basename, windowSize, maxRefCount, minIntervalLength, zetaK, flags, threads)
 * (BVGraph.java:1679-1688, 2436-2650) for a graph given as CSR (off[n+1], succ[off[n]], each list strictly
 * increasing).  threads > 1 splits nodes into ranges compressed independently with an empty window each and
 * concatenated bit-exactly, as the reference does (:2471-2477, 2498-2550).  maxref < 0 means unbounded (-m -1).
 * Writes <basename>.graph/.offsets/.properties.  Returns 0 or a negative error (-1 invalid argument incl.
 * non-increasing list (BVGraph.java:2201), -4 I/O). */
int bvgt_store_csr(const char* basename, int32_t n, const int64_t* off, const int32_t* succ,
                   int32_t window, int32_t maxref, int32_t minlen, int32_t zetak, uint32_t flags,
                   int threads, bvgt_store_stats* stats);

/* Writes `count` values back to back in one of the instantaneous codes of dsiutils' OutputBitStream (coding ids above;
 * k = shrinking factor of zeta, modulus of Golomb, ignored otherwise), MSB first, zero-padded to a byte.  Returns the
 * number of bits written, -1 on a bad argument or when `cap` bytes do not suffice, -3 for skewed Golomb (which the
 * reference's BVGraph cannot read either).  For code-level tests of the readers. */
int64_t bvgt_write_codes(int coding, int32_t k, const uint64_t* values, int64_t count, uint8_t* out, int64_t cap);

typedef struct bvgt_gen_params {
    int32_t  n;            /* nodes */
    int64_t  target_arcs;  /* approximate number of arcs wanted */
    uint64_t seed;
    double   zipf_s;       /* exponent of the rank-size outdegree law, genzipf-style floor((n/r)^s) (reference c/genzipf.c:22-26) */
    double   p_copy;       /* probability that a node copies part of a prototype x-r, r in [1,7], same block */
    double   p_interval;   /* probability that a node carries runs of consecutive successors */
    double   p_local;      /* fraction of the remaining successors drawn near x instead of globally */
    int32_t  block;        /* nodes per generation block ("host"): prototypes never cross a block */
    int32_t  max_degree;   /* cap on the outdegree law (default 2^22; experiments use smaller caps) */
    /* shape of the copied part and of the local successors (defaults reproduce the round-1 benchmark graph) */
    double   copy_run;     /* mean length of a copied run of the prototype (default 8) */
    double   skip_run;     /* mean length of a skipped run (default 3) */
    int32_t  local_bits;   /* local successors sit at a log-uniform distance below 2^local_bits (default 16) */
    int32_t  interval_max; /* intervals per node: 1 .. interval_max (default 3) */
    double   p_same_degree;/* probability that a copying node takes its prototype's outdegree (+ 0..3) instead of the law's: pages of a site (default 0) */
} bvgt_gen_params;

void bvgt_gen_defaults(bvgt_gen_params* p, int32_t n, int64_t target_arcs, uint64_t seed);

/* Generates the graph and compresses it straight to <basename>.* with `threads` workers (node ranges aligned
 * to generation blocks).  If out_off/out_succ are non-NULL the CSR is also returned (out_off: n+1 entries;
 * out_succ: capacity succ_cap entries; -1 if it does not fit).  The graph is a function of the parameters
 * only, not of `threads`; the encoding (reference choices at range starts) depends on `threads`. */
int bvgt_generate_store(const char* basename, const bvgt_gen_params* p,
                        int32_t window, int32_t maxref, int32_t minlen, int32_t zetak, uint32_t flags,
                        int threads, int64_t* out_off, int32_t* out_succ, int64_t succ_cap,
                        bvgt_store_stats* stats);

/* Arc labels: the three Label classes of the reference that a BitStreamArcLabelledImmutableGraph can carry.
 * GAMMA = GammaCodedIntLabel (toBitStream = writeGamma(value), GammaCodedIntLabel.java:58-61), FIXED = FixedWidthIntLabel
 * (writeInt(value, width), FixedWidthIntLabel.java:75-78), FIXED_LIST = FixedWidthIntListLabel (writeGamma(length) then
 * length x writeInt(element, width), FixedWidthIntListLabel.java:80-85). */
enum { BVGT_LABEL_GAMMA = 0, BVGT_LABEL_FIXED = 1, BVGT_LABEL_FIXED_LIST = 2 };

/* Writes <basename>.labels (the labels of node 0's arcs in successor order, then node 1's, ... with no separators),
 * <basename>.labeloffsets (gamma(0), then for every node the gamma-coded number of bits its labels took) and
 * <basename>.properties (graphclass, labelspec = <class>(<key>[,<width>]), underlyinggraph = `underlying`), the layout
 * BitStreamArcLabelledImmutableGraph.load reads (BitStreamArcLabelledImmutableGraph.java:385-470; written the way
 * BitStreamArcLabelledGraphTest.java:131-203 writes its fixtures).  off[n+1] = CSR row offsets of the underlying graph.
 * GAMMA / FIXED: values[off[n]] holds one label per arc, list_off is ignored.  FIXED_LIST: arc j carries
 * values[list_off[j] .. list_off[j+1]).  Values must be >= 0 and, for the fixed-width kinds, < 2^width (0 <= width <= 31).
 * label_bits (may be NULL) receives the length of the label stream.  Returns 0, -1 (bad argument / value), -4 (I/O). */
int bvgt_store_labels(const char* basename, const char* underlying, const char* key, int32_t n, const int64_t* off,
                      const int64_t* list_off, const int32_t* values, int kind, int width, int threads, int64_t* label_bits);

/* EFGraph.store(graph, upperBound, basename, log2Quantum, cacheSize, byteOrder, pl) (reference EFGraph.java:812-888): the
 * quasi-succinct format -- per node gamma(outdegree) (LSB-first long-word stream, LongWordOutputBitStream :298-418), then the
 * Elias-Fano encoding of the successors plus the terminator upperBound (Accumulator :420-556): skip pointers to zeros
 * (numberOfPointers x pointerSize bits), lower bits ((outdegree + 1) x l), upper bits in unary.  Writes <basename>.graph (long
 * words in the given byte order, one trailing word as LongWordOutputBitStream.close() does), .offsets (delta-coded gaps,
 * MSB-first OutputBitStream) and .properties.  upper_bound <= 0 means n.  Returns 0, -1 (bad argument), -4 (I/O); graph_bits
 * (may be NULL) receives the bits written before the final padding. */
int bvgt_store_ef(const char* basename, int32_t n, const int64_t* off, const int32_t* succ, int32_t upper_bound,
                  int log2_quantum, int big_endian, int threads, int64_t* graph_bits);

#ifdef __cplusplus
}
#endif
#endif
