/*
 * bvg_oracle.h -- CPU oracle for the BVGraph adjacency decoder.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * decode path (vigna/webgraph, Java) used as the checker for the CUDA path.
 * Nothing in the product (webgraph_b200/, include/) links, imports or executes
 * it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do.
 *
 * Parity status: PINNED for the default codings (gamma outdegrees / blocks /
 * block counts / offsets, unary references, zeta_k residuals, intervals) by the
 * reference's own golden pair cnr-2000.{graph,offsets,properties} <->
 * cnr-2000.graph-txt.gz, the pair BVGraphTest.testLarge asserts equal
 * (reference test/it/unimi/dsi/webgraph/BVGraphTest.java:101-119).
 * UNPINNED ("parity unpinned") for the non-default codings (delta anywhere,
 * gamma / Golomb / nibble residuals, gamma references, unary block counts,
 * zeta_k with k != 3): no reference test or fixture covers them and no JVM
 * exists in this image to make one.  Skewed Golomb has no reader in the
 * reference (BVGraph.java:791-816) and is rejected (BVGO_EUNSUPPORTED).
 *
 * The bit-level codes live in dsiutils (it.unimi.dsi:dsiutils, pinned only as
 * `latest.release` in the reference's ivy.xml:18, not vendored); they are
 * restated from their published definition and anchored on the call sites
 * BVGraph.java:631-816 and on the golden pair above.
 */
#ifndef BVG_ORACLE_H
#define BVG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    BVGO_OK = 0,
    BVGO_EINVAL = -1,       /* IllegalArgumentException  (BVGraph.java:860,900,1037,1165) */
    BVGO_ESTATE = -2,       /* IllegalStateException     (BVGraph.java:705,869)           */
    BVGO_EUNSUPPORTED = -3, /* UnsupportedOperationException (BVGraph.java:901, coders)   */
    BVGO_EIO = -4,          /* IOException / truncated stream (BVGraph.java:876,1131)      */
    BVGO_EFORMAT = -5,      /* malformed .properties / impossible record                   */
    BVGO_ENOMEM = -6
};

/* CompressionFlags.java:26-44 */
enum { BVGO_DELTA = 1, BVGO_GAMMA = 2, BVGO_GOLOMB = 3, BVGO_SKEWED_GOLOMB = 4,
       BVGO_UNARY = 5, BVGO_ZETA = 6, BVGO_NIBBLE = 7 };

typedef struct orc_graph {
    int32_t  n;            /* nodes              (BVGraph.java:1536-1538) */
    int64_t  m;            /* arcs               (:1539) */
    int32_t  window;       /* windowsize         (:1540) */
    int32_t  maxref;       /* maxrefcount        (:1541) */
    int32_t  minlen;       /* minintervallength  (:1542), 0 = no intervals */
    int32_t  zetak;        /* zetak              (:1543), default 3 */
    uint32_t flags;        /* compressionflags   (:1532, 1317-1325) */
    int outdegree_coding, block_coding, residual_coding, reference_coding,
        block_count_coding, offset_coding;
    uint8_t*  graph;       /* .graph bytes followed by 16 zero bytes of padding */
    uint64_t  graph_bytes; /* file size */
    uint64_t* offsets;     /* n+1 bit offsets, or NULL when not loaded */
} orc_graph;

/* Mirrors BVGraph.loadInternal (BVGraph.java:1516-1609); load_offsets != 0 <=> offsetType > 0. */
int  orc_load(const char* basename, int load_offsets, orc_graph** out);
/* Builds a graph from in-memory buffers (copies them); offsets may be NULL. */
int  orc_from_memory(const uint8_t* graph, uint64_t graph_bytes, const uint64_t* offsets,
                     int32_t n, int64_t m, int32_t window, int32_t maxref, int32_t minlen,
                     int32_t zetak, uint32_t flags, orc_graph** out);
void orc_free(orc_graph* g);

/* BVGraph.outdegree(x), BVGraph.java:857-879. */
int  orc_outdegree(const orc_graph* g, int32_t x, int32_t* d);

/* BVGraph.successors(x) drained into out[0..d) (random access, recursive along the
 * reference chain, BVGraph.java:896-904 + 1032-1133).  Returns d (>= 0) or an error. */
int64_t orc_successors(const orc_graph* g, int32_t x, int32_t* out, int64_t cap);

/* BVGraph.nodeIterator(from) drained for nodes [from, to): sequential decode with the cyclic
 * window of W+1 lists, the window being seeded by random access when from > 0
 * (BVGraph.java:1136-1213).  out_off has to-from+1 entries (out_off[0] = 0).
 * out may be NULL (count only); returns total arcs or an error.  When offsets were not
 * loaded from must be 0 (BVGraph.java:1174). */
int64_t orc_decode_range(const orc_graph* g, int32_t from, int32_t to,
                         int64_t* out_off, int32_t* out, int64_t cap);

/* Consume-only scan of [from, to) (what SpeedTest's sequential loop does,
 * src/it/unimi/dsi/webgraph/test/SpeedTest.java:157-185) plus an order-independent checksum:
 * XOR over arcs (x,y) of (x * 0x9E3779B97F4A7C15 + y) mod 2^64. */
int  orc_scan_range(const orc_graph* g, int32_t from, int32_t to, int64_t* arcs, uint64_t* checksum);

/* Re-derives the n+1 offsets by a sequential pass over .graph alone (what
 * BVGraph.writeOffsets does, BVGraph.java:2662-2676). out has n+1 entries. */
int  orc_rebuild_offsets(const orc_graph* g, uint64_t* out);

/* Algorithmic bits needed to answer successors(x) by random access: the record of x plus the
 * records of its reference-chain ancestors. Returns bits or an error. */
int64_t orc_chain_bits(const orc_graph* g, int32_t x, int32_t* depth);
/* Root of x's reference chain (x itself when it has no reference); negative = error. */
int32_t orc_chain_root(const orc_graph* g, int32_t x);

/* Raw code readers, exported for known-answer tests of the dsiutils restatement. */
uint64_t orc_read_code(const uint8_t* buf, uint64_t nbytes, uint64_t* bitpos, int coding, int k);

/* ---- arc labels (labelling/BitStreamArcLabelledImmutableGraph.java) -- PARITY UNPINNED: no .labels fixture exists in the
 * reference tree; see bvg_oracle.c. ---- */
enum { ORC_LABEL_GAMMA = 0, ORC_LABEL_FIXED = 1, ORC_LABEL_FIXED_LIST = 2 };
typedef struct orc_labels {
    int kind, width;
    int32_t n;
    uint8_t* labels;        /* .labels bytes + 16 bytes of padding */
    uint64_t label_bytes;
    uint64_t* offsets;      /* n+1 bit offsets */
} orc_labels;
int  orc_labels_load(const char* basename, int32_t n, orc_labels** out);
void orc_labels_free(orc_labels* l);
int64_t orc_labels_node(const orc_labels* l, int32_t x, int32_t d, int64_t* list_off, int32_t* values, int64_t cap);
int64_t orc_labels_range(const orc_labels* l, int32_t from, int32_t to, const int64_t* row_off, int32_t* values, int64_t cap, uint64_t* sum);

/* ---- EFGraph (EFGraph.java) -- PARITY UNPINNED: no EFGraph fixture exists in the reference tree; see bvg_oracle.c. ---- */
typedef struct orc_efgraph {
    int32_t n;
    int64_t m;
    int32_t upper_bound, log2_quantum;
    uint64_t* words;        /* the .graph long words in host order + two zero words */
    uint64_t nwords;
    uint64_t* offsets;      /* n+1 bit offsets */
} orc_efgraph;
int  orc_ef_load(const char* basename, orc_efgraph** out);
void orc_ef_free(orc_efgraph* g);
int  orc_ef_outdegree(const orc_efgraph* g, int32_t x, int32_t* d);
int64_t orc_ef_successors(const orc_efgraph* g, int32_t x, int32_t* out, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif
