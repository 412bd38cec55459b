/*
 * bvgraph_b200.h -- C ABI of libbvgraph_b200.so: B200-native BVGraph adjacency decode.
 *
 * This is the drop-in boundary: what a JNI / cffi / ctypes shim binds in place of the
 * reference's Java decode path.  Plain pointers and sizes only.  Each entry point cites the reference
 * interface it replaces (paths relative to the reference root, src/it/unimi/dsi/webgraph/).
 *
 * The .graph bit stream, the offsets and a small decode index live in HBM; all decoding is done by
 * hand-written sm_100a kernels.  There is NO CPU decode path: without a CUDA device every compute
 * call fails with BVG_ECUDA.
 *
 * Conventions: every function returns BVG_OK (0) or a negative bvg_status; out-params are untouched
 * on error; the caller owns every buffer it passes.  `on_device` != 0 means the out / xs pointers
 * are device pointers on the graph's device (e.g. torch tensors' data_ptr()), and the call is
 * asynchronous on the graph's stream (bvg_set_stream); otherwise they are host pointers and the
 * call returns after the results have been copied back.
 * Threading (reference contract: an ImmutableGraph instance is not thread-safe, copy() is, and copies share the
 * immutable data, ImmutableGraph.java:157-165,411-420): the data of a bvg_graph is immutable after open and shared
 * by everything opened on it.  The entry points that take a bvg_graph run on the graph's one stream and report into
 * its one error word; they lock the graph, so concurrent callers are serialised, never corrupted (a binding's copy()
 * may hand out the same handle).  A bvg_cursor (== BVGraphNodeIterator) owns a stream, an error word and its batch
 * buffers: cursors of one graph run concurrently, one per host thread, and that is the way to scale the sequential
 * route across host threads (ImmutableGraph.splitNodeIterators).  A single cursor is single-threaded.
 */
#ifndef BVGRAPH_B200_H
#define BVGRAPH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bvg_graph  bvg_graph;
typedef struct bvg_cursor bvg_cursor;

enum bvg_status {
    BVG_OK = 0,
    BVG_EINVAL = -1,        /* IllegalArgumentException: node out of range (BVGraph.java:860,900,1037,1165) */
    BVG_ESTATE = -2,        /* IllegalStateException: no offsets (:869,1174), ref > window (:705), cursor before next (:1222) */
    BVG_EUNSUPPORTED = -3,  /* UnsupportedOperationException: random access without offsets (:901); unsupported coding */
    BVG_EIO = -4,           /* IOException -> RuntimeException: missing/truncated file or stream (:876,1131) */
    BVG_EFORMAT = -5,       /* malformed .properties, wrong graphclass/version (:1528-1534), impossible record */
    BVG_ENOMEM = -6,        /* host or device allocation failed / caller buffer too small */
    BVG_ECUDA = -7,         /* no CUDA device / CUDA runtime error */
    BVG_EEND = -8           /* NoSuchElementException: cursor past the end (:1202) */
};

/* ---- loading: BVGraph.load / loadMapped / loadOffline -> loadInternal (BVGraph.java:1380-1500, 1516-1609) ----
 * offset_type as in BVGraph (:439-441): 2 mapped, 1 standard, 0 sequential (no random access), -1 offline.
 * On the GPU all four place the bit stream and offsets in HBM; offset_type <= 0 only switches the random-access
 * entry points to the reference's errors.  As in the reference, offset_type <= 0 does not need the .offsets file
 * (loadInternal never opens it, :1581-1609; the iterator just keeps reading, :1201-1213): when it is missing the record
 * boundaries are found from the .graph stream itself, in parallel on the device (what BVGraph.writeOffsets, :2662-2676,
 * does with one sequential pass).  offset_type > 0 without .offsets is BVG_EIO (FileNotFoundException).
 * devices/ndev: CUDA ordinals to use (NULL/0 = current device); the
 * first one holds the graph (multi-GPU runs open one shard per process, see bvg_open_shard). */
int  bvg_open(const char* basename, int offset_type, const int* devices, int ndev, bvg_graph** out);

/* Opens only nodes [from, to) of the graph for range-sharded scans (the reference's own range split is
 * ImmutableGraph.splitNodeIterators, ImmutableGraph.java:379-409).  Loads the bits of [from - halo, to) where the
 * halo covers the reference chains leaving the shard (window * maxrefcount nodes, as BVGraphNodeIterator's ctor
 * re-reads the window before `from`, BVGraph.java:1173-1183). Node ids stay global. */
int  bvg_open_shard(const char* basename, int device, int32_t from, int32_t to, bvg_graph** out);

/* Opens a graph held in host memory: `graph` is the .graph byte stream, `offsets_stream` the .offsets byte stream
 * (gamma/delta coded gaps).  offsets_stream may be NULL when offset_type <= 0: the record boundaries are then found
 * from the graph stream on the device (see bvg_open); NULL with offset_type > 0 is BVG_EINVAL. */
int  bvg_open_memory(const uint8_t* graph, uint64_t graph_bytes, const uint8_t* offsets_stream, uint64_t offsets_bytes,
                     int32_t nodes, int64_t arcs, int32_t window, int32_t maxref, int32_t minlen, int32_t zetak,
                     uint32_t flags, int offset_type, int device, bvg_graph** out);
/* Same, holding only nodes [from, to) (+ halo) on the device: what a rank of a multi-GPU run opens when the file is
 * already in (pinned) host memory. */
int  bvg_open_memory_shard(const uint8_t* graph, uint64_t graph_bytes, const uint8_t* offsets_stream, uint64_t offsets_bytes,
                           int32_t nodes, int64_t arcs, int32_t window, int32_t maxref, int32_t minlen, int32_t zetak,
                           uint32_t flags, int offset_type, int device, int32_t from, int32_t to, bvg_graph** out);
/* Host-only: bounds[0..nshards] of contiguous node ranges holding equal shares of the .graph bits, every cut moved to
 * the nearest node no reference crosses (within 4096 nodes) so that the shards owe each other no boundary lists.  The range
 * split itself is ImmutableGraph.splitNodeIterators' (ImmutableGraph.java:379-409). */
int  bvg_plan_shards(const char* basename, int nshards, int32_t* bounds);
/* The same from measured costs: old_cost[j] = what shard [old_bounds[j], old_bounds[j+1]) cost (e.g. device seconds per scan);
 * the new cuts sit at equal shares of that cost, taken as uniform over the bits of each old shard. */
int  bvg_replan_shards(const char* basename, int nshards, const int32_t* old_bounds, const double* old_cost, int32_t* bounds);
void bvg_close(bvg_graph* g);

/* numNodes / numArcs / windowSize / maxRefCount / minIntervalLength / zetaK / flags (BVGraph.java:579-625). */
int  bvg_info(const bvg_graph* g, int32_t* nodes, int64_t* arcs, int32_t* window, int32_t* maxref,
              int32_t* minlen, int32_t* zetak, uint32_t* flags, int64_t* graph_bits);
/* Shard extent [from, to) (whole graph: 0, n), longest reference chain and largest outdegree found at open. */
int  bvg_extent(const bvg_graph* g, int32_t* from, int32_t* to, int32_t* max_chain, int32_t* max_outdegree);
/* randomAccess() (BVGraph.java:592-594). */
int  bvg_random_access(const bvg_graph* g);
/* CUDA stream (cudaStream_t) the graph's kernels and copies are issued on; NULL = default stream. */
int  bvg_set_stream(bvg_graph* g, void* cuda_stream);
int  bvg_device(const bvg_graph* g);

/* ---- random access: BVGraph.outdegree(x) :857-879, successors(x) :896-904, successorArray(x) ImmutableGraph.java:329-333 ---- */
int  bvg_outdegree(const bvg_graph* g, int32_t x, int32_t* d);
int  bvg_successors(const bvg_graph* g, int32_t x, int32_t* out, int32_t cap, int32_t* d);
/* Batched successors(x) for nx nodes: out_off[nx+1] (arc offsets into out, out_off[0]=0), out[cap].
 * If out == NULL only out_off is produced (sizing call). */
int  bvg_successors_batch(const bvg_graph* g, const int32_t* xs, int64_t nx, int64_t* out_off,
                          int32_t* out, int64_t cap, int on_device);
/* Batched outdegree(x): d[i] = outdegree(xs[i]); xs == NULL means the node range [from, from+nx). */
int  bvg_outdegree_batch(const bvg_graph* g, const int32_t* xs, int32_t from, int64_t nx, int32_t* d, int on_device);

/* ---- sequential access: BVGraph.nodeIterator(from) :1292-1301 drained over [from, to) ----
 * out_off[to-from+1], out[cap]: the successor lists of from..to-1 back to back (what nextInt()+successorArray() yield). */
int  bvg_range_arcs(const bvg_graph* g, int32_t from, int32_t to, int64_t* arcs);
int  bvg_decode_range(const bvg_graph* g, int32_t from, int32_t to, int64_t* out_off, int32_t* out, int64_t cap,
                      int on_device);
/* Consume-only scan of [from, to), the loop of the reference's SpeedTest (test/SpeedTest.java:157-185):
 * every successor is decoded on the GPU and folded into arcs and checksum = XOR over arcs (x,y) of
 * (x * 0x9E3779B97F4A7C15 + y) mod 2^64; no successor array is written to HBM by the caller. */
int  bvg_scan_range(const bvg_graph* g, int32_t from, int32_t to, int64_t* arcs, uint64_t* checksum);
/* The same over a graph held in HOST memory that is used once (the reference's offline / sequential access,
 * BVGraph.loadOffline / loadSequential, BVGraph.java:1380-1500, scanned as by test/SpeedTest.java:157-185): nothing stays
 * on the device.  The node range is cut into `pieces` bit-balanced pieces; while the device indexes and scans piece p,
 * the bytes of piece p + 1 cross PCIe (give pinned memory for that overlap).  pieces = 1 is open + scan + close.
 * [from, to) is the node range to scan (a shard of a multi-GPU scan, or 0, nodes).  offsets_stream is required here
 * (a graph without .offsets is opened with bvg_open / bvg_open_memory and scanned with bvg_scan_range). */
int  bvg_scan_memory(const uint8_t* graph, uint64_t graph_bytes, const uint8_t* offsets_stream, uint64_t offsets_bytes,
                     int32_t nodes, int64_t arcs, int32_t window, int32_t maxref, int32_t minlen, int32_t zetak,
                     uint32_t flags, int device, int32_t from, int32_t to, int pieces, int64_t* arcs_out, uint64_t* checksum_out);
/* Asynchronous variant for benchmarking: enqueues the scan on the graph's stream and leaves
 * {arcs, checksum} in the two-word device buffer d_result (int64, uint64). */
int  bvg_scan_range_async(const bvg_graph* g, int32_t from, int32_t to, void* d_result);

/* ---- fused consumers of the decode path: the successors never leave the device ----
 * bvg_indegrees: the counting pass of a transposition (Transform.java:977-987, numPred[a[d]]++): counts[y] += number of arcs
 * (x, y) with x in [from, to).  counts holds counts_len entries (host, or device with on_device); successors >= counts_len are
 * ignored.  *arcs = arcs scanned.  Ranks of a sharded graph add their counts up (an all-reduce).
 * bvg_bfs: breadth-first visit from `source` (algo/ParallelBreadthFirstVisit.java:155-181): dist[x] = distance, -1 when
 * unreachable (n entries); *levels = eccentricity of the source, *reached = nodes reached.  Whole graphs with offsets only. */
int  bvg_indegrees(const bvg_graph* g, int32_t from, int32_t to, uint32_t* counts, int64_t counts_len, int on_device, int64_t* arcs);
int  bvg_bfs(const bvg_graph* g, int32_t source, int32_t* dist, int on_device, int32_t* levels, int64_t* reached);
/* One HyperBall iteration over nodes [from, to) (reference algo/HyperBall.java:875-915, the branch in which every node
 * enumerates its successors): out[x] = register-wise max of in[x] and in[s] for every successor s of x.  A counter is
 * 2^log2m registers of ONE BYTE each (4 <= log2m <= 9; the reference packs 5-7-bit registers into longs), counters of all
 * numNodes() nodes back to back in `in` / `out` (distinct buffers, 16-byte aligned; out is written for [from, to) only).
 * modified (may be NULL) receives the number of nodes whose counter changed (the reference's modified-counter count).
 * Rows are decoded on the device chunk by chunk and consumed there.  on_device != 0: device pointers, stream-ordered. */
int  bvg_hyperball_step(const bvg_graph* g, int32_t from, int32_t to, int log2m, const uint8_t* in, uint8_t* out, int on_device,
                        int64_t* modified);

/* ---- NodeIterator: BVGraphNodeIterator :1136-1281 (nextInt, outdegree, successorArray, copy(upperBound)) ---- */
int  bvg_cursor_open(const bvg_graph* g, int32_t from, int32_t upper, bvg_cursor** out);
/* succ stays valid until the next call on this cursor (NodeIterator.successorArray() aliasing, BVGraph.java:1228-1233). */
int  bvg_cursor_next(bvg_cursor* c, int32_t* node, int32_t* d, const int32_t** succ);
/* The whole batch that holds the next node, zero-copy: nodes first .. first + count - 1, their successors are
 * succ[off[i] .. off[i + 1]) for i in [0, count) (pinned host memory owned by the cursor, valid until the next call on it).
 * The cursor moves past the batch.  A binding wraps the two arrays (JNI NewDirectByteBuffer, numpy) and iterates without
 * further native calls; batches are decoded on the device one ahead of the caller. */
int  bvg_cursor_next_batch(bvg_cursor* c, int32_t* first, int32_t* count, const int64_t** off, const int32_t** succ);
int  bvg_cursor_copy(const bvg_cursor* c, int32_t upper, bvg_cursor** out);
void bvg_cursor_close(bvg_cursor* c);
/* The inner loop of a binding, in C: up to max_nodes calls of bvg_cursor_next (all that is left when max_nodes < 0),
 * every successor consumed into arcs and the checksum of bvg_scan_range -- what the reference's SpeedTest does with
 * nodeIterator() (test/SpeedTest.java:157-185).  Measures the NodeIterator route end to end (device decode of batches,
 * copies to pinned host memory, host-side iteration) without a per-node FFI call.  Returns BVG_OK also at the end. */
int  bvg_cursor_drain(bvg_cursor* c, int64_t max_nodes, int64_t* nodes, int64_t* arcs, uint64_t* checksum);

/* ---- range sharding across GPUs ----
 * A shard's first nodes may copy from lists of the previous shard.  Either the shard re-decodes that halo from its
 * own replicated bits (default, what BVGraphNodeIterator's ctor does), or the previous shard exports its last
 * boundary lists and they are all-gathered (NCCL) and imported:
 * bvg_boundary_count = number of trailing nodes whose lists a successor shard can reference (<= window*maxref);
 * bvg_boundary_export decodes them into out_off[count+1] / out[cap] (device pointers);
 * bvg_halo_import installs `count` lists ending at node from-1 so that scans/decodes of this shard use them
 * instead of re-decoding.  bvg_halo_needed tells whether any chain actually crosses the shard's start. */
int  bvg_boundary_count(const bvg_graph* g, int32_t* count);
int  bvg_boundary_export(const bvg_graph* g, int64_t* out_off, int32_t* out, int64_t cap, int on_device);
int  bvg_halo_needed(const bvg_graph* g, int32_t* first_needed_node);
int  bvg_halo_import(bvg_graph* g, int32_t count, const int64_t* off, const int32_t* lists, int on_device);

/* ---- arc labels: BitStreamArcLabelledImmutableGraph ----
 * A labelled graph is <basename>.properties (graphclass, labelspec, underlyinggraph), <basename>.labels (the labels of
 * node 0's arcs in successor order, then node 1's, ... written by Label.toBitStream with no separators) and
 * <basename>.labeloffsets (gamma-coded gaps) over an underlying BVGraph (reference labelling/
 * BitStreamArcLabelledImmutableGraph.java:139-145, 385-470).  The three Label classes the reference ships are decoded on
 * the device: GammaCodedIntLabel (readGamma, GammaCodedIntLabel.java:52-56), FixedWidthIntLabel (readInt(width),
 * FixedWidthIntLabel.java:69-73), FixedWidthIntListLabel (readGamma length, then length x readInt(width),
 * FixedWidthIntListLabel.java:72-78).  Anything else in labelspec is BVG_EUNSUPPORTED.
 * A bvg_labels is opened ON a bvg_graph (the underlying graph, or a shard of it: only the shard's stretch of the label
 * stream is loaded), uses its device, stream and call lock, and must be closed before it. */
typedef struct bvg_labels bvg_labels;
enum { BVG_LABEL_GAMMA = 0, BVG_LABEL_FIXED = 1, BVG_LABEL_FIXED_LIST = 2 };

/* Resolves the underlyinggraph property of <basename>.properties against the labelled graph's directory
 * (BitStreamArcLabelledImmutableGraph.java:391-395) into buf.  BVG_EIO when the file or the key is missing. */
int  bvg_labels_underlying(const char* basename, char* buf, int cap);
/* load(): parses labelspec, decodes .labeloffsets on the device, uploads the label stream.  BVG_EIO: a file or the
 * labelspec key is missing (IOException, :409), or the offsets run past the stream; BVG_EFORMAT: unparsable labelspec;
 * BVG_EINVAL: width outside 0..31 (IllegalArgumentException, FixedWidthIntLabel.java:41). */
int  bvg_labels_open(const bvg_graph* g, const char* basename, bvg_labels** out);
/* The same from caller-owned buffers (the whole .labels and .labeloffsets files) with the label class given. */
int  bvg_labels_open_memory(const bvg_graph* g, const uint8_t* labels, uint64_t label_bytes, const uint8_t* label_offsets,
                            uint64_t offsets_bytes, int kind, int width, bvg_labels** out);
void bvg_labels_close(bvg_labels* l);
/* kind, width, length of the whole label stream in bits, bytes of HBM held (any pointer may be NULL). */
int  bvg_labels_info(const bvg_labels* l, int* kind, int* width, int64_t* label_bits, int64_t* loaded_bytes);
/* The labels of all arcs of nodes [from, to), in the order bvg_decode_range returns the successors: the label of the k-th
 * successor of x is entry (row offset of x) + k, what nodeIterator().labelArray() / successors(x).label() give
 * (BitStreamArcLabelledImmutableGraph.java:225-262).  Integer labels: values[arc]; list_off (may be NULL) gets 0..arcs.
 * List labels: arc j carries values[list_off[j] .. list_off[j+1]).  list_off has arcs + 1 entries, values `cap` entries
 * (BVG_ENOMEM when that is too few; call with both NULL to get the sizes); nvalues (may be NULL) receives the number of
 * values.  on_device != 0: the pointers are device pointers on the graph's device, filled in stream order on the graph's
 * stream.  BVG_EFORMAT when the stretch between two label offsets does not hold one label per arc. */
int  bvg_labels_decode_range(const bvg_labels* l, int32_t from, int32_t to, int64_t* list_off, int32_t* values, int64_t cap,
                             int on_device, int64_t* nvalues);
/* Consume-only pass over the labels of [from, to): sum over arcs j = 0, 1, ... of the range of
 * (2j + 1) * (0x9E3779B97F4A7C15 * len_j + sum_i (v_ji + 1) * (2i + 1)) mod 2^64 (len = 1 for integer labels). */
int  bvg_labels_scan_range(const bvg_labels* l, int32_t from, int32_t to, int64_t* arcs, int64_t* nvalues, uint64_t* checksum);

/* ---- EFGraph (decode half): the reference's quasi-succinct format ----
 * <basename>.graph is an LSB-first stream of 64-bit words (byteorder property): per node gamma(outdegree), then the
 * Elias-Fano encoding of the successors and the terminator upperbound -- skip pointers, lower bits, upper bits in unary
 * (reference EFGraph.java:420-556 written, :1064-1145 read); <basename>.offsets holds delta-coded gaps (:641-672).
 * bvg_ef_open mirrors EFGraph.loadInternal (:709-790): BVG_EIO for a missing file, a wrong graphclass or a version > 0
 * (IOException), BVG_EINVAL for a quantum that is no power of two or an unknown byteorder (IllegalArgumentException),
 * BVG_EFORMAT when the outdegrees do not add up to the arcs property.  outdegree / successors / decode_range / scan_range
 * have the meaning of their bvg_* namesakes (same CSR layout, same (arcs, XOR) checksum), so that an EFGraph and a
 * BVGraph of the same graph can be compared by their scans. */
typedef struct bvg_efgraph bvg_efgraph;
int  bvg_ef_open(const char* basename, int device, bvg_efgraph** out);
int  bvg_ef_open_memory(const uint8_t* graph, uint64_t graph_bytes, const uint8_t* offsets_stream, uint64_t offsets_bytes,
                        int32_t nodes, int64_t arcs, int32_t upper_bound, int32_t quantum, int big_endian, int device,
                        bvg_efgraph** out);
void bvg_ef_close(bvg_efgraph* g);
int  bvg_ef_info(const bvg_efgraph* g, int32_t* nodes, int64_t* arcs, int32_t* upper_bound, int32_t* quantum, int64_t* graph_bits);
int  bvg_ef_outdegree(const bvg_efgraph* g, int32_t x, int32_t* d);                                   /* EFGraph.java:1054-1060 */
int  bvg_ef_successors(const bvg_efgraph* g, int32_t x, int32_t* out, int32_t cap, int32_t* d);      /* :1223-1225 */
int  bvg_ef_range_arcs(const bvg_efgraph* g, int32_t from, int32_t to, int64_t* arcs);
int  bvg_ef_decode_range(const bvg_efgraph* g, int32_t from, int32_t to, int64_t* out_off, int32_t* out, int64_t cap, int on_device);
int  bvg_ef_scan_range(const bvg_efgraph* g, int32_t from, int32_t to, int64_t* arcs, uint64_t* checksum);
int  bvg_ef_last_error_node(const bvg_efgraph* g, int32_t* node, int64_t* bitpos);
/* EFGraph.store on the device (EFGraph.java:812-888, Accumulator :420-556), the compress half for this format: the stream is
 * written element-parallel (every field is a function of one successor of one list) from a CSR (off[n + 1], succ[off[n]];
 * host or device pointers).  graph_out (host, graph_cap bytes) receives the long words in little-endian order including the
 * trailing word LongWordOutputBitStream.close() writes, *graph_bytes their size (set also on BVG_ENOMEM: call again with a
 * larger buffer), node_bits (host, n + 1 entries) the bit offset of every node -- the gaps the caller delta-codes into
 * .offsets.  upper_bound <= 0 means n.  BVG_EINVAL for a list that is not strictly increasing or reaches the upper bound
 * (IllegalArgumentException in Accumulator.add).  device_ms (may be NULL): time of the kernels. */
int  bvg_ef_compress(const int64_t* off, const int32_t* succ, int32_t n, int32_t upper_bound, int log2_quantum, int on_device,
                     int device, uint8_t* graph_out, uint64_t graph_cap, uint64_t* graph_bytes, int64_t* node_bits, double* device_ms);

/* ---- BVGraph.store on the device (compress half) ----
 * The reference's compressor (CompressionThread / diffComp / intervalize, BVGraph.java:2049-2386) for the default codings
 * (gamma outdegrees, blocks, block counts and intervals, unary references, zeta_k residuals; compressionflags empty), from a
 * CSR (off[n + 1], succ[off[n]]; host or device pointers).  The node range is cut into ranges of `range_nodes` nodes (<= 0:
 * 256) compressed with an empty window each, as the reference's own multi-threaded store cuts it (:2471-2550): with one range
 * the bytes are those of the single-threaded reference (cnr-2000.graph is reproduced), with the ranges of a T-threaded host
 * writer the bytes of that writer.  graph_out (host, graph_cap bytes) receives .graph, *graph_bytes its size (also on
 * BVG_ENOMEM: call again), node_bits (host, n + 1) the bit position of every node -- the gaps the caller gamma-codes into
 * .offsets.  window <= 31; maxref < 0 = unbounded.  BVG_EINVAL for a list that is not strictly increasing
 * (IllegalArgumentException, :2201). */
int  bvg_bv_compress(const int64_t* off, const int32_t* succ, int32_t n, int32_t window, int32_t maxref, int32_t minlen,
                     int32_t zetak, int32_t range_nodes, int on_device, int device, uint8_t* graph_out, uint64_t graph_cap,
                     uint64_t* graph_bytes, int64_t* node_bits, double* device_ms);

/* ---- diagnostics ---- */
const char* bvg_strerror(int status);
/* Node and bit position of the first record a kernel rejected (BVGraph.java:1129-1131 logs the same pair). */
int  bvg_last_error_node(const bvg_graph* g, int32_t* node, int64_t* bitpos);
/* Number of kernels this library has launched so far in the process (bench.py's gpu_launches). */
int64_t bvg_kernel_launches(void);
/* Device memory of closed graphs and of temporaries stays in a library-owned cache (so that open / scan / close cycles
 * and repeated scans cost no cudaMalloc; at most BVG_CACHE_GB gigabytes, default 32, 0 = no caching).  This gives it back
 * to the driver: every idle block of `device`, or of all devices when device < 0.  Returns the bytes released. */
int64_t bvg_release_cached_memory(int device);
/* Per-kernel device timing for bench.py's roofline object: while enabled every kernel of this graph is bracketed by CUDA
 * events on the launching stream; bvg_profile_read drains them into a JSON object {"kernel": {"launches": n, "ms": t}}. */
int  bvg_profile(const bvg_graph* g, int enable);
int  bvg_profile_read(const bvg_graph* g, char* buf, int cap);
/* Bytes of HBM held by the graph: bit stream, offsets, decode index (header arrays, long-record index and, once a range decode or
 * a scan has asked for them, the length-bucketed schedules). */
int  bvg_memory_footprint(const bvg_graph* g, int64_t* stream_bytes, int64_t* offsets_bytes, int64_t* index_bytes);
/* What a scan of the graph's extent reads of the stream, for roofline arithmetic (out[6]): [0] stream bits of the extent,
 * [1] long records, [2] bits of their residual runs (decoded by every scan), [3] bits of their copy-block and interval
 * sections (expanded once at open, not read again), [4] their arcs, [5] 1 when the schedules are built. */
int  bvg_scan_bits(const bvg_graph* g, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif
