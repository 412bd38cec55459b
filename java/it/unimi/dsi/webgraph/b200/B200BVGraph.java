/*
 * B200BVGraph -- ImmutableGraph whose successor lists are decoded on a B200 by libbvgraph_b200.so.
 *
 * Reference-side binding for the drop-in boundary (include/bvgraph_b200.h).  NOT compiled in this repository's
 * image (no JDK, no webgraph/dsiutils/fastutil jars); it is the stub a WebGraph maintainer would add next to
 * it.unimi.dsi.webgraph.BVGraph.  It keeps BVGraph's loader contract (static load/loadMapped/loadOffline/
 * loadSequential/loadOnce looked up by reflection, ImmutableGraph.java:89-104, 232-241, 647-685) so that
 *   -g it.unimi.dsi.webgraph.b200.B200BVGraph
 * on any WebGraph CLI, or graphclass=it.unimi.dsi.webgraph.b200.B200BVGraph in .properties, routes every
 * traversal (SpeedTest, Transform.transpose, HyperBall, ParallelBreadthFirstVisit, ...) through the GPU decoder.
 * Files whose graphclass is it.unimi.dsi.webgraph.BVGraph are accepted as they are.
 */
package it.unimi.dsi.webgraph.b200;

import java.io.IOException;
import java.util.NoSuchElementException;

import it.unimi.dsi.logging.ProgressLogger;
import it.unimi.dsi.webgraph.ImmutableGraph;
import it.unimi.dsi.webgraph.LazyIntIterator;
import it.unimi.dsi.webgraph.LazyIntIterators;
import it.unimi.dsi.webgraph.NodeIterator;

public class B200BVGraph extends ImmutableGraph {
	static { System.loadLibrary("bvgraph_b200_jni"); }

	/** Native bvg_graph*; shared by all copies (the native graph is immutable, ImmutableGraph.java:157-165). */
	private final long handle;
	private final CharSequence basename;
	private final int n;
	private final long m;
	private final boolean randomAccess;

	// ---- native methods: one per C-ABI entry point used here (java/jni/bvg_jni.c) ----
	private static native long nativeOpen(String basename, int offsetType) throws IOException;  // bvg_open
	private static native void nativeClose(long handle);                                        // bvg_close
	private static native int nativeNumNodes(long handle);                                      // bvg_info
	private static native long nativeNumArcs(long handle);                                      // bvg_info
	private static native int nativeOutdegree(long handle, int x);                              // bvg_outdegree
	private static native int[] nativeSuccessorArray(long handle, int x);                       // bvg_successors
	/** Decodes nodes [from, to) in one call: returns offsets (to-from+1 longs) and fills succ[0]. bvg_decode_range */
	private static native long[] nativeDecodeRange(long handle, int from, int to, int[][] succ);
	private static native long nativeRangeArcs(long handle, int from, int to);                  // bvg_range_arcs

	private B200BVGraph(final long handle, final CharSequence basename, final boolean randomAccess) {
		this.handle = handle;
		this.basename = basename;
		this.n = nativeNumNodes(handle);
		this.m = nativeNumArcs(handle);
		this.randomAccess = randomAccess;
	}

	// ---- loaders, same signatures as BVGraph (BVGraph.java:1380-1500) ----
	public static B200BVGraph load(final CharSequence basename, final ProgressLogger pl) throws IOException { return new B200BVGraph(nativeOpen(basename.toString(), 1), basename, true); }
	public static B200BVGraph load(final CharSequence basename) throws IOException { return load(basename, null); }
	public static B200BVGraph loadMapped(final CharSequence basename, final ProgressLogger pl) throws IOException { return new B200BVGraph(nativeOpen(basename.toString(), 2), basename, true); }
	public static B200BVGraph loadMapped(final CharSequence basename) throws IOException { return loadMapped(basename, null); }
	public static B200BVGraph loadOffline(final CharSequence basename, final ProgressLogger pl) throws IOException { return new B200BVGraph(nativeOpen(basename.toString(), -1), basename, false); }
	public static B200BVGraph loadOffline(final CharSequence basename) throws IOException { return loadOffline(basename, null); }
	@Deprecated public static B200BVGraph loadSequential(final CharSequence basename, final ProgressLogger pl) throws IOException { return new B200BVGraph(nativeOpen(basename.toString(), 0), basename, false); }
	@Deprecated public static B200BVGraph loadSequential(final CharSequence basename) throws IOException { return loadSequential(basename, null); }

	@Override public int numNodes() { return n; }
	@Override public long numArcs() { return m; }
	@Override public boolean randomAccess() { return randomAccess; }
	@Override public boolean hasCopiableIterators() { return true; }
	@Override public CharSequence basename() { return basename; }
	@Override public B200BVGraph copy() { return this; }  // native graph is immutable and re-entrant

	@Override public int outdegree(final int x) {  // BVGraph.java:857-879; JNI maps BVG_EINVAL/ESTATE to IAE/ISE
		return nativeOutdegree(handle, x);
	}

	@Override public int[] successorArray(final int x) {  // ImmutableGraph.java:329-333
		return nativeSuccessorArray(handle, x);
	}

	@Override public LazyIntIterator successors(final int x) {  // BVGraph.java:896-904
		final int[] a = nativeSuccessorArray(handle, x);
		return LazyIntIterators.wrap(a, a.length);
	}

	/** Sequential iterator: the GPU decodes BATCH nodes per JNI call (a per-successor or per-node JNI call would dominate). */
	@Override public NodeIterator nodeIterator(final int from) {  // BVGraph.java:1292-1301
		if (from < 0 || from > n) throw new IllegalArgumentException("Node index out of range: " + from);
		return new BatchedIterator(from, n);
	}

	private final class BatchedIterator extends NodeIterator {
		private static final int BATCH = 1 << 16;
		private final int from, upper;
		private int curr, lo, hi;     // curr: last node returned; [lo, hi): nodes held
		private long[] off;
		private int[] succ;

		BatchedIterator(final int from, final int upper) { this.from = from; this.upper = Math.min(upper, n); this.curr = from - 1; }

		@Override public boolean hasNext() { return curr < upper - 1; }

		@Override public int nextInt() {
			if (!hasNext()) throw new NoSuchElementException();
			if (++curr >= hi || off == null) {
				lo = curr;
				hi = (int)Math.min((long)upper, (long)lo + BATCH);
				final int[][] s = new int[1][];
				off = nativeDecodeRange(handle, lo, hi, s);
				succ = s[0];
			}
			return curr;
		}

		@Override public int outdegree() {
			if (curr == from - 1) throw new IllegalStateException();  // BVGraph.java:1237
			return (int)(off[curr - lo + 1] - off[curr - lo]);
		}

		@Override public int[] successorArray() {  // a fresh array: callers may keep it (stricter than BVGraph.java:1228-1233)
			if (curr == from - 1) throw new IllegalStateException();
			final int a = (int)off[curr - lo], b = (int)off[curr - lo + 1];
			return java.util.Arrays.copyOfRange(succ, a, b);
		}

		@Override public LazyIntIterator successors() {
			final int[] a = successorArray();
			return LazyIntIterators.wrap(a, a.length);
		}

		@Override public NodeIterator copy(final int upperBound) {  // BVGraph.java:1252-1260
			return new BatchedIterator(curr + 1, upperBound);
		}
	}
}
