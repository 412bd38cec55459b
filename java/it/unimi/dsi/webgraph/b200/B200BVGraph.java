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
 *
 * Sequential route: every NodeIterator owns a native cursor (its own CUDA stream, error word and two pinned batches:
 * the device decodes the next batch while Java iterates the current one).  A batch reaches Java as two direct buffers
 * over that pinned memory -- no JNI call per node, no array copy per batch; successors() reads straight from the buffer,
 * successorArray() makes the one copy its int[] contract asks for.  Iterators obtained from splitNodeIterators() run
 * concurrently, one per thread (measured in C with bvg_cursor_drain: 8.0 / 10.5 G edges/s with 8 / 16 threads on a 16-core host).
 */
package it.unimi.dsi.webgraph.b200;

import java.io.IOException;
import java.lang.ref.Cleaner;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.IntBuffer;
import java.nio.LongBuffer;
import java.util.NoSuchElementException;

import it.unimi.dsi.logging.ProgressLogger;
import it.unimi.dsi.webgraph.ImmutableGraph;
import it.unimi.dsi.webgraph.LazyIntIterator;
import it.unimi.dsi.webgraph.LazyIntIterators;
import it.unimi.dsi.webgraph.NodeIterator;

public class B200BVGraph extends ImmutableGraph implements AutoCloseable {
	static { System.loadLibrary("bvgraph_b200_jni"); }
	private static final Cleaner CLEANER = Cleaner.create();

	/** Owner of the native bvg_graph*: freed by close() or, failing that, when the last copy is collected. */
	private static final class Native implements Runnable {
		long handle;
		Native(final long handle) { this.handle = handle; }
		@Override public synchronized void run() { if (handle != 0) { nativeClose(handle); handle = 0; } }
	}

	/** Shared by all copies: the native graph's data is immutable; its entry points are serialised by the library, the
	 *  cursors behind NodeIterators are independent of each other (include/bvgraph_b200.h, "Threading"). */
	private final Native nat;
	private final Cleaner.Cleanable cleanable;
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
	private static native long nativeCursorOpen(long handle, int from, int upper);              // bvg_cursor_open
	private static native void nativeCursorClose(long cursor);                                  // bvg_cursor_close
	/** meta = {first node, nodes}; bufs = {offsets (nodes + 1 longs), successors (ints)}: direct buffers over the cursor's
	 *  pinned batch, valid until the next call on this cursor.  bvg_cursor_next_batch */
	private static native boolean nativeCursorNextBatch(long cursor, int[] meta, ByteBuffer[] bufs);

	private B200BVGraph(final long handle, final CharSequence basename, final boolean randomAccess) {
		this.handle = handle;
		this.nat = new Native(handle);
		this.cleanable = CLEANER.register(this, nat);
		this.basename = basename;
		this.n = nativeNumNodes(handle);
		this.m = nativeNumArcs(handle);
		this.randomAccess = randomAccess;
	}

	/** Frees the HBM held by the graph (stream, offsets, index) now instead of at collection. */
	@Override public void close() { cleanable.clean(); }

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
	/** The same object: graph-level native calls are serialised by the library (a lock per native graph), iterators carry
	 *  their own native state.  Threads that want random access in parallel should batch (bvg_successors_batch). */
	@Override public B200BVGraph copy() { return this; }

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

	/** Sequential iterator over a native cursor; from != 0 on a graph loaded without offsets is an IllegalStateException, as
	 *  in the reference (BVGraph.java:1174; the JNI shim maps BVG_ESTATE). */
	@Override public NodeIterator nodeIterator(final int from) {  // BVGraph.java:1292-1301
		if (from < 0 || from > n) throw new IllegalArgumentException("Node index out of range: " + from);
		return new BatchedIterator(from, n);
	}

	private static final class CursorOwner implements Runnable {
		long cursor;
		CursorOwner(final long cursor) { this.cursor = cursor; }
		@Override public synchronized void run() { if (cursor != 0) { nativeCursorClose(cursor); cursor = 0; } }
	}

	private final class BatchedIterator extends NodeIterator implements AutoCloseable {
		private final int from, upper;
		private final CursorOwner owner;
		private final Cleaner.Cleanable cursorCleanable;
		private final int[] meta = new int[2];
		private final ByteBuffer[] bufs = new ByteBuffer[2];
		private int curr, lo, hi;     // curr: last node returned; [lo, hi): nodes of the batch held
		private LongBuffer off;
		private IntBuffer succ;

		BatchedIterator(final int from, final int upper) {
			this.from = from;
			this.upper = Math.min(upper, n);
			this.curr = from - 1;
			this.owner = new CursorOwner(nativeCursorOpen(handle, from, this.upper));
			this.cursorCleanable = CLEANER.register(this, owner);
		}

		@Override public void close() { cursorCleanable.clean(); }

		@Override public boolean hasNext() { return curr < upper - 1; }

		@Override public int nextInt() {
			if (!hasNext()) throw new NoSuchElementException();
			if (++curr >= hi || off == null) {
				if (!nativeCursorNextBatch(owner.cursor, meta, bufs)) throw new NoSuchElementException();
				lo = meta[0];
				hi = lo + meta[1];
				off = bufs[0].order(ByteOrder.nativeOrder()).asLongBuffer();
				succ = bufs[1].order(ByteOrder.nativeOrder()).asIntBuffer();
			}
			return curr;
		}

		@Override public int outdegree() {
			if (curr == from - 1) throw new IllegalStateException();  // BVGraph.java:1237
			return (int)(off.get(curr - lo + 1) - off.get(curr - lo));
		}

		@Override public int[] successorArray() {  // a fresh array: callers may keep it (stricter than BVGraph.java:1228-1233)
			if (curr == from - 1) throw new IllegalStateException();
			final int a = (int)off.get(curr - lo), b = (int)off.get(curr - lo + 1);
			final int[] out = new int[b - a];
			final IntBuffer v = succ.duplicate();
			v.position(a);
			v.get(out);
			return out;
		}

		/** Reads the pinned batch in place; valid until the iterator moves to the next batch (the reference's iterators are
		 *  likewise invalidated by nextInt(), BVGraph.java:1228-1233). */
		@Override public LazyIntIterator successors() {
			if (curr == from - 1) throw new IllegalStateException();
			final int a = (int)off.get(curr - lo), b = (int)off.get(curr - lo + 1);
			final IntBuffer v = succ;
			return new LazyIntIterator() {
				private int i = a;
				@Override public int nextInt() { return i < b ? v.get(i++) : -1; }  // -1 forever after the end (LazyIntIterator.java:35)
				@Override public int skip(final int k) { final int s = Math.min(k, b - i); i += s; return s; }
			};
		}

		@Override public NodeIterator copy(final int upperBound) {  // BVGraph.java:1252-1260
			return new BatchedIterator(curr + 1, upperBound);
		}
	}
}
