"""webgraph_b200 -- B200-native BVGraph adjacency decompressor behind WebGraph's ImmutableGraph surface.

Layout:  csrc/cuda (sm_100a kernels + the C ABI of include/bvgraph_b200.h), csrc/tools (host-side compressor
and generator), bvgraph.py (Python mirror of ImmutableGraph / NodeIterator / LazyIntIterator over the C ABI).
"""
__all__ = ["build", "tools"]
