"""Range sharding of a whole-graph scan across ranks: the host-side logic around the C ABI's
bvg_plan_shards / bvg_boundary_export / bvg_halo_import.

A shard [lo, hi) depends on earlier nodes only through reference chains, at most window x chain-depth nodes back
(BVGraph.java:705, 2315).  When a chain crosses a cut, every rank publishes the decoded lists of its last `bc` nodes in
one fixed-size message and rank r imports rank r-1's: ONE all-gather per step, no other collective on the data path.
The reference's precedent is BVGraphNodeIterator's constructor, which re-reads the window before `from`
(BVGraph.java:1173-1183); ImmutableGraph.splitNodeIterators (ImmutableGraph.java:379-409) is the range split.

Message layout (int64 words): [bc + 1 arc offsets | ceil(bcap / 2) words holding bcap int32 successors].
The same functions run on CUDA tensors over NCCL (bench.py) and on CPU tensors over gloo (tests/test_multi_rank_gloo.py).
"""
import torch
import torch.distributed as dist


def message_words(bc, bcap):
    return (bc + 1) + (bcap + 1) // 2


def message_views(buf, bc, bcap):
    """(offsets int64[bc + 1], successors int32[bcap]) views of one message."""
    off = buf[:bc + 1]
    lists = buf[bc + 1:bc + 1 + (bcap + 1) // 2].view(torch.int32)[:bcap]
    return off, lists


def agree_capacity(local_arcs, device, group=None):
    """Largest boundary (in arcs) over the ranks: every message is padded to it."""
    t = torch.tensor([int(local_arcs)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def any_rank(flag, device, group=None):
    t = torch.tensor([1 if flag else 0], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return bool(t.item())


def exchange(send, recv, group=None):
    """The one collective of a step: recv[r * len(send) : (r + 1) * len(send)] = rank r's message."""
    try:
        dist.all_gather_into_tensor(recv, send, group=group)
    except (RuntimeError, NotImplementedError):  # backends without the flat variant
        parts = list(recv.view(dist.get_world_size(group), -1).unbind(0))
        dist.all_gather(parts, send, group=group)


def previous_rank_message(recv, rank, words):
    return recv[(rank - 1) * words: rank * words]
