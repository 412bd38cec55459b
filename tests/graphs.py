"""Small structured / random graphs used by the parity tests (restating the reference's test fixtures:
ArrayListMutableGraph.newCompleteGraph / newCompleteBinaryIntree / newCompleteBinaryOuttree, reference
src/it/unimi/dsi/webgraph/ArrayListMutableGraph.java:157-181; ErdosRenyiGraph in examples/)."""
import numpy as np


def csr_from_lists(lists):
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    for i, l in enumerate(lists):
        off[i + 1] = off[i] + len(l)
    succ = np.fromiter((v for l in lists for v in l), dtype=np.int32, count=int(off[-1]))
    return off, succ


def complete_graph(n, loops=False):
    return csr_from_lists([[j for j in range(n) if loops or j != i] for i in range(n)])


def binary_intree(height):
    n = (1 << (height + 1)) - 1
    return csr_from_lists([[(i - 1) // 2] if i > 0 else [] for i in range(n)])


def binary_outtree(height):
    n = (1 << (height + 1)) - 1
    return csr_from_lists([[j for j in (2 * i + 1, 2 * i + 2) if j < n] for i in range(n)])


def erdos_renyi(n, p, seed):
    rng = np.random.default_rng(seed)
    m = rng.random((n, n)) < p
    np.fill_diagonal(m, False)
    return csr_from_lists([np.nonzero(m[i])[0].tolist() for i in range(n)])


def copy_heavy(n, seed, maxdeg=40, universe=None):
    """Lists that mostly copy from a recent predecessor: long reference chains, many copy blocks, intervals."""
    rng = np.random.default_rng(seed)
    universe = universe or max(4 * n, 64)
    lists = []
    for x in range(n):
        cur = set()
        if x and rng.random() < 0.8:
            proto = lists[x - 1 - int(rng.integers(0, min(x, 7)))]
            for v in proto:
                if rng.random() < 0.7:
                    cur.add(v)
        for _ in range(int(rng.integers(0, maxdeg // 4 + 1))):
            cur.add(int(rng.integers(0, universe)))
        if rng.random() < 0.3:
            s = int(rng.integers(0, universe - 12))
            cur.update(range(s, s + int(rng.integers(2, 12))))
        if rng.random() < 0.1:
            cur = set()
        lists.append(sorted(cur))
    off, succ = csr_from_lists(lists)
    return off, succ, universe


def rows(off, succ):
    return [succ[off[i]:off[i + 1]] for i in range(len(off) - 1)]
