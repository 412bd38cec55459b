#!/usr/bin/env python
"""Writes tests/golden/synthetic/c1_100k.json: the identity of BASELINE config C1/C2's input (100 k-node synthetic
power-law graph, BVGraph defaults W=7 R=3 minLen=4 zeta_3) as produced by this repo's generator + compressor
(webgraph_b200.tools.generate_store, seed 0x5EED, one compression range): node/arc counts, XOR checksum of the arcs, bit
counts and the sha256 of the .graph / .offsets bytes.  The reference ships no power-law generator, so this is
OUR input, pinned so that the benchmark graph cannot drift unnoticed between rounds; the decode ground truth for it is
the generator's own CSR, which tests/test_writer_roundtrip.py::test_c1_graph_is_pinned checks the oracle against.
Run from the repo root:  python tests/golden/synthetic/make_fixture.py"""
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
from webgraph_b200 import tools  # noqa: E402

PARAMS = {"n": 100000, "target_arcs": 3000000, "seed": 0x5EED, "threads": 1}


def describe(base, st):
    out = {k: int(st[k]) for k in ("nodes", "arcs", "graph_bits", "offsets_bits", "copied_arcs", "intervalised_arcs",
                                   "residual_arcs", "max_outdegree", "max_ref_chain", "xor_checksum", "sum_successors")}
    for ext in ("graph", "offsets"):
        out["sha256_" + ext] = hashlib.sha256(open(base + "." + ext, "rb").read()).hexdigest()
    return out


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as d:
        base = os.path.join(d, "c1")
        st = tools.generate_store(base, PARAMS["n"], PARAMS["target_arcs"], seed=PARAMS["seed"], threads=PARAMS["threads"])
        fx = {"params": PARAMS, "expect": describe(base, st)}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_100k.json")
    with open(path, "w") as f:
        json.dump(fx, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", path, fx["expect"]["arcs"], "arcs")
