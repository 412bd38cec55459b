"""Arc-label stream on the device: bench.py's arc_labels measurement on its own.
Usage: python profiles/labels_bench.py [nodes] [arcs] > profiles/r02_labels.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 125_000_000
    print(json.dumps(bench.arc_labels(os.environ.get("BVG_BENCH_DIR", "/tmp/bvg_bench"), n, m), indent=1))
