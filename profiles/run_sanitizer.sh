#!/bin/bash
# usage: run_sanitizer.sh TAG  -- compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over a subset of the GPU parity
# suite that covers every kernel family: cnr-2000 scan / random access / ranges, the synthetic 100 k graph (long records, shards
# with imported halos), a corrupted stream, the tile kernel and the stream-position kernel.  Summaries go to gpurun_out/TAG_*.txt.
TAG=${1:-san}
mkdir -p gpurun_out
SUBSET="test_cnr2000_scan_checksum or test_cnr2000_ranges_and_split_iterators or test_cnr2000_single_node_calls or test_synthetic_100k_full_decode or test_shards_halo_redecode_and_import or test_corrupt_stream_reports_error or test_copy_heavy_chains_and_unbounded_refcount or test_alternative_scan_kernels"
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check no"
  timeout 1500 compute-sanitizer --tool $tool $extra --print-limit 20 --error-exitcode 99 \
      python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SUBSET" > gpurun_out/${TAG}_${tool}.txt 2>&1
  echo "[$tool] exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/${TAG}_${tool}.txt | tr '\n' ' ')"
  # the label stream (gamma / fixed / list labels, shards, device buffers), the HyperBall iteration, EFGraph, random access and cursors
  timeout 1500 compute-sanitizer --tool $tool $extra --print-limit 20 --error-exitcode 99 \
      python -m pytest tests/test_labels.py tests/test_efgraph.py tests/test_gpu_parity.py -x -q -m gpu -k "test_gpu_labels_match_the_oracle or test_gpu_labels_on_shards or test_fused_consumer_hyperball_step or test_gpu_efgraph_matches_csr or test_gpu_efgraph_loader_errors or test_cnr2000_random_access_all_nodes or test_cnr2000_node_iterator" > gpurun_out/${TAG}_${tool}_f.txt 2>&1
  echo "[$tool, labels + hyperball + efgraph] exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/${TAG}_${tool}_f.txt | tr '\n' ' ')"
done
