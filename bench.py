#!/usr/bin/env python
"""bench.py -- decoded edges/s of the BVGraph decode path on B200 (BASELINE.json's metric).

  python bench.py --gpus N --steps K --warmup W           # this repo's CUDA path
  python bench.py --impl reference --gpus N ...           # the reference's algorithm on the host cores

A "step" is one whole-graph consume-only scan (what the reference's SpeedTest sequential loop does,
reference src/it/unimi/dsi/webgraph/test/SpeedTest.java:157-185) of a synthetic power-law BVGraph
(1 B arcs, zeta_3, W=7, R=3 at N=1: BASELINE config C3).  At N>1 the same graph is range-sharded into
bit-balanced contiguous node ranges, one per rank (config C5, strong scaling); every step the shards
exchange their boundary reference lists with one NCCL all-gather and each rank scans its shard.

value      : arcs decoded by all ranks / max-over-ranks device time, inputs resident in HBM
e2e        : same metric through the public call with HOST buffers: bvg_scan_memory -- every step uploads every byte
             of the rank's .graph range and .offsets from pinned memory, decodes the offsets, builds the index, scans and
             brings (arcs, checksum) back; the node range goes in --e2e-pieces pieces over two streams so that piece
             p + 1 crosses PCIe while piece p is indexed and scanned (1 piece = open + scan + close)
roofline   : dominant kernel's algorithmic bytes (.graph bits of the scanned range / 8) / its CUDA-event time,
             against MEASURED_PEAKS.json's hbm_gbs
cpu_baseline: the oracle (C restatement of BVGraph.nodeIterator(), kind "port": no JVM exists in the image) on one
             host core, bounded sample
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WEBLIKE = dict(zipf_s=0.45, p_copy=0.97, copy_run=40.0, skip_run=1.5, p_interval=0.4, p_local=0.97, local_bits=9, max_degree=3000,
               interval_max=2, p_same_degree=0.95)
METRIC = "decoded_edges_per_second"
UNIT = "edges/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes", type=int, default=32_000_000)
    ap.add_argument("--arcs", type=int, default=1_070_000_000)  # dedup shortfall ~6 %: lands on ~1.0e9 arcs
    ap.add_argument("--seed", type=int, default=0x5EED)
    ap.add_argument("--workload", default="powerlaw", choices=["powerlaw", "weblike"],
                    help="powerlaw: the no-locality power-law graph BASELINE's metric is quoted on; weblike: a copy-heavy graph with cnr-2000's mix")
    ap.add_argument("--max-degree", type=int, default=1 << 22, help="experiments only: cap on the generator's outdegree law")
    ap.add_argument("--workdir", default=os.environ.get("BVG_BENCH_DIR", "/tmp/bvg_bench"))
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-pieces", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational random-access / materialise legs")
    ap.add_argument("--random-nodes", type=int, default=10_000_000)
    ap.add_argument("--cpu-sample-arcs", type=float, default=3.0e8)
    return ap.parse_args()


def graph_files(args, rank, world, barrier):
    """Rank 0 generates + compresses the synthetic graph once per box (host tools, all cores); others wait."""
    from webgraph_b200 import tools
    kw = dict(max_degree=args.max_degree)
    tag = "pl"
    if args.workload == "weblike":
        # lands near the reference's own fixture (cnr-2000: 3.56 bits/arc, 66 % copied arcs, avgref 1.38): ~3.6 bits/arc, 63 % copied, avgref 1.26
        kw = dict(WEBLIKE)
        tag = "web"
    base = os.path.join(args.workdir, "%s_n%d_m%d_s%x_d%d" % (tag, args.nodes, args.arcs, args.seed, args.max_degree), "g")
    meta = base + ".meta.json"
    if rank == 0 and not os.path.exists(meta):
        os.makedirs(os.path.dirname(base), exist_ok=True)
        t = time.time()
        st = tools.generate_store(base, args.nodes, args.arcs, seed=args.seed, window=7, maxref=3, minlen=4, zetak=3,
                                  threads=os.cpu_count() or 1, **kw)
        st["generate_seconds"] = time.time() - t
        with open(meta + ".tmp", "w") as f:
            json.dump(st, f)
        os.replace(meta + ".tmp", meta)
    barrier()
    with open(meta) as f:
        return base, json.load(f)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed `ncu --set full` capture of
    this same workload (profiles/r01_traffic.json, written by profiles/summarise.py); None when there is no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)
        return t["kernels"].get(kernel)
    except Exception:
        return None


def cpu_port_baseline(base, sample_arcs, threads):
    """The oracle's sequential scan (C restatement of BVGraph.nodeIterator()) on the host: the only place besides tests
    where bench.py executes oracle/."""
    from tests import oracle_binding as ob
    og = ob.load().load(base)
    frac = min(1.0, sample_arcs / max(og.m, 1))
    hi = max(1, int(og.n * frac))
    t = time.time()
    arcs, cs = og.scan_range(0, hi, threads=threads)
    dt = time.time() - t
    og.close()
    return arcs, dt, hi


def workload_name(n_gpus):
    """config.workload, the same string on both arms (BASELINE configs C3 / C5)."""
    if n_gpus <= 1:
        return "C3: consume-only sequential scan of a 1 B-arc synthetic power-law BVGraph (zeta_3, W=7, R=3, minLen=4)"
    return "C5: the C3 graph range-sharded over %d GPUs (bit-balanced node ranges), NCCL all-gather of boundary reference lists per step" % n_gpus


def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm on the box's host cores.  The reference is Java and no JVM/JAR exists
    in this image, so this is the oracle port (oracle/, pinned on the reference's cnr-2000 golden pair) split over all
    host threads exactly like ImmutableGraph.splitNodeIterators."""
    if rank != 0:
        return
    base, st = graph_files(args, 0, 1, lambda: None)
    threads = os.cpu_count() or 1
    # each step: a bounded sample sized for ~a few seconds with all threads
    sample = min(float(st["arcs"]), args.cpu_sample_arcs * max(1, threads // 4))
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_port_baseline(base, sample / 4, threads)
    tot_arcs, tot_t = 0, 0.0
    for _ in range(args.steps):
        arcs, dt, hi = cpu_port_baseline(base, sample, threads)
        tot_arcs += arcs
        tot_t += dt
    value = tot_arcs / tot_t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(args.gpus),
                       "nodes": st["nodes"], "arcs": st["arcs"], "bits_per_arc": st["graph_bits"] / max(st["arcs"], 1),
                       "graph_bytes": (int(st["graph_bits"]) + 7) // 8, "max_outdegree": st["max_outdegree"], "seed": args.seed,
                       "generator": "webgraph_b200.tools.generate_store (SURVEY 8d)",
                       "mode": "the same graph scanned by the CPU port of BVGraph.nodeIterator() (every successor consumed, arcs + XOR checksum)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "first %d nodes (%d arcs) per step, all host threads, node ranges split like splitNodeIterators" % (hi, arcs)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from webgraph_b200 import bvgraph, sharding
    from webgraph_b200.bvgraph import BVGraph

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the decode path is CUDA-only (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    base, st = graph_files(args, rank, world, barrier)
    L = bvgraph.lib()
    n_total, m_total = int(st["nodes"]), int(st["arcs"])

    # ---- host copies of the files in pinned memory (the e2e leg uploads them every step) ----
    graph_np = np.fromfile(base + ".graph", dtype=np.uint8)
    offs_np = np.fromfile(base + ".offsets", dtype=np.uint8)
    graph_pin = torch.from_numpy(graph_np).pin_memory()
    offs_pin = torch.from_numpy(offs_np).pin_memory()
    del graph_np, offs_np

    bounds = bvgraph.plan_shards(base, world)
    lo, hi = bounds[rank], bounds[rank + 1]

    def open_shard():
        h = C.c_void_p()
        bvgraph._check(L.bvg_open_memory_shard(graph_pin.data_ptr(), graph_pin.numel(), offs_pin.data_ptr(), offs_pin.numel(),
                                               n_total, m_total, 7, 3, 4, 3, 0, 1, local_rank, lo, hi, C.byref(h)))
        return BVGraph(h, base)

    g = open_shard()
    stream = torch.cuda.current_stream()
    g.setStream(stream.cuda_stream)
    result = torch.zeros(2, dtype=torch.int64, device=dev)

    # ---- boundary reference lists: one NCCL all-gather per step when any chain crosses a shard cut ----
    need_halo = False
    if world > 1:
        first = C.c_int32()
        bvgraph._check(L.bvg_halo_needed(g.handle, C.byref(first)))
        flag = torch.tensor([1 if first.value < lo else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        need_halo = bool(flag.item())
    bcount = C.c_int32(0)
    if need_halo:
        bvgraph._check(L.bvg_boundary_count(g.handle, C.byref(bcount)))
        bc = bcount.value
        barcs = g.rangeArcs(hi - bc, hi)
        cap_t = torch.tensor([barcs], device=dev)
        dist.all_reduce(cap_t, op=dist.ReduceOp.MAX)
        bcap = int(cap_t.item())
        # fixed-size message: [count+1 offsets as int64 | bcap successors as int32 padded to int64 words] (webgraph_b200/sharding.py)
        msg_words = sharding.message_words(bc, bcap)
        send = torch.zeros(msg_words, dtype=torch.int64, device=dev)
        recv = torch.zeros(world * msg_words, dtype=torch.int64, device=dev)

    def exchange_halo():
        if not need_halo:
            return
        bc = bcount.value
        off_ptr = send.data_ptr()
        lists_ptr = send.data_ptr() + 8 * (bc + 1)
        bvgraph._check(L.bvg_boundary_export(g.handle, off_ptr, lists_ptr, bcap, 1))
        sharding.exchange(send, recv)
        if rank > 0:
            src = sharding.previous_rank_message(recv, rank, msg_words).data_ptr()
            bvgraph._check(L.bvg_halo_import(g.handle, bc, src, src + 8 * (bc + 1), 1))

    def step():
        exchange_halo()
        bvgraph._check(L.bvg_scan_range_async(g.handle, lo, hi, result.data_ptr()))

    # ---- correctness of what is being timed: arcs and checksum against the generator's own ----
    step()
    torch.cuda.synchronize()
    chk = result.clone()
    if world > 1:
        arcs_t = chk[0:1].clone()
        dist.all_reduce(arcs_t, op=dist.ReduceOp.SUM)
        parts = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(parts, chk)
        cs = 0
        for p in parts:
            cs ^= int(p[1].item()) & 0xFFFFFFFFFFFFFFFF
        arcs_all = int(arcs_t.item())
    else:
        arcs_all, cs = int(chk[0].item()), int(chk[1].item()) & 0xFFFFFFFFFFFFFFFF
    if (arcs_all != m_total or cs != int(st["xor_checksum"])) and not os.environ.get("BVG_BENCH_NOCHECK"):  # NOCHECK: timing experiments with BVG_DEBUG_* only
        raise SystemExit("decode mismatch: arcs %d vs %d, checksum %#x vs %#x" % (arcs_all, m_total, cs, int(st["xor_checksum"])))

    for _ in range(max(args.warmup - 1, 0)):
        step()
    torch.cuda.synchronize()
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = bvgraph.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = bvgraph.kernel_launches() - launches0
    clocks = sampler.result()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    value = m_total * args.steps / (ms * 1e-3)

    # ---- per-kernel timing of the same step (CUDA events on the launching stream), for the roofline object ----
    g.profile(True)
    for _ in range(min(args.steps, 5)):
        step()
    torch.cuda.synchronize()
    g.profile(False)
    prof = g.profileRead()
    shard_bits = None
    off_np = None
    roof = None
    if prof:
        dom = max(prof.items(), key=lambda kv: kv[1]["ms"])
        per_step_launches = {k: v["launches"] / min(args.steps, 5) for k, v in prof.items()}
        dom_ms = dom[1]["ms"] / dom[1]["launches"]
        total_ms = sum(v["ms"] for v in prof.values()) / min(args.steps, 5)
        # algorithmic bytes of one launch of the dominant kernel: the .graph bits of the nodes it decodes
        bits = shard_graph_bits(base, bounds, rank, st)
        peak, which = hbm_peak()
        ach = bits / 8 / (dom_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom[0], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": None, "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs, burst copy)" if which == "measured" else "fallback",
                "kernel_ms": dom_ms, "kernel_share_of_step": dom[1]["ms"] / max(sum(v["ms"] for v in prof.values()), 1e-9),
                "step_kernels_ms": {k: v["ms"] / min(args.steps, 5) for k, v in prof.items()},
                "launches_per_step": per_step_launches,
                "whole_step_frac": (bits / 8 / (ms / args.steps * 1e-3) / 1e9) / peak}
        roof["traffic"] = ncu_traffic(dom[0])

    # ---- e2e: host buffers in, result out, every step (open from pinned host memory + scan + close) ----
    e2e = None
    es = max(1, args.e2e_steps)
    foot = g.memoryFootprint()
    g.close()
    torch.cuda.synchronize()
    barrier()
    for _ in range(1):  # one untimed open-scan-close cycle: allocator pools and page tables warm, as for `value`
        g2 = open_shard()
        g2.setStream(stream.cuda_stream)
        g2.scanRange(lo, hi)
        g2.close()
    torch.cuda.synchronize()
    barrier()
    pieces = args.e2e_pieces
    a_out, c_out = C.c_int64(), C.c_uint64()

    def scan_from_host():
        bvgraph._check(L.bvg_scan_memory(graph_pin.data_ptr(), graph_pin.numel(), offs_pin.data_ptr(), offs_pin.numel(),
                                         n_total, m_total, 7, 3, 4, 3, 0, local_rank, lo, hi, pieces, C.byref(a_out), C.byref(c_out)))
        return a_out.value, c_out.value

    for _ in range(2):  # two more untimed calls: the piece-sized blocks of the device memory cache
        scan_from_host()
    torch.cuda.synchronize()
    barrier()
    e2e_parts = []
    t0 = time.perf_counter()
    for _ in range(es):
        e2e_parts.append(scan_from_host())
    torch.cuda.synchronize()
    te = time.perf_counter() - t0
    if world == 1 and any(r != (m_total, int(st["xor_checksum"])) for r in e2e_parts):
        raise SystemExit("e2e decode mismatch: %r" % (e2e_parts,))
    tt = torch.tensor([te], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    te = float(tt.item())
    e2e = {"value": m_total * es / te, "unit": UNIT,
           "h2d_bytes_per_step": int(foot["stream_bytes"] + foot["offsets_bytes"]), "d2h_bytes_per_step": 16 + 24,
           "steps": es, "warmup": 1,
           "pieces": pieces,
           "what": "bvg_scan_memory(pinned host .graph/.offsets, %d pieces): per step H2D of every byte, offsets decode, index build and scan of each piece (piece p + 1 crosses PCIe while piece p is indexed and scanned), result back to the host" % pieces}

    # ---- the other BASELINE configs on the same graph, outside the timed region (N = 1 only): C4 random access to 10 M
    # uniformly random nodes (seeded, as SpeedTest -r, reference test/SpeedTest.java:98-111) and the materialising decode ----
    extra = None
    if world == 1 and not args.no_extras:
        extra = {}
        g3 = open_shard()
        g3.setStream(stream.cuda_stream)
        try:
            nq = args.random_nodes
            gen = torch.Generator(device=dev)
            gen.manual_seed(args.seed)
            xs = torch.randint(0, n_total, (nq,), device=dev, dtype=torch.int32, generator=gen)
            qoff = torch.zeros(nq + 1, dtype=torch.int64, device=dev)
            bvgraph._check(L.bvg_successors_batch(g3.handle, xs.data_ptr(), nq, qoff.data_ptr(), None, 0, 1))
            torch.cuda.synchronize()
            qarcs = int(qoff[-1].item())
            qout = torch.empty(max(qarcs, 1), dtype=torch.int32, device=dev)
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for rep in range(2):  # first repetition warms the allocator
                r0.record()
                bvgraph._check(L.bvg_successors_batch(g3.handle, xs.data_ptr(), nq, qoff.data_ptr(), qout.data_ptr(), qarcs, 1))
                r1.record()
                torch.cuda.synchronize()
            rms = r0.elapsed_time(r1)
            extra["random_access"] = {"config": "C4", "nodes": nq, "arcs": qarcs, "ms": rms, "nodes_per_s": nq / (rms * 1e-3),
                                      "edges_per_s": qarcs / (rms * 1e-3), "what": "bvg_successors_batch, device buffers, sizes + decode"}
            del qout, qoff, xs
            moff = torch.zeros(n_total + 1, dtype=torch.int64, device=dev)
            mout = torch.empty(m_total, dtype=torch.int32, device=dev)
            for rep in range(2):
                r0.record()
                bvgraph._check(L.bvg_decode_range(g3.handle, 0, n_total, moff.data_ptr(), mout.data_ptr(), m_total, 1))
                r1.record()
                torch.cuda.synchronize()
            mms = r0.elapsed_time(r1)
            extra["materialise"] = {"ms": mms, "edges_per_s": m_total / (mms * 1e-3), "bytes_written": 4 * m_total + 8 * (n_total + 1),
                                    "what": "bvg_decode_range of the whole graph into device CSR (int64 offsets + int32 successors)"}
            del moff, mout
            # the NodeIterator route (bvg_cursor_*: batches decoded on the device, copied to pinned host memory, iterated in C)
            cur = C.c_void_p()
            bvgraph._check(L.bvg_cursor_open(g3.handle, 0, 2 ** 31 - 1, C.byref(cur)))
            cn, ca, cc = C.c_int64(), C.c_int64(), C.c_uint64()
            bvgraph._check(L.bvg_cursor_drain(cur, 500_000, C.byref(cn), C.byref(ca), C.byref(cc)))  # warm: pinned buffers
            t0 = time.perf_counter()
            bvgraph._check(L.bvg_cursor_drain(cur, 8_000_000, C.byref(cn), C.byref(ca), C.byref(cc)))
            dt = time.perf_counter() - t0
            L.bvg_cursor_close(cur)
            extra["node_iterator"] = {"nodes": cn.value, "arcs": ca.value, "ms": dt * 1e3, "edges_per_s": ca.value / dt,
                                      "what": "bvg_cursor_drain: bvg_cursor_next over 8 M nodes, every successor consumed on one host thread"}
        except Exception as e:  # informational legs must never take the headline down
            extra["error"] = repr(e)
        g3.close()
        # a graph that comes without .offsets (loadSequential / loadOffline in the reference, BVGraph.java:1581-1609): the
        # record boundaries are found from the .graph stream on the device (bvg_boundaries.cuh); timed against the same
        # open with .offsets, result checked by a scan
        try:
            def timed_open(offs_ptr, offs_len):
                h = C.c_void_p()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                bvgraph._check(L.bvg_open_memory(graph_pin.data_ptr(), graph_pin.numel(), offs_ptr, offs_len, n_total, m_total,
                                                 7, 3, 4, 3, 0, 0, local_rank, C.byref(h)))
                torch.cuda.synchronize()
                return BVGraph(h), (time.perf_counter() - t0) * 1e3
            g4, with_ms = timed_open(offs_pin.data_ptr(), offs_pin.numel())
            g4.close()
            g4, without_ms = timed_open(None, 0)
            ok = g4.scanRange(0, n_total) == (m_total, int(st["xor_checksum"]))
            g4.close()
            extra["open_without_offsets"] = {"ms": without_ms, "ms_with_offsets": with_ms, "scan_matches": bool(ok),
                                             "what": "bvg_open_memory of the pinned .graph with offsets = NULL (offset_type 0): record boundaries from the stream alone, then the usual index build; beside the same open given .offsets"}
        except Exception as e:
            extra["open_without_offsets"] = {"error": repr(e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arcs_c, dt_c, hi_c = cpu_port_baseline(base, args.cpu_sample_arcs, 1)
        cpu = {"value": arcs_c / dt_c, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "oracle sequential scan of the first %d nodes (%d arcs), 1 thread, %.1f s; host has %d cores" % (hi_c, arcs_c, dt_c, os.cpu_count() or 0)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "int32", "data": "synthetic",
                "config": {"workload": workload_name(world),
                           "nodes": n_total, "arcs": m_total, "bits_per_arc": st["graph_bits"] / max(m_total, 1),
                           "graph_bytes": (int(st["graph_bits"]) + 7) // 8, "avg_ref_chain": st["tot_ref"] / max(n_total, 1),
                           "max_outdegree": st["max_outdegree"], "seed": args.seed, "generator": "webgraph_b200.tools.generate_store (SURVEY 8d)",
                           "l2": "input stream (%.2f GB) is far larger than L2; no flush needed" % (st["graph_bits"] / 8e9),
                           "halo_exchange": bool(need_halo), "mode": "consume-only scan (arcs + XOR checksum verified against the generator)"},
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(lt.item()), "clocks": clocks, "extra": extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def shard_graph_bits(base, bounds, rank, st):
    """.graph bits of this rank's node range (algorithmic bytes of one scan).  Whole graph: the file's bit length."""
    if len(bounds) == 2:
        return float(st["graph_bits"])
    # equal-bit shards by construction (bvg_plan_shards)
    return float(st["graph_bits"]) / (len(bounds) - 1)


if __name__ == "__main__":
    main()
