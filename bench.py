#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

  metric   : "compress+decompress MB/s on 1 GiB datagen, bit-exact vs reference"
  workload : configs[1] = level -10 compress + decompress, 1 GiB `datagen -P50`, 128 KiB independent blocks
             (--level 21 / 41 select configs[2] / [3]; they are parity cases, not the default bench line)
  step     : one pass of the hot path over the batch: compress all 8192 blocks, then decompress them.
  value    : uncompressed MB (10^6 B) per second of that round trip, inputs resident in HBM, CUDA-event timed,
             whole job over all ranks (weak scaling: every rank owns its own 1 GiB shard, no data-path collective).
  e2e      : same round trip through the reference-facing frame API on HOST buffers (LizardF_compressFrame +
             LizardF_decompress, 128 KiB independent blocks, pinned memory): H2D + kernels + D2H inside the timed
             region, wall clock; per-call split, the box's PCIe copy rates and the NUMA node the process was bound to
             are reported next to it.

`--impl reference` times the UNMODIFIED reference (oracle/_ref/liblizard_ref_speed.so, default flags) on the box's
host cores through the pthread harness in oracle/liboracle.so, same workload, all hardware threads.

One JSON line on stdout (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BS = 1 << 17
# clean-state compressed totals of `datagen -g1G -P50` (seed 0), 8192 x 128 KiB, cap = BS-1
KNOWN_TOTALS_1G = {10: 670259129, 21: 616060194, 41: 385653946}
ALGO_BYTES_PER_BYTE = None  # computed from the measured ratio: 1 + 1/ratio
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` (profiles/r01_v14_ncu_*_level10.txt; 1 GiB).
# The decode figure is of the shipped kernel.  The encode kernel was last captured one change before the shipped one
# (same parser without candidate tags, launch shape 14,11,2): that figure is reported as `traffic_previous_build`, and
# `traffic` stays null until the shipped kernel has its own capture.
NCU_TRAFFIC_BYTES = {(10, "lizard_decode_units_kernel"): 3.17e9}
NCU_TRAFFIC_PREVIOUS_BUILD = {(10, "lizard_encode_units_kernel"): 10.08e9}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--level", type=int, default=10)
    ap.add_argument("--size-mib", type=int, default=1024)
    ap.add_argument("--cpu-sample-mib", type=int, default=64)
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                f = [x.strip() for x in line.split(",")]
                if len(f) < 6:
                    continue
                try:
                    self.samples.append(float(f[0]))
                    self.max_mhz = float(f[1])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def load_checker_libs():
    ref = os.path.join(ROOT, "oracle", "_ref", "liblizard_ref_speed.so")
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(orc):
        raise RuntimeError("oracle/liboracle.so missing: run __graft_entry__.build()")
    O = ctypes.CDLL(orc)
    O.oracle_time_compress.restype = ctypes.c_double
    O.oracle_time_compress.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    O.oracle_time_decompress.restype = ctypes.c_double
    O.oracle_time_decompress.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                         ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    if os.path.exists(ref):
        R = ctypes.CDLL(ref)
        cfn = ctypes.cast(R.Lizard_compress, ctypes.c_void_p)
        dfn = ctypes.cast(R.Lizard_decompress_safe, ctypes.c_void_p)
        kind = "reference"
    else:   # the reference did not travel: fall back to our restatement (level coverage is the same)
        cfn = ctypes.cast(O.oracle_Lizard_compress, ctypes.c_void_p)
        dfn = ctypes.cast(O.oracle_Lizard_decompress_safe, ctypes.c_void_p)
        kind = "port"
    return O, cfn, dfn, kind


def cpu_round_trip(O, cfn, dfn, src_ptr, nbytes, level, threads, iters):
    """Times compress and decompress of nbytes (128 KiB blocks) on `threads` host threads. Returns seconds."""
    n = (nbytes + BS - 1) // BS
    stride = BS + 64
    comp = ctypes.create_string_buffer(n * stride)
    back = ctypes.create_string_buffer(n * BS)
    sizes = (ctypes.c_int * n)()
    ctypes.memset(comp, 1, n * stride)      # pre-touch, as programs/bench.c:195,225,260 does
    ctypes.memset(back, 1, n * BS)
    tc = O.oracle_time_compress(cfn, src_ptr, nbytes, BS, level, comp, stride, sizes, threads, iters)
    td = O.oracle_time_decompress(dfn, comp, stride, sizes, n, back, BS, threads, iters)
    ok = ctypes.string_at(back, min(nbytes, 1 << 20)) == ctypes.string_at(src_ptr, min(nbytes, 1 << 20))
    return tc, td, sum(sizes), ok


# ------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    import lizard_b200 as lz
    nbytes = args.size_mib << 20
    buf = ctypes.create_string_buffer(nbytes)
    lz.datagen_into(ctypes.addressof(buf), nbytes, 50.0, 0)
    O, cfn, dfn, kind = load_checker_libs()
    threads = os.cpu_count() or 1
    for _ in range(max(args.warmup, 1) - 1):
        cpu_round_trip(O, cfn, dfn, ctypes.addressof(buf), min(nbytes, 256 << 20), args.level, threads, 1)
    t0 = time.time()
    tc, td, csum, ok = cpu_round_trip(O, cfn, dfn, ctypes.addressof(buf), nbytes, args.level, threads, max(args.steps, 1))
    mb = nbytes / 1e6
    value = mb / (tc + td)
    line = {
        "impl": "reference", "metric": "compress+decompress MB/s on 1 GiB datagen, bit-exact vs reference",
        "value": round(value, 1), "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round((tc + td) * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "level -%d compress + decompress, %d MiB datagen -P50, 128 KiB independent blocks, host CPU"
                               % (args.level, args.size_mib), "level": args.level, "block": BS,
                   "compress_MBps": round(mb / tc, 1), "decompress_MBps": round(mb / td, 1), "compressed_bytes": csum,
                   "round_trip_ok": bool(ok), "timing": "best of %d passes, CLOCK_MONOTONIC, pre-touched buffers" % max(args.steps, 1)},
        "cpu_baseline": {"value": round(value, 1), "unit": "MB/s", "cores": threads, "kind": kind,
                         "sample": "whole %d MiB buffer, one Lizard_compress/Lizard_decompress_safe call per 128 KiB block, "
                                   "%d pthreads" % (args.size_mib, threads)},
        "e2e": {"value": round(value, 1), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.time() - t0, 2),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
def bind_near_gpu(torch, local_rank):
    """Run this process (and therefore place its pinned host buffers, first touch) on the NUMA node the GPU's PCIe
    root hangs off.  Returns the node number or None when the topology cannot be read."""
    try:
        prop = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def pcie_probe(torch, dev, h_buf):
    """Pinned-memory copy rates of this box's link, GB/s: (H2D alone, D2H alone, both directions at once)."""
    n = min(h_buf.numel(), 256 << 20)
    h_a = h_buf[:n]
    h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.zeros(n, dtype=torch.uint8, device=dev)
    s2 = torch.cuda.Stream(device=dev)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t

    def both():
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)
        d_a.copy_(h_a, non_blocking=True)

    t_h2d = timed(lambda: d_a.copy_(h_a, non_blocking=True))
    t_d2h = timed(lambda: h_b.copy_(d_b, non_blocking=True))
    t_both = timed(both)
    return round(n / t_h2d / 1e9, 1), round(n / t_d2h / 1e9, 1), round(2 * n / t_both / 1e9, 1)


def run_ours(args, rank, world, local_rank):
    import torch
    import lizard_b200 as lz
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the codec has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_near_gpu(torch, local_rank)
    L = lz.lib()
    st = L.LizardB200_setDevice(local_rank)
    if st != 0:
        raise SystemExit("LizardB200_setDevice failed: %d %s" % (st, L.LizardB200_lastError().decode()))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    nbytes = args.size_mib << 20
    n = nbytes // BS
    level = args.level
    # ---- synthetic input: every rank owns one shard (seed = rank), generated straight into pinned memory ----
    h_src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    lz.datagen_into(h_src.data_ptr(), nbytes, 50.0, rank)
    d_src = h_src.to(dev, non_blocking=True)
    stride = (L.Lizard_compressBound(BS) + 15) // 16 * 16
    d_comp = torch.empty(n * stride, dtype=torch.uint8, device=dev)
    d_back = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    d_src_off = idx * BS
    d_comp_off = idx * stride
    d_src_len = torch.full((n,), BS, dtype=torch.int32, device=dev)
    d_cap = torch.full((n,), BS - 1, dtype=torch.int32, device=dev)       # the frame layer's capacity (lizard_frame.c:459)
    d_back_cap = torch.full((n,), BS, dtype=torch.int32, device=dev)
    d_csize = torch.zeros(n, dtype=torch.int32, device=dev)
    d_dsize = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)

    def compress():
        s = L.LizardB200_compress_device(d_src.data_ptr(), d_src_off.data_ptr(), d_src_len.data_ptr(), d_comp.data_ptr(),
                                         d_comp_off.data_ptr(), d_cap.data_ptr(), d_csize.data_ptr(), n, level, sp)
        if s != 0:
            raise SystemExit("compress_device failed: %d %s" % (s, L.LizardB200_lastError().decode()))

    def decompress():
        s = L.LizardB200_decompress_device(d_comp.data_ptr(), d_comp_off.data_ptr(), d_csize.data_ptr(), d_back.data_ptr(),
                                           d_src_off.data_ptr(), d_back_cap.data_ptr(), d_dsize.data_ptr(), n, sp)
        if s != 0:
            raise SystemExit("decompress_device failed: %d %s" % (s, L.LizardB200_lastError().decode()))

    # ---- warm-up + correctness of what will be timed ----
    for _ in range(max(args.warmup, 3)):
        compress()
        decompress()
    torch.cuda.synchronize()
    csize = d_csize.cpu()
    if int((csize <= 0).sum()) != 0:
        raise SystemExit("bench.py: %d blocks failed to compress" % int((csize <= 0).sum()))
    comp_total = int(csize.sum())
    if not torch.equal(d_back, d_src) or int((d_dsize != BS).sum()) != 0:
        raise SystemExit("bench.py: round trip mismatch")
    if rank == 0 and nbytes == (1 << 30) and level in KNOWN_TOTALS_1G and comp_total != KNOWN_TOTALS_1G[level]:
        raise SystemExit("bench.py: compressed total %d != reference clean-state total %d" % (comp_total, KNOWN_TOTALS_1G[level]))

    launches0 = L.LizardB200_launchCount()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    # ---- timed region: exactly K steps, CUDA events on the launching stream, barrier + sync on both sides ----
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    for k in range(args.steps):
        ev[k][0].record(stream)
        compress()
        ev[k][1].record(stream)
        decompress()
        ev[k][2].record(stream)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t_c = sum(ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)) / 1e3
    t_d = sum(ev[k][1].elapsed_time(ev[k][2]) for k in range(args.steps)) / 1e3
    launches = L.LizardB200_launchCount() - launches0

    # ---- end to end through the host-buffer C-ABI (pinned host memory in, host memory out) ----
    e2e = None
    frame_size = 0
    if not args.no_e2e:
        # the reference-facing call a user makes: LizardF_compressFrame / LizardF_decompress on HOST buffers
        lz.bind_frame_api(L)
        prefs = lz.make_prefs(level, 1, True, False, 0)          # 128 KiB independent blocks, no content checksum
        cap = L.LizardF_compressFrameBound(nbytes, ctypes.byref(prefs))
        h_frame = torch.empty(cap, dtype=torch.uint8).pin_memory()
        h_back = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        dctx = ctypes.c_void_p()
        L.LizardF_createDecompressionContext(ctypes.byref(dctx), 100)

        split = [0.0, 0.0]

        def e2e_step():
            t_a = time.perf_counter()
            fs = L.LizardF_compressFrame(h_frame.data_ptr(), cap, h_src.data_ptr(), nbytes, ctypes.byref(prefs))
            if L.LizardF_isError(fs):
                raise SystemExit("LizardF_compressFrame: " + L.LizardF_getErrorName(fs).decode())
            so, si = ctypes.c_size_t(nbytes), ctypes.c_size_t(fs)
            t_b = time.perf_counter()
            r = L.LizardF_decompress(dctx, h_back.data_ptr(), ctypes.byref(so), h_frame.data_ptr(), ctypes.byref(si), None)
            split[0] += t_b - t_a
            split[1] += time.perf_counter() - t_b
            if r != 0 or so.value != nbytes or si.value != fs:
                raise SystemExit("LizardF_decompress: result %d, out %d, in %d of %d" % (r, so.value, si.value, fs))
            return fs

        # warm-up; the timed loop follows immediately (an idle gap lets the GPU drop to its idle clocks and the first
        # kernel afterwards runs ~10x slower for tens of ms); the result is verified after the timed loop
        for _ in range(max(args.warmup, 3)):
            frame_size = e2e_step()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        split[0] = split[1] = 0.0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e = time.perf_counter() - t0
        if not torch.equal(h_back, h_src):
            raise SystemExit("bench.py: e2e round trip mismatch")
        L.LizardF_freeDecompressionContext(dctx)
    clocks = sampler.stop()
    link = pcie_probe(torch, dev, h_src) if e2e is not None else None

    # ---- max over ranks ----
    times = torch.tensor([t_c, t_d, e2e if e2e is not None else 0.0], dtype=torch.float64, device=dev)
    totals = torch.tensor([float(comp_total)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(totals, op=dist.ReduceOp.SUM)
    t_c, t_d, t_e = [float(x) for x in times.cpu()]
    comp_all = float(totals.cpu()[0])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    K = args.steps
    job_bytes = float(nbytes) * world
    mb = job_bytes / 1e6
    value = mb * K / (t_c + t_d)
    ratio = job_bytes / comp_all
    algo_per_launch = float(nbytes) * (1.0 + 1.0 / ratio)       # one rank's launch: read 1 + write 1/ratio (and the reverse)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    def roof(t_total, kernel):
        ach = algo_per_launch / (t_total / K) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures of this
        # workload (profiles/r01_SUMMARY.md section 3 / 9); only known for the level-10 1 GiB launch
        traffic = NCU_TRAFFIC_BYTES.get((level, kernel)) if nbytes == (1 << 30) else None
        prev = NCU_TRAFFIC_PREVIOUS_BUILD.get((level, kernel)) if nbytes == (1 << 30) else None
        r = {"kernel": kernel, "bound": "hbm", "achieved": round(ach, 1), "peak": hbm_peak, "unit": "GB/s",
             "frac": round(ach / hbm_peak, 4), "traffic": traffic, "peak_source": peak_src,
             "algorithmic_bytes_per_launch": int(algo_per_launch), "avg_launch_ms": round(t_total / K * 1e3, 3)}
        if prev is not None:
            r["traffic_previous_build"] = prev
        return r

    line = {
        "metric": "compress+decompress MB/s on 1 GiB datagen, bit-exact vs reference",
        "value": round(value, 1), "unit": "MB/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": round((t_c + t_d) / K * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "level -%d compress + decompress, %d MiB datagen -P50 per GPU (seed = rank), 128 KiB independent blocks"
                               % (level, args.size_mib), "level": level, "block": BS, "blocks_per_gpu": n,
                   "dst_capacity": BS - 1, "l2": "inputs (1 GiB) larger than L2 (126 MB); no flush needed",
                   "compress_MBps": round(mb * K / t_c, 1), "decompress_MBps": round(mb * K / t_d, 1),
                   "compressed_bytes": int(comp_all), "ratio": round(ratio, 4),
                   "parity": "round trip equal; compressed total == reference clean-state total (SURVEY 8c)"
                             if (nbytes == (1 << 30) and level in KNOWN_TOTALS_1G) else "round trip equal"},
        "roofline": roof(t_c, "lizard_encode_units_kernel"),
        "roofline_decode": roof(t_d, "lizard_decode_units_kernel"),
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if e2e is not None:
        line["e2e"] = {"value": round(mb * K / t_e, 1), "unit": "MB/s",
                       "h2d_bytes_per_step": int(nbytes + frame_size), "d2h_bytes_per_step": int(frame_size + nbytes),
                       "api": "LizardF_compressFrame + LizardF_decompress (128 KiB independent blocks), pinned host buffers, "
                              "wall clock, chunked H2D / kernels / D2H overlap", "frame_bytes": int(frame_size),
                       "compress_ms_rank0": round(split[0] / K * 1e3, 2), "decompress_ms_rank0": round(split[1] / K * 1e3, 2),
                       "pcie_GBps_rank0": {"h2d": link[0], "d2h": link[1], "both_directions_total": link[2]},
                       "host_numa_node_rank0": numa_node}
    # ---- CPU side by side (rank 0, N = 1 only): the reference's own code on one host thread, bounded sample ----
    if world == 1:
        try:
            O, cfn, dfn, kind = load_checker_libs()
            sample = min(nbytes, args.cpu_sample_mib << 20)
            tc, td, _, ok = cpu_round_trip(O, cfn, dfn, h_src.data_ptr(), sample, level, 1, 2)
            line["cpu_baseline"] = {"value": round(sample / 1e6 / (tc + td), 1), "unit": "MB/s", "cores": 1, "kind": kind,
                                    "sample": "first %d MiB of the same buffer, per-128-KiB-block calls, best of 2 passes"
                                              % (sample >> 20),
                                    "compress_MBps": round(sample / 1e6 / tc, 1), "decompress_MBps": round(sample / 1e6 / td, 1),
                                    "round_trip_ok": bool(ok)}
        except Exception as ex:   # never lose the GPU line because the checker is missing
            line["cpu_baseline"] = {"value": None, "unit": "MB/s", "cores": 0, "kind": "unavailable", "sample": str(ex)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
