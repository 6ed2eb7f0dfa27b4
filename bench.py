#!/usr/bin/env python
"""bench.py -- headline benchmark: ALAC encode + decode throughput on B200, bit-exact vs libalac.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
  python bench.py --impl reference ...                     (the reference's CPU path on host cores)

One "step" = encode the whole synthetic corpus, then decode the packets back, on every rank.
Workload (BASELINE.json configs[1]): synthetic 1-hour 16-bit / 44.1 kHz stereo PCM, full
EncodeStereo search, per rank (weak scaling: every rank owns its own hour, sharded by frame
range; there is no data-path collective).  `value` times the device-resident path; `e2e` times
the same step through the C ABI with pinned HOST buffers (H2D + D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SAMPLE_RATE = 44100
CHANNELS = 2
DEPTH = 16
FRAME = 4096
SECONDS = 3600
K_SEGMENT = 1                      # encoder-reset schedule, same on GPU and CPU arms
METRIC = "encode+decode round-trip MSamples/s (sample-frames through encode AND decode per second)"
UNIT = "MSamples/s"


def workload_name(seconds: int) -> str:
    return (f"synthetic {seconds / 3600:g}-hour {DEPTH}-bit/{SAMPLE_RATE / 1000:g} kHz stereo PCM, full EncodeStereo "
            f"mixRes/numU/numV search, frames_per_segment={K_SEGMENT}")


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = f"/tmp/alac_bench_clocks_{os.getpid()}.csv"

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(pcm_host: np.ndarray, frames_per_thread: int, threads: int, want_packets=None):
    """Time the CPU oracle (oracle/_ref = the reference's own primitives when present) encoding and
    decoding `frames_per_thread` packets on each of `threads` host threads (disjoint frame ranges).
    Returns (round-trip sample-frames/s, encode rate, decode rate, kind, packets of thread 0)."""
    from oracle import oracle as O
    O.build()
    ref = O.have_reference()
    bpf = CHANNELS * DEPTH // 8
    total_frames = pcm_host.nbytes // bpf // FRAME
    frames_per_thread = max(1, min(frames_per_thread, total_frames // max(threads, 1)))
    res = [None] * threads

    def work(i):
        a = i * frames_per_thread * FRAME * bpf
        chunk = pcm_host[a:a + frames_per_thread * FRAME * bpf]
        enc = O.Encoder(CHANNELS, DEPTH, SAMPLE_RATE, FRAME, reference=ref)
        t0 = time.perf_counter()
        es = enc.encode_stream(chunk, K_SEGMENT)
        t1 = time.perf_counter()
        dec = O.Decoder(es.cookie, reference=ref)
        back, st = dec.decode_stream(es.packets, es.sizes)
        t2 = time.perf_counter()
        ok = (not st.any()) and np.array_equal(back, chunk)
        res[i] = (t1 - t0, t2 - t1, ok, es)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    w0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    wall = time.perf_counter() - w0
    assert all(r is not None and r[2] for r in res), "CPU oracle round trip failed"
    sf = threads * frames_per_thread * FRAME
    enc_wall = max(r[0] for r in res)
    dec_wall = max(r[1] for r in res)
    return sf / wall, sf / enc_wall, sf / dec_wall, ("reference" if ref else "port"), res[0][3], frames_per_thread


def gpu_local_affinity(index: int):
    """Bind this process to the CPUs NVML reports as local to GPU `index`; returns the previous mask (or None)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        before = os.sched_getaffinity(0)
        cpus &= before
        if cpus:
            os.sched_setaffinity(0, cpus)
            return before
    except Exception:
        pass
    return None


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from tests import synth
    threads = host_threads()
    # bounded sample: ~3 s of CPU work per step per thread (about 4.5 M sample-frames/s/thread)
    frames_per_thread = max(8, min(3200, (args.seconds * SAMPLE_RATE // FRAME) // threads))
    pcm = synth.corpus_torch(0, frames_per_thread * threads * FRAME, CHANNELS, DEPTH, "cpu", seed=0).numpy()
    rates = []
    for i in range(args.warmup + args.steps):
        rt, er, dr, kind, _, fpt = cpu_oracle_rate(pcm, frames_per_thread, threads)
        if i >= args.warmup:
            rates.append((rt, er, dr))
    rt = float(np.mean([r[0] for r in rates]))
    sf = frames_per_thread * threads * FRAME
    line = {
        "impl": "reference", "metric": METRIC, "value": rt / 1e6, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sf / rt, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_name(args.seconds), "sample": f"{sf} sample-frames per step "
                   f"({sf / SAMPLE_RATE / 60:.1f} min of audio), {threads} host threads x {frames_per_thread} packets"},
        "x_realtime": rt / SAMPLE_RATE,
        "encode_msamples_s": float(np.mean([r[1] for r in rates])) / 1e6,
        "decode_msamples_s": float(np.mean([r[2] for r in rates])) / 1e6,
        "cpu_baseline": {"value": rt / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{sf} sample-frames of the workload per step, one oracle encoder+decoder per host thread; "
                                   "reference dp_enc/dp_dec/ag_enc/ag_dec objects under the restated frame drivers" if kind == "reference"
                                   else f"{sf} sample-frames per step, oracle port"},
        "e2e": {"value": rt / 1e6, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    import alac_b200
    from tests import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    frames_total = args.seconds * SAMPLE_RATE                    # sample-frames per rank
    bpf = CHANNELS * DEPTH // 8
    cfg = alac_b200.EncoderConfig(channels=CHANNELS, bit_depth=DEPTH, sample_rate=SAMPLE_RATE, frame_size=FRAME,
                                  frames_per_segment=K_SEGMENT)
    eng = alac_b200.Engine(local)

    # ---- corpus: rank r owns frame range [r*T, (r+1)*T) of the global synthetic stream -------------
    parts = []
    step_frames = 1 << 24
    for a in range(0, frames_total, step_frames):
        n = min(step_frames, frames_total - a)
        parts.append(synth.corpus_torch(rank * frames_total + a, n, CHANNELS, DEPTH, dev, seed=0))
    pcm_d = torch.cat(parts)
    del parts
    npk = (frames_total + FRAME - 1) // FRAME
    bound = alac_b200.encode_bound(cfg, frames_total)
    pk_d = torch.empty(bound, dtype=torch.uint8, device=dev)
    sz_d = torch.empty(npk, dtype=torch.int32, device=dev)
    out_d = torch.empty(pcm_d.numel(), dtype=torch.uint8, device=dev)

    def step_device():
        enc = eng.encode(pcm_d, cfg, out=pk_d, out_sizes=sz_d)
        dec = eng.decode(enc.cookie, enc.packets, enc.sizes, out=out_d)
        return enc, dec

    # ---- parity in the same run: round-trip identity on the full corpus, packets vs oracle on a sample
    enc, dec = step_device()
    assert dec.status == 0 and torch.equal(dec.pcm, pcm_d), "GPU round trip is not the identity"
    payload = enc.nbytes
    ratio = payload / pcm_d.numel()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled from before the warm-up to the end of the timed region (the timed region alone is
    # only tens of milliseconds, shorter than nvidia-smi's start-up)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    for _ in range(max(args.warmup - 1, 0)):
        step_device()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    ms_search = ms_final = ms_asm = ms_dec = ms_ent = ms_fin = ms_fus = ms_enc_k = ms_dec_k = 0.0
    ev0.record()
    for _ in range(args.steps):
        e_, d_ = step_device()
        launches += e_.stats["kernel_launches"] + d_.stats["kernel_launches"]
        ms_search += e_.stats["ms_search"]
        ms_final += e_.stats["ms_final"]
        ms_asm += e_.stats["ms_assemble"]
        ms_dec += d_.stats["ms_decode"]
        ms_ent += d_.stats["ms_entropy"]
        ms_fin += d_.stats["ms_finish"]
        ms_fus += d_.stats["ms_fused"]
        ms_enc_k += e_.stats["ms_kernels"]
        ms_dec_k += d_.stats["ms_kernels"]
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total, ms_enc_k, ms_dec_k], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_enc_max, ms_dec_max = [float(x) for x in t.tolist()]
    ms_step = ms_total / args.steps
    value = world * frames_total / (ms_step / 1e3)

    # ---- e2e: same step through the C ABI with pinned HOST buffers ---------------------------------
    # pinned buffers are first touched (and the calls issued) from the CPUs next to this rank's GPU, as a
    # production host would do; the full mask is restored before the CPU baseline runs
    full_mask = gpu_local_affinity(local)
    pcm_h = torch.empty(pcm_d.numel(), dtype=torch.uint8).pin_memory()
    pcm_h.copy_(pcm_d)
    pk_h = torch.empty(bound, dtype=torch.uint8).pin_memory()
    sz_h = torch.empty(npk, dtype=torch.int32).pin_memory()
    out_h = torch.empty(pcm_d.numel(), dtype=torch.uint8).pin_memory()
    pcm_np, pk_np, sz_np, out_np = pcm_h.numpy(), pk_h.numpy(), sz_h.numpy().view(np.uint32), out_h.numpy()

    def step_host():
        e_ = eng.encode(pcm_np, cfg, out=pk_np, out_sizes=sz_np)
        d_ = eng.decode(e_.cookie, e_.packets, e_.sizes, out=out_np)
        return e_, d_

    e_, d_ = step_host()
    assert np.array_equal(d_.pcm, pcm_np), "host-buffer round trip is not the identity"
    e2e_steps = max(1, min(args.steps, 5))
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e_, d_ = step_host()
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    t = torch.tensor([w1 - w0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item()) / e2e_steps
    e2e_value = world * frames_total / e2e_s
    if full_mask:
        os.sched_setaffinity(0, full_mask)
    h2d = pcm_np.nbytes + payload + 4 * npk + 12 * npk
    d2h = payload + 4 * npk + pcm_np.nbytes + 8 * npk

    # ---- N > 1: the one cross-GPU step, outside the timed region: packet-offset scan over the ranks' totals and one
    #      point-to-point copy per rank into rank 0's buffer (alac_b200.shard.concat_packets_to), timed on the device
    concat = None
    if world > 1:
        from alac_b200 import shard
        enc, _ = step_device()
        szt = enc.sizes.to(torch.int32)
        shard.concat_packets_to(0, enc.packets, szt)                    # first call sets up the NCCL P2P channels
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        res = shard.concat_packets_to(0, enc.packets, szt)
        c1.record()
        torch.cuda.synchronize()
        t = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            concat = {"ms": float(t.item()), "bytes": int(res[0].numel()), "how": "send/recv (NVLink P2P), no collective"}
        del res
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on this box's host cores (bounded sample of the same workload) ---------------
    threads = host_threads()
    fpt_multi = max(8, min(1600, npk // threads))
    need = fpt_multi * threads * FRAME * bpf
    sample_np = pcm_np[:need]
    rt1, er1, dr1, kind, es0, fpt1 = cpu_oracle_rate(sample_np, min(2400, npk), 1)
    rtN, erN, drN, kind, es0N, fptN = cpu_oracle_rate(sample_np, fpt_multi, threads)
    # bit-exactness of the GPU packets against the oracle on that sample (K=1: packets are independent)
    n_chk = len(es0.sizes)
    assert np.array_equal(sz_np[:n_chk], es0.sizes), "GPU packet sizes differ from the CPU oracle"
    assert np.array_equal(pk_np[:es0.packets.nbytes], es0.packets), "GPU packet bytes differ from the CPU oracle"

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    # per-kernel CUDA-event times (recorded by the engine on the launching stream), ms per step
    kernels = {"enc_search_split": ms_search / args.steps, "enc_final": ms_final / args.steps, "enc_assemble": ms_asm / args.steps, "dec_fused": ms_fus / args.steps,
               "dec_entropy": ms_ent / args.steps, "dec_finish": ms_fin / args.steps}
    # algorithmic bytes per launch: what the kernel must read and write once
    pcm_bytes, chan_bytes = pcm_d.numel(), 4 * CHANNELS * frames_total
    alg = {"enc_search_split": pcm_bytes // 8 * 5 + pcm_bytes // 8,     # stage A reads n/8 five times, stage B re-reads n/8 (cache hits: counted once each)
           "enc_final": pcm_bytes + payload,                            # PCM once, Golomb streams once
           "enc_assemble": 2 * payload,
           "dec_fused": payload + chan_bytes + pcm_bytes,               # packets once, the U channel once out and back, PCM once
           "dec_entropy": payload + chan_bytes,                          # packets once, one int32 residual per channel-sample
           "dec_finish": chan_bytes + pcm_bytes}
    dominant = max(kernels, key=kernels.get)
    dom_ms = kernels[dominant]
    alg_bytes = alg[dominant]
    achieved = alg_bytes / (dom_ms / 1e3) / 1e9
    traffic = None          # DRAM bytes per launch from the committed ncu --set full capture (same workload only)
    issue = None
    try:
        if args.seconds == SECONDS:
            for k in json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["kernels"]:
                if dominant in k["kernel"]:
                    traffic = int(k["dram_read_bytes"] + k["dram_write_bytes"])
                    issue = k.get("issue_active_pct")
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_name(args.seconds), "sample_frames_per_gpu": frames_total,
                   "packets_per_gpu": npk, "compression_ratio": round(ratio, 4),
                   "l2": "inputs larger than L2 (635 MB PCM per step), no flush needed",
                   "parallelism": f"frame-range shards x{world}, no collective on the data path"},
        "x_realtime": value / SAMPLE_RATE,
        "encode_msamples_s": world * frames_total / (ms_enc_max / args.steps / 1e3) / 1e6,
        "decode_msamples_s": world * frames_total / (ms_dec_max / args.steps / 1e3) / 1e6,
        "kernel_ms_per_step": {k: round(v, 4) for k, v in kernels.items()},
        "clocks": clocks,
        "e2e": {"value": e2e_value / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_s * 1e3, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dominant + "_kernel<16,stereo>", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_kind,
                     "algorithmic_bytes_per_launch": int(alg_bytes), "ms_per_launch": dom_ms,
                     "issue_active_pct_ncu": issue,
                     "note": "serial integer chains (one per packet x channel): bound by dependent-issue latency, not by HBM; "
                             "see DESIGN.md section 4 and profiles/ for issue-slot utilisation"},
        "cpu_baseline": {"value": rtN / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
                         "single_thread_value": rt1 / 1e6,
                         "encode_msamples_s": erN / 1e6, "decode_msamples_s": drN / 1e6,
                         "sample": f"first {fptN * threads} packets ({fptN * threads * FRAME} sample-frames) of the workload, "
                                   f"{threads} threads x {fptN} packets; single thread: {fpt1} packets; "
                                   "GPU packets byte-compared with this oracle output in the same run"},
        "bit_exact": True,
    }
    if concat:
        line["concat_to_rank0"] = concat
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--seconds", type=int, default=SECONDS, help="audio seconds per rank (default: the 1-hour workload)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
