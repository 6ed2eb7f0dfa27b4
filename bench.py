#!/usr/bin/env python
"""bench.py -- headline benchmark: ALAC encode + decode throughput on B200, bit-exact vs libalac.

  python bench.py --gpus N --steps K --warmup W [--config c2|c3]      (N > 1: launched under torchrun)
  python bench.py --impl reference ...                                (the reference's CPU path on host cores)

One "step" = encode the synthetic corpus, then decode the packets back.

  --config c2 (default, BASELINE.json configs[1]): synthetic 1-hour 16-bit / 44.1 kHz stereo PCM, full EncodeStereo
      search.  N > 1: every rank owns its own hour (weak scaling).
  --config c3 (BASELINE.json configs[2], the north-star corpus): synthetic 10-hour 24-bit / 96 kHz stereo PCM sharded by
      frame range over the N GPUs (strong scaling); at N = 1 the whole corpus runs on one GPU.

N > 1 (either config): the cross-GPU step is INSIDE the timed step.  Every rank encodes its frame range with
alac_b200_encode_placed -- packet-size scan, exchange of the byte totals through a 1 KB block in rank 0's memory, and the
assemble kernel storing each packet at its final offset of ONE buffer on rank 0 over NVLink (no NCCL on the data path);
rank 0's call returns when every rank's block is in place.  Each rank then decodes its packet range straight out of that
buffer (peer loads) into its own PCM shard.  step time = slowest device + exchange + placement.

`value` times the device-resident path; `e2e` times the same step through the C ABI with pinned HOST buffers
(H2D + D2H inside the timed region).
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

CHANNELS = 2
FRAME = 4096
K_SEGMENT = 1                      # encoder-reset schedule, same on GPU and CPU arms
METRIC = "encode+decode round-trip MSamples/s (sample-frames through encode AND decode per second)"
UNIT = "MSamples/s"
CONFIGS = {
    # name: (bit depth, sample rate, seconds of audio, scaling at N > 1)
    "c2": (16, 44100, 3600, "weak"),
    "c3": (24, 96000, 36000, "strong"),
}


class Work:
    def __init__(self, name: str, seconds=None):
        self.name = name
        self.depth, self.rate, self.seconds, self.scaling = CONFIGS[name]
        if seconds:
            self.seconds = seconds
        self.bps = {16: 2, 20: 3, 24: 3, 32: 4}[self.depth]
        self.bpf = self.bps * CHANNELS
        self.frames = self.seconds * self.rate                     # sample-frames of the config

    def label(self) -> str:
        return (f"synthetic {self.seconds / 3600:g}-hour {self.depth}-bit/{self.rate / 1000:g} kHz stereo PCM, full EncodeStereo "
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
def cpu_oracle_rate(W: Work, pcm_host: np.ndarray, frames_per_thread: int, threads: int):
    """Time the CPU oracle (oracle/_ref = the reference's own primitives when present) encoding and
    decoding `frames_per_thread` packets on each of `threads` host threads (disjoint frame ranges).
    Returns (round-trip sample-frames/s, encode rate, decode rate, kind, packets of thread 0, packets per thread)."""
    from oracle import oracle as O
    O.build()
    ref = O.have_reference()
    total_frames = pcm_host.nbytes // W.bpf // FRAME
    frames_per_thread = max(1, min(frames_per_thread, total_frames // max(threads, 1)))
    res = [None] * threads

    def work(i):
        a = i * frames_per_thread * FRAME * W.bpf
        chunk = pcm_host[a:a + frames_per_thread * FRAME * W.bpf]
        enc = O.Encoder(CHANNELS, W.depth, W.rate, FRAME, reference=ref)
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
    import torch  # noqa: F401
    from tests import synth
    W = Work(args.config, args.seconds)
    threads = host_threads()
    # bounded sample: ~3 s of CPU work per step per thread (about 4.5 M sample-frames/s/thread)
    frames_per_thread = max(8, min(3200 if W.depth == 16 else 2400, (W.frames // FRAME) // threads))
    pcm = synth.corpus_torch(0, frames_per_thread * threads * FRAME, CHANNELS, W.depth, "cpu", seed=0).numpy()
    rates = []
    for i in range(args.warmup + args.steps):
        rt, er, dr, kind, _, fpt = cpu_oracle_rate(W, pcm, frames_per_thread, threads)
        if i >= args.warmup:
            rates.append((rt, er, dr))
    rt = float(np.mean([r[0] for r in rates]))
    sf = frames_per_thread * threads * FRAME
    line = {
        "impl": "reference", "metric": METRIC, "value": rt / 1e6, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sf / rt, "higher_is_better": True,
        "scaling": W.scaling, "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": W.label(), "name": W.name, "sample": f"{sf} sample-frames per step "
                   f"({sf / W.rate / 60:.1f} min of audio), {threads} host threads x {frames_per_thread} packets"},
        "x_realtime": rt / W.rate,
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
# ncu symbols of the kernels a step launches, by the name the engine's per-kernel timers use
def kernel_symbols(depth: int, final_form: int = 2, search_dense: int = 0) -> dict:
    """final_form / search_dense: which kernels the engine picked (alac_b200_stats): the final pass is enc_final2_kernel
    <depth, stereo, wrap, two-warp> for dense streams, the search kernel carries a fifth `dense` argument for the
    block-ring form (20/24/32-bit and mono)."""
    final = (f"alacb::enc_final_kernel<{depth}, true, true, false>" if not final_form
             else f"alacb::enc_final2_kernel<{depth}, true, false, {'true' if final_form == 2 else 'false'}>")
    return {"enc_search_split": f"alacb::enc_search_split_kernel<{depth}, true, true, false, {'true' if search_dense else 'false'}>",
            "enc_final": final,
            "enc_assemble": f"alacb::enc_assemble_kernel<{depth}>",
            "dec_fused": f"alacb::dec_fused_kernel<{depth}>",
            "dec_entropy": f"alacb::dec_entropy_kernel<{depth}>",
            "dec_finish": f"alacb::dec_finish_kernel<{depth}>"}


def load_ncu_profile(config: str):
    """Committed ncu --set full numbers of this workload (profiles/): per-kernel DRAM bytes and warp instructions."""
    for name in (f"r02_traffic_{config}.json", "r01_traffic.json" if config == "c2" else ""):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))["kernels"]
        except Exception:
            continue
    return []


def pcie_floor_ms(torch, dev, h2d_a, d2h_a, h2d_b, d2h_b, world, dist):
    """Time of the step's host<->device bytes alone, both directions active, every rank at once: encode leg
    (H2D PCM || D2H packets) then decode leg (H2D packets || D2H PCM), pinned memory, CUDA events."""
    n = max(h2d_a, d2h_a, h2d_b, d2h_b, 1)
    hbuf_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    hbuf_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    dbuf_in = torch.empty(n, dtype=torch.uint8, device=dev)
    dbuf_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for up, down in ((h2d_a, d2h_a), (h2d_b, d2h_b)):
            s_in.wait_stream(torch.cuda.current_stream())
            s_out.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s_in):
                dbuf_in[:up].copy_(hbuf_in[:up], non_blocking=True)
            with torch.cuda.stream(s_out):
                hbuf_out[:down].copy_(dbuf_out[:down], non_blocking=True)
            torch.cuda.current_stream().wait_stream(s_in)
            torch.cuda.current_stream().wait_stream(s_out)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        best = ms if best is None else min(best, ms)
    return best


class Ctx:
    pass


def measure_device(W: Work, ctx, steps: int, warmup: int, sampler=None):
    """Device-resident round trip of config W on this rank's share of the job; returns a dict of measurements plus
    the buffers (for the parity checks that follow)."""
    import torch
    import torch.distributed as dist
    import alac_b200
    from alac_b200 import shard
    from tests import synth
    rank, world, dev, eng = ctx.rank, ctx.world, ctx.dev, ctx.eng
    cfg = alac_b200.EncoderConfig(channels=CHANNELS, bit_depth=W.depth, sample_rate=W.rate, frame_size=FRAME,
                                  frames_per_segment=K_SEGMENT)
    cookie = alac_b200.magic_cookie(cfg)
    m = Ctx()
    m.cfg = cfg
    # ---- this rank's frame range of the job
    per_pk = (W.frames + FRAME - 1) // FRAME
    if W.scaling == "weak":
        m.job_frames = world * W.frames                         # every rank owns its own copy of the config
        first_frame, m.frames_rank = rank * W.frames, W.frames
        m.job_packets, first_packet = world * per_pk, rank * per_pk
    else:
        m.job_frames = W.frames
        first_frame, m.frames_rank = shard.plan_frame_shards(W.frames, FRAME, world, K_SEGMENT)[rank]
        m.job_packets = per_pk
        first_packet = shard.plan_packet_shards(per_pk, world, K_SEGMENT)[rank][0]
    parts = []
    step_frames = 1 << 24
    for a in range(0, m.frames_rank, step_frames):
        n = min(step_frames, m.frames_rank - a)
        parts.append(synth.corpus_torch(first_frame + a, n, CHANNELS, W.depth, dev, seed=0))
    m.pcm_d = torch.cat(parts) if len(parts) > 1 else parts[0]
    del parts
    m.npk = (m.frames_rank + FRAME - 1) // FRAME
    m.sz_d = torch.empty(max(m.npk, 1), dtype=torch.int32, device=dev)
    out_d = torch.empty(m.pcm_d.numel(), dtype=torch.uint8, device=dev)
    m.job = m.pk_d = None
    m.phase_wall = [0.0, 0.0, 0.0, 0]
    if world > 1:
        # staged placement (ALAC_B200_PLACE=direct selects the in-kernel peer-store form for comparison)
        slots = None
        if os.environ.get("ALAC_B200_PLACE", "staged") != "direct":
            if W.scaling == "weak":
                slots = [alac_b200.encode_bound(cfg, W.frames)] * world
            else:
                slots = [alac_b200.encode_bound(cfg, nn) for _, nn in shard.plan_frame_shards(W.frames, FRAME, world, K_SEGMENT)]
        m.placement_form = "staged" if slots else "direct"
        m.job = shard.SharedJob(eng, dev, alac_b200.encode_bound(cfg, m.job_frames, world), m.job_packets, slot_bytes=slots)
    else:
        m.pk_d = torch.empty(alac_b200.encode_bound(cfg, m.frames_rank), dtype=torch.uint8, device=dev)

    def step_device():
        """-> (encode stats, decode result, payload bytes of this rank)"""
        if m.job is None:
            enc = eng.encode(m.pcm_d, cfg, out=m.pk_d, out_sizes=m.sz_d)
            dec = eng.decode(cookie, enc.packets, enc.sizes, out=out_d)
            return enc.stats, dec, enc.nbytes
        t0 = time.perf_counter()
        sizes, _, nb, base, mine, st = eng.encode_placed(m.pcm_d, cfg, m.job.placement(first_packet, defer_finish=os.environ.get('ALAC_B200_DEFER', '1') != '0'), out_sizes=m.sz_d)
        t1 = time.perf_counter()
        # every rank decodes its own packet range (its own copy of the block: the concatenation on GPU 0 is the job's
        # only cross-GPU step, BASELINE.json north_star).  On GPU 0 "wait for every rank, close the gaps" runs on the
        # device meanwhile; the step ends when the job's buffer is complete.
        dec = eng.decode(cookie, mine, sizes, out=out_d)
        t2 = time.perf_counter()
        m.job.finish()
        t3 = time.perf_counter()
        m.phase_wall = [m.phase_wall[0] + t1 - t0, m.phase_wall[1] + t2 - t1, m.phase_wall[2] + t3 - t2, m.phase_wall[3] + 1]
        return st, dec, nb

    # ---- parity in the same run: round-trip identity on this rank's whole range
    _, dec, m.payload = step_device()
    assert dec.status == 0 and torch.equal(dec.pcm, m.pcm_d), "GPU round trip is not the identity"
    m.ratio = m.payload / m.pcm_d.numel()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled from before the warm-up to the end of the timed region (the timed region alone is
    # only tens of milliseconds, shorter than nvidia-smi's start-up)
    if sampler is not None and rank == 0:
        sampler.start()
        time.sleep(0.3)
    for _ in range(max(warmup - 1, 0)):
        step_device()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    acc = {k: 0.0 for k in ("ms_search", "ms_final", "ms_assemble", "ms_entropy", "ms_finish", "ms_fused")}
    ms_enc_k = ms_dec_k = 0.0
    m.phase_wall = [0.0, 0.0, 0.0, 0]
    ev0.record()
    for _ in range(steps):
        es, d_, _ = step_device()
        launches += es["kernel_launches"] + d_.stats["kernel_launches"]
        for k in ("ms_search", "ms_final", "ms_assemble"):
            acc[k] += es[k]
        for k in ("ms_entropy", "ms_finish", "ms_fused"):
            acc[k] += d_.stats[k]
        ms_enc_k += es["ms_kernels"]
        ms_dec_k += d_.stats["ms_kernels"]
    ev1.record()
    barrier()
    # ---- the same K steps with the calls of neighbouring steps overlapped (one GPU): a second engine decodes step i on
    #      its own stream while the first engine already encodes step i + 1 (alac_b200_encode_submit / _decode_submit).
    #      Every step is still a full encode followed by the decode of ITS packets; only the idle issue slots of one
    #      call are filled by the other (the chain kernels of the 1-hour workload keep 4 warps per sub-partition busy).
    m.overlap = None
    bound = alac_b200.encode_bound(cfg, m.frames_rank)
    if world == 1 and 3 * bound + 2 * m.pcm_d.numel() < 40e9 and os.environ.get("ALAC_BENCH_OVERLAP", "1") != "0":
        eng2 = alac_b200.Engine(dev.index)
        s_enc, s_dec = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        pks = [m.pk_d, torch.empty_like(m.pk_d), torch.empty_like(m.pk_d)]
        szs = [m.sz_d, torch.empty_like(m.sz_d), torch.empty_like(m.sz_d)]

        def pipeline_device(n):
            with torch.cuda.stream(s_enc):
                w_enc = eng.encode_submit(m.pcm_d, cfg, out=pks[0], out_sizes=szs[0])
            w_dec, dec = None, None
            for i in range(n):
                enc = w_enc()
                if i + 1 < n:
                    with torch.cuda.stream(s_enc):
                        w_enc = eng.encode_submit(m.pcm_d, cfg, out=pks[(i + 1) % 3], out_sizes=szs[(i + 1) % 3])
                if w_dec is not None:
                    dec = w_dec()
                with torch.cuda.stream(s_dec):
                    w_dec = eng2.decode_submit(cookie, enc.packets, enc.sizes, out=out_d)
            return w_dec()

        out_d.zero_()
        dec = pipeline_device(max(warmup, 2))
        torch.cuda.synchronize()
        assert dec.status == 0 and torch.equal(dec.pcm, m.pcm_d), "overlapped GPU round trip is not the identity"
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        o0.record()
        pipeline_device(steps)
        torch.cuda.synchronize()
        o1.record()
        torch.cuda.synchronize()
        oms = o0.elapsed_time(o1) / steps
        m.overlap = {"value": m.job_frames / (oms / 1e3) / 1e6, "unit": UNIT, "ms_per_step": oms, "steps": steps,
                     "how": "two engines on two streams: alac_b200_encode_submit of step i+1 runs while step i decodes "
                            "(alac_b200_decode_submit); every step is a full encode followed by the decode of its own packets"}
        eng2.close()
        del pks, szs
    m.clocks = sampler.stop() if (sampler is not None and rank == 0) else {}
    t = torch.tensor([ev0.elapsed_time(ev1), ms_enc_k, ms_dec_k], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_enc_max, ms_dec_max = [float(x) for x in t.tolist()]
    m.ms_step = ms_total / steps
    m.value = m.job_frames / (m.ms_step / 1e3)
    m.enc_rate = m.job_frames / (ms_enc_max / steps / 1e3)
    m.dec_rate = m.job_frames / (ms_dec_max / steps / 1e3)
    tot = torch.tensor([m.payload, launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    m.job_payload, m.job_launches = int(tot[0].item()), int(tot[1].item())
    m.phases = None
    if world > 1:
        # host wall time per step of the three calls of a step, per rank: [encode_placed, decode, placed_finish] in ms
        ph = torch.tensor([x / max(m.phase_wall[3], 1) * 1e3 for x in m.phase_wall[:3]], dtype=torch.float64, device=dev)
        allph = [torch.zeros_like(ph) for _ in range(world)]
        dist.all_gather(allph, ph)
        m.phases = [[round(float(v), 2) for v in t.tolist()] for t in allph]
    if m.job is not None:
        # the placed call runs its chunks on several streams, so its per-kernel timers overlap; the per-kernel figures (and the
        # roofline) of a multi-GPU run come from one plain single-stream pass over the same shard, outside the timed region
        pk_l = torch.empty(alac_b200.encode_bound(cfg, m.frames_rank), dtype=torch.uint8, device=dev)
        for k in acc:
            acc[k] = 0.0
        for _ in range(2):
            enc = eng.encode(m.pcm_d, cfg, out=pk_l, out_sizes=m.sz_d)
            dd = eng.decode(cookie, enc.packets, enc.sizes, out=out_d)
        for k in ("ms_search", "ms_final", "ms_assemble"):
            acc[k] = enc.stats[k] * steps
        for k in ("ms_entropy", "ms_finish", "ms_fused"):
            acc[k] = dd.stats[k] * steps
        es = enc.stats
        del pk_l
    m.final_form, m.search_dense = int(es["final_form"]), int(es["search_dense"])
    m.kernels = {"enc_search_split": acc["ms_search"] / steps, "enc_final": acc["ms_final"] / steps,
                 "enc_assemble": acc["ms_assemble"] / steps, "dec_fused": acc["ms_fused"] / steps,
                 "dec_entropy": acc["ms_entropy"] / steps, "dec_finish": acc["ms_finish"] / steps}
    del out_d
    return m


def check_against_oracle(W: Work, m, es0):
    """GPU packets of this rank == the CPU oracle's on the sample it encoded (K=1: packets are independent)."""
    import torch
    n_chk = len(es0.sizes)
    got_sizes = m.sz_d[:n_chk].cpu().numpy().view(np.uint32)
    got_pk = (m.job.packets if m.job is not None else m.pk_d)[:es0.packets.nbytes].cpu().numpy()
    assert np.array_equal(got_sizes, es0.sizes), "GPU packet sizes differ from the CPU oracle"
    assert np.array_equal(got_pk, es0.packets), "GPU packet bytes differ from the CPU oracle"
    if m.job is not None:
        # the job's single buffer on this GPU: every rank's block at its place, size table complete
        all_sizes = m.job.sizes[:m.job_packets].to(torch.int64)
        assert int(all_sizes.sum().item()) == m.job_payload, "shared buffer: size table does not add up to the ranks' totals"


def run_cuda(args):
    import torch
    import torch.distributed as dist
    import alac_b200

    W = Work(args.config, args.seconds)
    ctx = Ctx()
    ctx.rank = rank = int(os.environ.get("RANK", "0"))
    ctx.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    ctx.dev = dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx.eng = eng = alac_b200.Engine(local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    m = measure_device(W, ctx, args.steps, args.warmup, ClockSampler(local))
    cfg, pcm_d, npk, payload = m.cfg, m.pcm_d, m.npk, m.payload

    # ---- e2e: the same round trip through the C ABI with pinned HOST buffers ------------------------------
    # pinned buffers are first touched (and the calls issued) from the CPUs next to this rank's GPU, as a
    # production host would do; the full mask is restored before the CPU baseline runs.  The corpus goes through in
    # slices of at most one hour of audio (the pinned staging buffers are reused from slice to slice).
    full_mask = gpu_local_affinity(local)
    slice_frames = min(m.frames_rank, 3600 * W.rate)
    n_slices = (m.frames_rank + slice_frames - 1) // slice_frames
    slice_pk = (slice_frames + FRAME - 1) // FRAME
    pcm_h = torch.empty(slice_frames * W.bpf, dtype=torch.uint8).pin_memory()
    # engines per direction in the pipelined leg.  One each is the measured optimum: a second pair gains 2 % on the 1-hour
    # 16-bit slices (28.1 against 28.7 ms) and loses 14 % on the 2 GB slices of the 24/96 corpus (903 against 790 ms)
    D = max(1, int(os.environ.get("ALAC_BENCH_E2E_DEPTH", "1")))
    nbuf = 2 * D + 1
    pk_hs = [torch.empty(alac_b200.encode_bound(cfg, slice_frames), dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
    sz_hs = [torch.empty(slice_pk, dtype=torch.int32).pin_memory() for _ in range(nbuf)]
    out_hs = [torch.empty(slice_frames * W.bpf, dtype=torch.uint8).pin_memory() for _ in range(D)]
    pcm_np, out_nps = pcm_h.numpy(), [t.numpy() for t in out_hs]
    out_np = out_nps[0]
    pk_np = [t.numpy() for t in pk_hs]
    sz_np = [t.numpy().view(np.uint32) for t in sz_hs]
    pcm_h.copy_(pcm_d[:slice_frames * W.bpf])
    enc_engs = [eng] + [alac_b200.Engine(local) for _ in range(D - 1)]
    dec_engs = [alac_b200.Engine(local) for _ in range(D)]      # further engines on the same GPU: the decode leg of the pipeline

    def step_host():
        """one slice, strictly alternating synchronous calls: host PCM -> host packets -> host PCM"""
        e_ = eng.encode(pcm_np, cfg, out=pk_np[0], out_sizes=sz_np[0])
        d_ = eng.decode(e_.cookie, e_.packets, e_.sizes, out=out_np)
        return e_, d_

    def pipeline_host(n):
        """n slices through D encode engines and D decode engines, all kept busy with the asynchronous forms of the C ABI
        (alac_b200_encode_submit / alac_b200_decode_submit, one call in flight per engine): slice i is encoded by engine
        i mod D and decoded by decode engine i mod D; 2 D + 1 packet buffers rotate between them.  The tail of one call (its
        last kernels, its last bytes coming down) overlaps the head of the next, so both directions of the link stay
        busy.  Every slice still goes host -> device -> host (packets) -> device -> host (PCM), all inside the timed
        region."""
        enc_wait, dec_wait = {}, {}
        for j in range(min(D, n)):
            enc_wait[j] = enc_engs[j % D].encode_submit(pcm_np, cfg, out=pk_np[j % nbuf], out_sizes=sz_np[j % nbuf])
        d_ = None
        for i in range(n):
            e_ = enc_wait.pop(i)()
            j = i + D                               # (its packet buffer was last read by the decode of slice i - D - 1: done)
            if j < n:
                enc_wait[j] = enc_engs[j % D].encode_submit(pcm_np, cfg, out=pk_np[j % nbuf], out_sizes=sz_np[j % nbuf])
            if (i - D) in dec_wait:                 # decode engine i mod D is free after this
                d_ = dec_wait.pop(i - D)()
            dec_wait[i] = dec_engs[i % D].decode_submit(e_.cookie, e_.packets, e_.sizes, out=out_nps[i % D])
        for i in sorted(dec_wait):
            d_ = dec_wait[i]()
        return e_, d_

    e_, d_ = step_host()
    assert np.array_equal(d_.pcm, pcm_np), "host-buffer round trip is not the identity"
    out_np[:] = 0
    for o in out_nps:
        o[:] = 0
    e_, d_ = pipeline_host(2 * D)
    assert all(np.array_equal(o, pcm_np) for o in out_nps), "pipelined host-buffer round trip is not the identity"
    slice_payload = e_.nbytes
    e2e_steps = max(1, min(args.steps, 20 if n_slices == 1 else 1))

    def timed(fn):
        barrier()
        w0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        tt = torch.tensor([w1 - w0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()) / e2e_steps

    # (the last slice of a shard may be shorter; slices are timed at full length: <= 1 % more work than the config)
    e2e_serial_s = timed(lambda: [step_host() for _ in range(e2e_steps * n_slices)])
    e2e_s = timed(lambda: pipeline_host(e2e_steps * n_slices))
    e2e_value = m.job_frames / e2e_s
    h2d = n_slices * (pcm_np.nbytes + slice_payload + 4 * slice_pk)
    d2h = n_slices * (slice_payload + 4 * slice_pk + pcm_np.nbytes + 8 * slice_pk)
    # floors: the step's bytes alone over the link, every rank at once -- alternating calls (encode leg then decode leg) and
    # fully overlapped (all H2D bytes against all D2H bytes)
    floor_serial_ms = n_slices * pcie_floor_ms(torch, dev, pcm_np.nbytes, slice_payload, slice_payload, pcm_np.nbytes, world, dist)
    floor_ms = n_slices * pcie_floor_ms(torch, dev, pcm_np.nbytes + slice_payload, pcm_np.nbytes + slice_payload, 0, 0, world, dist)
    if full_mask:
        os.sched_setaffinity(0, full_mask)
    for x in enc_engs[1:] + dec_engs:
        x.close()
    del pcm_h, out_hs, pk_hs

    # ---- CPU baseline on this box's host cores (bounded sample of the same workload), rank 0 ----------------
    line = None
    threads = host_threads()
    if rank == 0:
        fpt_multi = max(8, min(1600 if W.depth == 16 else 1200, npk // threads))
        need = fpt_multi * threads * FRAME * W.bpf
        sample_np = pcm_d[:need].cpu().numpy()
        rt1, er1, dr1, kind, es0, fpt1 = cpu_oracle_rate(W, sample_np, min(2400 if W.depth == 16 else 1800, npk), 1)
        rtN, erN, drN, kind, es0N, fptN = cpu_oracle_rate(W, sample_np, fpt_multi, threads)
        check_against_oracle(W, m, es0)

        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        kernels = m.kernels         # per-kernel CUDA-event times (recorded by the engine on the launching stream), ms per step, this rank
        # Algorithmic bytes per launch: PCM bytes + packet bytes, each counted once -- no scratch round trips.
        # The dominant kernel is reported against that whole total of its direction; its own share is given next to it:
        #   search: the first eighth of the PCM (what stages A / B read);  final: the PCM;  assemble: the packets;
        #   fused decode: packets + PCM;  entropy: the packets;  finish: the PCM.
        pcm_bytes = pcm_d.numel()
        share = {"enc_search_split": pcm_bytes // 8, "enc_final": pcm_bytes, "enc_assemble": payload,
                 "dec_fused": payload + pcm_bytes, "dec_entropy": payload, "dec_finish": pcm_bytes}
        dominant = max(kernels, key=kernels.get)
        dom_ms = kernels[dominant]
        alg_bytes = pcm_bytes + payload
        achieved = alg_bytes / (dom_ms / 1e3) / 1e9
        symbols = kernel_symbols(W.depth, m.final_form, m.search_dense)
        traffic = issue_frac = warp_inst = None
        if W.seconds == CONFIGS[W.name][2] and world == 1:
            for k in load_ncu_profile(W.name):
                if dominant + "_kernel" in k["kernel"] or dominant.replace("enc_final", "enc_final2") + "_kernel" in k["kernel"]:
                    traffic = int(k["dram_read_bytes"] + k["dram_write_bytes"])
                    warp_inst = k.get("inst_executed")
        sm_clock = (m.clocks.get("sm_mhz") or 1965.0) * 1e6
        if warp_inst:
            # issue-slot bound: one warp instruction per SM sub-partition per clock (148 SMs x 4)
            issue_frac = warp_inst / (148 * 4 * sm_clock * dom_ms / 1e3)
        line = {
            "metric": METRIC, "value": m.value / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": m.ms_step, "higher_is_better": True, "scaling": W.scaling if world > 1 else "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": W.label(), "name": W.name, "sample_frames_per_job": m.job_frames, "sample_frames_this_gpu": m.frames_rank,
                       "packets_this_gpu": npk, "compression_ratio": round(m.ratio, 4),
                       "l2": f"inputs larger than L2 ({pcm_bytes / 1e6:.0f} MB PCM per GPU and step), no flush needed",
                       "parallelism": (f"frame-range shards x{world}; packet-offset exchange + {m.placement_form} placement of every rank's packets into "
                                       "one contiguous buffer on GPU 0 over NVLink inside the timed step (alac_b200_encode_placed); "
                                       "every rank decodes its own packet range; no collective" if world > 1 else "one GPU")},
            "x_realtime": m.value / W.rate,
            "encode_msamples_s": m.enc_rate / 1e6, "decode_msamples_s": m.dec_rate / 1e6,
            "kernel_ms_per_step": {symbols[k]: round(v, 4) for k, v in kernels.items()},
            "clocks": m.clocks,
            "e2e": {"value": e2e_value / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "slices_per_step": n_slices,
                    "how": f"{D} encode + {D} decode engines on the GPU driven through alac_b200_encode_submit / alac_b200_decode_submit: the "
                           "calls of neighbouring steps overlap (full-duplex PCIe), every step still host -> device -> host -> device -> host",
                    "pcie_floor_ms": floor_ms, "e2e_over_floor": e2e_s * 1e3 / floor_ms if floor_ms else None,
                    "alternating_calls_ms_per_step": e2e_serial_s * 1e3, "alternating_calls_pcie_floor_ms": floor_serial_ms},
            "gpu_launches": int(m.job_launches),
            **({"overlapped_steps": m.overlap} if m.overlap else {}),
            **({"call_ms_per_rank": {"order": ["alac_b200_encode_placed", "alac_b200_decode", "alac_b200_placed_finish"], "ranks": m.phases}} if m.phases else {}),
            "roofline": {"bound": "hbm", "kernel": symbols[dominant], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_kind,
                         "algorithmic_bytes_per_launch": int(alg_bytes), "ms_per_launch": dom_ms,
                         "traffic_over_algorithmic": (traffic / alg_bytes) if traffic else None,
                         "issue_frac": issue_frac, "warp_instructions_per_launch_ncu": warp_inst,
                         "kernel_share_of_algorithmic_bytes": int(share[dominant]),
                         "note": "algorithmic bytes = PCM + packet bytes, each once (SURVEY 8d); scratch traffic is not counted.  The kernels are "
                                 "serial integer chains (one per packet x channel) bound by the issue rate, not by HBM: issue_frac = "
                                 "warp instructions (ncu) / (148 SMs x 4 sub-partitions x SM clock x kernel time)"},
            "cpu_baseline": {"value": rtN / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
                             "single_thread_value": rt1 / 1e6,
                             "encode_msamples_s": erN / 1e6, "decode_msamples_s": drN / 1e6,
                             "sample": f"first {fptN * threads} packets ({fptN * threads * FRAME} sample-frames) of the workload, "
                                       f"{threads} threads x {fptN} packets; single thread: {fpt1} packets; "
                                       "GPU packets byte-compared with this oracle output in the same run"},
            "x_all_host_cores": m.value / rtN,
            "bit_exact": True,
        }
    if m.job is not None:
        dist.barrier()
        m.job.close()
    del m, pcm_d

    # ---- N > 1, default config: the north-star corpus (config 3: 10 h 24/96 stereo) sharded by frame range over the same
    #      ranks, strong scaling, same timed step (encode + exchange + placement + decode); reported as `config3`
    if world > 1 and W.name == "c2" and not args.no_config3:
        torch.cuda.empty_cache()
        W3 = Work("c3")
        m3 = measure_device(W3, ctx, 3, 2)
        if rank == 0:
            fpt = max(8, min(800, m3.npk // threads))
            sample_np = m3.pcm_d[:fpt * threads * FRAME * W3.bpf].cpu().numpy()
            rtN3, erN3, drN3, kind3, es03, fptN3 = cpu_oracle_rate(W3, sample_np, fpt, threads)
            check_against_oracle(W3, m3, es03)
            line["config3"] = {"workload": W3.label(), "scaling": "strong", "value": m3.value / 1e6, "unit": UNIT, "ms_per_step": m3.ms_step,
                               "steps": 3, "warmup": 2, "x_realtime": m3.value / W3.rate, "encode_msamples_s": m3.enc_rate / 1e6,
                               "decode_msamples_s": m3.dec_rate / 1e6, "compression_ratio": round(m3.ratio, 4),
                               "sample_frames_per_job": m3.job_frames, "packets_this_gpu": m3.npk,
                               "call_ms_per_rank": {"order": ["alac_b200_encode_placed", "alac_b200_decode", "alac_b200_placed_finish"], "ranks": m3.phases},
                               "kernel_ms_per_step": {kernel_symbols(24, m3.final_form, m3.search_dense)[k]: round(v, 4) for k, v in m3.kernels.items()},
                               "cpu_baseline": {"value": rtN3 / 1e6, "unit": UNIT, "cores": threads, "kind": kind3,
                                                "sample": f"{threads} threads x {fptN3} packets of the same corpus"},
                               "x_all_host_cores": m3.value / rtN3, "bit_exact": True}
        dist.barrier()
        m3.job.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="c2: 1 h 16/44.1 (default); c3: 10 h 24/96, the north-star corpus")
    ap.add_argument("--seconds", type=int, default=None, help="override the audio seconds of the config (development)")
    ap.add_argument("--no-config3", action="store_true", help="N > 1, default config: skip the extra 10-hour 24/96 strong-scaling record")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
