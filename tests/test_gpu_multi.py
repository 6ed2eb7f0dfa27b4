"""New C-ABI surface of round 2 on the GPU: asynchronous forms, placed encode (one-rank jobs on any box; multi-rank
jobs and the multi-device engine when the box has two or more GPUs)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import alac_b200
from alac_b200 import shard
from tests import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("depth,ch", [(16, 2), (24, 2), (32, 1)])
def test_async_forms_match_the_synchronous_calls(engine, depth, ch):
    pcm = synth.make("music", 4096 * 40 + 999, ch, depth, seed=4)
    cfg = alac_b200.EncoderConfig(channels=ch, bit_depth=depth, frames_per_segment=1)
    want = engine.encode(pcm, cfg)
    other = alac_b200.Engine()
    try:
        # two engines in flight at once: an encode on one, a decode on the other (what bench.py's e2e leg does)
        wait_e = engine.encode_submit(pcm, cfg)
        wait_d = other.decode_submit(want.cookie, want.packets, want.sizes)
        got, dec = wait_e(), wait_d()
    finally:
        other.close()
    assert np.array_equal(got.packets, want.packets) and np.array_equal(got.sizes, want.sizes)
    assert dec.status == 0 and np.array_equal(dec.pcm, pcm)
    # one call per engine: a second submit before the wait is refused
    w = engine.encode_submit(pcm, cfg)
    with pytest.raises(alac_b200.AlacError):
        engine.encode_submit(pcm, cfg)
    w()


@pytest.mark.parametrize("form", ["direct", "staged"])
def test_placed_encode_single_rank_job(engine, form):
    """alac_b200_encode_placed with n_ranks = 1: the exchange, the placement and the home rank's wait all run; the job's
    buffer must equal the plain call's output."""
    import ctypes as C
    dev = torch.device("cuda", torch.cuda.current_device())
    ch, depth = 2, 24
    cfg = alac_b200.EncoderConfig(channels=ch, bit_depth=depth, frames_per_segment=1)
    pcm = synth.corpus_torch(0, 4096 * 300 + 17, ch, depth, dev, seed=2)
    want = engine.encode(pcm, cfg)
    cap = alac_b200.encode_bound(cfg, pcm.numel() // cfg.bytes_per_frame)
    dst = torch.zeros(cap, dtype=torch.uint8, device=dev)
    dsz = torch.zeros(want.num_packets, dtype=torch.int32, device=dev)
    xchg = torch.zeros(alac_b200.EXCHANGE_BYTES, dtype=torch.uint8, device=dev)
    stg = torch.zeros(cap, dtype=torch.uint8, device=dev)
    slots = (C.c_uint64 * 1)(0)
    for epoch in (1, 2, 3):
        pl = alac_b200.Placement(dst.data_ptr(), cap, dsz.data_ptr(), 0, xchg.data_ptr(), 0, 1, 0, epoch,
                                 stg.data_ptr() if form == "staged" else None, slots if form == "staged" else None,
                                 1 if (form == "staged" and epoch == 2) else 0)
        sizes, npk, nb, base, mine, _ = engine.encode_placed(pcm, cfg, pl)
        assert engine.placed_finish() in (0, want.nbytes)
        torch.cuda.synchronize()
        assert (npk, nb, base) == (want.num_packets, want.nbytes, 0)
        assert torch.equal(dst[:nb], want.packets) and torch.equal(dsz, torch.as_tensor(want.sizes, device=dev).to(torch.int32))
        assert torch.equal(sizes, dsz)
        dec = engine.decode(want.cookie, mine, sizes)
        assert dec.status == 0 and torch.equal(dec.pcm, pcm)
        dst.zero_()


def _run(cmd):
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    return p.returncode, p.stdout.decode("latin1")[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_multi_device_engine_matches_single_gpu():
    rc, out = _run([sys.executable, "scripts/multi_engine_check.py", "2"])
    assert rc == 0, out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("form", ["staged", "direct"])
def test_placed_encode_two_ranks(form):
    rc, out = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                    "--master-port", "29533" if form == "staged" else "29534", "scripts/placed_check.py", "24", form])
    assert rc == 0, out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_alacconvert_cli_on_two_gpus(tmp_path):
    """`alacconvert -g 2` (alac_b200_engine_create_multi inside the CLI): the same file as on one GPU, and it decodes back."""
    from tests import caf_ref
    exe = os.path.join(os.path.dirname(alac_b200.library_path()), "alacconvert")
    ch, depth, sr = 2, 24, 96000
    pcm = synth.make("music", 4096 * 700 + 123, ch, depth, seed=11)
    wav = str(tmp_path / "in.wav")
    open(wav, "wb").write(caf_ref.wav_bytes(sr, ch, depth, pcm.tobytes()))
    outs = []
    for g in ("1", "2"):
        caf, back = str(tmp_path / f"g{g}.caf"), str(tmp_path / f"g{g}.wav")
        subprocess.run([exe, "-k", "1", "-g", g, wav, caf], check=True, stdout=subprocess.DEVNULL)
        subprocess.run([exe, "-g", g, caf, back], check=True, stdout=subprocess.DEVNULL)
        assert open(back, "rb").read() == open(wav, "rb").read()
        outs.append(open(caf, "rb").read())
    assert outs[0] == outs[1]


def test_stats_name_the_kernel_forms(engine):
    """alac_b200_stats.final_form / search_dense: a dense stereo stream takes the block-ring kernels (two-warp final pass for
    a launch that does not fill the GPU), a stream at an odd byte offset takes the generic ones -- with the same bytes."""
    dev = torch.device("cuda", torch.cuda.current_device())
    ch, depth = 2, 24
    cfg = alac_b200.EncoderConfig(channels=ch, bit_depth=depth, frames_per_segment=1)
    frames = 4096 * 64 + 100
    buf = torch.zeros(frames * 6 + 16, dtype=torch.uint8, device=dev)
    pcm = synth.corpus_torch(0, frames, ch, depth, dev, seed=5)
    buf[:pcm.numel()] = pcm
    dense = engine.encode(buf[:pcm.numel()], cfg)
    assert dense.stats["final_form"] == 2 and dense.stats["search_dense"] == 1
    buf[6:6 + pcm.numel()] = pcm.clone()                    # one sample-frame (6 bytes) further: packets off the 16-byte grid
    odd = engine.encode(buf[6:6 + pcm.numel()], cfg)
    assert odd.stats["final_form"] == 0 and odd.stats["search_dense"] == 0
    assert odd.nbytes == dense.nbytes and torch.equal(odd.packets, dense.packets)
    dec = engine.decode(odd.cookie, odd.packets, odd.sizes)
    assert dec.status == 0 and torch.equal(dec.pcm, pcm)
