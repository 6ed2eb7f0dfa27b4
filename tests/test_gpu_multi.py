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
