"""Batch of independent captures (BASELINE.json configs[3]): mixed Ultralight / Classic sessions, hi_val per capture,
decoded concurrently through separate nfc_streams; every capture must equal the oracle's decode of that capture."""
import numpy as np
import pytest

from oracle import oracle
from usrp_nfc_b200 import _cabi, batch, synth

pytestmark = pytest.mark.gpu


def _captures(rate, n, seed0):
    sess = synth.load_sessions()
    p = synth.rate_params(rate)
    caps, params = [], []
    for i in range(n):
        name = "ultralight" if i % 3 == 0 else "classic1k"
        hi = [1.05, 1.06, 1.07, 1.08, 1.09, 1.10][i % 6]
        ch = synth.Channel(pause=0.02 + 0.005 * (i % 4), tag_high=1.06 + 0.01 * (i % 3), fade=0.02 * (i % 2))
        pcm = synth.capture(sess[name], rate, seed0 + i, channel=ch, av_window=p["av_window"])
        caps.append(synth.envelope(synth.pcm_to_float(pcm)))
        params.append(dict(hi_val=hi, **p))
    return caps, params


@pytest.mark.parametrize("rate,n", [(2e6, 12), (13.56e6, 6)])
def test_batch_matches_oracle_per_capture(rate, n):
    caps, params = _captures(rate, n, 3000)
    got = batch.decode_batch(caps, rate, params, workers=4)
    assert len(got) == n
    total = 0
    for x, prm, (fr, bits) in zip(caps, params, got):
        want = oracle.decode_capture(x, rate, **prm)
        assert len(fr) == len(want["frames"])
        for f in ("pos", "nbits", "type"):
            assert np.array_equal(fr[f], want["frames"][f]), f
        for (pos, typ, b), wb in zip(batch.frames_as_lists(fr, bits), want["frame_bits"]):
            assert np.array_equal(b, wb)
        total += len(fr)
    assert total > 0


def _uniform_captures(rate, n, n_items, seed0):
    """n captures of exactly n_items items: session traffic cut where it falls (some captures end inside a frame)."""
    sess = synth.load_sessions()
    p = synth.rate_params(rate)
    caps, his, los = [], [], []
    for i in range(n):
        name = "ultralight" if i % 3 == 0 else "classic1k"
        ch = synth.Channel(pause=0.02 + 0.005 * (i % 4), tag_high=1.06 + 0.01 * (i % 3), fade=0.02 * (i % 2))
        pcm = synth.capture(sess[name], rate, seed0 + i, channel=ch, av_window=p["av_window"], sessions=2)
        x = synth.envelope(synth.pcm_to_float(pcm))
        if i % 5 == 0:  # the capture starts inside the traffic: pauses and load modulation in its warm-up window
            x = x[p["av_window"] + 6000 + 997 * (i % 7):]
        reps = int(np.ceil(n_items / x.size))
        caps.append(np.tile(x, reps)[:n_items] if reps > 1 else x[:n_items])
        his.append([1.05, 1.06, 1.07, 1.08, 1.09, 1.10][i % 6])
        los.append([0.1, 0.08, 0.12][i % 3])
    return np.stack(caps).astype(np.float32), np.array(los), np.array(his), p


def _check_batch(x2d, rate, p, los, his, res):
    got = batch.split_captures(res, x2d.shape[0])
    total = 0
    for i, (fr, bits) in enumerate(got):
        want = oracle.decode_capture(x2d[i], rate, lo_val=float(los[i]), hi_val=float(his[i]), **p)
        assert len(fr) == len(want["frames"]), (i, len(fr), len(want["frames"]))
        for f in ("pos", "nbits", "type"):
            assert np.array_equal(fr[f], want["frames"][f]), (i, f)
        for (pos, typ, b), wb in zip(batch.frames_as_lists(fr, bits), want["frame_bits"]):
            assert np.array_equal(b, wb), i
        total += len(fr)
    return total


def test_batch_one_pass_matches_oracle_per_capture():
    """64 mixed captures (Ultralight / Classic, own lo_val and hi_val, some cut inside a frame, some starting inside one) in
    one pass of the device; every capture equals the oracle's decode of that capture alone."""
    rate, n, n_items = 13.56e6, 64, 262144 + 1000
    x2d, los, his, p = _uniform_captures(rate, n, n_items, 7000)
    res = batch.decode_batch_onepass(x2d, rate, p, lo_vals=los, hi_vals=his)
    assert res["pitch"] is not None and res["pitch"] % 4096 == 0 and res["pitch"] >= n_items
    assert (res["frames"]["pos"] % res["pitch"] >= p["av_window"]).all() and (res["frames"]["pos"] % res["pitch"] < n_items).all()
    assert _check_batch(x2d, rate, p, los, his, res) > 500
    st = res["stream"].stats()
    assert st["slicer_kernel_launches"] == 1  # one launch of the streaming slicer for the whole batch
    import os
    assert st["pipe_tiles"] > 0 or os.environ.get("NFC_SLICER_PIPE") == "0"  # (the pipelined mode, unless it is switched off)
    # the same stream again, device-resident input, other thresholds
    import torch
    xd = torch.from_numpy(x2d).cuda()
    res2 = batch.decode_batch_onepass(xd, rate, p, lo_vals=los[::-1].copy(), hi_vals=his[::-1].copy(), stream=res["stream"])
    assert _check_batch(x2d, rate, p, los[::-1], his[::-1], res2) > 500
    res["stream"].close()


def test_batch_one_pass_spans_slabs():
    """More captures than one slab of 2^30 positions holds: slabs are cut at capture boundaries."""
    rate, n_items = 13.56e6, 3_000_000
    n = 5  # pitch 3002368: 357 captures per slab would be needed to overflow; force small slabs through a long pitch instead
    x2d, los, his, p = _uniform_captures(rate, n, n_items, 7100)
    res = batch.decode_batch_onepass(x2d, rate, p, lo_vals=los, hi_vals=his)
    assert _check_batch(x2d, rate, p, los, his, res) > 50
    res["stream"].close()


def test_batch_one_pass_refuses_what_needs_the_sequential_path():
    rate, n, n_items = 13.56e6, 4, 131072
    x2d, los, his, p = _uniform_captures(rate, n, n_items, 7200)
    x2d[2, 60000] = -1.0  # a negative envelope sample can only be admitted on the sequential path's terms
    x2d[2, 60001:60200] = 0.0
    res = batch.decode_batch_onepass(x2d, rate, p, lo_vals=los, hi_vals=his)
    got = batch.split_captures(res, n)
    for i, (fr, bits) in enumerate(got):
        want = oracle.decode_capture(x2d[i], rate, lo_val=float(los[i]), hi_val=float(his[i]), **p)
        assert len(fr) == len(want["frames"]) and np.array_equal(fr["pos"], want["frames"]["pos"]), i
    res["stream"].close()


def test_batch_in_several_slabs_matches_oracle_per_capture():
    """The batch's extraction / runs / line code in slabs of a few captures each (slab_len shortens them), their chains queued
    back to back: what a slab inherits from the one before must not show in any capture's frames."""
    rate, n, n_items = 13.56e6, 23, 262144 + 1000
    x2d, los, his, p = _uniform_captures(rate, n, n_items, 9100)
    s = _cabi.Stream(rate, outputs=_cabi.OUT_FRAMES, **p)
    res0 = batch.decode_batch_onepass(x2d, rate, p, lo_vals=los, hi_vals=his, stream=s)  # one slab: also sets the sizing rates
    pitch = res0["pitch"]
    total0 = _check_batch(x2d, rate, p, los, his, res0)
    s.release_frames()
    s.set_tuning(slab_len=3 * pitch)  # eight slabs
    res = batch.decode_batch_onepass(x2d, rate, p, lo_vals=los, hi_vals=his, stream=s)
    assert res["pitch"] == pitch
    assert _check_batch(x2d, rate, p, los, his, res) == total0 > 150
    s.close()
