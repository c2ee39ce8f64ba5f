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
