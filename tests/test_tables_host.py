"""The line-code tables built by csrc/tables.cpp (host code of the product, no GPU needed) drive a
table-only automaton that must reproduce the reference decoders' outputs on the decoder known answers."""
import numpy as np
import pytest

from oracle import oracle
from tests import helpers as H
from usrp_nfc_b200 import _cabi


def _run_tables(evs, samp_rate, max_len):
    dcg, tg = _cabi.build_tables(samp_rate, max_len, 0)
    dcm, tm = _cabi.build_tables(samp_rate, max_len, 1)
    ms = gs = 0
    out = []
    for v, d, t in evs:
        if t == 1:
            e = int(tm[dcm[d], v + 1, ms])
            ms = e & 15
            n = (e >> 4) & 3
            if n > 0:
                out.append((1, (e >> 6) & 7))
            if n > 1:
                out.append((1, (e >> 9) & 7))
        elif t == 0:
            e = int(tg[dcg[d], v + 1, gs])
            gs = e & 7
            if (e >> 4) & 3:
                out.append((0, (e >> 6) & 7))
    return out


def test_tables_reproduce_decoder_kats():
    z = H.load_case("decoder_kat")
    for ci in range(40):
        evs = z["ev%d" % ci]
        got = _run_tables(evs.tolist(), 2e6, 50)
        assert got == [tuple(r) for r in z["sym%d" % ci].tolist()], ci


@pytest.mark.parametrize("rate,mx", [(2e6, 50), (13.56e6, 339), (20e6, 500), (1e6, 7), (2e6, 10)])
def test_tables_match_oracle_on_random_events(rate, mx):
    rng = np.random.default_rng(int(rate) % 1000 + mx)
    n = 3000
    evs = np.stack([rng.integers(-1, 3, n), rng.integers(1, mx + 1, n), rng.choice([-1, 0, 1], n)], 1)
    # bias durations towards the interesting region (a few bit periods)
    bitp = 9.44 * rate / 1e6
    short = rng.random(n) < 0.8
    evs[short, 1] = np.clip(rng.integers(1, max(2, int(2.2 * bitp)), short.sum()), 1, mx)
    ev = np.zeros(n, oracle.EVENT_DTYPE)
    ev["v"], ev["d"], ev["type"], ev["pos"] = evs[:, 0], evs[:, 1], evs[:, 2], np.arange(n)
    dec = oracle.Decoders(True, True)
    dec.feed(ev, 1e6 / rate)
    sym = dec.symbols()
    want = list(zip(sym["type"].tolist(), sym["val"].tolist()))
    assert _run_tables(evs.tolist(), rate, mx) == want
    dcm, tm = _cabi.build_tables(rate, mx, 1)
    assert tm.shape[0] <= 24
