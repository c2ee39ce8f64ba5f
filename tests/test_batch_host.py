"""Host logic of the batch driver (no GPU)."""
from usrp_nfc_b200 import batch


def test_rank_share_is_a_partition():
    for world in (1, 2, 3, 8):
        seen = sorted(i for r in range(world) for i in batch.rank_share(4096, r, world))
        assert seen == list(range(4096))
        sizes = [len(batch.rank_share(4096, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_bench_same_frames_compares_records_and_bits():
    """bench.py's full-size self-check compares two decodes field by field."""
    import numpy as np
    import bench
    from usrp_nfc_b200 import _cabi
    fr = np.zeros(3, dtype=_cabi.FRAME_DTYPE)
    fr["pos"], fr["nbits"], fr["type"], fr["bit_off"] = [10, 20, 30], [2, 1, 2], [0, 1, 0], [0, 0, 2]
    b0, b1 = np.array([1, 0, 1, 1], np.uint8), np.array([1], np.uint8)
    assert bench.same_frames((fr, b0, b1), (fr.copy(), b0.copy(), b1.copy()))
    other = fr.copy()
    other["pos"][1] += 1
    assert not bench.same_frames((fr, b0, b1), (other, b0, b1))
    flipped = b0.copy()
    flipped[2] ^= 1
    assert not bench.same_frames((fr, b0, b1), (fr, flipped, b1))
    assert not bench.same_frames((fr, b0, b1), (fr[:2], b0, b1))
