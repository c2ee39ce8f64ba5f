"""Host logic of the batch driver (no GPU)."""
from usrp_nfc_b200 import batch


def test_rank_share_is_a_partition():
    for world in (1, 2, 3, 8):
        seen = sorted(i for r in range(world) for i in batch.rank_share(4096, r, world))
        assert seen == list(range(4096))
        sizes = [len(batch.rank_share(4096, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
