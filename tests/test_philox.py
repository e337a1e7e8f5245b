"""CPU: the Philox 4x32-10 restatement (oracle/philox.py) against the Random123 known-answer vectors (kat_vectors of the
Random123 distribution, `philox4x32 10` lines), and the layout of the NeuTex jitter stream built on it."""
import numpy as np

from oracle import philox as P

KAT = [
    ((0x00000000,) * 4, (0x00000000,) * 2, (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers():
    for ctr, key, want in KAT:
        got = P.philox4x32_10(ctr, key)
        assert tuple(int(x) for x in got) == want


def test_jitter_stream_layout():
    # number idx is word idx & 3 of block idx >> 2; seed 0, block 0 is the first known answer
    u = P.jitter_uniform(0, np.arange(4, dtype=np.uint64))
    want = np.array([w >> 8 for w in KAT[0][2]], dtype=np.float32) / np.float32(16777216.0)
    assert np.array_equal(u, want)
    a = P.neutex_noise(9, 0, 8, 48)
    b = P.neutex_noise(9, 5, 3, 48)
    assert a.shape == (8, 48) and np.array_equal(a[5:8], b)            # rays index the stream by frame position, 64 slots each
    big = P.neutex_noise(3, 0, 4096, 64)
    assert 0.0 <= big.min() and big.max() < 1.0 and abs(big.mean() - 0.5) < 2e-3
