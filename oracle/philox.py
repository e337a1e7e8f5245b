"""Test infrastructure (never imported by the product): numpy restatement of the counter-based generator the seeded NeuTex
render draws its jitter from (csrc/ngf_neutex.cu: jitter_uniform) — Philox 4x32 with 10 rounds (Salmon, Moraes, Dror, Shaw:
"Parallel random numbers: as easy as 1, 2, 3", SC'11), pinned by the Random123 known-answer vectors in
tests/test_philox.py.  The reference itself draws torch.rand(R, 64) inside cube_ray_generation
(UV-Mapping/model/renderer.py:113-118); any U[0,1) stream is a valid stand-in, this one is reproducible across devices."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: 4 arrays (or scalars) of uint32, key: 2 -> 4 arrays of uint32."""
    c = [np.asarray(x, dtype=np.uint64) & MASK for x in ctr]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        h0, l0, h1, l1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [h1 ^ c[1] ^ np.uint64(k0), l1, h0 ^ c[3] ^ np.uint64(k1), l0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in c]


def jitter_uniform(seed: int, idx):
    """U[0,1) number `idx` (uint64 array) of stream `seed`: word idx & 3 of block idx >> 2, top 24 bits."""
    idx = np.asarray(idx, dtype=np.uint64)
    blk = idx >> np.uint64(2)
    out = philox4x32_10([blk & MASK, blk >> np.uint64(32), np.zeros_like(blk), np.zeros_like(blk)],
                        [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF])
    sel = (idx & np.uint64(3)).astype(np.int64)
    x = np.choose(sel, out)
    return (x >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def neutex_noise(seed: int, first_ray: int, n_rays: int, sample_num: int):
    """The [n_rays, sample_num] numbers a seeded render of frame rays first_ray.. draws (ngf_neutex_noise)."""
    ray = np.arange(first_ray, first_ray + n_rays, dtype=np.uint64)[:, None]
    i = np.arange(sample_num, dtype=np.uint64)[None, :]
    return jitter_uniform(seed, ray * np.uint64(64) + i)
