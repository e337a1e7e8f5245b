"""CPU: the row permutation the TriPlane colour kernel stores its layer-1 operand with (csrc/ngf_mlp.cuh: tile_row /
tile_sample) is a permutation of every aligned group of 8 rows, its inverse is the one the epilogue uses, and it makes the
eight 16-byte stores of every quarter warp of the gather hit eight different shared-memory banks (bank = chunk + row mod 8
with the padded K-group stride)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "neural-gauge-fields_b200", "csrc", "ngf_mlp.cuh")).read()


def tile_row(g):
    return (g & ~7) | ((((g >> 2) & 1) + 6 * (g & 3)) & 7)


def tile_sample(row):
    res = row & 7
    p = res & 1
    return (row & ~7) | (p << 2) | (((8 - (res - p)) >> 1) & 3)


def test_python_mirror_matches_the_source():
    assert "return PERM ? (g & ~7) | ((((g >> 2) & 1) + 6 * (g & 3)) & 7) : g;" in SRC
    assert "return (row & ~7) | (p << 2) | (((8 - (res - p)) >> 1) & 3);" in SRC
    assert re.search(r"kLboA = kTileM \* 16 \+ 16;", SRC)          # one 16-byte slot of padding per K group


def test_permutation_and_inverse():
    rows = [tile_row(g) for g in range(128)]
    assert sorted(rows) == list(range(128))
    assert all(tile_sample(tile_row(g)) == g for g in range(128))
    assert all(tile_row(g) >> 3 == g >> 3 for g in range(128))       # stays inside its 8-row core matrix


def test_quarter_warp_stores_are_bank_conflict_free():
    lbo = 128 * 16 + 16
    for plane in range(3):
        for q in range(0, 768, 8):                                   # a quarter warp = 8 consecutive (sample, chunk) items
            banks = set()
            for it in range(q, q + 8):
                m, chunk = it // 6, it % 6
                addr = (plane * 6 + chunk) * lbo + tile_row(m) * 16
                banks.add((addr // 16) % 8)
            assert len(banks) == 8, (plane, q)
    # the identity mapping (rows in sample order) collides: that is what the permutation removes
    worst = 0
    for q in range(0, 768, 8):
        banks = [(((it % 6) * lbo + (it // 6) * 16) // 16) % 8 for it in range(q, q + 8)]
        worst = max(worst, 8 - len(set(banks)))
    assert worst >= 2
