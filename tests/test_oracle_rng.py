"""The oracle's Philox4x32-10 against the Random123 known-answer vectors, and
the stream conventions of oracle/og_rng.hpp."""
import numpy as np


def test_philox_kat(og):
    # Random123 kat_vectors, philox4x32 10 rounds
    assert og.philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert og.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert og.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_uniform_range_and_moments(og):
    u, lanes = og.rng_stream(7, 0, 1, 3, 9, 200000)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 5e-3 and abs(u.var() - 1 / 12) < 2e-3
    # every draw owns 52 bits; the lane carries them on top (what Random.int multiplies)
    assert np.array_equal(u, (lanes >> np.uint64(12)).astype(np.float64) * 2.0 ** -52)
    assert not np.any(lanes & np.uint64(0xFFF))


def test_eleven_draws_per_five_blocks(og):
    """The stream layout of oracle/og_rng.hpp: draws 0..9 of a group are the lanes of five Philox blocks, draw 10 is
    assembled from the 12 top bits of the first words of blocks 0..2 (bits no other draw uses)."""
    seed, epoch, purpose, g, step = 11, 3, 1, 123456789, 77
    _, lanes = og.rng_stream(seed, epoch, purpose, g, step, 33)
    key = og.philox([epoch & 0xFFFFFFFF, epoch >> 32, 0x6d636d63, 0], [seed & 0xFFFFFFFF, seed >> 32])[:2]
    c3 = ((g >> 32) & 0xFFFF) | ((purpose & 0xFF) << 16) | (((step >> 32) & 0xFF) << 24)
    for m in range(3):
        blocks = [og.philox([5 * m + k, step & 0xFFFFFFFF, g & 0xFFFFFFFF, c3], key) for k in range(5)]
        want = []
        for k in range(5):
            w = blocks[k]
            want.append(((w[0] & 0xFFFFF) << 32) | w[1])
            want.append(((w[2] & 0xFFFFF) << 32) | w[3])
        s = [blocks[0][0] >> 20, blocks[0][2] >> 20, blocks[1][0] >> 20, blocks[1][2] >> 20, blocks[2][0] >> 20]
        want.append((s[0] << 40) | (s[1] << 28) | (s[2] << 16) | (s[3] << 4) | (s[4] >> 8))
        got = [int(x) >> 12 for x in lanes[11 * m:11 * m + 11]]
        assert got == want


def test_streams_are_distinct(og):
    a, _ = og.rng_stream(7, 0, 1, 3, 9, 16)
    for args in [(8, 0, 1, 3, 9), (7, 1, 1, 3, 9), (7, 0, 2, 3, 9), (7, 0, 1, 4, 9), (7, 0, 1, 3, 10)]:
        b, _ = og.rng_stream(*args, 16)
        assert not np.any(a == b)
    # draws are addressed, not sequential state: a prefix of a longer request is identical
    c, _ = og.rng_stream(7, 0, 1, 3, 9, 5)
    assert np.array_equal(a[:5], c)
