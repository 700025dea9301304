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
    # u52 = low 52 bits of the lane
    assert np.array_equal(u, (lanes & np.uint64((1 << 52) - 1)).astype(np.float64) * 2.0 ** -52)


def test_streams_are_distinct(og):
    a, _ = og.rng_stream(7, 0, 1, 3, 9, 16)
    for args in [(8, 0, 1, 3, 9), (7, 1, 1, 3, 9), (7, 0, 2, 3, 9), (7, 0, 1, 4, 9), (7, 0, 1, 3, 10)]:
        b, _ = og.rng_stream(*args, 16)
        assert not np.any(a == b)
    # draws are addressed, not sequential state: a prefix of a longer request is identical
    c, _ = og.rng_stream(7, 0, 1, 3, 9, 5)
    assert np.array_equal(a[:5], c)
