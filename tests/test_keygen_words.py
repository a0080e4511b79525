"""CPU: the numpy restatement of Philox4x32-10 against the Random123 known-answer vectors (kat_vectors, philox4x32 10)."""
import numpy as np

from philox_ref import philox4x32_10, words


def test_philox_known_answers():
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        got = philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
        assert tuple(int(v) for v in got) == want


def test_word_stream_layout():
    w = words(0, 0, 9).view(np.uint32)
    assert tuple(int(v) for v in w[:4]) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert np.array_equal(words(7, 3, 9)[:5], words(7, 3, 5))
    assert not np.array_equal(words(7, 3, 8), words(7, 4, 8)) and not np.array_equal(words(7, 3, 8), words(8, 3, 8))
