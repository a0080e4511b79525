"""Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) in numpy — the
host restatement of the generator the device-side encrypt / keygen kernels use (tfhe.jl_b200/csrc/keygen.cuh), pinned
on the known-answer vectors of the Random123 distribution (tests/test_keygen_words.py)."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr, key):
    """ctr: uint32 [..., 4], key: uint32 [..., 2] -> uint32 [..., 4]"""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k = [np.asarray(key[..., i], dtype=np.uint64) for i in range(2)]
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0]
        k = [(k[0] + np.uint64(W0)) & mask, (k[1] + np.uint64(W1)) & mask]
    return np.stack(c, axis=-1).astype(np.uint32)


def words(seed: int, stream: int, count: int) -> np.ndarray:
    """word i of stream (seed, stream) = lane i & 3 of the block with counter i >> 2 (keygen.cuh, philox_block)"""
    blocks = np.arange((count + 3) // 4, dtype=np.uint64)
    ctr = np.stack([blocks & np.uint64(0xFFFFFFFF), blocks >> np.uint64(32),
                    np.full_like(blocks, stream & 0xFFFFFFFF), np.full_like(blocks, stream >> 32)], axis=-1).astype(np.uint32)
    key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32), (blocks.size, 2))
    return philox4x32_10(ctr, key).reshape(-1)[:count].view(np.int32)
