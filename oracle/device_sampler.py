"""Oracle: the DEVICE pair stream restated in numpy.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  The reference samples its (anchor, reference) pairs with
numpy's global generator (`cellulus/datasets/zarr_dataset.py:177-251`); the device sampler
(`cellulus_b200/csrc/pair_stream.cuh`) draws the SAME DISTRIBUTION from a counter-based Philox4x32-10
stream.  This file restates that stream independently of the CUDA code, so that the two kernels that
consume it (`sample_pairs_kernel`, `oce_loss_sampled_kernel`) can be held bit-exact to a CPU statement of
which pairs a `(seed, sequence)` names.  The distribution itself is pinned against the reference's sampler in
`tests/` (bounds, np.repeat run structure, open ball minus the origin, chi-square uniformity).

Philox4x32-10 is the published algorithm of Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"
(SC'11): multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key increments 0x9E3779B9 / 0xBB67AE85, ten rounds.
"""

from __future__ import annotations

import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter: np.ndarray, stream: int, seed: int) -> np.ndarray:
    """Vectorised over `counter` (uint64 array): returns (n, 4) uint32 words.
    Counter words (c0, c1) = counter lo / hi, (c2, c3) = stream lo / hi; key = seed lo / hi."""
    counter = np.asarray(counter, dtype=np.uint64)
    c0 = counter & _MASK
    c1 = counter >> np.uint64(32)
    c2 = np.full_like(c0, np.uint64(stream & 0xFFFFFFFF))
    c3 = np.full_like(c0, np.uint64((stream >> 32) & 0xFFFFFFFF))
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0  # 32 x 32 -> 64 bit products (operands < 2^32, no overflow of uint64)
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def _mulhi(word: np.ndarray, n: int) -> np.ndarray:
    """Bounded integer in [0, n) from a 32-bit word: the high half of word * n."""
    return ((word.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def offset_table(kappa: float, num_dims: int) -> np.ndarray:
    """Admissible offsets: integer points of [-trunc(kappa), trunc(kappa)]^D with sum o^2 < kappa^2 and o != 0
    (`in_circle` / `not_zero`, zarr_dataset.py:179-191), enumerated with column 0 fastest."""
    kap = int(kappa)
    side = 2 * kap + 1
    c = np.arange(side**num_dims)
    cols = []
    for _ in range(num_dims):
        cols.append(c % side - kap)
        c = c // side
    cand = np.stack(cols, axis=1)
    keep = ((cand**2).sum(1) < kappa**2) & (np.abs(cand).sum(1) > 0)
    return cand[keep]


def sample_pairs(batch: int, extent_xyz, kappa: float, num_anchors: int, num_references: int, seed: int,
                 sequence: int = 0):
    """(anchors, refs), each (batch, num_anchors * num_references, D) int64, columns (x, y[, z])."""
    D = len(extent_xyz)
    kap = int(kappa)
    A, R = int(num_anchors), int(num_references)
    table = offset_table(kappa, D)
    q = 2 if len(table) <= 1024 else 1  # bounded draws per 32-bit word (pair_stream.cuh)
    n_tg = (R + 4 * q - 1) // (4 * q)
    b = np.repeat(np.arange(batch, dtype=np.uint64), A)
    a = np.tile(np.arange(A, dtype=np.uint64), batch)
    ba = b * np.uint64(A) + a
    words = philox4x32_10(ba, 2 * sequence, seed)
    anchors = np.stack([kap + _mulhi(words[:, k], int(extent_xyz[k]) - 2 * kap + 1) for k in range(D)], axis=1)
    tg = np.arange(n_tg, dtype=np.uint64)
    counters = (ba[:, None] * np.uint64(n_tg) + tg[None, :]).reshape(-1)
    ow = philox4x32_10(counters, 2 * sequence + 1, seed).reshape(batch * A, n_tg, 4)
    if q == 2:  # second draw of a word: the low half of the first product is the next uniform word
        second = ((ow.astype(np.uint64) * np.uint64(len(table))) & _MASK).astype(np.uint32)
        ow = np.concatenate([ow, second], axis=2)
    ow = ow.reshape(batch * A, n_tg * 4 * q)[:, :R]
    offs = table[_mulhi(ow.reshape(-1), len(table))].reshape(batch * A, R, D)
    anchors_rep = np.repeat(anchors[:, None, :], R, axis=1)
    refs = anchors_rep + offs
    return anchors_rep.reshape(batch, A * R, D).astype(np.int64), refs.reshape(batch, A * R, D).astype(np.int64)
