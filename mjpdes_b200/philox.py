"""Philox4x32-10 in numpy: the host side of the substream contract of
``include/mjpdes_b200.h`` (draw #idx of consumer c of instance g = words (0,1) of
Philox4x32-10(key = seed, counter = {idx, c, g_lo, g_hi})).  Used by the host compile step
for per-replication ``:bounds`` parameters (rand-range, core.clj:547-549)."""
from __future__ import annotations

import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised over the counter words (uint32 arrays or scalars); returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.atleast_1d(np.asarray(c, dtype=np.uint64)) & _MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)) & _MASK, p1 & _MASK, \
                         ((p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)) & _MASK, p0 & _MASK
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


def uniform53(seed: int, instance_ids, consumer: int, idx: int) -> np.ndarray:
    """[0,1) doubles, one per instance id: the stand-in for clojure.core/rand."""
    g = np.asarray(instance_ids, dtype=np.uint64)
    o0, o1, _, _ = philox4x32_10(np.uint64(idx), np.uint64(consumer), g & _MASK, g >> np.uint64(32),
                                 seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    k = ((o0.astype(np.uint64) << np.uint64(32)) | o1.astype(np.uint64)) >> np.uint64(11)
    return k.astype(np.float64) * 2.0 ** -53
