"""Host model of the kernels' measurement-noise draw (csrc/philox.cuh), for the parity tests.

  philox4x32_10   Salmon et al. 2011, vectorised over counters (numpy uint64 arithmetic)
  icdf_table_eval the device algorithm restated step by step on the committed coefficient table (csrc/icdftab.inc)
  normals         z[state i] of env gid at time index t: word (i & 1) + 2 (gid & 1) of the Philox block with counter
                  (gid >> 1 lo, gid >> 1 hi, t, i >> 1) and key = seed, mapped by z = Phi^-1((w + 1/2) 2^-32)
The independent reference for the transform is scipy.special.ndtri, not the table.
"""
import os

import numpy as np
from scipy.special import ndtri

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, key):
    c = [np.asarray(v, dtype=np.uint64) & MASK for v in (c0, c1, c2, c3)]
    k0, k1 = int(key) & 0xFFFFFFFF, (int(key) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & MASK, (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & MASK]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c


def normal_from_word(w):
    """z = Phi^-1((w + 1/2) 2^-32), evaluated on the folded tail so both tails keep full relative accuracy."""
    w = np.asarray(w, dtype=np.uint64)
    neg = w < np.uint64(1 << 31)
    wp = np.where(neg, w, np.uint64(0xFFFFFFFF) - w)
    z = ndtri((wp.astype(np.float64) + 0.5) * 2.0 ** -32)
    return np.where(neg, z, -z)


def load_table():
    rows = []
    for ln in open(os.path.join(ROOT, "pcgym_b200", "csrc", "icdftab.inc")):
        ln = ln.strip()
        if ln and not ln.startswith("//"):
            rows.append([float.fromhex(v) for v in ln.rstrip(",").split(",")])
    return np.array(rows)                                  # [512, 6]


def icdf_table_eval(w, tab=None):
    """The device algorithm (normal_icdf_u32) in numpy: fold, leading-bit segment, centred remainder, Horner."""
    tab = load_table() if tab is None else tab
    w = np.asarray(w, dtype=np.uint64)
    flip = w >= np.uint64(1 << 31)
    wp = np.where(flip, np.uint64(0xFFFFFFFF) - w, w)
    m = wp * np.uint64(2) + np.uint64(1)
    b = np.floor(np.log2(m.astype(np.float64))).astype(np.int64)          # leading-one index (m < 2^32: exact in fp64)
    mn = (m << (31 - b).astype(np.uint64)) & MASK
    idx = b * 16 + ((mn >> np.uint64(27)) & np.uint64(15)).astype(np.int64)
    v = (((mn << np.uint64(5)) & MASK) ^ np.uint64(0x80000000)).astype(np.int64)
    v = np.where(v >= (1 << 31), v - (1 << 32), v).astype(np.float64)     # as signed 32-bit
    c = tab[idx]
    z = c[:, 5]
    for k in range(4, -1, -1):
        z = z * v + c[:, k]
    return np.where(flip, -z, z)


def normals(seed, gid, t, nx):
    """z [nx, len(gid)] for global env ids `gid` at time index t (the value the kernel multiplies into state i)."""
    gid = np.asarray(gid, dtype=np.uint64)
    pid = gid >> np.uint64(1)
    odd = (gid & np.uint64(1)).astype(bool)
    out = np.empty((nx, gid.shape[0]))
    for j in range((nx + 1) // 2):
        r = philox4x32_10(pid & MASK, pid >> np.uint64(32), np.full_like(pid, t), np.full_like(pid, j), seed)
        w0 = np.where(odd, r[2], r[0])
        w1 = np.where(odd, r[3], r[1])
        out[2 * j] = normal_from_word(w0)
        if 2 * j + 1 < nx:
            out[2 * j + 1] = normal_from_word(w1)
    return out
