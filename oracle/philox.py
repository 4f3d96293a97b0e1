"""Counter-based uniform draws (Philox4x32-10), numpy restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference draws its uniforms with ``torch.rand((B,P,9,H,W,M))``
(/root/reference/scripts/LDATI.py:169-171), whose stream depends on the batch
chunking and on M = max count.  "Identical injected draws" are therefore defined
here as a pure function of the *global* event coordinates, and the CUDA kernel
(v2ce_toolbox_b200/csrc/ldati.cu: ``philox_uniform``) evaluates the same function:

    idx   = ((frame*2 + p)*9 + c) * (H*W) + (h*W + w)          (uint64)
    ctr   = (lo32(idx), hi32(idx), j >> 2, 0)
    key   = (lo32(seed), hi32(seed))
    word  = Philox4x32-10(ctr, key)[j & 3]
    u     = float32(word >> 8) * 2**-24                          in [0, 1)

Philox4x32-10 is the published Random123 generator (Salmon et al., SC'11):
multipliers 0xD2511F53 / 0xCD9E8D57, Weyl constants 0x9E3779B9 / 0xBB67AE85.
Known-answer vectors from the Random123 distribution (kat_vectors) are checked
in tests/test_oracle_vs_golden.py.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_SH32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32-valued arrays
    (held in uint64 for the 32x32->64 multiplies).  Returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _SH32, p0 & _MASK
        hi1, lo1 = p1 >> _SH32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def uniform_from_index(idx, j, seed, stream=0):
    """u[idx, j] as float32 in [0,1).  idx: uint64 array, j: int array (same shape).  `stream` is the fourth counter
    word: 0 for LDATI and the baseline samplers' integer-part draws, 1 / 2 for their fractional-part timestamp and
    Bernoulli draws (oracle/baseline_oracle.py)."""
    idx = np.asarray(idx, dtype=np.uint64)
    j = np.asarray(j, dtype=np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    w = philox4x32_10(idx & _MASK, idx >> _SH32, j >> np.uint64(2), np.uint64(int(stream)),
                      seed & 0xFFFFFFFF, seed >> 32)
    sel = (j & np.uint64(3)).astype(np.int64)
    word = np.choose(sel, w)
    return (word >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def event_index(frame, p, c, pix, hw):
    """Global pixel-bin index used as the Philox counter (all uint64 arrays)."""
    frame = np.asarray(frame, dtype=np.uint64)
    p = np.asarray(p, dtype=np.uint64)
    c = np.asarray(c, dtype=np.uint64)
    pix = np.asarray(pix, dtype=np.uint64)
    return ((frame * np.uint64(2) + p) * np.uint64(9) + c) * np.uint64(hw) + pix


def dense_draws(frame_base, B, H, W, M, seed):
    """The dense (B,2,9,H,W,M) tensor the reference's torch.rand call would hold
    if it served these draws -- used to inject draws into the reference."""
    hw = H * W
    f = (np.arange(B, dtype=np.uint64) + np.uint64(frame_base)).reshape(B, 1, 1, 1)
    p = np.arange(2, dtype=np.uint64).reshape(1, 2, 1, 1)
    c = np.arange(9, dtype=np.uint64).reshape(1, 1, 9, 1)
    pix = np.arange(hw, dtype=np.uint64).reshape(1, 1, 1, hw)
    idx = event_index(f, p, c, pix, hw)                     # (B,2,9,HW)
    out = np.empty((B, 2, 9, hw, M), dtype=np.float32)
    for j in range(M):
        out[..., j] = uniform_from_index(idx, np.full(idx.shape, j, dtype=np.uint64), seed)
    return out.reshape(B, 2, 9, H, W, M)


# Random123 known-answer vectors for philox4x32-10 (kat_vectors file of the
# Random123 distribution): (counter, key) -> output.
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]
