"""numpy twin of csrc/philox.cuh (Philox4x32-10 -> FP64 uniform), used to check on-device RNG bit-for-bit."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox_uniform_f64(seed, index):
    index = np.asarray(index, dtype=np.uint64)
    c0 = (index & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    c1 = (index >> np.uint64(32)).astype(np.uint32)
    c2 = np.zeros_like(c0)
    c3 = np.zeros_like(c0)
    k0 = np.uint32(seed & 0xFFFFFFFF)
    k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n1 = (p1 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            n3 = (p0 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    bits = ((c0.astype(np.uint64) << np.uint64(32)) | c1.astype(np.uint64)) >> np.uint64(11)
    return bits.astype(np.float64) * (1.0 / 9007199254740992.0)
