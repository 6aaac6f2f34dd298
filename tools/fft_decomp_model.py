"""NumPy model of the 3-pass register FFT used by the fused 128x128 kernels.

Verifies the index algebra (which sub-DFT each pass performs, twiddles, spectral index
held by each thread after the forward transform, exchange-buffer addressing and its
bank-conflict freedom) before it is written in CUDA.  Not part of the product path.
"""
import numpy as np

N = 128
w = np.exp(-2j * np.pi / N)


def forward(x):
    """x[r,c] -> X[kr,kc] via P1 (32-pt over c high bits), P2 (4-pt c low, 8-pt r high), P3 (16-pt r low)."""
    # c = 4*n1c + n2c ; kc = k1c + 32*k2c
    a = x.reshape(N, 32, 4)  # [r, n1c, n2c]
    a = np.fft.fft(a, axis=1)  # P1 -> [r, k1c, n2c]
    k1c = np.arange(32)[None, :, None]
    n2c = np.arange(4)[None, None, :]
    a = a * w ** (n2c * k1c)  # P2 step 1
    a = np.fft.fft(a, axis=2)  # P2 step 2 -> [r, k1c, k2c]
    # r = 16*n1r + n2r ; kr = k1r + 8*k2r
    a = a.reshape(8, 16, 32, 4)  # [n1r, n2r, k1c, k2c]
    a = np.fft.fft(a, axis=0)  # P2 step 3 -> [k1r, n2r, k1c, k2c]
    k1r = np.arange(8)[:, None, None, None]
    n2r = np.arange(16)[None, :, None, None]
    a = a * w ** (n2r * k1r)  # P2 step 4
    a = np.fft.fft(a, axis=1)  # P3 -> [k1r, k2r, k1c, k2c]
    X = np.empty((N, N), complex)
    k1r, k2r, k1c, k2c = np.meshgrid(np.arange(8), np.arange(16), np.arange(32), np.arange(4), indexing="ij")
    X[k1r + 8 * k2r, k1c + 32 * k2c] = a
    return X


def t12_idx(r, n2c, k1c):
    return ((k1c * 32 + (r >> 2)) * 16) + ((r & 3) << 2) + (n2c ^ (k1c & 3))


def t23_idx(k1c, n2r, k2c, k1r):
    hi = ((k1r * 4 + k2c) * 8 + (k1c >> 2)) * 4 + (n2r >> 2)
    return hi * 16 + (k1c & 3) + (((n2r & 3) ^ ((k1c >> 2) & 3)) << 2)


def nat_idx(r, c):
    return r * 128 + (c ^ ((r & 3) << 2))


def check_bijection_and_banks():
    # T12 bijection
    r, n2c, k1c = np.meshgrid(np.arange(128), np.arange(4), np.arange(32), indexing="ij")
    assert len(np.unique(t12_idx(r, n2c, k1c))) == 128 * 128
    k1c_, n2r, k2c, k1r = np.meshgrid(np.arange(32), np.arange(16), np.arange(4), np.arange(8), indexing="ij")
    assert len(np.unique(t23_idx(k1c_, n2r, k2c, k1r))) == 128 * 128
    # bank checks: half-warp (16 lanes) 64-bit accesses must hit 16 distinct idx mod 16
    def ok(idxs):
        return len(set(int(i) % 16 for i in idxs)) == 16
    # P1 lanes: l -> n2c = l&3, r10 = (l>>2)&3
    for k in range(32):
        for rhi in range(32):
            lanes = [t12_idx(rhi * 4 + ((l >> 2) & 3), l & 3, k) for l in range(16)]
            assert ok(lanes)
            lanes = [nat_idx(rhi * 4 + ((l >> 2) & 3), 4 * k + (l & 3)) for l in range(16)]
            assert ok(lanes)
    # P2 lanes: k1c[1:0] = l&3, n2r[1:0] = (l>>2)&3
    for n2c in range(4):
        for n1r in range(8):
            for k1hi in range(8):
                for n2hi in range(4):
                    lanes = [t12_idx(16 * n1r + n2hi * 4 + ((l >> 2) & 3), n2c, k1hi * 4 + (l & 3)) for l in range(16)]
                    assert ok(lanes)
    for k2c in range(4):
        for k1r in range(8):
            for k1hi in range(8):
                for n2hi in range(4):
                    lanes = [t23_idx(k1hi * 4 + (l & 3), n2hi * 4 + ((l >> 2) & 3), k2c, k1r) for l in range(16)]
                    assert ok(lanes)
    # P3 lanes: k1c[3:0] = l
    for n2r in range(16):
        for k2c in range(4):
            for k1r in range(8):
                for k1c4 in range(2):
                    lanes = [t23_idx(k1c4 * 16 + l, n2r, k2c, k1r) for l in range(16)]
                    assert ok(lanes)
    print("bijections and bank checks OK")


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N))
    X = forward(x)
    print("fwd err", np.abs(X - np.fft.fft2(x)).max())
    check_bijection_and_banks()


def ex_idx(k1c, rho, q):
    """Unified exchange layout: T12 uses (rho, q) = (r, n2c); T23 uses (16*k1r + n2r, k2c).
    P2 and P3 then work in place (they write exactly the addresses they read)."""
    return ((k1c * 32 + (rho >> 2)) * 16) + ((((rho & 3) ^ ((k1c >> 2) & 3))) << 2) + (q ^ (k1c & 3))


def check_unified():
    k1c, rho, q = np.meshgrid(np.arange(32), np.arange(128), np.arange(4), indexing="ij")
    assert len(np.unique(ex_idx(k1c, rho, q))) == 128 * 128

    def ok(idxs):
        return len(set(int(i) % 16 for i in idxs)) == 16

    # P1 writer: lanes l -> n2c = l&3, r[1:0] = (l>>2)&3 ; fixed k1c, r[6:2]
    for k in range(32):
        for rhi in range(32):
            assert ok([ex_idx(k, rhi * 4 + ((l >> 2) & 3), l & 3) for l in range(16)])
    # P2: lanes -> k1c[1:0] = l&3, n2r[1:0] = (l>>2)&3 ; fixed (q, n1r/k1r), k1c[4:2], n2r[3:2]
    for q_ in range(4):
        for hi in range(8):
            for k1hi in range(8):
                for n2hi in range(4):
                    assert ok([ex_idx(k1hi * 4 + (l & 3), 16 * hi + n2hi * 4 + ((l >> 2) & 3), q_) for l in range(16)])
    # P3: lanes -> k1c[3:0] = l ; fixed n2r, k2c, k1r, k1c[4]
    for n2r in range(16):
        for k2c in range(4):
            for k1r in range(8):
                for k4 in range(2):
                    assert ok([ex_idx(k4 * 16 + l, 16 * k1r + n2r, k2c) for l in range(16)])
    print("unified exchange layout OK")


if __name__ == "__main__":
    check_unified()


def nat2_idx(r, c):
    """Natural (spatial) layout used between steps: row-major rows of 128 float2, the four
    columns of a 4-chunk split into two 16-byte pairs (pair 0 in the first half-row, pair 1 in
    the second) so that a warp reading one row with two 128-bit loads per lane is conflict-free,
    XOR-swizzled so that the strided P1 gather (lanes = 4 n2c x 4 rows) is conflict-free too."""
    c1 = (c >> 1) & 1
    pos = (c & 1) | ((c >> 2) << 1) | (c1 << 6)
    pos = pos ^ ((r & 3) << 1) ^ (c1 << 3)
    return r * 128 + pos


def check_nat2():
    r, c = np.meshgrid(np.arange(128), np.arange(128), indexing="ij")
    assert len(np.unique(nat2_idx(r, c))) == 128 * 128
    # S layout: lane l reads pairs (4l,4l+1) and (4l+2,4l+3) with one 128-bit access each.
    for row in range(128):
        for half in range(2):
            for qw in range(4):
                slots = []
                for l in range(8 * qw, 8 * qw + 8):
                    i0 = nat2_idx(row, 4 * l + 2 * half)
                    i1 = nat2_idx(row, 4 * l + 2 * half + 1)
                    assert i1 == i0 + 1 and i0 % 2 == 0  # one aligned 16-byte pair
                    slots.append((i0 // 2) % 8)
                assert len(set(slots)) == 8
    # P1 gather (64-bit): lanes -> n2c = l&3, r[1:0] = (l>>2)&3
    for n1c in range(32):
        for rhi in range(32):
            idx = [nat2_idx(rhi * 4 + ((l >> 2) & 3), 4 * n1c + (l & 3)) for l in range(16)]
            assert len(set(i % 16 for i in idx)) == 16
    print("natural layout OK")


if __name__ == "__main__":
    check_nat2()
