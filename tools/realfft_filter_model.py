"""NumPy model of the round-2 plan for the 128x128 kernel: ONE real field per CTA, packed along the
columns into a 128 x 64 complex field, so that two CTAs (two independent fields) fit one SM and their
FFT-arithmetic and shared-memory phases overlap.

Checks the closed form that replaces untangle -> multiply by a real even multiplier -> re-tangle:

    z[r][m]  = u[r][2m] + i u[r][2m+1]                       (m = 0..63)
    Z        = fft2(z)                                         (128 x 64 complex)
    ZG[kr][km] = a[kr][km] Z[kr][km] + i b[kr][km] conj(Z[-kr][-km])
    a = P - Q sin(theta),  b = Q cos(theta),  theta = 2 pi km / 128,
    P = (M[kr][km] + M[kr][km+64]) / 2,  Q = (M[kr][km] - M[kr][km+64]) / 2
    g[r][2m] + i g[r][2m+1] = ifft2(ZG)[r][m],   g = irfft2(M * rfft2(u))

i.e. two real tables and the partner element (-kr, -km): 4 FP32 pipe slots per complex element, no
Hermitian untangling pass.  Also checks that the 8192 spectrum elements can be dealt to 256 threads x 32
registers so that every partner pair lives in ONE thread (no exchange for the spectral step): a thread
owns {k1r, 16-k1r} x {km, 64-km} x (8 values of k2r), kr = k1r + 16 k2r, with the self-conjugate classes
{0, 8} and {0, 32} paired with each other.
"""
import numpy as np

N, H = 128, 64
rng = np.random.default_rng(0)
u = rng.normal(size=(N, N))
k = np.fft.fftfreq(N, 0.01)
k2 = (2 * np.pi) ** 2 * (k[:, None] ** 2 + k[None, :] ** 2)
M = 1.0 / (1.0 + 1e-6 * 0.5 * 0.002 * k2 * k2)  # real, even in both wavenumbers

g_ref = np.fft.ifft2(M * np.fft.fft2(u)).real

z = u[:, 0::2] + 1j * u[:, 1::2]
Z = np.fft.fft2(z)
kr = np.arange(N)[:, None]
km = np.arange(H)[None, :]
theta = 2 * np.pi * km / N
P = 0.5 * (M[:, :H] + M[:, H:])
Q = 0.5 * (M[:, :H] - M[:, H:])
a = P - Q * np.sin(theta)
b = Q * np.cos(theta)
Zp = np.conj(Z[(-kr) % N, (-km) % H])
ZG = a * Z + 1j * b * Zp
zg = np.fft.ifft2(ZG)
g = np.empty_like(u)
g[:, 0::2], g[:, 1::2] = zg.real, zg.imag
err = np.abs(g - g_ref).max() / np.abs(g_ref).max()
print("closed-form filter vs rfft2 reference: max rel err", err)
assert err < 1e-13

# partner pairs inside one thread
owner = -np.ones((N, H), int)
t = 0
k1r_classes = [(0, 8)] + [(j, 16 - j) for j in range(1, 8)]
km_classes = [(0, 32)] + [(j, 64 - j) for j in range(1, 32)]
for ca in k1r_classes:
    for cb in km_classes:
        for k1r in ca:
            for m in cb:
                for k2r in range(8):
                    assert owner[k1r + 16 * k2r, m] == -1
                    owner[k1r + 16 * k2r, m] = t
        t += 1
assert t == 256 and (owner >= 0).all()
assert (np.bincount(owner.ravel()) == 32).all()
assert (owner == owner[(-kr) % N, (-km) % H]).all()
print("256 threads x 32 registers, every (kr, km) / (-kr, -km) pair in one thread: ok")


# ---- three register passes of the 128 x 64 complex transform (13 index bits, 256 threads x 32 values) ----
#   A  radix-32 over the high 5 bits of m            thread (r, m0),            m = 2 n + m0
#   B  twiddle w64^(m0 k1m); radix-2 over m0; radix-16 over the high 4 bits of r; twiddle w128^(n2r k1r)
#                                                     thread (k1m, n2r),         r = 8 n1r + n2r
#   C  radix-8 over n2r, four (k1r, km) slots per thread
#                                                     thread (class of k1r, class of km), kr = k1r + 16 k2r, km = k1m + 32 k2m
def forward_three_pass(z):
    w128, w64 = np.exp(-2j * np.pi / 128), np.exp(-2j * np.pi / 64)
    a = z.reshape(N, 32, 2)                       # [r, n, m0]
    a = np.fft.fft(a, axis=1)                     # A -> [r, k1m, m0]
    k1m = np.arange(32)[None, :, None]
    m0 = np.arange(2)[None, None, :]
    a = a * w64 ** (m0 * k1m)
    a = np.fft.fft(a, axis=2)                     # B, radix 2 -> [r, k1m, k2m]
    a = a.reshape(16, 8, 32, 2)                   # [n1r, n2r, k1m, k2m]
    a = np.fft.fft(a, axis=0)                     # B, radix 16 -> [k1r, n2r, k1m, k2m]
    k1r = np.arange(16)[:, None, None, None]
    n2r = np.arange(8)[None, :, None, None]
    a = a * w128 ** (n2r * k1r)
    a = np.fft.fft(a, axis=1)                     # C -> [k1r, k2r, k1m, k2m]
    Zt = np.empty((N, H), complex)
    k1r, k2r, k1m, k2m = np.meshgrid(np.arange(16), np.arange(8), np.arange(32), np.arange(2), indexing="ij")
    Zt[k1r + 16 * k2r, k1m + 32 * k2m] = a
    return Zt


Z3 = forward_three_pass(z)
err3 = np.abs(Z3 - Z).max() / np.abs(Z).max()
print("three-pass 128 x 64 transform vs fft2: max rel err", err3)
assert err3 < 1e-13
# per-thread work of the passes: A 32-point DFT (80 butterflies), B 16 radix-2 + 2 x 16-point DFTs + 1 + 15 x 2
# inter-pass twiddles, C 4 x 8-point DFTs: 5 N log2 N = 5 * 8192 * 13 flops per real field (7 % below half of the
# 128 x 128 complex transform that serves two fields today)


# ---- exchange layouts (64 KB buffer = 8192 8-byte slots) and their bank behaviour -------------------------
# A 64-bit shared-memory access is served per half-warp: conflict free iff its 16 lanes hit 16 distinct slots
# modulo 16.  Thread index bits: pass A  t = (r[6:3] | r[2:0] | m0) with the low 4 bits = (r[2:0], m0);
# pass B  t = (k1m[4:1] | n2r | k1m[0]) with the low 4 bits = (n2r, k1m[0]);  pass C  t = (k1r class | km class)
# with the low 4 bits = km class [3:0].
def slot_ab(r, m0, k1m):          # written by A (fixed k1m per instruction), read by B (fixed (m0, n1r))
    return k1m * 256 + 2 * r + (m0 ^ (k1m & 1))


def slot_bc(k1r, k2m, n2r, k1m):  # written by B (fixed (k2m, k1r)), read by C (fixed (slot of the class pair, n2r))
    return ((k1r * 2 + k2m) * 8 + n2r) * 32 + (k1m ^ (n2r << 1))  # XOR: B's half-warp runs over (n2r, k1m[0])


def distinct16(slots):
    return len({int(s) % 16 for s in slots}) == 16


r_, m0_, k_ = np.meshgrid(np.arange(N), np.arange(2), np.arange(32), indexing="ij")
assert len(np.unique(slot_ab(r_, m0_, k_))) == 8192
a_, b_, c_, d_ = np.meshgrid(np.arange(16), np.arange(2), np.arange(8), np.arange(32), indexing="ij")
assert len(np.unique(slot_bc(a_, b_, c_, d_))) == 8192
for k1m in range(32):                       # A writes: lanes = (r[2:0], m0)
    for rhi in range(16):
        assert distinct16([slot_ab(8 * rhi + (l >> 1), l & 1, k1m) for l in range(16)])
for m0 in range(2):                         # B reads: lanes = (n2r, k1m[0])
    for n1r in range(16):
        for khi in range(16):
            assert distinct16([slot_ab(8 * n1r + (l >> 1), m0, 2 * khi + (l & 1)) for l in range(16)])
for k1r in range(16):                       # B writes: same lanes as its reads, (n2r, k1m[0])
    for k2m in range(2):
        for khi in range(16):
            assert distinct16([slot_bc(k1r, k2m, l >> 1, 2 * khi + (l & 1)) for l in range(16)])
for k1r in range(16):                       # C reads: lanes = km class [3:0]; members km = c and km = 64 - c (c = 0: 32)
    for n2r in range(8):
        for chi in range(2):
            first = [(16 * chi + l) for l in range(16)]
            assert distinct16([slot_bc(k1r, 0, n2r, c) for c in first])
            second = [((64 - c) if c else 32) for c in first]
            assert distinct16([slot_bc(k1r, 1, n2r, km - 32) for km in second])
print("exchange layouts A->B and B->C: bijective and bank-conflict free")
