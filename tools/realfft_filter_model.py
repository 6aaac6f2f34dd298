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
