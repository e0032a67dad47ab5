"""Prototype (numpy, no GPU) of the "exchange one, keep one" 16-point FFT network planned for the 8-lane Burgers team
(DESIGN.md 9.2): 8 lanes x 2 registers, every butterfly is computed IN REGISTERS by one lane, and between stages the
lanes of a pair swap ONE register (a 2 x 2 transpose between the lane bit and the register bit) instead of exchanging
both -- half the shuffles of the current network, no multiply by (1, 0) on the lower lanes.
Layout in: register r of lane t holds x[8 r + t].  Layout out: register r of lane t holds X[8 r + bitrev3(t)].
Run:  python tools/fft_network_proto.py   (asserts against numpy.fft)."""
import numpy as np


def bitrev3(t):
    return ((t & 1) << 2) | (t & 2) | ((t >> 2) & 1)


def swap_bit(z, bit):
    """lane (bit = 0) keeps register 0 and receives the partner's register 0 into register 1;
    lane (bit = 1) keeps register 1 and receives the partner's register 1 into register 0."""
    out = z.copy()
    for t in range(8):
        p = t ^ bit
        if t & bit:
            out[t, 0] = z[p, 1]
        else:
            out[t, 1] = z[p, 0]
    return out


def forward(x):
    z = np.empty((8, 2), dtype=complex)
    for t in range(8):
        z[t, 0], z[t, 1] = x[t], x[8 + t]
    w16 = np.exp(-2j * np.pi * np.arange(16) / 16)
    t = np.arange(8)
    # (lane-dependent twiddle exponent, lane bit swapped into the register AFTER the butterfly)
    for tw, bit in ((w16[t], 4), (w16[2 * (t & 3)], 2), (w16[4 * (t & 1)], 1), (np.ones(8), 0)):
        a, b = z[:, 0].copy(), z[:, 1].copy()
        z[:, 0] = a + b                       # DIF butterfly, in registers, every lane
        z[:, 1] = (a - b) * tw
        if bit:
            z = swap_bit(z, bit)
    return z


def inverse(z):
    """Transposed network: swap, then the in-register DIT butterfly (conjugate twiddle on the second input)."""
    w16 = np.exp(-2j * np.pi * np.arange(16) / 16)
    t = np.arange(8)
    z = z.copy()
    for tw, bit in ((np.ones(8), 0), (w16[4 * (t & 1)], 1), (w16[2 * (t & 3)], 2), (w16[t], 4)):
        if bit:
            z = swap_bit(z, bit)              # the swap is an involution
        a, b = z[:, 0].copy(), z[:, 1] * np.conj(tw)
        z[:, 0], z[:, 1] = a + b, a - b
    return z


def check():
    rng = np.random.default_rng(0)
    x = rng.normal(size=16) + 1j * rng.normal(size=16)
    X = np.fft.fft(x)
    z = forward(x)
    for t in range(8):
        for r in range(2):
            assert abs(z[t, r] - X[8 * r + bitrev3(t)]) < 1e-12, (t, r)
    y = inverse(z)
    for t in range(8):
        assert abs(y[t, 0] - 16 * x[t]) < 1e-11 and abs(y[t, 1] - 16 * x[8 + t]) < 1e-11
    return True


if __name__ == "__main__":
    check()
    print("exchange-one-keep-one network: forward layout k = 8 r + bitrev3(t) and inverse verified against numpy.fft")
