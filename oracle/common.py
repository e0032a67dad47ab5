"""Shared pieces of the CPU oracle (test infrastructure only, see oracle/__init__.py).

Every function cites the reference lines it restates (paths relative to
/root/reference/python/_model).
"""
import numpy as np
from scipy.fft import fft, ifft, fftfreq  # same pocketfft backend scipy.fftpack routes to


def grid(L, N):
    """x_j = j L / N  (Burger.py:85-86, KS.py:53-54, Diffusion.py:38-39)."""
    return np.linspace(0.0, float(L), N, endpoint=False)


def wavenumbers(L, N):
    """FFT-order dimensional wavenumbers 2*pi*n/L, Nyquist stored as -N/2
    (Burger.py:161, KS.py:113)."""
    return fftfreq(N, float(L) / (2.0 * np.pi * N))


def hat_row(x, centre, width):
    """Piecewise-linear tent, 1 at ``centre``, 0 at +-``width`` (Burger.py:12-15)."""
    rising = np.clip((x + width - centre) / width, 0.0, 1.0)
    falling = np.clip((width - x + centre) / width, 0.0, 1.0)
    return rising + falling - 1.0


def action_basis(x, L, M, kind):
    """[M, N] matrix mapping M actions to N grid values (Burger.py:177-203, KS.py:139-164).

    'uniform': block indicators (needs N % M == 0); 'hat': tents with nodes at
    i*L/(M-1), i.e. NOT periodic (last node sits at x=L); M == 1: all ones.
    """
    N = x.shape[0]
    if M <= 1:
        B = np.ones((M, N))
    elif kind == "uniform":
        if N % M:
            raise AssertionError("uniform basis needs N % M == 0")
        B = np.zeros((M, N))
        blk = N // M
        for i in range(M):
            B[i, i * blk:(i + 1) * blk] = 1.0
    elif kind == "hat":
        h = float(L) / (M - 1)
        B = np.stack([hat_row(x, i * h, h) for i in range(M)])
    else:
        raise ValueError(f"unknown basis kind {kind!r}")
    np.testing.assert_allclose(B.sum(axis=0), 1.0)  # Burger.py:203
    return B


def laplacian_fd(u, dx):
    """(u_{j-1} - 2 u_j + u_{j+1}) / dx^2, periodic (Burger.py:613-615, 342-346)."""
    return (np.roll(u, 1, axis=-1) - 2.0 * u + np.roll(u, -1, axis=-1)) / dx ** 2


def upwind_fd(u, dx):
    """(u_j - u_{j-1}) / dx, periodic (Burger.py:345)."""
    return (u - np.roll(u, 1, axis=-1)) / dx


def energy_row_f32(v, N, dx):
    """One row of Ek_kt exactly as the reference gets it from its complex64
    history: 1/2 * Re(conj(vv) vv / N) * dx with vv = complex64(v)
    (Burger.py:152,498,562; KS.py:105,273,343) -> float32 [.., N]."""
    vv = np.asarray(v).astype(np.complex64)
    return 1.0 / 2.0 * np.real(vv.conj() * vv / N) * dx


class RunningSpectrum:
    """Running time-average of the energy spectrum, Ek_ktt[i] = cumsum_f32(Ek_kt)[i]/(i+1)
    (Burger.py:555, KS.py:336).  The reference recomputes the whole history each
    call; the float32 sequential cumsum makes the running form bit-identical."""

    def __init__(self, v0, N, dx):
        self.N, self.dx = N, dx
        self.acc = energy_row_f32(v0, N, dx).astype(np.float32)
        self.count = 1

    def push(self, v):
        self.acc = (self.acc + energy_row_f32(v, self.N, self.dx)).astype(np.float32)
        self.count += 1

    def mean(self):
        return self.acc.astype(np.float64) / float(self.count)


def spectral_rel_err(Ek_dns_row, Ek_sgs_row, gridSize):
    """kRelErr of burger_environment.py:174 / ks_environment.py:98:
    mean over k=1..gridSize/2-1 of ((|E_dns - E_sgs|)/E_dns)^2."""
    a = Ek_dns_row[..., 1:gridSize // 2]
    b = Ek_sgs_row[..., 1:gridSize // 2]
    return np.mean((np.abs(a - b) / a) ** 2, axis=-1)


def truncate_spectrum(v_fine, g):
    """DNS -> LES spectral hand-off (burger_environment.py:111, ks_environment.py:52):
    keep modes 0..(g+1)//2-1 and the last (g-1)//2... of the fine spectrum, scaled by
    g / N_fine.  For even g the coarse Nyquist slot receives fine mode -g/2."""
    Nf = v_fine.shape[-1]
    head = v_fine[..., :(g + 1) // 2]
    tail = v_fine[..., -(g - 1) // 2:]
    return np.concatenate((head, tail), axis=-1) * g / Nf


def agent_windows(row, A):
    """Per-agent halo windows of a length-N row: indices (a*N/A-1 .. (a+1)*N/A) mod N
    (Burger.py:657-660).  Returns [.., A, N/A+2]."""
    N = row.shape[-1]
    idx = np.stack([np.arange(a * N // A - 1, (a + 1) * N // A + 1) % N for a in range(A)])
    return row[..., idx]
