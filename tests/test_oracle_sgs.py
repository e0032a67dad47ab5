"""oracle/sgs_oracle.py (compute_Sgs, Burger.py:677-736 / KS.py:385-409) against goldens recorded from the reference classes."""
import numpy as np
import pytest

from oracle.sgs_oracle import compute_sgs


@pytest.mark.parametrize("tag", ["b512", "b256_forced", "b1024"])
def test_burgers_sgs_oracle(golden, tag):
    g = golden("sgs.npz")
    N, nURG, forcing, nsteps = (int(x) for x in g[f"{tag}/cfg"])
    sgs, alt, alt2 = compute_sgs(g[f"{tag}/uu"], g[f"{tag}/k"], 2 * np.pi / N, 1e-3, 0.02, nURG)
    for got, name in ((sgs, "sgs"), (alt, "alt"), (alt2, "alt2")):
        ref = g[f"{tag}/{name}"]
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) <= 1e-11 * np.max(np.abs(ref)), (tag, name)


def test_ks_sgs_oracle(golden):
    g = golden("sgs.npz")
    sgs = compute_sgs(g["ks256/uu"], g["ks256/k"], 22.0 / 256, 0.25, 1.0, 32, ks=True)
    # KS.uu is float32 (complex64 history, quirk Q7) and scipy.fftpack transforms float32 input in single precision, so the
    # reference's own result carries float32 round-off: tolerance of that chain, not 1e-11
    assert np.max(np.abs(sgs - g["ks256/sgs"])) <= 2e-4 * np.max(np.abs(g["ks256/sgs"]))
