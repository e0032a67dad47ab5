"""oracle/mt19937_oracle.py against NumPy's legacy RandomState: raw doubles, normals, and the forcing-table entries the
reference reads after np.random.seed(seed) (Burger.py:66, 94-95)."""
import numpy as np
import pytest

from oracle.mt19937_oracle import MT19937, forcing_table_entries


@pytest.mark.parametrize("seed", [0, 42, 1337, 2 ** 31 + 5])
def test_stream_matches_numpy(seed):
    rs = np.random.RandomState(seed)
    g = MT19937(seed)
    assert [g.double() for _ in range(700)] == list(rs.random_sample(700))
    rs = np.random.RandomState(seed)
    g = MT19937(seed)
    want = rs.normal(size=1001)
    got = np.array([g.normal() for _ in range(1001)])
    assert np.max(np.abs(got - want)) <= 4e-16 * np.max(np.abs(want))         # libm log / sqrt: within an ulp


def test_forcing_table_entries_match_the_reference_tables():
    seed, nsteps, stepper = 42, 50, 4
    np.random.seed(seed)
    r1 = np.random.normal(loc=0., scale=1., size=(32, nsteps))
    r2 = np.random.normal(loc=0., scale=1., size=(32, nsteps))
    a, b = forcing_table_entries(seed, nsteps, stepper)
    np.testing.assert_allclose(np.array(a), r1[1:4, :stepper], rtol=1e-15)
    np.testing.assert_allclose(np.array(b), r2[1:4, :stepper], rtol=1e-15)
