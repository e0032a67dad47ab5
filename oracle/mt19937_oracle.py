"""TEST INFRASTRUCTURE ONLY.  Restatement of the random stream behind the reference's stochastic-forcing tables
(/root/reference/python/_model/Burger.py:66, 94-95: ``np.random.seed(seed)``; ``randfac1 = np.random.normal(size=(32, nsteps))``;
``randfac2 = ...``): NumPy's legacy ``RandomState`` = MT19937 seeded by ``init_genrand`` + ``legacy_gauss`` (polar
Box-Muller with one cached deviate) on 53-bit doubles built from two 32-bit outputs.  The algorithm lives in a third-party
dependency (NumPy, reference pin numpy==1.20.1; NumPy here 2.3 -- the legacy stream is frozen by NumPy's compatibility
policy).  Pinned against ``np.random.RandomState`` in tests/test_oracle_mt19937.py.  Groundwork for generating the
forcing tables on the device: the rejection loop makes the stream inherently sequential per seed."""
import math

N_, M_ = 624, 397
MASK32 = 0xFFFFFFFF


class MT19937:
    def __init__(self, seed):
        self.mt = [0] * N_
        self.mt[0] = seed & MASK32
        for i in range(1, N_):                                  # init_genrand
            self.mt[i] = (1812433253 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 30)) + i) & MASK32
        self.pos = N_
        self.has_gauss, self.gauss = False, 0.0

    def _generate(self):
        mt = self.mt
        for kk in range(N_):
            y = (mt[kk] & 0x80000000) | (mt[(kk + 1) % N_] & 0x7FFFFFFF)
            mt[kk] = mt[(kk + M_) % N_] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        self.pos = 0

    def next32(self):
        if self.pos >= N_:
            self._generate()
        y = self.mt[self.pos]
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & MASK32

    def double(self):                                           # random_double: 53 bits from two outputs
        a, b = self.next32() >> 5, self.next32() >> 6
        return (a * 67108864.0 + b) / 9007199254740992.0

    def normal(self):                                           # legacy_gauss
        if self.has_gauss:
            self.has_gauss = False
            g, self.gauss = self.gauss, 0.0
            return g
        while True:
            x1 = 2.0 * self.double() - 1.0
            x2 = 2.0 * self.double() - 1.0
            r2 = x1 * x1 + x2 * x2
            if r2 < 1.0 and r2 != 0.0:
                break
        f = math.sqrt(-2.0 * math.log(r2) / r2)
        self.gauss, self.has_gauss = f * x1, True
        return f * x2


def forcing_table_entries(seed, nsteps, stepper):
    """The only entries of randfac1 / randfac2 the solver ever reads (rows 1..3, columns < stepper, Burger.py:416-419),
    by walking the stream: element [k, c] of a (32, nsteps) table is draw number k * nsteps + c."""
    g = MT19937(seed)
    want = {k * nsteps + c for k in (1, 2, 3) for c in range(stepper)}
    out = []
    for t in range(2):
        vals = {}
        for i in range(32 * nsteps):
            x = g.normal()
            if i in want:
                vals[i] = x
        out.append([[vals[k * nsteps + c] for c in range(stepper)] for k in (1, 2, 3)])
    return out[0], out[1]
