"""Poseidon over BN254 Fr -- oracle restatement.  TEST INFRASTRUCTURE ONLY.

Follows snark-verifier/src/util/hash/poseidon.rs:
  * `OptimizedPoseidonSpec::new` (:230-245), `calculate_optimized_constants` (:247-297),
    `calculate_sparse_matrices` (:299-315), `MDSMatrix::factorise` (:172-225)
  * `State` (:323-410), `Poseidon::{update,squeeze,permutation}` (:449-501)
Round constants / MDS come from the un-vendored `poseidon-circuit@50015b7`
(Cargo.lock:2826-2828; `Spec::constants()` at poseidon.rs:231-232), restated here from the
Poseidon paper's Grain-LFSR procedure (eprint 2019/458, as in zcash halo2_gadgets).
Pinned by the reference KATs (poseidon/tests.rs:7-32, :35-85) in tests/test_oracle_poseidon.py.
"""
from .bn254 import R


# ---------------------------------------------------------------- Grain LFSR (poseidon-circuit `grain.rs`)
class Grain:
    def __init__(self, t, r_f, r_p, field_bits=254):
        bits = []

        def app(v, n):
            for i in reversed(range(n)):
                bits.append((v >> i) & 1)

        app(1, 2)  # prime field
        app(0, 4)  # sbox x^alpha
        app(field_bits, 12)
        app(t, 12)
        app(r_f, 10)
        app(r_p, 10)
        app((1 << 30) - 1, 30)
        assert len(bits) == 80
        self.s = bits
        for _ in range(160):
            self._next()

    def _next(self):
        s = self.s
        b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(b)
        return b

    def bit(self):
        """shrinking generator: of each pair, output the 2nd bit iff the 1st is 1"""
        while True:
            b1 = self._next()
            b2 = self._next()
            if b1:
                return b2

    def _int(self, nbits):
        v = 0
        for _ in range(nbits):
            v = (v << 1) | self.bit()  # MSB first
        return v

    def field_element(self, nbits=254):
        while True:
            v = self._int(nbits)
            if v < R:
                return v

    def field_element_without_rejection(self, nbits=254):
        return self._int(nbits) % R


def _mat_inv(m):
    n = len(m)
    a = [list(row) + [1 if i == j else 0 for j in range(n)] for i, row in enumerate(m)]
    for c in range(n):
        p = next(i for i in range(c, n) if a[i][c] % R)
        a[c], a[p] = a[p], a[c]
        inv = pow(a[c][c], R - 2, R)
        a[c] = [x * inv % R for x in a[c]]
        for i in range(n):
            if i != c and a[i][c]:
                f = a[i][c]
                a[i] = [(x - f * y) % R for x, y in zip(a[i], a[c])]
    return [row[n:] for row in a]


def generate_constants(t, r_f, r_p, secure_mds=0):
    """`Poseidon128Pow5Gen::constants()` -> (round_constants[r_f+r_p][t], mds, mds_inv)"""
    g = Grain(t, r_f, r_p)
    rc = [[g.field_element() for _ in range(t)] for _ in range(r_f + r_p)]
    select = secure_mds
    while True:
        while True:
            vals = [g.field_element_without_rejection() for _ in range(2 * t)]
            if len(set(vals)) == len(vals):
                break
        if select:
            select -= 1
            continue
        xs, ys = vals[:t], vals[t:]
        mds = [[pow((x + y) % R, R - 2, R) for y in ys] for x in xs]
        break
    return rc, mds, _mat_inv(mds)


# ---------------------------------------------------------------- optimised spec (poseidon.rs:230-315)
def _mul_vec(m, v):
    return [sum(m[i][j] * v[j] for j in range(len(v))) % R for i in range(len(m))]


def _mat_mul(a, b):
    n = len(a)
    return [[sum(a[i][k] * b[k][j] for k in range(n)) % R for j in range(n)] for i in range(n)]


def _transpose(m):
    return [list(r) for r in zip(*m)]


def _factorise(m):
    """poseidon.rs:172-225: M = M' * M''; returns (M', (row, col_hat))."""
    t = len(m)
    w = [m[i][0] for i in range(1, t)]
    m_hat = [[m[i + 1][j + 1] for j in range(t - 1)] for i in range(t - 1)]
    w_hat = _mul_vec(_mat_inv(m_hat), w)  # == Cramer's rule of :207-216
    m_prime = [[1 if i == j else 0 for j in range(t)] for i in range(t)]
    for i in range(t - 1):
        for j in range(t - 1):
            m_prime[i + 1][j + 1] = m_hat[i][j]
    m_pp = [[1 if i == j else 0 for j in range(t)] for i in range(t)]
    m_pp[0] = list(m[0])
    for i in range(t - 1):
        m_pp[i + 1][0] = w_hat[i]
    row = [m_pp[i][0] for i in range(t)]
    col_hat = m_pp[0][1:]
    return m_prime, (row, col_hat)


class OptimizedPoseidonSpec:
    """`OptimizedPoseidonSpec<F,T,RATE>::new::<R_F,R_P,SECURE_MDS>()`"""

    def __init__(self, t, rate, r_f, r_p, secure_mds=0):
        assert rate + 1 == t
        self.t, self.rate, self.r_f, self.r_p = t, rate, r_f, r_p
        rc, mds, mds_inv = generate_constants(t, r_f, r_p, secure_mds)
        self.round_constants = rc
        self.mds = mds
        self.mds_inv = mds_inv
        half = r_f // 2
        # calculate_optimized_constants (:247-297)
        start = [rc[0]] + [_mul_vec(mds_inv, rc[i]) for i in range(1, half)]
        acc = list(rc[half + r_p])
        partial = [0] * r_p
        for k in reversed(range(r_p)):
            tmp = _mul_vec(mds_inv, acc)
            partial[k] = tmp[0]
            tmp[0] = 0
            acc = [(a + b) % R for a, b in zip(tmp, rc[half + k])]
        start.append(_mul_vec(mds_inv, acc))
        end = [_mul_vec(mds_inv, rc[i]) for i in range(half + r_p + 1, r_f + r_p)]
        assert len(end) == half - 1
        self.start, self.partial, self.end = start, partial, end
        # calculate_sparse_matrices (:299-315)
        mt = _transpose(mds)
        acc_m = [list(r) for r in mt]
        sparse = []
        for _ in range(r_p):
            m_prime, m_pp = _factorise(acc_m)
            acc_m = _mat_mul(mt, m_prime)
            sparse.append(m_pp)
        sparse.reverse()
        self.sparse = sparse
        self.pre_sparse_mds = _transpose(acc_m)


_SPEC_CACHE = {}


def spec(t=3, rate=2, r_f=8, r_p=57, secure_mds=0):
    """SDK parameters: T=3, RATE=2, R_F=8, R_P=57, SECURE_MDS=0 (snark-verifier-sdk/src/halo2.rs:52-56)."""
    key = (t, rate, r_f, r_p, secure_mds)
    if key not in _SPEC_CACHE:
        _SPEC_CACHE[key] = OptimizedPoseidonSpec(*key)
    return _SPEC_CACHE[key]


# ---------------------------------------------------------------- permutations on int lists
def permutation_optimized(sp, state, inputs):
    """`Poseidon::permutation` (poseidon.rs:469-501) on a list of ints; returns new state."""
    t = sp.t
    s = list(state)
    half = sp.r_f // 2
    # absorb_with_pre_constants (:362-384)
    assert len(inputs) < t
    pre = sp.start[0]
    s[0] = (s[0] + pre[0]) % R
    for i, x in enumerate(inputs):
        s[i + 1] = (s[i + 1] + x + pre[i + 1]) % R
    for idx, i in enumerate(range(1 + len(inputs), t)):
        s[i] = (s[i] + pre[i] + (1 if idx == 0 else 0)) % R
    for c in sp.start[1:half]:
        s = [(pow(x, 5, R) + k) % R for x, k in zip(s, c)]
        s = _mul_vec(sp.mds, s)
    s = [(pow(x, 5, R) + k) % R for x, k in zip(s, sp.start[-1])]
    s = _mul_vec(sp.pre_sparse_mds, s)
    for c, (row, col_hat) in zip(sp.partial, sp.sparse):
        s[0] = (pow(s[0], 5, R) + c) % R
        s0 = sum(r * x for r, x in zip(row, s)) % R
        s = [s0] + [(ch * s[0] + x) % R for ch, x in zip(col_hat, s[1:])]
    for c in sp.end:
        s = [(pow(x, 5, R) + k) % R for x, k in zip(s, c)]
        s = _mul_vec(sp.mds, s)
    s = [pow(x, 5, R) for x in s]
    s = _mul_vec(sp.mds, s)
    return s


def permutation_textbook(sp, state):
    """Plain Hades permutation (eprint 2019/458) -- independent cross-check of the optimised form."""
    s = list(state)
    half = sp.r_f // 2
    rc = sp.round_constants
    k = 0
    for _ in range(half):
        s = [pow((x + c) % R, 5, R) for x, c in zip(s, rc[k])]
        s = _mul_vec(sp.mds, s)
        k += 1
    for _ in range(sp.r_p):
        s = [(x + c) % R for x, c in zip(s, rc[k])]
        s[0] = pow(s[0], 5, R)
        s = _mul_vec(sp.mds, s)
        k += 1
    for _ in range(half):
        s = [pow((x + c) % R, 5, R) for x, c in zip(s, rc[k])]
        s = _mul_vec(sp.mds, s)
        k += 1
    return s


class Poseidon:
    """`Poseidon<F, L, T, RATE>` sponge (poseidon.rs:414-467) on ints."""

    def __init__(self, sp=None):
        self.spec = sp or spec()
        self.default_state = [1 << 64] + [0] * (self.spec.t - 1)  # :335-342
        self.state = list(self.default_state)
        self.buf = []
        self.n_perm = 0

    def clear(self):
        self.state = list(self.default_state)
        self.buf = []

    def update(self, elements):
        self.buf.extend(int(e) % R for e in elements)

    def squeeze(self):
        buf, self.buf = self.buf, []
        rate = self.spec.rate
        exact = len(buf) % rate == 0
        for i in range(0, len(buf), rate):
            self.state = permutation_optimized(self.spec, self.state, buf[i : i + rate])
            self.n_perm += 1
        if exact:
            self.state = permutation_optimized(self.spec, self.state, [])
            self.n_perm += 1
        return self.state[1]
