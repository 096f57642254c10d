"""Shared helpers for the parity tests (CPU side only: numpy + the oracle)."""
import numpy as np

ABS_TOL = 1e-12     # BASELINE.json north_star: amplitudes within 1e-12 absolute ...
L2_TOL = 1e-10      # ... and L2 <= 1e-10, f64


def random_state(n, seed, normalize=True):
    rng = np.random.default_rng(seed)
    s = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    if normalize:
        s /= np.linalg.norm(s)
    return s.astype(np.complex128)


def random_matrix(k, seed, unitary=False):
    rng = np.random.default_rng(seed)
    m = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
    if unitary:
        m, _ = np.linalg.qr(m)
    else:
        m /= np.sqrt(1 << k)
    return m.astype(np.complex128)


def assert_close(got, want, what=""):
    got = np.asarray(got).reshape(-1)
    want = np.asarray(want).reshape(-1)
    d = np.abs(got - want)
    assert d.max() <= ABS_TOL, f"{what}: max abs diff {d.max():.3e} > {ABS_TOL}"
    l2 = float(np.sqrt((d * d).sum()))
    assert l2 <= L2_TOL, f"{what}: L2 diff {l2:.3e} > {L2_TOL}"


def tree_probabilities(state, target):
    """The engine's reduction restated in numpy: |a|^2 = fl(fl(re*re) + fl(im*im)) leaves in the bucket of
    bit `target`, summed by a balanced adjacent-pair binary tree over the amplitude index."""
    s = np.asarray(state, dtype=np.complex128).reshape(-1)
    p = s.real * s.real + s.imag * s.imag
    idx = np.arange(s.shape[0])
    if target is None:
        bit = np.zeros_like(idx)
    else:
        bit = (idx >> target) & 1
    v0 = np.where(bit == 0, p, 0.0)
    v1 = np.where(bit == 1, p, 0.0)
    while v0.shape[0] > 1:
        v0 = v0[0::2] + v0[1::2]
        v1 = v1[0::2] + v1[1::2]
    return float(v0[0]), float(v1[0])
