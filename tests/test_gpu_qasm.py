"""GPU: run_openqasm on the CUDA engine against the same interpreter loop on the CPU oracle, shared uniform
stream: classical registers identical, final quantum registers within tolerance."""
import numpy as np
import pytest

from tests import qasm_programs as P
from tests.test_qasm import run_on_oracle
from tests.util import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["BELL3", "QFT4", "TELEPORT", "MIXED"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_program_matches_oracle(engine, name, seed):
    src = getattr(P, name)
    us = list(np.random.default_rng(seed).random(16))
    got = engine.openqasm.run_openqasm_source(src, us=us)
    want = run_on_oracle(src, us)
    assert got.cregs == want.cregs
    assert set(got.qregs) == set(want.qregs)
    for k in want.qregs:
        assert_close(got.qregs[k].state, want.qregs[k].state, f"{name} register {k}")


def test_run_openqasm_from_file(engine, tmp_path):
    f = tmp_path / "bell.qasm"
    f.write_text(P.BELL3)
    res = engine.openqasm.run_openqasm(str(f), seed=3)
    p = np.abs(res.qregs["q"].state) ** 2
    assert abs(p[0] - 0.5) < 1e-7 and abs(p[3] - 0.5) < 1e-7
