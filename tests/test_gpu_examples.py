"""The example drivers (examples/*.py, counterparts of the reference's examples/*.rs) run end to end on the GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    out = subprocess.run([sys.executable] + list(args), capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


def test_grovers_target_64(engine):
    out = _run("examples/grovers.py", "--target", "64", "--seed", "1")
    assert "Result: 64" in out and "Iterations of oracle + diffuser: 8" in out


def test_grovers_decomposed_beyond_reference_reach(engine):
    out = _run("examples/grovers.py", "--target", "12345", "--qubits", "16", "--decomposed", "--seed", "2")
    assert "Result: 12345" in out


@pytest.mark.parametrize("flags", [[], ["--fast"], ["--use-circuit"]])
def test_shors_15(engine, flags):
    out = _run("examples/shors.py", "--number", "15", "--seed", "3", *flags)
    assert "The factors of 15 are 3 and 5" in out or "The factors of 15 are 5 and 3" in out


def test_shors_21_fast(engine):
    out = _run("examples/shors.py", "--number", "21", "--fast", "--seed", "4")
    assert "The factors of 21 are" in out


def test_shors_working(engine):
    """The reference's own test of the driver (examples/shors.rs:424-431): every composite number up to 15 through
    shors(n, fast = true, use_circuit = false); only 15 gets past the classical checks and onto the register."""
    import numpy as np

    sys.path.insert(0, ROOT)
    from examples import shors as driver

    rng = np.random.default_rng(15)
    for n in [4, 6, 8, 9, 10, 12, 14, 15]:
        r = driver.shors(n, True, False, rng)
        assert n % r == 0 and 1 < r < n, (n, r)


def test_openqasm_example_and_protocols(engine):
    out = _run("examples/openqasm_example.py")
    assert "000: 50" in out.replace("49.99999999999999", "50") or "000: 50.0" in out or "000: 49.9" in out
    out = _run("examples/protocols.py", "teleport")
    assert "Correct" in out
    out = _run("examples/protocols.py", "teleport", "--circuit")
    assert "Correct" in out
    out = _run("examples/protocols.py", "superdense", "-a", "1", "-b", "0")
    assert "received (True, False)" in out
