"""pytest plugin (test infrastructure): runs `-m gpu` test files against the HOST build of the API layer (tests/cpu_api,
see tests/test_cpu_api.py) on a machine without a GPU, to check the C ABI's host logic after a change to qb_api.cu /
qb_gate.cu before GPU time is spent:

    python -m pytest tests/test_cpu_api.py -q                       # builds tests/cpu_api/_build/libquaru_hostapi.so
    PYTHONPATH=tests/cpu_api python -m pytest -p hostapi_plugin tests/test_gpu_parity.py tests/test_gpu_reference_suite.py \
        tests/test_gpu_qasm.py -q -m gpu --deselect tests/test_gpu_parity.py::test_probabilities_bit_exact_tree

Expected to fail there, by construction: whatever depends on the kernels themselves (bit-exact reduction tree, sweep counts of
the fused path, the state-file format and checksum, subprocess examples and the C++ mirror, which load the CUDA library).
It only redirects the Python mirror's loader inside the pytest process; the product has no such switch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from quaru_b200 import _ffi  # noqa: E402

_ffi.LIB_PATH = os.path.join(ROOT, "tests", "cpu_api", "_build", "libquaru_hostapi.so")
_ffi._lib = None
