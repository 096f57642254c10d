#!/usr/bin/env python
"""tools/jit_inspect.py -- what the specialised tile kernels of a circuit look like, without a GPU.

    python tools/jit_inspect.py [n=33] [variant=2] [workload=bench|c3|qft] [out_dir]

Plans the workload exactly as the library would (qb_rounds_debug), generates the CUDA source of every pass
(qb_jit_source_debug), compiles it for sm_100a with NVRTC (qb_jit_compile_debug) and reports per pass: rounds, ops, source
lines, compile time, cubin size, registers, spills and the SASS instruction mix (cuobjdump).  With out_dir the sources and
cubins are kept."""
import collections
import ctypes
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from quaru_b200 import _ffi  # noqa: E402
from tests import rounds_emulator as em  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 33
    variant = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    workload = sys.argv[3] if len(sys.argv) > 3 else "bench"
    out_dir = sys.argv[4] if len(sys.argv) > 4 else None
    if workload == "c3":
        ops = bench.build_circuit(n, layers=200, qft=False)
    elif workload == "qft":
        ops = bench.build_circuit(n, layers=0, qft=True)
    else:
        ops = bench.build_circuit(n)
    os.environ.setdefault("QB_SWAP_RELABEL", "1")
    L = _ffi.lib()
    if os.environ.get("QB_PEEPHOLE", "1") != "0":      # what flush() does for registers of >= 28 local qubits
        from tests.test_peephole import _rewrite
        before = len(ops)
        rewrites, ops = _rewrite(n, ops)
        print(f"peephole: {rewrites} rewrites, {before} -> {len(ops)} gates")
    progs = [p for p in em.dump_programs(n, ops) if p[0] == "prog"]
    print(f"{len(ops)} gates, {len(progs)} register-blocked passes")
    tot = collections.Counter()
    tmp = out_dir or tempfile.mkdtemp()
    os.makedirs(tmp, exist_ok=True)
    for i, p in enumerate(progs):
        pb = p[3]["bytes"]
        cap = 16 << 20
        buf = ctypes.create_string_buffer(cap)
        persistent = 2 if (variant >> 6) & 1 else (variant >> 4) & 1      # bit 6: the prefetching ring
        got = L.qb_jit_source_debug(pb, len(pb), variant & 15, persistent, (variant >> 5) & 1, 0, 0, None, buf, cap)
        assert got > 0, _ffi.last_error()
        src = buf.raw[:got]
        ms = ctypes.c_double(0)
        cub = ctypes.create_string_buffer(8 << 20)
        size = L.qb_jit_compile_debug(src, ctypes.byref(ms), cub, len(cub))
        assert size > 0, _ffi.last_error()
        cpath = os.path.join(tmp, f"pass{i:02d}.cubin")
        open(cpath, "wb").write(cub.raw[:size])
        open(os.path.join(tmp, f"pass{i:02d}.cu"), "wb").write(src)
        res = subprocess.run(["cuobjdump", "-res-usage", cpath], capture_output=True, text=True).stdout
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", res.replace("\n", " "))
        sass = subprocess.run(["cuobjdump", "-sass", cpath], capture_output=True, text=True).stdout
        mix = collections.Counter()
        for ln in sass.splitlines():
            mm = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if mm:
                mix[mm.group(1).split(".")[0]] += 1
        tot.update(mix)
        n_inst = sum(mix.values())
        fp64 = mix["DFMA"] + mix["DMUL"] + mix["DADD"]
        print(f"pass {i:2d}: gates {len(p[1]):3d} rounds {len(p[3]['rounds'])} ops {len(p[3]['ops']):3d} src {src.count(b'\\n'):5d} lines  "
              f"nvrtc {ms.value:7.0f} ms  cubin {size >> 10:4d} KB  regs {m.group(1) if m else '?'} local {m.group(3) if m else '?'}  "
              f"SASS {n_inst:5d} (FP64 {fp64}, MOV {mix['MOV']}, SEL {mix['SEL']}, LDL {mix['LDL']}, STL {mix['STL']}, LDS {mix['LDS']}, STS {mix['STS']}, "
              f"LDG {mix['LDG']}, BRA {mix['BRA']})")
    n_inst = sum(tot.values())
    print("total SASS", n_inst, {k: v for k, v in tot.most_common(14)})


if __name__ == "__main__":
    main()
