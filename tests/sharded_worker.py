"""Worker for the multi-GPU parity test; launched by tests/test_gpu_sharded.py as
   python -m torch.distributed.run --nproc-per-node P --master-addr 127.0.0.1 tests/sharded_worker.py
Every rank builds the same seeded circuit, runs it on its shard and compares its slice (and measurement
outcomes under a shared uniform stream) with the CPU oracle."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from oracle.oracle import OracleRegister
from quaru_b200 import Operation, Register, _ffi
from tests import circuits
from tests.util import ABS_TOL, L2_TOL


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    def new_uid():
        """A ncclUniqueId is single-use: one per communicator, made on rank 0 and broadcast."""
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            assert _ffi.lib().qb_nccl_unique_id(buf) == 0, _ffi.last_error()
            uid = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        return bytes(uid.cpu().tolist())

    failures = []
    remote_passes = [0]
    jit_launches = [0]
    g = int(np.log2(world))

    def check(name, n, ops, fusion, measure=()):
        reg = Register.sharded(n, 0, rank, world, new_uid(), device=local)
        reg.set_option(_ffi.QB_OPT_FUSION, fusion)
        reg.set_option(_ffi.QB_OPT_JIT, 2)               # fused passes of shards with >= 12 local qubits run through their
        reg.set_option(_ffi.QB_OPT_JIT_MIN_QUBITS, 12)   # specialised kernels (local and remote variants), compiled synchronously
        ora = OracleRegister.new([False] * n)
        rng = np.random.default_rng(1234)
        for m, t in ops:
            reg.apply(Operation(m, t))
            ora.apply((m, t))
        for t in measure:
            u = float(rng.random())
            a, b = reg.measure(t, u=u), ora.measure(t, u)
            if a != b:
                failures.append(f"{name}: outcome of qubit {t} differs ({a} vs {b})")
        nl = n - g
        lo = rank << nl
        mine = reg.read(0, 1 << n)[lo:lo + (1 << nl)]          # canonicalises the placement first
        want = ora.state[lo:lo + (1 << nl)]
        d = np.abs(mine - want)
        l2 = float(np.sqrt((d * d).sum()))
        if not (d.max() <= ABS_TOL and l2 <= L2_TOL):
            failures.append(f"{name}: rank {rank} max|diff| {d.max():.3e} l2 {l2:.3e}")
        nrm = reg.norm2()
        if abs(nrm - 1.0) > 1e-10:
            failures.append(f"{name}: norm2 {nrm}")
        swaps = reg.stat(_ffi.QB_STAT_SWAPS)
        remote_passes[0] += reg.stat(_ffi.QB_STAT_REMOTE_PASSES)
        jit_launches[0] += reg.stat(_ffi.QB_STAT_JIT_LAUNCHES)
        reg.close()
        return swaps

    quick = world >= 8 or os.environ.get("QB_SHARDED_QUICK") in ("1", "2")      # bound the 8-GPU run (GPU-minutes are x8)
    minimal = os.environ.get("QB_SHARDED_QUICK") == "2"                          # two registers only
    for n in (() if minimal else (3 * g + 1, 14) if quick else (3 * g + 1, 10, 14)):
        for fusion in ((1,) if quick else (0, 1)):
            s = check(f"layers n={n} fusion={fusion}", n, circuits.random_circuit(n, 6, seed=n), fusion,
                      measure=(n - 1, 0, n // 2))
            check(f"qft n={n} fusion={fusion}", n, circuits.qft_circuit(n), fusion)
            ops = [o for o in circuits.mixed_circuit(n, 80, seed=7 + n) if len(o[1]) <= n - g]
            check(f"mixed n={n} fusion={fusion}", n, ops, fusion, measure=(1, n - 1))
    if minimal:
        ops = [o for o in circuits.mixed_circuit(15, 80, seed=22) if len(o[1]) <= 15 - g]
        check("mixed n=15 (minimal)", 15, ops, 1, measure=(1, 14))
    # a bigger one: 20 qubits, fused, 3 layers
    swaps = check("layers n=20", 20, circuits.random_circuit(20, 3, seed=20), 1, measure=(19,))
    if rank == 0:
        print(f"exchanges in the n=20 run: {swaps}; specialised-kernel launches so far: {jit_launches[0]}", flush=True)
    # Placement memory (qb_shard.cu shard_run): the same circuit applied again first moves the qubits back to the placement it
    # was first run from, so every pass program -- and its specialised kernel -- repeats.  Three applications against the oracle;
    # the third must not hand a single new program to the compiler.
    if os.environ.get("QB_SHARD_REMOTE") != "1":
        n = 14 + g
        ops = circuits.random_circuit(n, 4, seed=31) + circuits.qft_circuit(n)[:60]
        reg = Register.sharded(n, 0, rank, world, new_uid(), device=local)
        reg.set_option(_ffi.QB_OPT_JIT, 2)
        reg.set_option(_ffi.QB_OPT_JIT_MIN_QUBITS, 12)
        ora = OracleRegister.new([False] * n)
        eops = [Operation(m, t) for m, t in ops]
        compiles = []
        for rep in range(3):
            reg.apply_batch(eops)
            reg.sync()
            compiles.append(reg.stat(_ffi.QB_STAT_JIT_COMPILES))
            for o in ops:
                ora.apply(o)
        nl, lo = n - g, rank << (n - g)
        d = np.abs(reg.read(0, 1 << n)[lo:lo + (1 << nl)] - ora.state[lo:lo + (1 << nl)])
        if d.max() > ABS_TOL:
            failures.append(f"repeated circuit: rank {rank} max|diff| {d.max():.3e}")
        if compiles[2] != compiles[1]:
            failures.append(f"repeated circuit: rank {rank} compiled new programs on the third application {compiles}")
        fused_x = reg.stat(_ffi.QB_STAT_FUSED_EXCHANGES)
        fused_on = os.environ.get("QB_SHARD_FUSED_EXCHANGE", "1" if world == 2 else "0") == "1"      # the default: 2 GPUs only
        if fused_on and os.environ.get("QB_SHARD_TRANSPORT") != "nccl" and fused_x == 0:
            failures.append(f"repeated circuit: rank {rank}: no exchange was carried by a tile pass")
        if rank == 0:
            print(f"repeated circuit n={n}: programs compiled after each application {compiles}, exchanges {reg.stat(_ffi.QB_STAT_SWAPS)} "
                  f"({fused_x} carried by a pass's stores)", flush=True)
        reg.close()

    if os.environ.get("QB_SHARD_REMOTE") == "1":
        # experimental exchange-free scheduler: shards of >= 2^12 amplitudes, so that rank bits can be tile bits
        n = 13 + g
        check(f"remote: layers n={n}", n, circuits.random_circuit(n, 6, seed=3 * n), 1, measure=(n - 1, 0))
        check(f"remote: qft n={n}", n, circuits.qft_circuit(n), 1)
        ops = [o for o in circuits.mixed_circuit(n, 120, seed=n) if len(o[1]) <= 2]
        check(f"remote: mixed n={n}", n, ops, 1, measure=(1, n - 2))
        if rank == 0:
            print(f"remote passes: {remote_passes[0]}", flush=True)

    if os.environ.get("QB_SHARDED_EXTRA") == "1":
        # in-place reset and raw state files on a register whose placement has been permuted by exchanges
        import tempfile
        n = 3 * g + 4
        ops = circuits.random_circuit(n, 4, seed=5)
        reg = Register.sharded(n, 0, rank, world, new_uid(), device=local)
        for m, t in ops:
            reg.apply(Operation(m, t))
        reg.flush()
        nl, lo = n - g, rank << (n - g)
        path = os.path.join(tempfile.gettempdir(), f"qb_shard_{os.getpid()}_{rank}.qbstate")
        reg.save_state(path)
        saved = reg.read(0, 1 << n)[lo:lo + (1 << nl)]
        reg.reset(3)                                               # placement back to the identity, state |3>
        ora = OracleRegister.from_int(n, 3)
        for m, t in ops[:40]:
            reg.apply(Operation(m, t))
            ora.apply((m, t))
        d = np.abs(reg.read(0, 1 << n)[lo:lo + (1 << nl)] - ora.state[lo:lo + (1 << nl)])
        if d.max() > ABS_TOL:
            failures.append(f"reset after exchanges: rank {rank} max|diff| {d.max():.3e}")
        for m, t in ops[:25]:                                      # permute the placement again, then load over it
            reg.apply(Operation(m, t))
        reg.load_state(path)
        if not np.array_equal(reg.read(0, 1 << n)[lo:lo + (1 << nl)], saved):
            failures.append(f"save/load round trip differs on rank {rank}")
        os.remove(path)
        reg.close()

    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    for f in failures:
        print(f"[rank {rank}] FAIL {f}", flush=True)
    if rank == 0:
        print("SHARDED_OK" if int(flag.item()) == 0 else f"SHARDED_FAILED ({int(flag.item())})", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
