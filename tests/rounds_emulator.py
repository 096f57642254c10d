"""NumPy emulator of the register-blocked tile kernel's PROGRAM semantics (quaru_b200/csrc/qb_tile2.cu).

`qb_rounds_debug` serialises the pass programs the fused path would launch (rounds, ops, coefficient tables);
this module replays them on a full state vector, amplitude by amplitude, following the kernel's op definitions.
It is test infrastructure: it lets the CPU suite check the HOST side of the fused path (pass planning, round
splitting, chain accumulation, phase-table merging, layer packing, control splitting into register / thread /
outer bits) against the oracle without a GPU.  The arithmetic order differs from the kernel's (and the
oracle's), so comparisons use the usual 1e-12 tolerance.
"""
import ctypes
import os
import struct

import numpy as np

# op codes (enum RCode)
RC_D1C, RC_D1R, RC_DG1R = 0, 4, 8
RC_D1C_P, RC_D1R_P, RC_DG1R_P = 16, 20, 24
RC_D2, RC_PH_T, RC_PH_NEG, RC_PH_R, RC_DG, RC_DG1T = 32, 33, 34, 35, 36, 37
RC_H, RC_PH_ACC, RC_PH_FLUSH, RC_TAB, RC_TAB_ACC, RC_LAYER = 40, 44, 48, 52, 53, 54

ROP_FMT = "<BBBBHHI3sBQ"          # code a b flags rmask tmask mat src[3] k omask   (24 bytes)
ROUND_FMT = "<4s8sHHB3s16H"       # R tb op_begin op_end has_tph pad sro[16]          (52 bytes)
assert struct.calcsize(ROP_FMT) == 24 and struct.calcsize(ROUND_FMT) == 52


def dump_programs(n, ops):
    """ops: [(matrix, targets)] -> list of passes: ('gates', [tags]) or ('prog', [tags], tile_pos, program dict)."""
    from quaru_b200 import _ffi
    L = _ffi.lib()
    arr = (_ffi.QbGate * max(len(ops), 1))()
    keep = []
    for i, (m, t) in enumerate(ops):
        mm = np.ascontiguousarray(m, dtype=np.complex128)
        tt = (ctypes.c_uint32 * max(len(t), 1))(*t)
        keep += [mm, tt]
        arr[i].matrix = mm.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        arr[i].targets = tt
        arr[i].k = len(t)
    cap = 64 << 20
    buf = ctypes.create_string_buffer(cap)
    got = L.qb_rounds_debug(n, arr, len(ops), buf, cap)
    assert got >= 0, _ffi.last_error()
    raw = buf.raw[:got]
    passes, off = [], 0
    while off < len(raw):
        kind, ng = struct.unpack_from("<II", raw, off); off += 8
        tags = list(struct.unpack_from(f"<{ng}q", raw, off)); off += 8 * ng
        if kind == 1:
            passes.append(("gates", tags))
            continue
        tile_pos = list(raw[off:off + 12]); off += 12
        prog_begin = off
        lc, h, n_rounds, n_ops, n_mats = struct.unpack_from("<5I", raw, off); off += 20
        hp = list(raw[off:off + 8]); off += 8
        rounds = []
        for _ in range(n_rounds):
            f = struct.unpack_from(ROUND_FMT, raw, off); off += 52
            rounds.append({"R": list(f[0]), "tb": list(f[1]), "op_begin": f[2], "op_end": f[3], "has_tph": f[4], "sro": list(f[6:22])})
        rops = []
        for _ in range(n_ops):
            f = struct.unpack_from(ROP_FMT, raw, off); off += 24
            rops.append({"code": f[0], "a": f[1], "b": f[2], "flags": f[3], "rmask": f[4], "tmask": f[5], "mat": f[6],
                         "src": list(f[7]), "k": f[8], "omask": f[9]})
        mats = np.frombuffer(raw, dtype=np.complex128, count=n_mats, offset=off).copy(); off += 16 * n_mats
        gro_first = list(struct.unpack_from("<16Q", raw, off)); off += 128
        gro_last = list(struct.unpack_from("<16Q", raw, off)); off += 128
        passes.append(("prog", tags, tile_pos, {"lc": lc, "h": h, "hp": hp, "rounds": rounds, "ops": rops, "mats": mats,
                                                "gro_first": gro_first, "gro_last": gro_last, "bytes": raw[prog_begin:off]}))
    return passes


def _bits_to(g, positions):
    """pext: gather the bits of g at `positions` (ascending) into a compact integer."""
    out = np.zeros_like(g)
    for j, p in enumerate(positions):
        out |= ((g >> p) & 1) << j
    return out


def emulate_program(state, n, tile_pos, prog):
    """Applies one pass program to the full state vector (returns a new vector)."""
    s = np.array(state, dtype=np.complex128)
    N = 1 << n
    g = np.arange(N, dtype=np.int64)
    assert len(tile_pos) == 12 and tile_pos == sorted(tile_pos)
    assert tile_pos[:prog["lc"]] == list(range(prog["lc"])) and tile_pos[prog["lc"]:] == prog["hp"][:prog["h"]]
    outer_pos = [p for p in range(n) if p not in tile_pos]
    loc = _bits_to(g, tile_pos)                 # tile-local index of every amplitude
    blk = _bits_to(g, outer_pos)                # blockIdx.x of the tile that owns it
    mats = prog["mats"]
    assert mats[0] == 1 and mats[1] == 0 and mats[2] == 0 and mats[3] == 1, "mats[0..3] must hold the identity"
    n_rounds = len(prog["rounds"])
    for which, rd in (("gro_first", prog["rounds"][0]), ("gro_last", prog["rounds"][-1])):   # host-computed HBM offsets
        for j in range(16):
            want = sum(((j >> i) & 1) << tile_pos[rd["R"][i]] for i in range(4))
            assert prog[which][j] == want, f"{which}[{j}]"
    for ri, rd in enumerate(prog["rounds"]):
        R = rd["R"]
        assert sorted(set(R)) == R and len(R) == 4
        assert sorted(rd["tb"] + R) == list(range(12)), "register bits and thread bits must partition the tile bits"
        if ri in (0, n_rounds - 1):
            io_low = int(os.environ.get("QB_TILE_IO_LOW", "5"))      # experimental: 4 / 3 allow register bits 4 / 3 there too
            io_low = io_low if io_low in (3, 4, 5) else 5
            assert all(b >= io_low for b in R), "HBM-facing rounds keep lanes on the low tile bits"
            assert rd["tb"][:io_low] == list(range(io_low)), "a warp access covers whole 128-byte lines"
        for j in range(16):                     # host-computed swizzled offsets
            l = sum(((j >> i) & 1) << R[i] for i in range(4))
            assert rd["sro"][j] == l ^ (((l >> 3) ^ (l >> 6) ^ (l >> 9)) & 7)
        rmask_bits = sum(1 << b for b in R)
        j = _bits_to(loc, R)                    # register index of every amplitude
        tl = loc & ~rmask_bits                  # its thread's tile-local bits
        gbit = [1 << tile_pos[b] for b in R]    # global bit of register bit i
        tph = np.ones(N, dtype=np.complex128)

        def sel_bit(sc):
            return ((blk >> (sc & 0x3f)) & 1) if (sc & 0x80) else ((tl >> (sc & 0x3f)) & 1)

        def pair_update(a, m00, m01, m10, m11, cond):
            """2x2 update on register bit a where `cond` (per amplitude, evaluated on the bit=0 partner) holds."""
            i0 = g[(j >> a) & 1 == 0]
            i1 = i0 | gbit[a]
            c = cond[i0]
            a0, a1 = s[i0].copy(), s[i1].copy()
            m00, m01, m10, m11 = (np.broadcast_to(x, N)[i0] if np.ndim(x) else x for x in (m00, m01, m10, m11))
            s[i0] = np.where(c, m00 * a0 + m01 * a1, a0)
            s[i1] = np.where(c, m10 * a0 + m11 * a1, a1)

        for op in prog["ops"][rd["op_begin"]:rd["op_end"]]:
            code, rmask, tmask, mat = op["code"], op["rmask"], op["tmask"], op["mat"]
            live = (blk & op["omask"]) == op["omask"]          # `continue` in the kernel
            act = (tl & tmask) == tmask
            rsel = (j & rmask) == rmask
            if code < RC_D2 and (code & 12) != 8 and code not in range(12, 16) and code not in range(28, 32):
                a = code & 3                                       # D1C / D1R, plain or predicated
                pred = rsel if code >= 16 else np.ones(N, dtype=bool)
                pair_update(a, mats[mat], mats[mat + 1], mats[mat + 2], mats[mat + 3], live & act & pred)
            elif code in range(RC_DG1R, RC_DG1R + 4) or code in range(RC_DG1R_P, RC_DG1R_P + 4):
                a = code & 3
                pred = rsel if code >= 16 else np.ones(N, dtype=bool)
                f = np.where((j >> a) & 1 == 1, mats[mat + 1], mats[mat])
                s = np.where(live & act & pred, f * s, s)
            elif code == RC_D2:
                a, b = op["a"], op["b"]
                M = mats[mat:mat + 16].reshape(4, 4)
                base = g[((j >> a) & 1 == 0) & ((j >> b) & 1 == 0)]
                c = (live & rsel)[base]
                ix = [base, base | gbit[a], base | gbit[b], base | gbit[a] | gbit[b]]
                old = [s[i].copy() for i in ix]
                for r in range(4):
                    new = sum(M[r, q] * old[q] for q in range(4))
                    s[ix[r]] = np.where(c, new, old[r])
            elif code in (RC_PH_T, RC_PH_ACC):
                tph = np.where(live & act, mats[mat] * tph, tph)
            elif code == RC_PH_NEG:
                s = np.where(live & act & rsel, -s, s)
            elif code == RC_PH_R:
                s = np.where(live & act & rsel, mats[mat] * s, s)
            elif code == RC_DG1T:
                f = np.where(sel_bit(op["src"][0]) == 1, mats[mat + 1], mats[mat])
                tph = np.where(live & act, f * tph, tph)
            elif code == RC_DG:
                idx = np.zeros(N, dtype=np.int64)
                any_reg = False
                for i in range(op["k"]):
                    sc = op["src"][i]
                    if sc & 0xc0:
                        idx |= sel_bit(sc) << i
                    else:
                        idx |= ((j >> (sc & 3)) & 1) << i
                        any_reg = True
                f = mats[mat + idx]
                if not any_reg and rmask == 0:
                    tph = np.where(live & act, f * tph, tph)
                else:
                    s = np.where(live & act & rsel, f * s, s)
            elif code in range(RC_H, RC_H + 4):
                assert tmask == 0 and rmask == 0 and op["omask"] == 0
                pair_update(code - RC_H, 1.0, 1.0, 1.0, -1.0, np.ones(N, dtype=bool))      # unscaled: the scalar is in a table
            elif code in range(RC_PH_FLUSH, RC_PH_FLUSH + 4):
                b = code - RC_PH_FLUSH
                s = np.where((j >> b) & 1 == 1, tph * s, s)
                tph = np.ones(N, dtype=np.complex128)
            elif code in (RC_TAB, RC_TAB_ACC):
                assert tmask == 0 and rmask == 0 and op["omask"] == 0
                sels = op["src"] + [op["b"]]
                idx = np.zeros(N, dtype=np.int64)
                for i in range(op["k"]):
                    idx |= sel_bit(sels[i]) << i
                tph = mats[mat + idx] * tph
            elif code == RC_LAYER:
                assert op["omask"] == 0
                present, hmask = op["k"] & 15, op["k"] >> 4
                tms = [tmask, rmask, op["src"][0] | (op["src"][1] << 8), op["src"][2] | (op["b"] << 8)]
                for sl in range(4):
                    if not (present >> sl) & 1:
                        continue
                    if (hmask >> sl) & 1:
                        pair_update(sl, 1.0, 1.0, 1.0, -1.0, np.ones(N, dtype=bool))
                    else:
                        c0, c1 = mats[mat + 2 * sl], mats[mat + 2 * sl + 1]
                        pair_update(sl, c0.real, c0.imag, c1.real, c1.imag, (tl & tms[sl]) == tms[sl])
            else:
                raise AssertionError(f"unknown op code {code}")
        if rd["has_tph"]:
            s = tph * s
        else:
            assert np.all(tph == 1), "a round wrote the thread phase without flagging has_tph"
    return s
