"""Callers of the hot path: PRP, Lucas-Lehmer (classic and LL-safe) and P-1 (stage 1, V-trace
stage 2) loops over the engine API.

Host loops only -- every arithmetic step is an engine call (square_mul, mul, set_multiplicand, copy,
sub), exactly as src/modes does it; nothing here touches digits except the final residue export.
They take any object with the engine methods (the CUDA `Engine`, or the kernel emulator in the CPU
tests).  Restated from: RunPrpOrLlMarin.cpp:213-462 (PRP type 1 base 3 with Gerbicz-Li checking and
-erroriter injection; classic LL), AlgoUtils.hpp:165-224 (residue words, /9, res64 / res2048),
RunPM1.cpp:7016-7135,7263-7299 (stage 1, E = 2 p prod q^floor(log_q B1), "fast3" squarings).
"""
import math

import numpy as np

from . import hostmath
from .engine import digits_to_int

# register map of RunPrpOrLlMarin.cpp:213
R0, R1, R2, R3, R4, R5, RBASE, RTMP = range(8)


def _digits(eng, r):
    return eng.get(r) if hasattr(eng, "get") else eng.get_digits(r)


def _value(eng, r, p):
    v = digits_to_int(_digits(eng, r))
    return 0 if v == (1 << p) - 1 else v


def _same(eng, a, b, p):
    """engine::is_equal (engine.h:148-157): compared on the device when the engine can (no 10^8-bit integers on the
    host per Gerbicz check), through the canonical values otherwise."""
    if hasattr(eng, "is_equal"):
        return eng.is_equal(a, b)
    return _value(eng, a, p) == _value(eng, b, p)


def pack_words(enc, p):
    """pack_words_from_eng_digits (AlgoUtils.hpp:165-187): the residue as little-endian 32-bit words."""
    v = digits_to_int(enc)
    nwords = (p + 31) // 32
    return np.frombuffer(v.to_bytes(nwords * 4 + 8, "little")[:nwords * 4], dtype=np.uint32).copy()


def words_to_int(words):
    return int.from_bytes(np.asarray(words, dtype=np.uint32).tobytes(), "little")


def prp3_div9(p, r):
    """prp3_div9 (AlgoUtils.hpp:195-206): residue / 9 modulo 2^p - 1 (two exact divisions by 3)."""
    m = (1 << p) - 1
    for _ in range(2):
        r = (r + ((3 - r % 3) % 3) * m) // 3 if m % 3 else r
    return r % m


def format_res64(r):
    return f"{r & 0xFFFFFFFFFFFFFFFF:016X}"


def format_res2048(r):
    return f"{r & ((1 << 2048) - 1):0512x}"


def run_prp(eng, p, gerbicz=True, check_level=0, error_iter=0, log=None, max_restores=4, stop_after_checks=0, known_factors=(),
            checkpoint_path=None, backup_every=0, stop_at=None):
    """PRP test of 2^p - 1 (type 1, base 3): p squarings from 3, prime iff the residue is 9.

    Gerbicz-Li checking as the reference: every B = floor(sqrt(p)) iterations the accumulator
    d <- d * x, and every `check_level` blocks (and at the end) d is checked against
    (d_prev)^(2^B) * 3 recomputed from the previous accumulator; on mismatch the last good state is
    restored (RunPrpOrLlMarin.cpp:338-412).  `error_iter` injects a fault (x -= 2) after that
    iteration, like -erroriter (RunPrpOrLlMarin.cpp:326-336).  `stop_after_checks` ends the run after that
    many passed checks (the reference's size-ladder smoke test stops at the first "Check passed",
    unit_test_all.sh:54-79); the result then carries `stopped_at` instead of a verdict.  `known_factors` (-factors)
    turns the verdict into the cofactor test of src/math/Cofactor.cpp:33-63: with KF their product and N = (2^p - 1) / KF,
    N is a probable prime iff residue / 9 == 3^(KF - 1) (mod N); the reported residue stays the full-modulus one
    (src/io/JsonBuilder.cpp:182-217).
    Checkpoints (RunPrpOrLlMarin.cpp:153-226,296-320): with `checkpoint_path` the run resumes from that file (or its
    `.old`) when it holds this exponent, this mode and THIS backend's tag, saves every `backup_every` iterations (the
    reference saves on a timer, -t seconds) and, at iteration `stop_at` (a stand-in for the reference's SIGINT
    handler), saves and returns {"interrupted": iteration}.  The restored state becomes the last verified one, as the
    reference does (`copy(R4, R0); copy(R5, R1)` after the restore).
    """
    from . import checkpoint as ckpt
    say = log or (lambda s: None)
    total = p
    B = max(int(math.isqrt(p)), 1)
    level_auto = int(1000 * 600.0 / B) or max((total // B) // max(int(math.isqrt(B)), 1), 1)
    level = check_level if check_level > 0 else max(level_auto, 1)
    ri = 0
    if checkpoint_path:
        got = ckpt.load_latest(checkpoint_path, eng, p, ckpt.MODE_PRP)
        if got:
            ri = got[0]
            say(f"Resuming from a checkpoint at iteration {ri}.")
    if ri == 0:
        eng.set(R1, 1)
        eng.set(R0, 3)
    eng.copy(R4, R0); eng.copy(R5, R1)
    eng.set(RBASE, 3)
    eng.set_multiplicand(RTMP, RBASE)
    itersave, jsave = (ri - 1, total - ri) if ri else (0, total - 1)
    checkpass, errors, restores, errordone, passed = 0, 0, 0, False, 0
    it, j = ri, total - 1 - ri
    while it < total:
        if checkpoint_path and it > ri and ((stop_at is not None and it == stop_at) or (backup_every and it % backup_every == 0)):
            ckpt.save(checkpoint_path, eng, p, it, mode=ckpt.MODE_PRP)
            if stop_at is not None and it == stop_at:
                say(f"Interrupted by user, state saved at iteration {it}")
                return {"exponent": p, "worktype": "PRP-3", "interrupted": it, "errors_gerbicz": errors, "fft_length": eng.get_size()}
        eng.square_mul(R0)
        if error_iter > 0 and it + 1 == error_iter and not errordone:
            errordone = True
            eng.sub(R0, 2)
            say(f"Injected error at iteration {it + 1}")
        if gerbicz and ((j != 0 and j % B == 0) or it == total - 1):
            checkpass += 1
            eng.copy(R3, R1)
            eng.set_multiplicand(R2, R0)
            eng.mul(R1, R2)
            if checkpass == level or it == total - 1:
                checkpass = 0
                modB = B if p % B == 0 else p % B
                for _ in range(B - modB - 1 if B > modB else 0):
                    eng.square_mul(R3)
                if p % B == 0:
                    eng.mul(R3, RTMP)
                else:
                    eng.square_mul(R3, 3)
                for _ in range(modB):
                    eng.square_mul(R3)
                if not _same(eng, R3, R1, p):
                    say(f"[Gerbicz Li] Check FAILED! iter={it + 1}")
                    say(f"[Gerbicz Li] Restore iter={itersave} (j={jsave})")
                    errors += 1
                    restores += 1
                    if restores > max_restores:
                        raise RuntimeError("Gerbicz-Li: too many restores")
                    eng.copy(R0, R4); eng.copy(R1, R5)
                    it, j, checkpass = itersave, jsave, 0
                    if it == 0 and ri == 0:
                        continue            # restart from the very beginning (iter = -1 + 1)
                    it += 1; j -= 1
                    continue
                say(f"[Gerbicz Li] Check passed! iter={it + 1}")
                eng.copy(R4, R0); eng.copy(R5, R1)
                itersave, jsave = it, j
                passed += 1
                if stop_after_checks and passed >= stop_after_checks:
                    return {"exponent": p, "worktype": "PRP-3", "stopped_at": it + 1, "checks_passed": passed,
                            "errors_gerbicz": errors, "fft_length": eng.get_size()}
        it += 1; j -= 1
    enc = _digits(eng, R0)
    x = digits_to_int(enc)
    m = (1 << p) - 1
    x = 0 if x == m else x
    r = prp3_div9(p, x)
    is_prp, rtype = x == 9 % m, 1
    if known_factors:
        kf = 1
        for f in known_factors:
            kf *= int(f)
        if m % kf:
            raise ValueError("known factors do not divide 2^p - 1")
        cof = m // kf
        is_prp, rtype = r % cof == pow(3, kf - 1, cof), 5
    return {"exponent": p, "worktype": "PRP-3", "is_prp": is_prp, "res64": format_res64(r), "res2048": format_res2048(r),
            "residue_type": rtype, "errors_gerbicz": errors, "fft_length": eng.get_size(), "iterations": total}


def result_json(r, known_factors=(), program_version="b200"):
    """The result line of a finished run with the reference's field names, order and values (JsonBuilder.cpp:342-470: status,
    exponent, worktype, factors / known-factors, b1, b2, res64, res2048, residue-type, errors, shift-count, curve fields,
    fft-length, program) -- the parity surface of the drivers.  `r`: the dict a driver of this module or
    of ecm.py returns.  The proof, os, user, timestamp and checksum fields of the reference belong to its I/O shell."""
    import json
    wt = r.get("worktype", "")
    out = {}
    if wt == "PRP-3":
        out["status"] = "P" if r["is_prp"] else "C"
    elif wt == "LL":
        out["status"] = "P" if r["is_prime"] else "C"
    else:
        out["status"] = "F" if r.get("factor") else "NF"
    out["exponent"] = r["exponent"]
    out["worktype"] = wt
    known = [str(int(f)) for f in known_factors]
    if wt in ("PRP-3", "LL"):
        if known:
            out["known-factors"] = known
    else:
        if r.get("factor"):
            out["factors"] = [str(r["factor"])]
        if known:
            out["known-factors"] = known
    if r.get("b1"):
        out["b1"] = r["b1"]
        if r.get("b2", 0) > r["b1"]:
            out["b2"] = r["b2"]
    if wt in ("PRP-3", "LL"):
        out["res64"] = r["res64"]
        if wt == "PRP-3":
            out["res2048"] = r["res2048"]
            out["residue-type"] = r["residue_type"]
        out["errors"] = {"gerbicz": r.get("errors_gerbicz", 0)}
        out["shift-count"] = 0
    elif wt == "ECM":
        if r.get("curves_tested_for_found"):
            out["curves"] = r["curves_tested_for_found"]
        out["curve-type"] = "Edwards"
        out["torsion-subgroup"] = {None: 0, "none": 0, "16": 16}.get(r.get("torsion"), 8)
        out["base-seed"] = r.get("seed")
        out["errors"] = {"invariant": sum(1 for c in r.get("curves", []) if c.get("status") == "invariant")}
    elif wt == "PM1":
        out["errors"] = {"gerbicz": r.get("errors_gerbicz", 0)}
    if r.get("fft_length"):
        out["fft-length"] = r["fft_length"]
    out["program"] = {"name": "prmers", "version": program_version}
    return json.dumps(out, separators=(",", ":"))


def run_ll(eng, p, error_iter=0):
    """Classic Lucas-Lehmer (-llunsafe, RunPrpOrLlMarin.cpp:229,322-324): s <- s^2 - 2, p - 2 times from 4."""
    eng.set(R0, 4)
    for it in range(p - 2):
        eng.square_mul(R0)
        eng.sub(R0, 2)
        if error_iter and it + 1 == error_iter:
            eng.sub(R0, 2)
    x = _value(eng, R0, p)
    return {"exponent": p, "worktype": "LL", "is_prime": x == 0, "res64": format_res64(x), "fft_length": eng.get_size()}


def primes_upto(n):
    sieve = bytearray([1]) * (n + 1)
    sieve[:2] = b"\0\0"
    for i in range(2, int(n ** 0.5) + 1):
        if sieve[i]:
            sieve[i * i::i] = bytearray(len(sieve[i * i::i]))
    return [i for i in range(n + 1) if sieve[i]]


def pm1_exponent(p, b1):
    """buildE (AlgoUtils.hpp:248-336): E = 2 p prod_{q <= B1} q^floor(log_q B1)."""
    e = 2 * p
    for q in primes_upto(b1):
        qk = q
        while qk * q <= b1:
            qk *= q
        e *= qk
    return e


def pm1_exponent_increment(b1_old, b1_new):
    """buildE_incremental_fast (RunPM1.cpp:5771-5870): E(B1_new) / E(B1_old) = the new primes and the prime powers that
    fit under the higher bound only (the factor 2 p is in both)."""
    if b1_new <= b1_old:
        return 1
    e = 1
    for q in primes_upto(b1_new):
        k_old = k_new = 0
        qk = q
        while qk <= b1_new:
            k_new += 1
            if qk <= b1_old:
                k_old += 1
            qk *= q
        e *= q ** (k_new - k_old)
    return e


def run_pm1_stage1_extend(eng, p, b1_old, b1_new, H_old, gcd=True, resume_dir=None, progress=None):
    """Stage 1 extended from a finished lower bound (`-b1old`, RunPM1.cpp:5901-5945, 6070-6122, 6271-6300, 6334-6345): the
    residue H_old of bound B1_old comes from its resume file (read_pm1_resume), H_new = H_old^(E(B1_new) / E(B1_old)) by
    square-and-multiply with H_old as the multiplicand (RBASE) -- the base is no longer 3, so this is not the fast3
    chain -- and the new residue is what a stage 1 at B1_new gives.  Writes the resume files of the new bound into
    `resume_dir` (RunPM1.cpp:7268-7274)."""
    m = (1 << p) - 1
    E = pm1_exponent_increment(b1_old, b1_new)
    bits = E.bit_length()
    rbase = PM1_RBASE if getattr(eng, "reg_count", 11) > PM1_RBASE else 1     # the reference's register map, or two registers
    _set_int(eng, rbase, H_old % m, p)
    eng.set_multiplicand(rbase, rbase)
    eng.set(PM1_RSTATE, 1)
    for k in range(bits - 1, -1, -1):
        eng.square_mul(PM1_RSTATE)
        if (E >> k) & 1:
            eng.mul(PM1_RSTATE, rbase)
        if progress and (bits - k) % 1000 == 0:
            progress(bits - k, bits)
    x = _value(eng, PM1_RSTATE, p)
    g = hostmath.gcd((x - 1) % m, m) if gcd else 1
    out = {"exponent": p, "worktype": "PM1", "b1": b1_new, "b1_old": b1_old, "bits": bits, "x": x,
           "factor": g if 1 < g < m else None, "fft_length": eng.get_size()}
    if resume_dir:
        out["resume_files"] = write_pm1_resume(resume_dir, p, b1_new, x)
    return out


# register map of runPM1Marin (RunPM1.cpp:6206-6216)
PM1_RSTATE, PM1_RACC_L, PM1_RACC_R, PM1_RCHK, PM1_RPOW, PM1_RTMP, PM1_RSTART, PM1_RSAVE_S, PM1_RSAVE_L, PM1_RSAVE_R, PM1_RBASE = range(11)


def _fast3_pow(eng, r, e):
    """r <- 3^e by the same fast3 squarings the main loop uses (RunPM1.cpp:7161-7167)."""
    eng.set(r, 1)
    for k in range(e.bit_length() - 1, -1, -1):
        eng.square_mul(r, 3 if (e >> k) & 1 else 1)


def run_pm1_stage1(eng, p, b1, reg=0, gcd=True, gerbicz=None, check_level=0, error_iter=0, log=None, max_restores=8,
                   checkpoint_path=None, backup_every=0, stop_at=None):
    """P-1 stage 1, "fast3" (RunPM1.cpp:7016-7300): X <- X^2 * (3 if bit else 1) over the bits of
    E = 2 p prod q^floor(log_q B1), MSB first; the factor is gcd(X - 1, 2^p - 1).

    Gerbicz-Li checking as the reference does it for an exponentiation (RunPM1.cpp:7102-7190): the bits are cut in
    blocks of B = floor(sqrt(bits)); for a lot of full blocks L <- L * (state at block start), R <- R * (state at block
    end), eacc += (bits of the block), and the check is L^(2^B) * 3^eacc == R; a shorter first block is checked on its
    own (start^(2^len) * 3^wbits == state).  A mismatch restores the state saved at the last passed check and redoes
    the blocks since.  On by default when the engine has the 11 registers of the reference's map; `error_iter` injects
    a fault like -erroriter (RunPM1.cpp:7134).  The reference compares its block counter with a variable that stays 0
    (RunPM1.cpp:7155 `gl_checkpass == checkpass`), so its lot check only ever runs at the last bit; here the computed
    level (every `check_level` full blocks, default the reference's 600 s formula) is honoured.
    Checkpoints (RunPM1.cpp:6219-6267; Gerbicz-Li runs only): with `checkpoint_path` the run resumes from that file --
    all registers plus the bit index and the block bookkeeping (checkpass, blocks since the check, bits in the block,
    block length, in_lot, eacc, wbits) -- saves every `backup_every` bits and, after `stop_at` bits (the reference's
    SIGINT path), saves and returns {"interrupted": bits done}."""
    from . import checkpoint as ckpt
    say = log or (lambda s: None)
    E = pm1_exponent(p, b1)
    bits = E.bit_length()
    regs = getattr(eng, "reg_count", None)
    if gerbicz is None:
        gerbicz = regs is not None and regs >= 11 and reg == 0
    m = (1 << p) - 1
    errors = 0
    if not gerbicz:
        eng.set(reg, 1)
        for k in range(bits - 1, -1, -1):
            eng.square_mul(reg, 3 if (E >> k) & 1 else 1)
            if error_iter and bits - k == error_iter:
                eng.sub(reg, 2)
        state = reg
    else:
        S, L, Rr, CHK, POW, TMP, START, SAVE_S = PM1_RSTATE, PM1_RACC_L, PM1_RACC_R, PM1_RCHK, PM1_RPOW, PM1_RTMP, PM1_RSTART, PM1_RSAVE_S
        B = max(1, math.isqrt(bits))
        level_auto = int(1000 * 600.0 / B) or max((bits // B) // max(math.isqrt(B), 1), 1)
        level = check_level if check_level > 0 else max(level_auto, 1)
        eng.set(PM1_RBASE, 3); eng.set(S, 1)
        eacc = wbits = 0
        gl_checkpass = blocks_since_check = bits_in_block = 0
        current_block_len = ((bits - 1) % B) + 1
        in_lot, errordone, restores = False, False, 0
        i = bits
        got = ckpt.load_pm1(checkpoint_path, eng, p) if checkpoint_path else None
        if got and got[1]["bits"] == bits:
            i, st = got[0], got[1]
            gl_checkpass, blocks_since_check, bits_in_block = st["checkpass"], st["blocks_since_check"], st["bits_in_block"]
            current_block_len, in_lot, eacc, wbits = st["current_block_len"], st["in_lot"], st["eacc"], st["wbits"]
            say(f"Resuming from a checkpoint at bit {bits - i}.")
        elif got:
            eng.set(PM1_RBASE, 3); eng.set(S, 1)         # a checkpoint of another bound: start over
        resumed_at = i
        while i > 0:
            if checkpoint_path and i < resumed_at and ((stop_at is not None and bits - i == stop_at) or (backup_every and (bits - i) % backup_every == 0)):
                ckpt.save_pm1(checkpoint_path, eng, p, i, {"checkpass": gl_checkpass, "blocks_since_check": blocks_since_check,
                              "bits_in_block": bits_in_block, "current_block_len": current_block_len, "in_lot": in_lot, "eacc": eacc,
                              "wbits": wbits, "processed_bits": bits - i, "bits": bits})
                if stop_at is not None and bits - i == stop_at:
                    say(f"Interrupted by user, state saved at bit {bits - i}")
                    return {"exponent": p, "worktype": "PM1", "b1": b1, "interrupted": bits - i, "bits": bits, "errors_gerbicz": errors,
                            "fft_length": eng.get_size()}
            if bits_in_block == 0:
                current_block_len = ((i - 1) % B) + 1
                if current_block_len == B:
                    if gl_checkpass == 0 and blocks_since_check == 0 and wbits == 0 and eacc == 0:
                        eng.set(L, 1); eng.set(Rr, 1); eng.copy(SAVE_S, S)
                        eng.set(PM1_RSAVE_L, 1); eng.set(PM1_RSAVE_R, 1)
                        in_lot = True
                else:
                    in_lot = False
                    gl_checkpass = blocks_since_check = 0
                    eacc = wbits = 0
                eng.copy(START, S)
            b = (E >> (i - 1)) & 1
            eng.square_mul(S, 3 if b else 1)
            wbits = (wbits << 1) | b
            bits_in_block += 1
            if error_iter > 0 and bits - i + 1 == error_iter and not errordone:
                errordone = True
                eng.sub(S, 2)
                say(f"Injected error at iteration {bits - i + 1}")
            if bits_in_block == current_block_len:
                if current_block_len == B:
                    eng.set_multiplicand(TMP, START); eng.mul(L, TMP)
                    eng.set_multiplicand(TMP, S); eng.mul(Rr, TMP)
                    eacc += wbits
                    blocks_since_check += 1
                    gl_checkpass += 1
                    if in_lot and (gl_checkpass == level or i == 1):
                        eng.copy(CHK, L)
                        for _ in range(B):
                            eng.square_mul(CHK)
                        _fast3_pow(eng, POW, eacc)
                        eng.set_multiplicand(TMP, POW); eng.mul(CHK, TMP)
                        if not _same(eng, CHK, Rr, p):
                            say("[Gerbicz Li] Mismatch : Last correct state will be restored")
                            errors += 1; restores += 1
                            if restores > max_restores:
                                raise RuntimeError("Gerbicz-Li: too many restores")
                            eng.copy(S, SAVE_S); eng.set(L, 1); eng.set(Rr, 1); eng.copy(START, S)
                            i += blocks_since_check * B - 1          # back to the first bit of the lot (the loop's --i is not taken)
                            eacc = wbits = 0
                            blocks_since_check = gl_checkpass = bits_in_block = 0
                            continue
                        say("[Gerbicz Li] Check passed")
                        eng.copy(SAVE_S, S); eng.set(L, 1); eng.set(Rr, 1)
                        eacc = 0; blocks_since_check = gl_checkpass = 0
                else:
                    eng.copy(CHK, START)
                    for _ in range(current_block_len):
                        eng.square_mul(CHK)
                    _fast3_pow(eng, POW, wbits)
                    eng.set_multiplicand(TMP, POW); eng.mul(CHK, TMP)
                    if not _same(eng, CHK, S, p):
                        say("[Gerbicz Li] Mismatch : Last correct state will be restored")
                        errors += 1; restores += 1
                        if restores > max_restores:
                            raise RuntimeError("Gerbicz-Li: too many restores")
                        eng.copy(S, START)
                        i += current_block_len - 1
                        wbits = bits_in_block = 0
                        continue
                    say("[Gerbicz Li] Check passed")
                    eng.copy(SAVE_S, S); eng.set(L, 1); eng.set(Rr, 1)
                    eacc = 0; blocks_since_check = gl_checkpass = 0
                bits_in_block = 0
                wbits = 0
            i -= 1
        state = S
    x = _value(eng, state, p)
    g = hostmath.gcd((x - 1) % m, m) if gcd else 1
    return {"exponent": p, "worktype": "PM1", "b1": b1, "factor": g if 1 < g < m else None, "bits": bits,
            "fft_length": eng.get_size(), "x": x, "errors_gerbicz": errors}


# ---- canonical P-1 interchange files (backend independent): GMP-ECM resume line and Prime95 stage-1 save file
# (AlgoUtils.hpp:465-508,600-733; written by the reference after stage 1, RunPM1.cpp:7263-7275)
_CHKSUMMOD = 4294967291


def pm1_resume_line(p, b1, x):
    """writeEcmResumeLine (AlgoUtils.hpp:487-508): one line a GMP-ECM / PrMers stage 2 can resume from."""
    chk = (b1 % _CHKSUMMOD) * (((1 << p) - 1) % _CHKSUMMOD) % _CHKSUMMOD * (x % _CHKSUMMOD) % _CHKSUMMOD
    return (f"METHOD=P-1; B1={b1}; N=2^{p}-1; X=0x{x:x}; CHECKSUM={chk}; "
            "PROGRAM=PrMers; X0=0x3; Y=0x0; Y0=0x0; WHO=; TIME=;\n")


def parse_pm1_resume_line(line):
    """The fields the reference reads back (AlgoUtils.hpp:585-612); verifies the checksum."""
    f = dict(kv.strip().split("=", 1) for kv in line.strip().rstrip(";").split(";") if "=" in kv)
    b1, p, x = int(f["B1"]), int(f["N"][2:].split("-")[0]), int(f["X"], 16)
    chk = (b1 % _CHKSUMMOD) * (((1 << p) - 1) % _CHKSUMMOD) % _CHKSUMMOD * (x % _CHKSUMMOD) % _CHKSUMMOD
    if "CHECKSUM" in f and int(f["CHECKSUM"]) != chk:
        raise ValueError("P-1 resume line: checksum mismatch")
    return {"p": p, "b1": b1, "x": x}


def pm1_p95_bytes(p, b1, x, date_start="", date_end="", version="b200"):
    """write_prime95_s1_from_bytes (AlgoUtils.hpp:641-733): Prime95's stage-1 save file for the residue X."""
    import struct, zlib
    h = f"{x:x}"
    h = "0" * ((8 - (len(h) & 7)) & 7) + h
    data = bytes.fromhex(h)[::-1]                                   # little-endian bytes, padded to whole 32-bit words
    s32 = sum(struct.unpack(f"<{len(data) // 4}I", data)) if data else 0
    chk = ((b1 << 1) + 6 + (len(data) >> 1) + s32) & 0xFFFFFFFF
    out = struct.pack("<IIdiIiBBHQdIiQQii", 830093643, 8, 1.0, 2, p, -1, ord("S"), ord("1"), 0, 0, 1.0, chk, 5, b1, b1, 1, len(data) >> 2)
    out += data
    js = (',"programs":[{"work":{"type":"PM1","stage":"1"},"program":{"name":"prmers","version":"' + version + '"},"os":{"os":"Linux",'
          '"architecture":"x86_64"},"date_start":"' + date_start + '","date_end":"' + date_end + '"}]').encode()
    out += b"MOREINFOJSONDATA" + struct.pack("<III", 8 + len(js), 1, zlib.crc32(js) & 0xFFFFFFFF) + js
    return out


def parse_pm1_p95(buf):
    """read_prime95_s1_to_bytes (AlgoUtils.hpp:735-800): exponent, B1 and residue of a Prime95 stage-1 file."""
    import struct
    fmt = "<IIdiIiBBHQdIiQQii"
    hd = struct.unpack_from(fmt, buf, 0)
    if hd[0] != 830093643:
        raise ValueError("not a Prime95 P-1 save file")
    p, chk, b1, nwords = hd[4], hd[11], hd[13], hd[16]
    off = struct.calcsize(fmt)
    data = buf[off:off + 4 * nwords]
    if len(data) != 4 * nwords:
        raise ValueError("truncated Prime95 P-1 save file")
    s32 = sum(struct.unpack(f"<{nwords}I", data)) if nwords else 0
    if ((b1 << 1) + 6 + (len(data) >> 1) + s32) & 0xFFFFFFFF != chk:
        raise ValueError("Prime95 P-1 save file: checksum mismatch")
    return {"p": p, "b1": b1, "x": int.from_bytes(data, "little")}


def write_pm1_resume(directory, p, b1, x):
    """resume_p<p>_B1_<b1>.save / .p95 as RunPM1.cpp:7268-7274 names them; returns the two paths."""
    import os
    base = os.path.join(directory, f"resume_p{p}_B1_{b1}")
    with open(base + ".save", "w") as f:
        f.write(pm1_resume_line(p, b1, x))
    with open(base + ".p95", "wb") as f:
        f.write(pm1_p95_bytes(p, b1, x))
    return base + ".save", base + ".p95"


# ---- P-1 stage 2, V-trace (RunPM1.cpp:1931-4335): register map of RunPM1.cpp:3192-3209 ------------------
RSTATE, RACC, RV1, RMUL_V1, RVD, RMUL_VD, RGPREV, RGCUR, RGNEXT, RBPREV, RBCUR, RBNEXT, RTMP2, RMUL_TMP = range(14)
BABY_BASE = 14


def _set_int(eng, r, x, p):
    if hasattr(eng, "set_mpz"):
        eng.set_mpz(r, x)
    else:
        from .engine import int_to_digits, _widths
        eng.set_digits(r, int_to_digits(x % ((1 << p) - 1), _widths(eng.get_size(), p)))


def vtrace_choose_D(b1, max_babies=480):
    """D = product of the smallest primes, all <= B1 (so every stage-2 prime is coprime to D), as large as the
    baby table allows; babies are the j <= D/2 coprime to D (phi(D)/2 registers)."""
    D, phi = 2, 1
    for q in (3, 5, 7, 11, 13):
        if q > b1 or phi * (q - 1) // 2 > max_babies:
            break
        D *= q; phi *= q - 1
    return max(D, 4)


def vtrace_regs_needed(D):
    return BABY_BASE + sum(1 for j in range(1, D // 2 + 1) if math.gcd(j, D) == 1)


def _lucas_jump(eng, k):
    """(RGPREV, RGCUR) <- (V_{(k-1) D}, V_{k D}) by the binary Lucas ladder on V_D:
    V_{2m} = V_m^2 - 2, V_{2m+1} = V_m V_{m+1} - V_D.  Lets a rank start its giant steps anywhere."""
    if k == 0:
        raise ValueError("k must be >= 1")
    A, B, T, MT = RGPREV, RGCUR, RGNEXT, RMUL_TMP
    eng.set(A, 2); eng.copy(B, RVD)                       # (V_0, V_1) of the sequence with parameter V_D
    for bit in bin(k - 1)[2:] if k > 1 else "":
        # T = A * B - V_D  (= V_{2m+1})
        eng.set_multiplicand(MT, A); eng.copy(T, B); eng.mul(T, MT); eng.sub_reg(T, RVD)
        if bit == "1":
            eng.square_mul(B); eng.sub(B, 2)               # B = V_{2m+2}
            eng.copy(A, T)                                 # A = V_{2m+1}
        else:
            eng.square_mul(A); eng.sub(A, 2)               # A = V_{2m}
            eng.copy(B, T)                                 # B = V_{2m+1}


def euler_phi(n):
    r, x, q = n, n, 2
    while q * q <= x:
        if x % q == 0:
            while x % q == 0:
                x //= q
            r -= r // q
        q += 1
    if x > 1:
        r -= r // x
    return r


VTRACE_BASE_REGS = BABY_BASE
VTRACE_D_CANDIDATES = (30, 60, 90, 210, 420, 630, 1260, 2310, 4620, 6930, 30030, 60060, 90090)


def vtrace_plan(p, b1, b2, n, free_bytes, D=None, baby_batch=None, max_regs=4096, mem_fraction=0.8):
    """Choose D and the number of baby traces kept on the GPU (RunPM1.cpp:2227-2620 in spirit, README.md:520-562 in
    letter): m(D) = phi(D)/2 baby traces, b of them resident, P = ceil(m/b) passes over the giant steps.  Cost model
    of the README: P * (B2 - B1) / D giant steps of ~4 engine ops, one term (4 ops) per prime pair, D baby steps of 4
    ops per pass.  `free_bytes` is what cudaMemGetInfo reports (the reference asks OpenCL, RunPM1.cpp:1994-2026); a
    register costs 8 n bytes.  A forced D and / or batch (-pm1-vtrace-d, -pm1-vtrace-baby-batch) are honoured."""
    budget = max(int(free_bytes * mem_fraction) // (8 * n) - VTRACE_BASE_REGS, 1)
    budget = min(budget, max_regs - VTRACE_BASE_REGS)
    span = max(b2 - b1, 1)
    nprimes = max(span / max(math.log(max(b2, 3)), 1.0), 1.0)
    best = None
    for d in ([D] if D else [d for d in VTRACE_D_CANDIDATES if d // 2 < b2]):
        if d % 2 or d < 4:
            raise ValueError("D must be even and >= 4")
        m_d = euler_phi(d) // 2
        b = min(baby_batch or m_d, m_d, budget) if not baby_batch else min(baby_batch, m_d)
        if b > budget and baby_batch:
            raise ValueError(f"baby batch {baby_batch} needs {VTRACE_BASE_REGS + b} registers of {8 * n} bytes: more than the device has free")
        passes = -(-m_d // b)
        # every prime pair (k D - j, k D + j) shares a term: the denser the residues, the more pairs
        pair_gain = 1.0 + 0.5 * (1.0 - m_d / (d / 2.0))
        cost = 4.0 * passes * span / d + 4.0 * nprimes / pair_gain + 4.0 * d * passes
        cand = {"D": d, "babies": m_d, "baby_batch": b, "passes": passes, "regs": VTRACE_BASE_REGS + b, "cost": cost}
        if best is None or cost < best["cost"]:
            best = cand
    return best


def run_pm1_stage2_vtrace(eng, p, b1, b2, H, D=None, k_range=None, sieve=None, baby_batch=None, v1_in_register=False, fused_terms=None):
    """P-1 stage 2 on traces V_n = H^n + H^-n: for every prime B1 < q <= B2 write q = k D +- j and
    accumulate (V_{kD} - V_j); both members of a pair share one term.  Returns the accumulator residue
    (an int) and statistics; the factor is gcd(acc, 2^p - 1) (use pm1_stage2_factor).  `k_range`
    restricts the giant steps to [k_lo, k_hi) so that ranks can split the prime range:
    each rank rebuilds the (deterministic) baby table, jumps to its first k by a Lucas ladder and
    accumulates its own product.  `baby_batch` keeps only that many baby traces resident and makes
    ceil(m(D) / baby_batch) passes over the giant steps, one per window of the baby table
    (-pm1-vtrace-baby-batch, README.md:520-562); the accumulator runs through all passes.
    `v1_in_register`: register RV1 already holds V_1 (then H is ignored: no 1/H on the host).
    `fused_terms` (default: whenever the engine offers `add_multiplicand` and one bit of carry headroom): the baby traces
    are kept as MULTIPLICANDS, V_kD is transformed once per giant step, and a term V_kD - V_j is the element-wise sum of
    two multiplicands (the transform is linear) -- `add_multiplicand; mul` per term instead of the reference's
    `copy; add; set_multiplicand; mul` (RunPM1.cpp:3790-3814); same accumulator, same registers."""
    m = (1 << p) - 1
    if fused_terms is None:
        fused_terms = hasattr(eng, "add_multiplicand") and eng.max_multiplier() >= 2
    D = D or vtrace_choose_D(b1)
    babies = [j for j in range(1, D // 2 + 1) if math.gcd(j, D) == 1]
    batch = min(baby_batch or len(babies), len(babies))
    regs = getattr(eng, "reg_count", None)
    if regs is not None and regs < BABY_BASE + batch:
        raise ValueError(f"V-trace stage 2 with {batch} resident baby traces needs {BABY_BASE + batch} registers, the engine has {regs}")
    if sieve is None:
        sieve = bytearray([1]) * (b2 + D + 2)
        sieve[:2] = b"\0\0"
        for i in range(2, int((b2 + D + 1) ** 0.5) + 1):
            if sieve[i]:
                sieve[i * i::i] = bytearray(len(sieve[i * i::i]))
    is_s2_prime = lambda q: b1 < q <= b2 and sieve[q]
    if not v1_in_register:
        V1 = (H + hostmath.invert_mod(H, m)) % m       # (a non-invertible H would itself expose a factor: ValueError)
        _set_int(eng, RV1, V1, p)
    eng.set_multiplicand(RMUL_V1, RV1)
    k_first, k_last = (b1 + 1 + D // 2) // D, (b2 + D // 2) // D
    k_lo, k_hi = k_range if k_range else (k_first, k_last + 1)
    k_lo, k_hi = max(k_lo, k_first), min(k_hi, k_last + 1)
    eng.set(RACC, 1)
    terms = primes = passes = 0
    have_vd = False
    for w0 in range(0, len(babies), batch):
        window = babies[w0:w0 + batch]
        slot = {j: BABY_BASE + i for i, j in enumerate(window)}
        passes += 1
        # baby steps: V_0 = 2, V_1, V_{n+1} = V_1 V_n - V_{n-1}; store -V_j so the term is one add (RunPM1.cpp:3551-3587)
        eng.set(RBPREV, 2); eng.copy(RBCUR, RV1)
        top = D if not have_vd else max(window)
        for n in range(1, top + 1):
            if n in slot:
                eng.set(slot[n], 0); eng.sub_reg(slot[n], RBCUR)
                if fused_terms:
                    eng.set_multiplicand(slot[n], slot[n])          # in place: the slot now holds the multiplicand of -V_j
            if n == top:
                if n == D:
                    eng.copy(RVD, RBCUR); have_vd = True
                break
            eng.copy(RBNEXT, RBCUR); eng.mul(RBNEXT, RMUL_V1); eng.sub_reg(RBNEXT, RBPREV)
            eng.copy(RBPREV, RBCUR); eng.copy(RBCUR, RBNEXT)
        eng.set_multiplicand(RMUL_VD, RVD)
        if k_lo >= k_hi:
            continue
        if k_lo == 0:
            eng.set(RGCUR, 2)                                  # V_0; V_{-D} = V_D
            eng.copy(RGPREV, RVD)
        else:
            _lucas_jump(eng, k_lo)
        for k in range(k_lo, k_hi):
            giant_ready = False
            for j in window:
                lo, hi = k * D - j, k * D + j
                a, b = lo > 0 and is_s2_prime(lo), is_s2_prime(hi)
                if not (a or b):
                    continue
                primes += int(a) + int(b)
                if fused_terms:
                    if not giant_ready:
                        eng.set_multiplicand(RTMP2, RGCUR); giant_ready = True      # multiplicand of V_kD, once per giant step
                    eng.add_multiplicand(RMUL_TMP, RTMP2, slot[j]); eng.mul(RACC, RMUL_TMP)
                else:
                    eng.copy(RTMP2, RGCUR); eng.add(RTMP2, slot[j])          # V_kD + (-V_j)
                    eng.set_multiplicand(RMUL_TMP, RTMP2); eng.mul(RACC, RMUL_TMP)
                terms += 1
            # next giant step: V_{(k+1)D} = V_D V_{kD} - V_{(k-1)D}  (RunPM1.cpp:3788-3797)
            eng.copy(RGNEXT, RGCUR); eng.mul(RGNEXT, RMUL_VD); eng.sub_reg(RGNEXT, RGPREV)
            eng.copy(RGPREV, RGCUR); eng.copy(RGCUR, RGNEXT)
    return {"acc": _value(eng, RACC, p), "terms": terms, "primes": primes, "D": D, "k_range": (k_lo, k_hi),
            "passes": passes, "baby_batch": batch, "fused_terms": bool(fused_terms)}


def run_pm1_stage2_sharded(engine_factory, p, b1, b2, H, D=None, baby_batch=None, known_factors=()):
    """Config 4's second half across the GPUs of a box: the giant-step range is cut in `world` equal
    shares, every rank (one process per GPU, own engine) rebuilds the baby table, jumps to its first giant step and
    accumulates its own product; the only exchange is the gather of the canonical accumulator residues (~p / 8 bytes
    each) at the end, multiplied together modulo 2^p - 1 for ONE gcd (RunPM1.cpp:4211-4212,4261-4262).  Every rank
    returns the merged result."""
    from . import multi
    rank, world = multi.rank_world()
    D = D or vtrace_choose_D(b1)
    share = split_k_range(b1, b2, D, world)[rank]
    nb = min(baby_batch or 10**9, sum(1 for j in range(1, D // 2 + 1) if math.gcd(j, D) == 1))
    eng = engine_factory(p, BABY_BASE + nb)
    try:
        r = run_pm1_stage2_vtrace(eng, p, b1, b2, H, D=D, k_range=share, baby_batch=baby_batch)
    finally:
        if hasattr(eng, "close"):
            eng.close()
    parts = multi.gather([{"rank": rank, "acc": r["acc"], "terms": r["terms"], "primes": r["primes"], "k_range": r["k_range"]}])
    parts.sort(key=lambda d: d["rank"])
    return {"exponent": p, "worktype": "PM1", "b1": b1, "b2": b2, "D": D, "world": world,
            "factor": pm1_stage2_factor(p, [d["acc"] for d in parts], known_factors),
            "terms": sum(d["terms"] for d in parts), "primes": sum(d["primes"] for d in parts), "parts": parts}


def pm1_stage2_factor(p, accs, known_factors=()):
    """gcd of the product of the per-rank accumulators with 2^p - 1 (RunPM1.cpp:4261-4262); factors already known
    (-factors, README.md:495-509) are divided out of it."""
    m = (1 << p) - 1
    acc = 1
    for a in accs:
        acc = hostmath.mulmod_mersenne(acc, a, p)
    g = hostmath.gcd(acc, m)
    for f in known_factors:
        while f > 1 and g % f == 0:
            g //= f
    return g if 1 < g < m else None


def pm1_ultralow_exponent(p, b1, b2, b2_start=0):
    """E2 = E(B1) * 2 p * Q, Q = product of the primes in (max(B1, b2start), B2] (RunPM1.cpp:1808-1813; buildE
    is the product of the prime powers only and the reference multiplies 2 p in once, which `pm1_exponent` includes)."""
    low = max(b1, b2_start)
    e = pm1_exponent(p, b1)
    for q in primes_upto(b2):
        if q > low:
            e *= q
    return e


def run_pm1_stage2_ultralow(eng, p, b1, b2, b2_start=0, gcd=True, progress=None):
    """`-pm1-ultralowmem` stage 2 (RunPM1.cpp:1795-1875, the README's MM31 example README.md:88-98): ONE register, no
    stage-1 residue needed -- 3^(E(B1) 2p Q) is recomputed with the fast3 squarings only (`square_mul(R, 3)` on a set
    bit), which equals H^Q for the stage-1 result H; the factor is gcd(X - 1, 2^p - 1)."""
    E2 = pm1_ultralow_exponent(p, b1, b2, b2_start)
    bits = E2.bit_length()
    R = 0
    eng.set(R, 1)
    for k in range(bits - 1, -1, -1):
        eng.square_mul(R, 3 if (E2 >> k) & 1 else 1)
        if progress and (bits - k) % 1000 == 0:
            progress(bits - k, bits)
    x = _value(eng, R, p)
    m = (1 << p) - 1
    g = hostmath.gcd((x - 1) % m, m) if gcd else 1
    return {"exponent": p, "worktype": "PM1", "b1": b1, "b2": b2, "bits": bits, "x": x,
            "factor": g if 1 < g < m else None, "fft_length": eng.get_size()}


def run_pm1_stage2_lowmem(eng, p, b1, b2, H, gcd=True, progress=None):
    """`-pm1-lowmem` stage 2 (RunPM1.cpp:1878-1926): THREE registers and no baby table -- for every prime q in (B1, B2] the
    stage-1 residue H is restored from the host, RQ = H^q by the engine's square-and-multiply (engine.h:160-170),
    RACC <- RACC (H^q - 1); one gcd at the end.  Returns the accumulator; the factor is gcd(acc, 2^p - 1)."""
    RH, RACC, RQ = 0, 1, 2
    m = (1 << p) - 1
    primes = [q for q in primes_upto(b2) if q > b1]
    eng.set(RACC, 1)
    _set_int(eng, RH, H % m, p)
    raw = None
    if hasattr(eng, "get_data") and hasattr(eng, "set_data"):
        raw = bytearray(eng.get_register_data_size())
        if not eng.get_data(raw, RH):
            raw = None
    powf = eng.pow if hasattr(eng, "pow") else (lambda d, s_, e: _pow(eng, d, s_, e))
    for i, q in enumerate(primes):
        if i:
            if raw is not None:
                eng.set_data(RH, raw)                    # pow turned RH into a multiplicand: the weighted digits come back from the host
            else:
                _set_int(eng, RH, H % m, p)
        powf(RQ, RH, q)
        eng.sub(RQ, 1)
        eng.set_multiplicand(RQ, RQ)
        eng.mul(RACC, RQ)
        if progress and (i + 1) % 100 == 0:
            progress(i + 1, len(primes))
    acc = _value(eng, RACC, p)
    g = hostmath.gcd(acc, m) if gcd else 1
    return {"exponent": p, "worktype": "PM1", "b1": b1, "b2": b2, "primes": len(primes), "acc": acc,
            "factor": g if 1 < g < m else None, "fft_length": eng.get_size()}


def read_pm1_resume(base_path, p=None, b1=None):
    """Stage-1 residue from `resume_p<p>_B1_<b1>.save` or `.p95` (RunPM1.cpp:1518-1560 tries the path as given, then
    `.save`, then `.p95`); the exponent and bound of the file must be the ones asked for."""
    import os
    for path, binary in ((base_path, base_path.endswith(".p95")), (base_path + ".save", False), (base_path + ".p95", True)):
        if not os.path.isfile(path):
            continue
        if binary:
            with open(path, "rb") as f:
                got = parse_pm1_p95(f.read())
        else:
            with open(path) as f:
                got = parse_pm1_resume_line(f.readline())
        if (p is not None and got["p"] != p) or (b1 is not None and got["b1"] != b1):
            raise ValueError(f"{path}: resume file of another exponent or B1")
        return got
    raise FileNotFoundError(base_path)


def run_pm1_stage2_resume2reg(eng, p, b1, b2, H, chunk_bits=200000, gcd=True, progress=None):
    """`-pm1-ultralowmem -pm1-s2-resume2reg` stage 2 (RunPM1.cpp:1507-1793): TWO registers, H from a stage-1 resume file
    (read_pm1_resume) -- the primes of (B1, B2] are multiplied into chunks Q of at most `chunk_bits` bits and
    RSTATE <- RSTATE^Q chunk by chunk (RBASE = the multiplicand of the state the chunk starts from, RSTATE restarts at 1);
    the factor is gcd(H^(prod Q) - 1, 2^p - 1)."""
    RSTATE_, RBASE = 0, 1
    m = (1 << p) - 1
    primes = [q for q in primes_upto(b2) if q > b1]
    _set_int(eng, RSTATE_, H % m, p)
    i, chunks, bits_done = 0, 0, 0
    chunk_bits = max(int(chunk_bits), 1024)
    while i < len(primes):
        Q, i0 = 1, i
        while i < len(primes):
            cand = Q * primes[i]
            if i > i0 and cand.bit_length() > chunk_bits:
                break
            Q = cand; i += 1
            if Q.bit_length() >= chunk_bits:
                break
        chunks += 1
        eng.set_multiplicand(RBASE, RSTATE_)
        eng.set(RSTATE_, 1)
        for k in range(Q.bit_length() - 1, -1, -1):
            eng.square_mul(RSTATE_)
            if (Q >> k) & 1:
                eng.mul(RSTATE_, RBASE)
        bits_done += Q.bit_length()
        if progress:
            progress(i, len(primes))
    x = _value(eng, RSTATE_, p)
    g = hostmath.gcd((x - 1) % m, m) if gcd else 1
    return {"exponent": p, "worktype": "PM1", "b1": b1, "b2": b2, "primes": len(primes), "chunks": chunks, "bits": bits_done, "x": x,
            "factor": g if 1 < g < m else None, "fft_length": eng.get_size()}


def _pow(eng, dst, src, e):
    """engine::pow (engine.h:160-170) for engines without their own: dst = src^e, src becomes a multiplicand."""
    eng.set_multiplicand(src, src)
    eng.set(dst, 1)
    for i in range(int(e).bit_length() - 1, -1, -1):
        eng.square_mul(dst)
        if (e >> i) & 1:
            eng.mul(dst, src)


def split_k_range(b1, b2, D, world):
    """Equal shares of the giant-step range [k_first, k_last] for `world` ranks."""
    k_first, k_last = (b1 + 1 + D // 2) // D, (b2 + D // 2) // D
    total = k_last + 1 - k_first
    cuts = [k_first + total * r // world for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


# ---- LL-safe: Lucas-Lehmer in Z[sqrt 3] with pair Gerbicz-Li (`prmers p -ll`, RunLlSafeMarin.cpp:479-694) ----
(LS_RES_A, LS_RES_B, LS_ACC_A, LS_ACC_B, LS_CHK_A, LS_CHK_B, LS_SAVE_R_A, LS_SAVE_R_B, LS_SAVE_F_A, LS_SAVE_F_B,
 LS_BASE_A, LS_BASE_B, LS_TA, LS_TB, LS_M0, LS_M1, LS_PREV_A, LS_PREV_B) = range(18)


def _pair_square(eng, A, B):
    """(A, B) <- (A^2 + 3 B^2, 2 A B), i.e. (A + B sqrt 3)^2 (RunLlSafeMarin.cpp:481-490)."""
    eng.copy(LS_TA, A); eng.copy(LS_TB, B)
    eng.square_mul(A); eng.square_mul(B, 3); eng.add(A, B)
    eng.copy(B, LS_TA); eng.set_multiplicand(LS_M0, LS_TB); eng.mul(B, LS_M0, 2)


def _pair_mul_by(eng, A, B, Cc, Dd):
    """(A, B) <- (A C + 3 B D, A D + B C) (RunLlSafeMarin.cpp:492-507)."""
    eng.set_multiplicand(LS_M0, Cc); eng.set_multiplicand(LS_M1, Dd)
    eng.copy(LS_TA, A); eng.copy(LS_TB, B)
    eng.copy(A, LS_TA); eng.mul(A, LS_M0)
    eng.copy(B, LS_TA); eng.mul(B, LS_M1)
    eng.copy(LS_TA, LS_TB); eng.mul(LS_TA, LS_M1, 3); eng.add(A, LS_TA)
    eng.copy(LS_TA, LS_TB); eng.mul(LS_TA, LS_M0); eng.add(B, LS_TA)


def run_ll_safe(eng, p, check_level=0, error_iter=0, log=None, max_restores=4, checkpoint_path=None, backup_every=0, stop_at=None):
    """`prmers p -ll`: (2 + sqrt 3)^(2^(p-1)) in Z[sqrt 3] mod 2^p - 1 is (-1, 0) iff 2^p - 1 is prime; the
    running product of the block-end pairs is checked Gerbicz-Li style and a mismatch restores the last
    verified pair (RunLlSafeMarin.cpp:539-694).  Needs 18 registers.
    Checkpoints (RunLlSafeMarin.cpp:418-477,577-603): with `checkpoint_path` the run resumes from that file (all 18
    registers, the iteration count and the last verified iteration), saves every `backup_every` iterations and, at
    iteration `stop_at` (the reference's SIGINT path), saves and returns {"interrupted": iteration}."""
    from . import checkpoint as ckpt
    say = log or (lambda s: None)
    m = (1 << p) - 1
    total = p - 1
    B = max(min(int(math.isqrt(p)), total), 1)
    level_auto = int(1000 * 600.0 / B) or 1
    level = check_level if check_level > 0 else max(level_auto, 1)
    got = ckpt.load_llsafe(checkpoint_path, eng, p) if checkpoint_path else None
    if got:
        ri, itersave, jsave = got[0], got[1], got[2]
        say(f"Resuming from a checkpoint at iteration {ri}.")
    else:
        ri, itersave, jsave = 0, 0, total - 1
        eng.set(LS_RES_A, 2); eng.set(LS_RES_B, 1)
        eng.set(LS_ACC_A, 1); eng.set(LS_ACC_B, 0)
        for d, s in ((LS_SAVE_R_A, LS_RES_A), (LS_SAVE_R_B, LS_RES_B), (LS_SAVE_F_A, LS_ACC_A), (LS_SAVE_F_B, LS_ACC_B)):
            eng.copy(d, s)
    eng.set(LS_BASE_A, 2); eng.set(LS_BASE_B, 1)
    checkpass, errors, restores, errordone = 0, 0, 0, False
    it, j = ri, total - ri - 1
    while it < total:
        if checkpoint_path and it > ri and ((stop_at is not None and it == stop_at) or (backup_every and it % backup_every == 0)):
            ckpt.save_llsafe(checkpoint_path, eng, p, it, itersave, jsave)
            if stop_at is not None and it == stop_at:
                say(f"Interrupted by user, state saved at iteration {it} j={j}")
                return {"exponent": p, "worktype": "LL", "interrupted": it, "errors_gerbicz": errors, "fft_length": eng.get_size()}
        if error_iter > 0 and it + 1 == error_iter and not errordone:
            errordone = True
            eng.sub(LS_RES_A, 2)
            say(f"Injected error at iteration {it + 1}")
        if it + 1 == total:
            eng.copy(LS_PREV_A, LS_RES_A); eng.copy(LS_PREV_B, LS_RES_B)
        _pair_square(eng, LS_RES_A, LS_RES_B)
        if (j != 0 and j % B == 0) or it == total - 1:
            checkpass += 1
            eng.copy(LS_CHK_A, LS_ACC_A); eng.copy(LS_CHK_B, LS_ACC_B)
            _pair_mul_by(eng, LS_ACC_A, LS_ACC_B, LS_RES_A, LS_RES_B)
            if checkpass >= level or it == total - 1:
                modB = B if total % B == 0 else total % B
                for _ in range(B - modB - 1 if B > modB else 0):
                    _pair_square(eng, LS_CHK_A, LS_CHK_B)
                if modB != B:
                    _pair_square(eng, LS_CHK_A, LS_CHK_B)
                _pair_mul_by(eng, LS_CHK_A, LS_CHK_B, LS_BASE_A, LS_BASE_B)
                for _ in range(modB):
                    _pair_square(eng, LS_CHK_A, LS_CHK_B)
                same = _same(eng, LS_CHK_A, LS_ACC_A, p) and _same(eng, LS_CHK_B, LS_ACC_B, p)
                checkpass = 0
                if not same:
                    say(f"[Gerbicz Li] Check FAILED! iter={it + 1}")
                    errors += 1; restores += 1
                    if restores > max_restores:
                        raise RuntimeError("Gerbicz-Li: too many restores")
                    for d, s in ((LS_RES_A, LS_SAVE_R_A), (LS_RES_B, LS_SAVE_R_B), (LS_ACC_A, LS_SAVE_F_A), (LS_ACC_B, LS_SAVE_F_B)):
                        eng.copy(d, s)
                    if itersave == 0:
                        it, j = 0, total - 1
                    else:
                        it, j = itersave + 1, jsave - 1
                    continue
                say(f"[Gerbicz Li] Check passed! iter={it + 1}")
                for d, s in ((LS_SAVE_R_A, LS_RES_A), (LS_SAVE_R_B, LS_RES_B), (LS_SAVE_F_A, LS_ACC_A), (LS_SAVE_F_B, LS_ACC_B)):
                    eng.copy(d, s)
                itersave, jsave = it, j
        it += 1; j -= 1
    a, b = _value(eng, LS_RES_A, p), _value(eng, LS_RES_B, p)
    # the classical LL residue s_{p-2} is twice the A component of the previous pair (RunLlSafeMarin.cpp:688-693)
    eng.add(LS_PREV_A, LS_PREV_A)
    ll = _value(eng, LS_PREV_A, p)
    return {"exponent": p, "worktype": "LL", "is_prime": a == m - 1 and b == 0, "res64": format_res64(ll),
            "res2048": format_res2048(ll), "errors_gerbicz": errors, "fft_length": eng.get_size()}
