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

from .engine import digits_to_int

# register map of RunPrpOrLlMarin.cpp:213
R0, R1, R2, R3, R4, R5, RBASE, RTMP = range(8)


def _digits(eng, r):
    return eng.get(r) if hasattr(eng, "get") else eng.get_digits(r)


def _value(eng, r, p):
    v = digits_to_int(_digits(eng, r))
    return 0 if v == (1 << p) - 1 else v


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


def run_prp(eng, p, gerbicz=True, check_level=0, error_iter=0, log=None, max_restores=4, stop_after_checks=0):
    """PRP test of 2^p - 1 (type 1, base 3): p squarings from 3, prime iff the residue is 9.

    Gerbicz-Li checking as the reference: every B = floor(sqrt(p)) iterations the accumulator
    d <- d * x, and every `check_level` blocks (and at the end) d is checked against
    (d_prev)^(2^B) * 3 recomputed from the previous accumulator; on mismatch the last good state is
    restored (RunPrpOrLlMarin.cpp:338-412).  `error_iter` injects a fault (x -= 2) after that
    iteration, like -erroriter (RunPrpOrLlMarin.cpp:326-336).  `stop_after_checks` ends the run after that
    many passed checks (the reference's size-ladder smoke test stops at the first "Check passed",
    unit_test_all.sh:54-79); the result then carries `stopped_at` instead of a verdict.
    """
    say = log or (lambda s: None)
    total = p
    B = max(int(math.isqrt(p)), 1)
    level_auto = int(1000 * 600.0 / B) or max((total // B) // max(int(math.isqrt(B)), 1), 1)
    level = check_level if check_level > 0 else max(level_auto, 1)
    eng.set(R1, 1)
    eng.set(R0, 3)
    eng.copy(R4, R0); eng.copy(R5, R1)
    eng.set(RBASE, 3)
    eng.set_multiplicand(RTMP, RBASE)
    itersave, jsave = 0, total - 1
    checkpass, errors, restores, errordone, passed = 0, 0, 0, False, 0
    it, j = 0, total - 1
    while it < total:
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
                if _value(eng, R3, p) != _value(eng, R1, p):
                    say(f"[Gerbicz Li] Check FAILED! iter={it + 1}")
                    say(f"[Gerbicz Li] Restore iter={itersave} (j={jsave})")
                    errors += 1
                    restores += 1
                    if restores > max_restores:
                        raise RuntimeError("Gerbicz-Li: too many restores")
                    eng.copy(R0, R4); eng.copy(R1, R5)
                    it, j, checkpass = itersave, jsave, 0
                    if it == 0:
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
    return {"exponent": p, "worktype": "PRP-3", "is_prp": x == 9 % m, "res64": format_res64(r), "res2048": format_res2048(r),
            "residue_type": 1, "errors_gerbicz": errors, "fft_length": eng.get_size(), "iterations": total}


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


def run_pm1_stage1(eng, p, b1, reg=0, gcd=True):
    """P-1 stage 1, "fast3" (RunPM1.cpp:7127): X <- X^2 * (3 if bit else 1) over the bits of E, MSB first;
    the factor is gcd(X - 1, 2^p - 1)."""
    E = pm1_exponent(p, b1)
    eng.set(reg, 1)
    for bit in bin(E)[2:]:
        eng.square_mul(reg, 3 if bit == "1" else 1)
    x = _value(eng, reg, p)
    m = (1 << p) - 1
    g = math.gcd((x - 1) % m, m) if gcd else 1          # (Python's gcd is quadratic: skip it for 10^8-bit smoke runs)
    return {"exponent": p, "worktype": "PM1", "b1": b1, "factor": g if 1 < g < m else None, "bits": E.bit_length(),
            "fft_length": eng.get_size(), "x": x}


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


def run_pm1_stage2_vtrace(eng, p, b1, b2, H, D=None, k_range=None, sieve=None):
    """P-1 stage 2 on traces V_n = H^n + H^-n: for every prime B1 < q <= B2 write q = k D +- j and
    accumulate (V_{kD} - V_j); both members of a pair share one term.  Returns the accumulator residue
    (an int) and statistics; the factor is gcd(acc, 2^p - 1) (use pm1_stage2_factor).  `k_range`
    restricts the giant steps to [k_lo, k_hi) so that ranks can split the prime range:
    each rank rebuilds the (deterministic) baby table, jumps to its first k by a Lucas ladder and
    accumulates its own product."""
    m = (1 << p) - 1
    D = D or vtrace_choose_D(b1)
    babies = [j for j in range(1, D // 2 + 1) if math.gcd(j, D) == 1]
    slot = {j: BABY_BASE + i for i, j in enumerate(babies)}
    if sieve is None:
        sieve = bytearray([1]) * (b2 + D + 2)
        sieve[:2] = b"\0\0"
        for i in range(2, int((b2 + D + 1) ** 0.5) + 1):
            if sieve[i]:
                sieve[i * i::i] = bytearray(len(sieve[i * i::i]))
    is_s2_prime = lambda q: b1 < q <= b2 and sieve[q]
    V1 = (H + pow(H, -1, m)) % m
    _set_int(eng, RV1, V1, p)
    eng.set_multiplicand(RMUL_V1, RV1)
    # baby steps: V_0 = 2, V_1, V_{n+1} = V_1 V_n - V_{n-1}; store -V_j so the term is one add (RunPM1.cpp:3551-3587)
    eng.set(RBPREV, 2); eng.copy(RBCUR, RV1)
    for n in range(1, D + 1):
        if n in slot:
            eng.set(slot[n], 0); eng.sub_reg(slot[n], RBCUR)
        if n == D:
            eng.copy(RVD, RBCUR)
            break
        eng.copy(RBNEXT, RBCUR); eng.mul(RBNEXT, RMUL_V1); eng.sub_reg(RBNEXT, RBPREV)
        eng.copy(RBPREV, RBCUR); eng.copy(RBCUR, RBNEXT)
    eng.set_multiplicand(RMUL_VD, RVD)
    k_first, k_last = (b1 + 1 + D // 2) // D, (b2 + D // 2) // D
    k_lo, k_hi = k_range if k_range else (k_first, k_last + 1)
    k_lo, k_hi = max(k_lo, k_first), min(k_hi, k_last + 1)
    eng.set(RACC, 1)
    terms = primes = 0
    if k_lo < k_hi:
        if k_lo == 0:
            eng.set(RGCUR, 2)                                  # V_0; V_{-D} = V_D
            eng.copy(RGPREV, RVD)
        else:
            _lucas_jump(eng, k_lo)
        for k in range(k_lo, k_hi):
            for j in babies:
                lo, hi = k * D - j, k * D + j
                a, b = lo > 0 and is_s2_prime(lo), is_s2_prime(hi)
                if not (a or b):
                    continue
                primes += int(a) + int(b)
                eng.copy(RTMP2, RGCUR); eng.add(RTMP2, slot[j])          # V_kD + (-V_j)
                eng.set_multiplicand(RMUL_TMP, RTMP2); eng.mul(RACC, RMUL_TMP)
                terms += 1
            # next giant step: V_{(k+1)D} = V_D V_{kD} - V_{(k-1)D}  (RunPM1.cpp:3788-3797)
            eng.copy(RGNEXT, RGCUR); eng.mul(RGNEXT, RMUL_VD); eng.sub_reg(RGNEXT, RGPREV)
            eng.copy(RGPREV, RGCUR); eng.copy(RGCUR, RGNEXT)
    return {"acc": _value(eng, RACC, p), "terms": terms, "primes": primes, "D": D, "k_range": (k_lo, k_hi)}


def pm1_stage2_factor(p, accs):
    """gcd of the product of the per-rank accumulators with 2^p - 1 (RunPM1.cpp:4261-4262)."""
    m = (1 << p) - 1
    acc = 1
    for a in accs:
        acc = acc * a % m
    g = math.gcd(acc, m)
    return g if 1 < g < m else None


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


def run_ll_safe(eng, p, check_level=0, error_iter=0, log=None, max_restores=4):
    """`prmers p -ll`: (2 + sqrt 3)^(2^(p-1)) in Z[sqrt 3] mod 2^p - 1 is (-1, 0) iff 2^p - 1 is prime; the
    running product of the block-end pairs is checked Gerbicz-Li style and a mismatch restores the last
    verified pair (RunLlSafeMarin.cpp:539-694).  Needs 18 registers."""
    say = log or (lambda s: None)
    m = (1 << p) - 1
    total = p - 1
    B = max(min(int(math.isqrt(p)), total), 1)
    level_auto = int(1000 * 600.0 / B) or 1
    level = check_level if check_level > 0 else max(level_auto, 1)
    eng.set(LS_RES_A, 2); eng.set(LS_RES_B, 1)
    eng.set(LS_ACC_A, 1); eng.set(LS_ACC_B, 0)
    for d, s in ((LS_SAVE_R_A, LS_RES_A), (LS_SAVE_R_B, LS_RES_B), (LS_SAVE_F_A, LS_ACC_A), (LS_SAVE_F_B, LS_ACC_B)):
        eng.copy(d, s)
    eng.set(LS_BASE_A, 2); eng.set(LS_BASE_B, 1)
    itersave, jsave, checkpass, errors, restores, errordone = 0, total - 1, 0, 0, 0, False
    it, j = 0, total - 1
    while it < total:
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
                same = (_value(eng, LS_CHK_A, p) == _value(eng, LS_ACC_A, p) and _value(eng, LS_CHK_B, p) == _value(eng, LS_ACC_B, p))
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
