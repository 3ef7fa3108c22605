"""Generates csrc/ebn_fastmath_tables.inc: the -ln(m) reduction table, the 2^(j/64) table, the
first-quadrant sin/cos cell table with its in-cell polynomials and the cell table of the inverse normal
CDF used by ebn_fastmath.cuh. Deterministic; rerun to regenerate.

    python tools/gen_fastmath_tables.py [log table bits, default 10] [sin/cos table bits, default 10]
"""
import os
import sys

import numpy as np

import mpmath

mpmath.mp.prec = 300

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "enhancedbayesiannetworks.jl_b200", "csrc", "ebn_fastmath_tables.inc")


def neglog_exact(x: float) -> float:
    """-ln(x) of an exactly representable double, correctly rounded (mpmath) or via longdouble."""
    if mpmath is not None:
        return float(-mpmath.log(mpmath.mpf(x)))
    return float(-np.log(np.longdouble(x)))


def log_table(bits):
    """2^bits cells over m in [1, 2) (top mantissa bits). Entry = (rc, ln(rc)) with rc ~ 1/centre
    rounded to 24 bits, so that r = fma(m, rc, -1) carries a single rounding and |r| <= 2^-(bits+1)."""
    rows = []
    n = 1 << bits
    for i in range(n):
        c = 1.0 + (i + 0.5) / n
        rc = float(np.float32(1.0 / c))
        rows.append((rc, -neglog_exact(rc)))  # ln(rc), correctly rounded
    return rows


def _cheb_fit(f, a, b, deg):
    """Polynomial interpolating f at the deg+1 Chebyshev nodes of [a, b] (near-minimax), mpmath."""
    mp = mpmath
    n = deg + 1
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for k in range(n):
            A[i, k] = x ** k
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [float(c[i]) for i in range(n)]


def _cheb_fit_centred(f, a, b, deg):
    """As _cheb_fit, in powers of (x - (a + b)/2)."""
    mp = mpmath
    n = deg + 1
    c = (a + b) / 2
    xs = [c + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for k in range(n):
            A[i, k] = (x - c) ** k
        y[i] = f(x)
    co = mp.lu_solve(A, y)
    return [float(co[i]) for i in range(n)]


def sincos_table(bits):
    """2^bits cells per quadrant over the FULL circle (4 * 2^bits entries): sqrt(2)*(sin, cos) of
    phi_j = (pi/2)(j + 1/2)/2^bits, correctly rounded, and the two short polynomials for the offset inside a cell,
    delta = (pi/2) t with |t| <= 2^-(bits+1):  sin(delta) = t*Sd(t^2), cos(delta) - 1 = t^2*Cd(t^2).
    7 bits: 3 + 2 coefficients; 10 bits: 2 + 2 (truncation below 1e-17 either way)."""
    mp = mpmath
    hp = mp.pi / 2
    n = 1 << bits
    rows = []
    for j in range(4 * n):
        phi = hp * (mp.mpf(j) + mp.mpf(1) / 2) / n
        rows.append((float(mp.sqrt(2) * mp.sin(phi)), float(mp.sqrt(2) * mp.cos(phi))))
    wmax = mp.mpf(2) ** (-2 * (bits + 1))
    Sd = lambda w: hp if w == 0 else mp.sin(hp * mp.sqrt(w)) / mp.sqrt(w)  # noqa: E731
    Cd = lambda w: -hp * hp / 2 if w == 0 else (mp.cos(hp * mp.sqrt(w)) - 1) / w  # noqa: E731
    return rows, _hi_word_top(_cheb_fit(Sd, mp.mpf(0), wmax, 2 if bits < 10 else 1)), _hi_word_top(_cheb_fit(Cd, mp.mpf(0), wmax, 1))


def _hi_word_top(coefs):
    """Rounds the highest-order coefficient to a double whose low 32 bits are zero (21 significant
    bits): such a value is an instruction immediate of DFMA, so the first Horner step
    fma(w, top, next) reads one immediate and one constant-bank operand instead of two registers that
    have to be re-loaded inside the loop. The term it multiplies is below 1e-12 of the result, so the
    2^-21 relative change of the coefficient moves the result by less than 1e-17."""
    import struct
    top = coefs[-1]
    bits = struct.unpack("<Q", struct.pack("<d", top))[0]
    bits = (bits + 0x80000000) & 0xFFFFFFFF00000000
    return coefs[:-1] + [struct.unpack("<d", struct.pack("<Q", bits))[0]]


ICDF_OCTAVES = 16      # q = min(p, 1-p) in [2^-(ICDF_OCTAVES+1), 1/2]; smaller q (3e-5 of the unit interval)
                       # takes the library routine
ICDF_SUB_BITS = 4      # 16 cells per octave of q
ICDF_DEG = 7


def icdf_table():
    """Phi^-1 for q = min(p, 1 - p) = 2^-e m, m in [1, 2): one degree-7 polynomial in (m - centre) per cell
    (16 cells per octave), near-minimax (Chebyshev nodes), max error 3e-16 relative to max(1, |z|).
    Row = 8 coefficients c0..c7 of z(q) <= 0. Rows run with DEcreasing q, so that the row number is a
    constant minus the top 16 bits of q's high word: row 0 is the cell that holds q = 1/2 exactly
    (e = 1, j = 0), and cell (e, j) of the octaves e = 2..ICDF_OCTAVES+1 is row 1 + (e - 2) * 16 + (15 - j)."""
    mp = mpmath
    sub = 1 << ICDF_SUB_BITS
    rows = []
    cells = [(1, 0)] + [(e, j) for e in range(2, ICDF_OCTAVES + 2) for j in reversed(range(sub))]
    for e, j in cells:
        a = 1 + mp.mpf(j) / sub
        b = 1 + mp.mpf(j + 1) / sub

        def f(m, e=e):
            q = m * mp.mpf(2) ** (-e)
            return -mp.sqrt(2) * mp.erfinv(1 - 2 * q) if q < mp.mpf(1) / 2 else mp.sqrt(2) * mp.erfinv(2 * q - 1)
        rows.append(_cheb_fit_centred(f, a, b, ICDF_DEG))
    return rows


PHI_CELLS_PER_UNIT = 16   # cells of width 1/16 centred on k/16
PHI_AMAX = 9.5            # Phi(-9.5) ~ 1e-21, below the 2^-60 the callers clamp to
PHI_DEG = 7


def phi_table():
    """Phi(-a), a = |z| in [0, PHI_AMAX]: cell k = round(16 a) is centred on c = k/16 and carries a degree-7
    polynomial P_k in dm = a - c of  Phi(-(c + dm)) * exp(c dm)  — the factor exp(-c dm), |c dm| <= 0.3, is
    what is left for the exponential routine, so neither a^2 nor a large argument enters it.
    Max error 4.4e-16 relative to Phi(-a) over the whole range (the tail keeps its relative accuracy)."""
    mp = mpmath
    W = PHI_CELLS_PER_UNIT
    rows = []
    n = PHI_DEG + 1
    for k in range(int(PHI_AMAX * W) + 2):
        c = mp.mpf(k) / W
        a = max(c - mp.mpf(1) / (2 * W), mp.mpf(0))
        b = c + mp.mpf(1) / (2 * W)
        xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * i + 1) / (2 * n)) for i in range(n)]
        A = mp.matrix(n, n)
        y = mp.matrix(n, 1)
        for i, x in enumerate(xs):
            for j in range(n):
                A[i, j] = (x - c) ** j
            y[i] = mp.ncdf(-x) * mp.exp(c * (x - c))
        co = mp.lu_solve(A, y)
        rows.append([float(co[i]) for i in range(n)])
    return rows


def main():
    log_bits = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    sc_bits = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    rows = log_table(log_bits)
    with open(OUT, "w") as f:
        f.write("// GENERATED by tools/gen_fastmath_tables.py — do not edit.\n")
        f.write(f"#define EBN_LOG_TAB_BITS {log_bits}\n#define EBN_SC_TAB_BITS {sc_bits}\n")
        f.write("// kLogTab[i] = {rc_i, ln(rc_i)}: -ln(m) = ln(rc_i) - log1p(m*rc_i - 1), m in [1,2)\n")
        f.write("static __device__ const double2 kLogTab[1 << EBN_LOG_TAB_BITS] = {\n")
        for rc, lrc in rows:
            f.write(f"  {{{rc.hex()}, {lrc.hex()}}},\n")
        f.write("};\n")
        f.write("// kExpTab[j] = 2^(j/64), correctly rounded\n")
        f.write("static __device__ const double kExpTab[64] = {\n")
        for j in range(64):
            v = float(mpmath.power(2, mpmath.mpf(j) / 64)) if mpmath is not None else float(np.exp2(np.longdouble(j) / 64))
            f.write(f"  {v.hex()},\n")
        f.write("};\n")
        sc_rows, sd, cd = sincos_table(sc_bits)
        f.write("// kSinCosTab[j] = sqrt(2)*{sin, cos}((pi/2)(j + 1/2)/2^bits), j < 4*2^bits (full circle); inside a cell, delta = (pi/2) t:\n")
        f.write("// sin(delta) = t*(d0 + d1 t^2 [+ d2 t^4]), cos(delta) - 1 = t^2*(e0 + e1 t^2)\n")
        f.write("static __device__ const double2 kSinCosTab[4 << EBN_SC_TAB_BITS] = {\n")
        for a, b in sc_rows:
            f.write(f"  {{{a.hex()}, {b.hex()}}},\n")
        f.write("};\n")
        f.write("#define EBN_SIND_COEFS " + ", ".join(float(c).hex() for c in sd) + "\n")
        f.write("#define EBN_COSD_COEFS " + ", ".join(float(c).hex() for c in cd) + "\n")
        f.write("// the highest-order coefficients again, as literals (immediates: their low words are zero)\n")
        f.write(f"#define EBN_SIND_TOP {float(sd[-1]).hex()}\n#define EBN_COSD_TOP {float(cd[-1]).hex()}\n")
        f.write("// kIcdfTab[row][k]: z = Phi^-1(q) = sum_k c_k (m - centre)^k for q = 2^-e m in cell j of octave e;\n")
        f.write("// row 0: q = 1/2, then row = 1 + (e-2)*16 + 15 - j for e = 2..EBN_ICDF_OCTAVES+1\n")
        f.write(f"#define EBN_ICDF_OCTAVES {ICDF_OCTAVES}\n#define EBN_ICDF_SUB_BITS {ICDF_SUB_BITS}\n")
        f.write(f"#define EBN_ICDF_ROWS {1 + (ICDF_OCTAVES << ICDF_SUB_BITS)}\n")
        f.write("static __device__ const __align__(16) double kIcdfTab[EBN_ICDF_ROWS][8] = {\n")
        for co in icdf_table():
            f.write("  {" + ", ".join(float(c).hex() for c in co) + "},\n")
        f.write("};\n")
        f.write("// kPhiTab[k][j]: Phi(-a) = exp(-c dm) * sum_j c_j dm^j for a = |z| in cell k = round(16 a), c = k/16, dm = a - c\n")
        prows = phi_table()
        f.write(f"#define EBN_PHI_ROWS {len(prows)}\n#define EBN_PHI_AMAX {PHI_AMAX}\n")
        f.write("static __device__ const __align__(16) double kPhiTab[EBN_PHI_ROWS][8] = {\n")
        for co in prows:
            f.write("  {" + ", ".join(float(c).hex() for c in co) + "},\n")
        f.write("};\n")
    print("wrote", OUT, "mpmath" if mpmath is not None else "longdouble")


if __name__ == "__main__":
    main()
