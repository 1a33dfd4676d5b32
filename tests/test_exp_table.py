"""The constants of exp_tab16 (bayesian_ensembling_b200/csrc/be_kernels.cuh), read from the source and checked on
the CPU: the table entries are the correctly rounded 2^(j/16), the split of ln2/16 leaves half an ulp of its low part, and the
kernel's instruction sequence, emulated with exactly rounded FMAs, stays within one ulp of exp over the whole
fast-path range.  (The device run of the same sequence is tests/test_gpu_parity.py::test_weights_exponential_whole_range.)"""
import math
import os
import random
import re
import struct
from decimal import Decimal, getcontext
from fractions import Fraction

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bayesian_ensembling_b200", "csrc",
                   "be_kernels.cuh")
HEX = r"-?0x1\.[0-9a-f]+p[+-]\d+"
MAGIC = 6755399441055744.0


def _constants():
    src = open(SRC).read()
    tab = re.search(r"EXP2_16TH\[16\]\s*=\s*\{(.*?)\};", src, re.S).group(1)
    tab = [float.fromhex(h) for h in re.findall(HEX, tab)]
    expc = re.search(r"ExpTabConsts EXPC\s*=\s*\{(.*?)\};", src, re.S).group(1)
    inv, lo, c7, c6, c5, c4, c3 = [float.fromhex(h) for h in re.findall(HEX, expc)]
    body = src[src.index("double exp_tab16_core"):src.index("bool exp_tab16_ok")]
    hi = -float.fromhex(re.search(r"fma\(kd,\s*(" + HEX + r"),\s*x\)", body).group(1))
    return tab, inv, hi, lo, (c7, c6, c5, c4, c3)


def _fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))  # one rounding, as the device's DFMA


def _exp_tab16(x, consts):
    tab, inv, hi, lo, (c7, c6, c5, c4, c3) = consts
    v = _fma(x, inv, MAGIC)
    low = struct.unpack("<q", struct.pack("<d", v))[0] & 0xFFFFFFFF
    k = low - (1 << 32) if low >= (1 << 31) else low
    kd = v - MAGIC
    r = _fma(kd, -hi, x)
    r = _fma(kd, -lo, r)
    q = _fma(c7, r, c6)
    for c in (c5, c4, c3, 0.5):
        q = _fma(q, r, c)
    p = _fma(q * r, r, r)
    t = tab[k & 15]
    return math.ldexp(_fma(t, p, t), k >> 4), r


def test_table_and_split_constants():
    getcontext().prec = 60
    tab, inv, hi, lo, coef = _constants()
    ln2 = Decimal(2).ln()
    assert len(tab) == 16
    for j, t in enumerate(tab):
        assert t == float((ln2 * j / 16).exp())  # Decimal -> float rounds correctly
    assert struct.unpack("<Q", struct.pack("<d", hi))[0] & 0xFFFFFFFF == 0  # an immediate operand: zero low word
    assert abs(Decimal(hi) + Decimal(lo) - ln2 / 16) <= Decimal(math.ulp(lo)) / 2  # lo is the rounded remainder
    assert inv == float(16 / ln2)
    assert coef == tuple(1.0 / math.factorial(n) for n in (7, 6, 5, 4, 3))


def test_emulated_sequence_within_one_ulp():
    getcontext().prec = 50
    consts = _constants()
    rng = random.Random(7)
    xs = [rng.uniform(-700.0, 700.0) for _ in range(2500)]
    xs += [rng.uniform(-1, 1) * 10 ** rng.uniform(-9, 1) for _ in range(1000)]
    xs += [0.0, 699.999999, -699.999999, math.log(2) / 32, -math.log(2) / 32, 1e-300]
    worst, rmax = 0.0, 0.0
    for x in xs:
        got, r = _exp_tab16(x, consts)
        want = Decimal(x).exp()
        worst = max(worst, float(abs((Decimal(got) - want) / want)))
        rmax = max(rmax, abs(r))
    assert rmax <= math.log(2) / 32 * (1 + 1e-9)
    assert worst <= 2.0**-52, worst  # one ulp


# ------------------------------------------------------------------------------------ log_tab64 (Normal-branch weights)
def _log_constants():
    src = open(SRC).read()
    inv = [float.fromhex(h) for h in re.findall(HEX, re.search(r"LOG_INV_C\[64\]\s*=\s*\{(.*?)\};", src, re.S).group(1))]
    nlg = [float.fromhex(h) for h in re.findall(HEX, re.search(r"LOG_NEG_LOG_INV_C\[64\]\s*=\s*\{(.*?)\};", src, re.S).group(1))]
    body = src[src.index("double log_tab64"):src.index("__global__ void k_normal_logprob")]
    coef = [float.fromhex(h) for h in re.findall(HEX, body)]  # 1/7, -1/6, 1/5, 1/3 (the others are literals)
    return inv, nlg, coef


def _log_tab64(x, inv, nlg):
    bits = struct.unpack("<Q", struct.pack("<d", x))[0]
    hi = bits >> 32
    e = (hi >> 20) - 1023
    j = (hi >> 14) & 63
    m = struct.unpack("<d", struct.pack("<Q", (bits & 0x000FFFFFFFFFFFFF) | (0x3FF << 52)))[0]
    r = _fma(m, inv[j], -1.0)
    q = _fma(r, 1.0 / 7.0, -1.0 / 6.0)
    for c in (0.2, -0.25, 1.0 / 3.0, -0.5):
        q = _fma(q, r, c)
    l1p = _fma(r * r, q, r)
    lg = _fma(float(e), 1.90821492927058770002e-10, l1p) + nlg[j]
    return _fma(float(e), 6.93147180369123816490e-01, lg), r


def test_log_table_constants():
    getcontext().prec = 60
    inv, nlg, coef = _log_constants()
    assert len(inv) == 64 and len(nlg) == 64
    for j in range(64):
        c = Decimal(1) + (Decimal(j) + Decimal(1) / 2) / 64
        assert inv[j] == float(1 / c)
        assert nlg[j] == float(-(Decimal(inv[j]).ln()))  # minus the log of the ROUNDED reciprocal
    assert coef[:4] == [1.0 / 7.0, -1.0 / 6.0, 0.2, 1.0 / 3.0]
    hi = 6.93147180369123816490e-01
    assert struct.unpack("<Q", struct.pack("<d", hi))[0] & 0xFFFFF == 0  # e * hi is exact for |e| < 2^11
    assert abs(Decimal(hi) + Decimal(1.90821492927058770002e-10) - Decimal(2).ln()) < Decimal(2) ** -84


def test_emulated_log_sequence():
    getcontext().prec = 50
    inv, nlg, _ = _log_constants()
    rng = random.Random(11)
    xs = [math.ldexp(rng.uniform(1.0, 2.0), rng.randint(-500, 500)) for _ in range(1500)]
    xs += [rng.uniform(0.5, 2.0) for _ in range(1500)]  # around 1, where log x is small
    xs += [1.0, 1.0 + 2.0**-52, 2.0 - 2.0**-52, 1.0 + 1.0 / 64, 1.0 + 1.0 / 128, 0.3, 0.01, 1e-3, 37.5]
    worst, rmax = 0.0, 0.0
    for x in xs:
        got, r = _log_tab64(x, inv, nlg)
        want = Decimal(x).ln()
        worst = max(worst, float(abs(Decimal(got) - want)) / (1.0 + abs(float(want))))
        rmax = max(rmax, abs(r))
    assert rmax <= 2.0**-7 * (1 + 1e-9)
    assert worst <= 2.5e-16, worst  # absolute, relative to 1 + |log x|: what an exponent needs
