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
