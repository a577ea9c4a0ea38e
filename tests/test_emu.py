"""CPU lock-step emulation of the CUDA kernel's phase functions vs the oracle -- no GPU.

tests/emu/emu.cpp steps the product's own __host__ __device__ code (pbs_team.cuh) for every thread
of one CTA.  This pins the NTT indexing, twiddle tables, lazy-reduction bounds (asserted), CRT and
digit quirks against oracle/tfhe_oracle.c bit for bit, at the full polynomial sizes.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import orc

HERE = os.path.dirname(os.path.abspath(__file__))
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

CFGS = {
    0: dict(glwe_dimension=2, glwe_poly_degree=9, pbs_log_base=4, pbs_levels=6),
    1: dict(glwe_dimension=1, glwe_poly_degree=10, pbs_log_base=8, pbs_levels=3),
    2: dict(glwe_dimension=1, glwe_poly_degree=11, pbs_log_base=8, pbs_levels=3),
}


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "emu", "emu.cpp")
    so = os.path.join(HERE, "emu", "libemu.so")
    deps = [src] + [os.path.join(HERE, "..", "tfhe-research_b200", "csrc", f) for f in ("tfhe_core.cuh", "pbs_team.cuh", "host_tables.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-fPIC", "-shared", "-o", so, src], check=True)
    L = C.CDLL(so)
    L.emu_transform_ggsw.argtypes = [C.c_int, u32p, u32p]
    L.emu_step.argtypes = [C.c_int, C.c_int, u32p, u32p, C.c_uint32]
    L.emu_mod_switch.argtypes = [C.c_uint32, C.c_int]
    L.emu_mod_switch.restype = C.c_uint32
    L.emu_decompose.argtypes = [C.c_uint32, C.c_int, C.c_int, i32p]
    L.emu_crt.argtypes = [C.c_uint32, C.c_uint32]
    L.emu_crt.restype = C.c_uint32
    return L


def test_scalar_helpers_match_oracle(emu):
    rng = np.random.default_rng(3)
    vals = [0, 1, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 0x0000F800, 0x0FF80000, 0xF8F8F8F8, 0x7FFFFF80, 0x80, 0xABCDEF12]
    vals += rng.integers(0, 1 << 32, 20000, dtype=np.uint64).tolist()
    O = orc.lib()
    for logn in (9, 10, 11):
        v = np.array(vals, dtype=np.uint32)
        out = orc.z(len(v))
        O.orc_switch_modulus(v, len(v), 32, logn + 1, out)
        assert [emu.emu_mod_switch(int(x), logn) for x in v[:3000]] == out[:3000].tolist()
    for lb, lv in ((4, 6), (4, 5), (8, 3), (2, 8), (8, 4), (4, 8)):
        a, b = np.zeros(lv, dtype=np.int32), orc.z(lv)
        for x in vals[:6000]:
            emu.emu_decompose(x, lb, lv, a)
            O.orc_decompose(x, lb, lv, b)
            assert a.tolist() == b.astype(np.int32).tolist(), (hex(x), lb, lv)


def test_crt_centred_lift(emu):
    q0, q1 = 165093377, 165142529
    rng = np.random.default_rng(4)
    vals = [0, 1, -1, (1 << 52), -(1 << 52), q0 * q1 // 2, -(q0 * q1 // 2)] + [int(x) - (1 << 52) for x in rng.integers(0, 1 << 53, 2000)]
    for v in vals:
        # lazily reduced inputs in [0, 2q)
        assert emu.emu_crt(v % q0 + q0 * (v & 1), v % q1 + q1 * ((v >> 1) & 1)) == v % (1 << 32)


@pytest.mark.parametrize("cfg", [0, 1, 2])
def test_emulated_cmux_step_bit_exact(emu, cfg):
    p = orc.params(**CFGS[cfg])
    N, k, l = p.N, p.k, p.pbs_levels
    rng = np.random.default_rng(100 + cfg)
    r32 = lambda *s: rng.integers(0, 1 << 32, s, dtype=np.uint64).astype(np.uint32)
    ggsw = r32((k + 1) * l, k + 1, N)
    ntt = orc.z(2 * (k + 1) * l * (k + 1) * N)
    assert emu.emu_transform_ggsw(cfg, ggsw.reshape(-1), ntt) == 0
    assert int(ntt.max()) < 165142529
    O = orc.lib()
    # external product (ggsw.rs:132) on uniformly random GLWE (exercises every digit incl. +B, H3)
    glwe = r32(k + 1, N)
    # plant the quirk/edge values explicitly
    glwe[0, :8] = [0xFFFFFFFF, 0x7FFFFF80, 0x0000F800, 0xF8F8F8F8, 0, 0x80000000, 0x0FF80000, 0x00FFFFFF]
    exp = orc.z((k + 1) * N)
    O.orc_external_product(C.byref(p), ggsw.reshape(-1), glwe.reshape(-1), exp)
    got = glwe.copy().reshape(-1)
    assert emu.emu_step(cfg, 1, ntt, got, 0) == 0
    assert got.tolist() == exp.tolist()
    # blind-rotate step: acc <- cmux(ggsw, acc, acc*X^a) for several rotations (bootstrapping.rs:94-104)
    for a in (1, N - 1, N, N + 5, 2 * N - 1, 777 % (2 * N)):
        acc = r32(k + 1, N)
        c1 = orc.z((k + 1) * N)
        O.orc_glwe_mul_monomial(C.byref(p), acc.reshape(-1), a, c1)
        exp = orc.z((k + 1) * N)
        O.orc_cmux(C.byref(p), ggsw.reshape(-1), acc.reshape(-1), c1, exp)
        got = acc.copy().reshape(-1)
        assert emu.emu_step(cfg, 0, ntt, got, a) == 0
        assert got.tolist() == exp.tolist(), a
    # a == 0: diff is identically zero, accumulator unchanged (justifies the skip in the kernel)
    acc = r32(k + 1, N)
    got = acc.copy().reshape(-1)
    emu.emu_step(cfg, 0, ntt, got, 0)
    assert got.tolist() == acc.reshape(-1).tolist()


# ------------------------------------------------------------------------------------------------------------------
# FP64-FFT path (fft_team.cuh): same checks, through tests/emu/emu_fft.cpp
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def emu_fft():
    src = os.path.join(HERE, "emu", "emu_fft.cpp")
    so = os.path.join(HERE, "emu", "libemu_fft.so")
    deps = [src] + [os.path.join(HERE, "..", "tfhe-research_b200", "csrc", f) for f in ("tfhe_core.cuh", "fft_team.cuh", "host_tables_fft.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src], check=True)
    L = C.CDLL(so)
    L.emu_fft_transform_ggsw.argtypes = [C.c_int, u32p, f64p]
    L.emu_fft_step.argtypes = [C.c_int, C.c_int, f64p, u32p, C.c_uint32, C.POINTER(C.c_double)]
    return L


@pytest.mark.parametrize("cfg", [0, 1, 2])
def test_emulated_fft_cmux_step_bit_exact(emu_fft, cfg):
    p = orc.params(**CFGS[cfg])
    N, k, l = p.N, p.k, p.pbs_levels
    rng = np.random.default_rng(200 + cfg)
    r32 = lambda *s: rng.integers(0, 1 << 32, s, dtype=np.uint64).astype(np.uint32)
    ggsw = r32((k + 1) * l, k + 1, N)
    # extreme key words: the centred limbs reach -2^15 and +2^15
    ggsw[0, 0, :6] = [0x7FFFFFFF, 0x80000000, 0x00008000, 0xFFFF8000, 0x7FFF7FFF, 0x80008000]
    key = np.zeros((k + 1) * l * 2 * (k + 1) * (N // 2) * 2, dtype=np.float64)
    assert emu_fft.emu_fft_transform_ggsw(cfg, ggsw.reshape(-1), key) == 0
    O = orc.lib()
    worst = 0.0
    mf = C.c_double(0.0)
    glwe = r32(k + 1, N)
    glwe[0, :8] = [0xFFFFFFFF, 0x7FFFFF80, 0x0000F800, 0xF8F8F8F8, 0, 0x80000000, 0x0FF80000, 0x00FFFFFF]
    exp = orc.z((k + 1) * N)
    O.orc_external_product(C.byref(p), ggsw.reshape(-1), glwe.reshape(-1), exp)
    got = glwe.copy().reshape(-1)
    assert emu_fft.emu_fft_step(cfg, 1, key, got, 0, C.byref(mf)) == 0
    assert got.tolist() == exp.tolist()
    worst = max(worst, mf.value)
    for a in (1, N - 1, N, N + 5, 2 * N - 1, 777 % (2 * N)):
        acc = r32(k + 1, N)
        c1 = orc.z((k + 1) * N)
        O.orc_glwe_mul_monomial(C.byref(p), acc.reshape(-1), a, c1)
        exp = orc.z((k + 1) * N)
        O.orc_cmux(C.byref(p), ggsw.reshape(-1), acc.reshape(-1), c1, exp)
        got = acc.copy().reshape(-1)
        assert emu_fft.emu_fft_step(cfg, 0, key, got, a, C.byref(mf)) == 0
        assert got.tolist() == exp.tolist(), a
        worst = max(worst, mf.value)
    acc = r32(k + 1, N)
    got = acc.copy().reshape(-1)
    emu_fft.emu_fft_step(cfg, 0, key, got, 0, C.byref(mf))
    assert got.tolist() == acc.reshape(-1).tolist()
    # distance of every pre-rounding value to the nearest integer: the a-priori bound is 2^-9, observed ~2^-16
    assert worst < 2.0 ** -10, worst


def test_emulated_fft_worst_case_inputs(emu_fft):
    """Adversarial magnitudes: every digit at its extreme (+B via the H3 quirk, -B/2) and every key limb at +-2^15."""
    cfg = 1
    p = orc.params(**CFGS[cfg])
    N, k, l = p.N, p.k, p.pbs_levels
    rng = np.random.default_rng(7)
    O = orc.lib()
    mf = C.c_double(0.0)
    for trial in range(3):
        sign = rng.integers(0, 2, ((k + 1) * l, k + 1, N)).astype(np.uint32)
        ggsw = np.where(sign == 1, np.uint32(0x7FFF7FFF), np.uint32(0x80008000)).astype(np.uint32)
        # 0xFF..F8 windows + carry give +B digits; 0x80 windows give -B/2
        glwe = np.where(rng.integers(0, 2, (k + 1, N)) == 1, np.uint32(0xFFFFFF80), np.uint32(0x80808080)).astype(np.uint32)
        key = np.zeros((k + 1) * l * 2 * (k + 1) * (N // 2) * 2, dtype=np.float64)
        assert emu_fft.emu_fft_transform_ggsw(cfg, ggsw.reshape(-1), key) == 0
        exp = orc.z((k + 1) * N)
        O.orc_external_product(C.byref(p), ggsw.reshape(-1), glwe.reshape(-1), exp)
        got = glwe.copy().reshape(-1)
        assert emu_fft.emu_fft_step(cfg, 1, key, got, 0, C.byref(mf)) == 0
        assert got.tolist() == exp.tolist()
        assert mf.value < 2.0 ** -10, mf.value


@pytest.mark.parametrize("cfg", [0, 1])
def test_emulated_fft_bmmp_step_bit_exact(emu_fft, cfg):
    """BMMP step: acc += ExtProd((X^(a+a')-1) G0 + (X^a-1) G1 + (X^a'-1) G2, acc), the monomial factors
    applied in the transform domain, vs the oracle's exact u32 bundle (oracle/tfhe_oracle.c orc_blind_rotate_bmmp)."""
    emu_fft.emu_fft_step_bmmp.argtypes = [C.c_int, u32p, u32p, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
    p = orc.params(**CFGS[cfg])
    N, k, l = p.N, p.k, p.pbs_levels
    rng = np.random.default_rng(300 + cfg)
    r32 = lambda *s: rng.integers(0, 1 << 32, s, dtype=np.uint64).astype(np.uint32)
    O = orc.lib()
    gg = (k + 1) * l * (k + 1) * N
    raw3 = r32(3 * gg)
    mf = C.c_double(0.0)
    for a0, a1 in ((1, 2), (N - 1, N + 7), (0, 5), (2 * N - 1, 2 * N - 1), (N, N), (0, 0)):
        acc = r32(k + 1, N)
        e = (a0 + a1, a0, a1)
        bundle = np.zeros(gg, dtype=np.uint32)
        rot = orc.z(N)
        for which in range(3):
            g = raw3[which * gg:(which + 1) * gg]
            for q in range(gg // N):
                O.orc_poly_mul_monomial(g[q * N:(q + 1) * N], N, e[which], rot)
                bundle[q * N:(q + 1) * N] += rot - g[q * N:(q + 1) * N]
        prod = orc.z((k + 1) * N)
        O.orc_external_product(C.byref(p), bundle, acc.reshape(-1), prod)
        exp = (acc.reshape(-1) + prod).astype(np.uint32)
        got = acc.copy().reshape(-1)
        assert emu_fft.emu_fft_step_bmmp(cfg, raw3, got, a0, a1, C.byref(mf)) == 0
        assert got.tolist() == exp.tolist(), (a0, a1)
        assert mf.value < 2.0 ** -8, mf.value
