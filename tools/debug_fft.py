#!/usr/bin/env python
"""Debug aid for the FFT path: GPU key transform vs the CPU emulation (bit-exact doubles), then one external product."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfhe_research_b200 as T
from oracle import orc

HERE = os.path.dirname(os.path.abspath(__file__))
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
E = C.CDLL(os.path.join(HERE, "..", "tests", "emu", "libemu_fft.so"))
E.emu_fft_transform_ggsw.argtypes = [C.c_int, u32p, f64p]
E.emu_fft_step.argtypes = [C.c_int, C.c_int, f64p, u32p, C.c_uint32, C.POINTER(C.c_double)]

p = T.TfheParams.preset("P1", lwe_dimension=3)
o = orc.params(**{f: getattr(p, f) for f, _ in T.TfheParams._fields_})
lwe_sk, glwe_sk, bsk, ksk = T.bootstrapping_key_gen(p, 0xB200)
ctx = T.Context(p, 0, path=T.PATH_FFT)
bk = ctx.upload_key(bsk, ksk)
gpu_key = bk.transformed()
gg = p.ggsw_words
per = gpu_key.size // p.n
for i in range(p.n):
    emu_key = np.zeros(per)
    E.emu_fft_transform_ggsw(1, bsk[i * gg:(i + 1) * gg], emu_key)
    g = gpu_key[i * per:(i + 1) * per]
    # the emulation stores rows (level, polynomial p); the device stores slots (level, d) with p = (c + d) mod P at column c
    Pn, Ln = p.k + 1, p.pbs_levels
    ek = emu_key.reshape(Ln, Pn, 2, Pn, -1)
    dk = np.empty_like(ek)
    for d in range(Pn):
        for c in range(Pn):
            dk[:, d, :, c] = ek[:, (c + d) % Pn, :, c]
    emu_key = dk.reshape(-1)
    bad = np.nonzero(g != emu_key)[0]
    print(f"ggsw {i}: transformed key mismatches: {bad.size} of {per}", bad[:8], g[bad[:4]], emu_key[bad[:4]])
rng = np.random.default_rng(2)
glwe = rng.integers(0, 1 << 32, (1, 2, 1024), dtype=np.uint64).astype(np.uint32)
got = ctx.external_product(bk, np.array([0], dtype=np.uint32), glwe)
exp = orc.z(2048)
orc.lib().orc_external_product(C.byref(o), bsk[:gg], glwe.reshape(-1), exp)
bad = np.nonzero(got.reshape(-1) != exp)[0]
print("external product mismatches:", bad.size, bad[:16])
print("margin", ctx.fft_rounding_margin())
emu_key = np.zeros(per); E.emu_fft_transform_ggsw(1, bsk[:gg], emu_key)
g2 = glwe.reshape(-1).copy(); mf = C.c_double(0)
E.emu_fft_step(1, 1, emu_key, g2, 0, C.byref(mf))
print("emu vs oracle mismatches:", int((g2 != exp).sum()))
d = (got.reshape(-1).astype(np.int64) - exp.astype(np.int64))
print("diff sample", d[:8], "diff mod 65536 zero?", int(((d % 65536) == 0).sum()))

# ---- unit-impulse probes: which GGSW row / column / rotation does the GPU pick up?
G = bsk[:gg].reshape(6, 2, 1024)


def negacyclic_shift(v, s):
    out = np.roll(v, s).astype(np.int64)
    out[:s] = -out[:s]
    return (out & 0xFFFFFFFF).astype(np.uint32)


for (pp, jj, val, name) in [(0, 0, 1 << 24, "p0 lev0"), (0, 0, 1 << 16, "p0 lev1"), (0, 0, 1 << 8, "p0 lev2"),
                            (1, 0, 1 << 24, "p1 lev0"), (0, 5, 1 << 24, "p0 lev0 X^5"), (0, 600, 1 << 24, "p0 lev0 X^600")]:
    x = np.zeros((1, 2, 1024), dtype=np.uint32)
    x[0, pp, jj] = val
    got = ctx.external_product(bk, np.array([0], dtype=np.uint32), x).reshape(2, 1024)
    exp = orc.z(2048)
    orc.lib().orc_external_product(C.byref(o), bsk[:gg], x.reshape(-1), exp)
    exp = exp.reshape(2, 1024)
    msg = []
    for c in range(2):
        hit = [(r, cc) for r in range(6) for cc in range(2) if np.array_equal(got[c], negacyclic_shift(G[r, cc], jj))]
        msg.append(f"col{c}: match_exp={np.array_equal(got[c], exp[c])} equals_row/col={hit} nz={int((got[c] != 0).sum())}")
    print(name, "|", " ; ".join(msg))

x = np.zeros((1, 2, 1024), dtype=np.uint32)
x[0, 0, 0] = 1 << 24
got = ctx.external_product(bk, np.array([0], dtype=np.uint32), x).reshape(2, 1024)
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/dbg_got.npy", got)
np.save("gpurun_out/dbg_G.npy", G)
