"""tfhe-research_b200 -- host-side mirror of the reference's API over the C-ABI library.

The Rust toolchain is absent from this image, so the host side above `include/tfhe_b200.h` is this
thin ctypes layer; function names and argument meaning follow the reference crate
(bootstrapping.rs, boolean.rs, test_vector.rs, lwe.rs).  All arithmetic happens inside
`libtfhe_b200.so` (hand-written CUDA for sm_100a + C++ host code).  There is NO CPU fallback: if the
library is missing, or no GPU is visible, the device entry points raise.

Import it as `tfhe_research_b200` (the root-level shim module) -- a directory name with a hyphen is
not importable directly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("TFHE_B200_LIB") or os.path.join(_HERE, "libtfhe_b200.so")   # override: A/B runs of kernel variants
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared", "-ccbin", "/usr/bin/g++", "-ldl"]
_SOURCES = ["tfhe_b200.cu", "host_api.cpp", "tfhe_mgpu.cpp"]
_DEPS = _SOURCES + ["kernels.cuh", "pbs_team.cuh", "tfhe_core.cuh", "host_tables.hpp", "api_internal.hpp",
                    "kernels_fft.cuh", "kernels_fft_latency.cuh", "kernels_fft_cluster.cuh", "kernels_ks_tcgen05.cuh", "fft_team.cuh", "fft_tmem.cuh", "host_tables_fft.hpp", "tfhe_mgpu.cpp",
                    os.path.join("..", "..", "include", "tfhe_b200.h")]

TFHE_OK, TFHE_E_PARAM, TFHE_E_CUDA, TFHE_E_OOM, TFHE_E_ASSERT, TFHE_E_NCCL = 0, -1, -2, -3, -4, -5
AND, OR, XOR, NAND, NOR, XNOR = range(6)
PATH_NTT, PATH_FFT = 0, 1   # arithmetic path of the external product (include/tfhe_b200.h TFHE_PATH_*)
KS_IMAD, KS_MMA, KS_TCGEN05 = 0, 1, 2   # arithmetic of the key-switching product (include/tfhe_b200.h TFHE_KS_*)


class TfheError(RuntimeError):
    def __init__(self, code, msg=""):
        names = {-1: "TFHE_E_PARAM", -2: "TFHE_E_CUDA", -3: "TFHE_E_OOM", -4: "TFHE_E_ASSERT", -5: "TFHE_E_NCCL"}
        super().__init__(f"{names.get(code, code)}: {msg}")
        self.code = code


class TfheParams(C.Structure):
    """lib.rs:23-34 (glwe_poly_degree holds log2 N, as in the reference)."""
    _fields_ = [("glwe_dimension", C.c_uint32), ("glwe_poly_degree", C.c_uint32), ("lwe_dimension", C.c_uint32),
                ("padding_bits", C.c_uint32), ("log_p", C.c_uint32), ("log_q", C.c_uint32),
                ("ks_log_base", C.c_uint32), ("ks_levels", C.c_uint32), ("pbs_log_base", C.c_uint32),
                ("pbs_levels", C.c_uint32), ("lwe_std_dev", C.c_double), ("glwe_std_dev", C.c_double)]

    @classmethod
    def default(cls, test_cfg: bool = False):
        p = cls()
        _check(lib().tfhe_params_default(1 if test_cfg else 0, C.byref(p)))
        return p

    @classmethod
    def preset(cls, name: str, **over):
        p = cls()
        _check(lib().tfhe_params_preset(name.encode(), C.byref(p)), f"unknown preset {name}")
        for k, v in over.items():
            setattr(p, k, v)
        return p

    N = property(lambda s: 1 << s.glwe_poly_degree)
    k = property(lambda s: s.glwe_dimension)
    n = property(lambda s: s.lwe_dimension)
    glwe_words = property(lambda s: (s.k + 1) * s.N)
    ggsw_words = property(lambda s: (s.k + 1) * s.pbs_levels * (s.k + 1) * s.N)
    bsk_words = property(lambda s: s.n * s.ggsw_words)
    ksk_words = property(lambda s: s.k * s.N * s.ks_levels * (s.n + 1))

    def validate(self):
        _check(lib().tfhe_params_validate(C.byref(self)), "unsupported parameter set")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libtfhe_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    deps = [os.path.join(_CSRC, d) for d in _DEPS]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(d) <= os.path.getmtime(LIB_PATH) for d in deps):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + [os.path.join(_CSRC, s) for s in _SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


_LIB = None
u32p = C.c_void_p  # host numpy pointer or device pointer


def lib():
    """Load the C-ABI library; fails loudly when it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    PP, VP, SZ = C.POINTER(TfheParams), C.c_void_p, C.c_size_t
    sig = {
        "tfhe_params_default": [C.c_int, PP], "tfhe_params_preset": [C.c_char_p, PP], "tfhe_params_validate": [PP],
        "tfhe_test_vector_from_lut": [PP, VP, SZ, VP], "tfhe_test_vector_identity": [PP, VP],
        "tfhe_test_vector_boolean": [PP, C.c_int, VP],
        "tfhe_lwe_encode": [PP, C.c_uint32, C.POINTER(C.c_uint32)], "tfhe_lwe_decode": [PP, C.c_uint32, C.POINTER(C.c_uint32)],
        "tfhe_lwe_encrypt": [PP, VP, SZ, C.c_uint32, C.c_uint64, C.c_uint64, VP],
        "tfhe_lwe_decrypt": [VP, SZ, VP, C.POINTER(C.c_uint32)],
        "tfhe_keygen": [PP, C.c_uint64, VP, VP, VP, VP], "tfhe_keygen_bmmp": [PP, C.c_uint64, VP, VP, VP, VP],
        "tfhe_bk_upload_bmmp": [VP, VP, VP, C.POINTER(VP)],
        "tfhe_ctx_create": [PP, C.c_int, C.POINTER(VP)], "tfhe_ctx_set_stream": [VP, VP],
        "tfhe_ctx_set_pbs_path": [VP, C.c_int], "tfhe_ctx_get_pbs_path": [VP], "tfhe_ctx_set_ks_path": [VP, C.c_int], "tfhe_ctx_set_fft_check": [VP, C.c_int], "tfhe_ctx_set_latency_config": [VP, C.c_int], "tfhe_ctx_set_fft_exchange": [VP, C.c_int],
        "tfhe_fft_rounding_margin": [VP, C.POINTER(C.c_double)],
        "tfhe_bk_upload": [VP, VP, VP, C.POINTER(VP)], "tfhe_bk_read_transformed": [VP, VP, SZ],
        "tfhe_bootstrap_batch": [VP, VP, VP, VP, SZ, VP, SZ, VP],
        "tfhe_bootstrap_batch_ks_first": [VP, VP, VP, VP, SZ, VP, SZ, VP],
        "tfhe_gate_k_batch": [VP, VP, C.c_uint32, C.c_uint32, VP, SZ, VP],
        "tfhe_file_write": [C.c_char_p, C.c_int, PP, VP, C.c_uint64], "tfhe_file_read": [C.c_char_p, VP, VP, C.c_uint64],
        "tfhe_gate_batch": [VP, VP, C.c_int, VP, VP, SZ, VP], "tfhe_gates_batch": [VP, VP, VP, VP, VP, SZ, VP],
        "tfhe_switch_modulus": [VP, VP, SZ, VP], "tfhe_decompose": [VP, C.c_int, VP, SZ, VP],
        "tfhe_glwe_mul_monomial": [VP, VP, VP, SZ, VP],
        "tfhe_negacyclic_mul": [VP, VP, VP, SZ, VP],
        "tfhe_external_product": [VP, VP, VP, VP, SZ, VP], "tfhe_cmux": [VP, VP, VP, VP, VP, SZ, VP],
        "tfhe_blind_rotate": [VP, VP, VP, VP, SZ, VP, SZ, VP],
        "tfhe_sample_extract": [VP, VP, SZ, VP], "tfhe_key_switch": [VP, VP, VP, SZ, VP],
        "tfhe_gate_linear": [VP, VP, VP, SZ, VP],
        "tfhe_bk_get_path": [VP],
        "tfhe_mgpu_create": [PP, C.c_int, VP, C.POINTER(VP)], "tfhe_mgpu_n_gpus": [VP],
        "tfhe_mgpu_bk_upload": [VP, VP, VP, C.POINTER(VP)], "tfhe_mgpu_bk_upload_bmmp": [VP, VP, VP, C.POINTER(VP)],
        "tfhe_mgpu_bootstrap_batch": [VP, VP, VP, VP, SZ, VP, SZ, VP], "tfhe_mgpu_gates_batch": [VP, VP, VP, VP, VP, SZ, VP],
        "tfhe_mgpu_last_timing": [VP, C.POINTER(C.c_double * 4)],
        "tfhe_measure_int_peak": [VP, C.POINTER(C.c_double * 8)], "tfhe_measure_fp64_peak": [VP, C.POINTER(C.c_double * 4)], "tfhe_last_timing": [VP, C.POINTER(C.c_double * 3)],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = C.c_int
    L.tfhe_ctx_destroy.argtypes = [VP]; L.tfhe_ctx_destroy.restype = None
    L.tfhe_bk_free.argtypes = [VP]; L.tfhe_bk_free.restype = None
    L.tfhe_last_error.argtypes = [VP]; L.tfhe_last_error.restype = C.c_char_p
    L.tfhe_ctx_launch_count.argtypes = [VP]; L.tfhe_ctx_launch_count.restype = C.c_uint64
    L.tfhe_bk_transformed_bytes.argtypes = [VP]; L.tfhe_bk_transformed_bytes.restype = C.c_size_t
    L.tfhe_ctx_get_stream.argtypes = [VP]; L.tfhe_ctx_get_stream.restype = C.c_void_p
    L.tfhe_mgpu_destroy.argtypes = [VP]; L.tfhe_mgpu_destroy.restype = None
    L.tfhe_mgpu_bk_free.argtypes = [VP]; L.tfhe_mgpu_bk_free.restype = None
    L.tfhe_mgpu_ctx.argtypes = [VP, C.c_int]; L.tfhe_mgpu_ctx.restype = C.c_void_p
    L.tfhe_mgpu_last_error.argtypes = [VP]; L.tfhe_mgpu_last_error.restype = C.c_char_p
    _LIB = L
    return L


EXPORTS = [
    "tfhe_params_default", "tfhe_params_preset", "tfhe_params_validate", "tfhe_test_vector_from_lut",
    "tfhe_test_vector_identity", "tfhe_test_vector_boolean", "tfhe_lwe_encode", "tfhe_lwe_decode", "tfhe_lwe_encrypt",
    "tfhe_lwe_decrypt", "tfhe_keygen", "tfhe_ctx_create", "tfhe_ctx_destroy", "tfhe_last_error", "tfhe_ctx_set_stream",
    "tfhe_ctx_launch_count", "tfhe_bk_upload", "tfhe_bk_free", "tfhe_bootstrap_batch", "tfhe_gate_batch",
    "tfhe_gates_batch", "tfhe_switch_modulus", "tfhe_decompose", "tfhe_glwe_mul_monomial", "tfhe_external_product",
    "tfhe_negacyclic_mul", "tfhe_cmux", "tfhe_blind_rotate", "tfhe_sample_extract", "tfhe_key_switch", "tfhe_gate_linear",
    "tfhe_measure_int_peak", "tfhe_last_timing", "tfhe_ctx_set_pbs_path", "tfhe_ctx_get_pbs_path", "tfhe_ctx_set_ks_path", "tfhe_fft_rounding_margin", "tfhe_ctx_set_fft_check", "tfhe_measure_fp64_peak",
    "tfhe_bk_transformed_bytes", "tfhe_bk_read_transformed", "tfhe_bootstrap_batch_ks_first", "tfhe_gate_k_batch",
    "tfhe_file_write", "tfhe_file_read", "tfhe_keygen_bmmp", "tfhe_bk_upload_bmmp", "tfhe_bk_get_path", "tfhe_ctx_get_stream",
    "tfhe_ctx_set_latency_config", "tfhe_ctx_set_fft_exchange", "tfhe_mgpu_create", "tfhe_mgpu_destroy", "tfhe_mgpu_n_gpus", "tfhe_mgpu_ctx", "tfhe_mgpu_last_error", "tfhe_mgpu_bk_upload",
    "tfhe_mgpu_bk_upload_bmmp", "tfhe_mgpu_bk_free", "tfhe_mgpu_bootstrap_batch", "tfhe_mgpu_gates_batch", "tfhe_mgpu_last_timing",
]


def _check(rc, msg=""):
    if rc != 0:
        raise TfheError(rc, msg)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x):
    """host numpy array or CUDA torch tensor -> raw pointer (kept alive by the caller)."""
    if x is None:
        return None
    if _is_torch(x):
        import torch
        if x.dtype not in (torch.int32, torch.uint32) or not x.is_contiguous():
            raise TypeError(f"expected a contiguous int32/uint32 tensor, got {x.dtype}, contiguous={x.is_contiguous()}")
        return x.data_ptr()
    if x.dtype != np.uint32 or not x.flags["C_CONTIGUOUS"]:
        raise TypeError(f"expected a C-contiguous uint32 array, got {x.dtype}")
    return x.ctypes.data


def _shape(x):
    return tuple(x.shape)


def _rows(x, width, what):
    """Batch size of x = [B, width] (or a single row [width]); anything else would be an out-of-bounds device access."""
    sh = _shape(x)
    if len(sh) == 1 and sh[0] == width:
        return 1
    if len(sh) == 2 and sh[1] == width:
        return sh[0]
    raise ValueError(f"{what}: expected shape [B, {width}] (or [{width}]), got {list(sh)}")


def _check_out(out, shape, like, what="out"):
    if _shape(out) != tuple(shape):
        raise ValueError(f"{what}: expected shape {list(shape)}, got {list(_shape(out))}")
    if _is_torch(out) != _is_torch(like):
        raise ValueError(f"{what}: must live where the input lives (both numpy or both CUDA tensors)")
    if _is_torch(out) and out.device != like.device:
        raise ValueError(f"{what}: on {out.device}, input on {like.device}")


def _u32(x):
    if _is_torch(x):
        return x if x.is_contiguous() else x.contiguous()
    return np.ascontiguousarray(x, dtype=np.uint32)


def _like(x, shape):
    if _is_torch(x):
        import torch
        return torch.empty(shape, dtype=x.dtype, device=x.device)
    return np.empty(shape, dtype=np.uint32)


# ------------------------------------------------------------------ host-side mirror (no GPU needed)
def construct_test_from_lut(params: TfheParams, lut) -> np.ndarray:
    """test_vector.rs:38-67."""
    lut = np.ascontiguousarray(lut, dtype=np.uint32)
    tv = np.empty(params.N, dtype=np.uint32)
    _check(lib().tfhe_test_vector_from_lut(C.byref(params), lut.ctypes.data, len(lut), tv.ctypes.data),
           "assert!(lut.len() == plaintext_modulus) test_vector.rs:41")
    return tv


def construct_identity_test_vector(params: TfheParams) -> np.ndarray:
    """test_vector.rs:23-35."""
    tv = np.empty(params.N, dtype=np.uint32)
    _check(lib().tfhe_test_vector_identity(C.byref(params), tv.ctypes.data))
    return tv


def construct_test_vector_boolean(params: TfheParams, gate: int) -> np.ndarray:
    """test_vector.rs:5-20 with f in {AND, OR, XOR}."""
    tv = np.empty(params.N, dtype=np.uint32)
    _check(lib().tfhe_test_vector_boolean(C.byref(params), gate, tv.ctypes.data))
    return tv


def encode_message(params: TfheParams, m: int) -> int:
    """LweCleartext::encode_message lwe.rs:83-88."""
    out = C.c_uint32()
    _check(lib().tfhe_lwe_encode(C.byref(params), m, C.byref(out)), "assert!(m < 1 << log_p) lwe.rs:84")
    return out.value


def decode(params: TfheParams, plaintext: int) -> int:
    """LwePlaintext::decode lwe.rs:102-107 (floor shift, no mask -- reference behaviour, H5)."""
    out = C.c_uint32()
    _check(lib().tfhe_lwe_decode(C.byref(params), plaintext, C.byref(out)))
    return out.value


def decode_rounded(params: TfheParams, plaintext: int) -> int:
    """Harness decoder: round to the nearest message and mask (NOT the reference's decode)."""
    shift = params.log_q - (params.log_p + params.padding_bits)
    return ((plaintext + (1 << (shift - 1))) >> shift) & ((1 << (params.log_p + params.padding_bits)) - 1)


def encrypt_lwe_plaintext(params: TfheParams, sk, plaintext: int, seed: int, index: int) -> np.ndarray:
    """lwe.rs:138-160 with a seeded RNG (stream = (seed, index))."""
    sk = np.ascontiguousarray(sk, dtype=np.uint32)
    ct = np.empty(len(sk) + 1, dtype=np.uint32)
    _check(lib().tfhe_lwe_encrypt(C.byref(params), sk.ctypes.data, len(sk), plaintext, seed, index, ct.ctypes.data))
    return ct


def decrypt_lwe(sk, ct) -> int:
    """lwe.rs:162-173."""
    sk = np.ascontiguousarray(sk, dtype=np.uint32)
    ct = np.ascontiguousarray(ct, dtype=np.uint32)
    out = C.c_uint32()
    _check(lib().tfhe_lwe_decrypt(sk.ctypes.data, len(sk), ct.ctypes.data, C.byref(out)))
    return out.value


def bootstrapping_key_gen(params: TfheParams, seed: int):
    """bootstrapping.rs:23-56 (+ the two secret keys): returns (lwe_sk, glwe_sk, bsk, ksk) flat u32 arrays."""
    lwe_sk = np.empty(params.n, dtype=np.uint32)
    glwe_sk = np.empty(params.k * params.N, dtype=np.uint32)
    bsk = np.empty(params.bsk_words, dtype=np.uint32)
    ksk = np.empty(params.ksk_words, dtype=np.uint32)
    _check(lib().tfhe_keygen(C.byref(params), seed, lwe_sk.ctypes.data, glwe_sk.ctypes.data, bsk.ctypes.data, ksk.ctypes.data))
    return lwe_sk, glwe_sk, bsk, ksk


def bootstrapping_key_gen_bmmp(params: TfheParams, seed: int):
    """BMMP key triples (notes/BMMP Bootstrapping.md:21-25): returns (lwe_sk, glwe_sk, bsk3, ksk); same secret keys and
    KSK as bootstrapping_key_gen(params, seed)."""
    lwe_sk = np.empty(params.n, dtype=np.uint32)
    glwe_sk = np.empty(params.k * params.N, dtype=np.uint32)
    bsk3 = np.empty(3 * (params.n // 2) * params.ggsw_words, dtype=np.uint32)
    ksk = np.empty(params.ksk_words, dtype=np.uint32)
    _check(lib().tfhe_keygen_bmmp(C.byref(params), seed, lwe_sk.ctypes.data, glwe_sk.ctypes.data, bsk3.ctypes.data, ksk.ctypes.data),
           "tfhe_keygen_bmmp (even lwe_dimension required)")
    return lwe_sk, glwe_sk, bsk3, ksk


# ------------------------------------------------------------------ flat wire / on-disk format (include/tfhe_b200.h)
FILE_LWE_BATCH, FILE_GLWE_BATCH, FILE_BSK, FILE_KSK, FILE_LWE_SK, FILE_GLWE_SK, FILE_TEST_VECTOR = range(1, 8)


class FileHeader(C.Structure):
    _fields_ = [("magic", C.c_char * 8), ("version", C.c_uint32), ("kind", C.c_uint32), ("params", TfheParams), ("count", C.c_uint64)]


def save_words(path: str, kind: int, params: TfheParams, words) -> None:
    """Write one array in the flat little-endian format (header + u32 words in the reference's ndarray layout)."""
    w = np.ascontiguousarray(words, dtype=np.uint32).reshape(-1)
    _check(lib().tfhe_file_write(os.fsencode(path), kind, C.byref(params), w.ctypes.data, w.size), f"cannot write {path}")


def load_words(path: str):
    """-> (kind, TfheParams, flat uint32 array)."""
    h = FileHeader()
    _check(lib().tfhe_file_read(os.fsencode(path), C.byref(h), None, 0), f"cannot read {path}")
    w = np.empty(h.count, dtype=np.uint32)
    _check(lib().tfhe_file_read(os.fsencode(path), C.byref(h), w.ctypes.data, w.size), f"cannot read {path}")
    p = TfheParams()
    C.memmove(C.byref(p), C.byref(h.params), C.sizeof(TfheParams))
    return int(h.kind), p, w


def lwe_secret_key_from_glwe(glwe_sk) -> np.ndarray:
    """LweSecretKey::from(&GlweSecretKey) lwe.rs:62-73: the row-major flattening (dimension kN)."""
    return np.ascontiguousarray(glwe_sk, dtype=np.uint32).reshape(-1).copy()


# ------------------------------------------------------------------ device side
class BootstrappingKey:
    """Device-resident BootstrappingKey (bootstrapping.rs:18-21): NTT-domain BSK + KSK."""

    def __init__(self, ctx, handle):
        self.ctx, self._h = ctx, handle
        ctx._keys.add(self)

    def free(self):
        if self._h:
            lib().tfhe_bk_free(self._h)
            self._h = None
            self.ctx._keys.discard(self)

    @property
    def path(self) -> int:
        """Arithmetic path the key was transformed for (fixed at upload; independent of later Context.set_pbs_path calls)."""
        return int(lib().tfhe_bk_get_path(self._h))

    def transformed(self) -> np.ndarray:
        """Host copy of the transformed BSK (u32 residues on the NTT path, float64 re/im pairs on the FFT path)."""
        nbytes = lib().tfhe_bk_transformed_bytes(self._h)
        out = np.empty(nbytes // 8, dtype=np.float64) if self.path == PATH_FFT else np.empty(nbytes // 4, dtype=np.uint32)
        self.ctx._ck(lib().tfhe_bk_read_transformed(self._h, out.ctypes.data, nbytes))
        return out

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One CUDA device.  Inputs may be numpy arrays (host, copied in/out) or CUDA torch tensors (in place)."""

    def __init__(self, params: TfheParams, device: int = 0, path=None):
        self.params = params
        h = C.c_void_p()
        rc = lib().tfhe_ctx_create(C.byref(params), device, C.byref(h))
        if rc == TFHE_E_CUDA:
            raise TfheError(rc, "no usable CUDA device: the PBS path has no CPU fallback")
        _check(rc, "tfhe_ctx_create")
        self._h = h
        self.device = device
        self._keys = weakref.WeakSet()   # live keys of this context: freed by close() before the context goes
        if path is not None:
            self.set_pbs_path(path)

    def set_pbs_path(self, path: int):
        """PATH_NTT (2-prime integer NTT) or PATH_FFT (exact FP64 FFT, limb-split key); call before upload_key."""
        self._ck(lib().tfhe_ctx_set_pbs_path(self._h, int(path)))

    def set_ks_path(self, path: int):
        """KS_IMAD (32-bit multiply-adds), KS_MMA (integer tensor cores via mma.sync, exact through byte planes) or KS_TCGEN05
        (the same product with tcgen05.mma + TMEM + TMA); same bits."""
        self._ck(lib().tfhe_ctx_set_ks_path(self._h, int(path)))

    @property
    def pbs_path(self) -> int:
        return int(lib().tfhe_ctx_get_pbs_path(self._h))

    def set_fft_check(self, on: bool = True):
        """FFT path: run the kernel variant that records the rounding margin (see fft_rounding_margin)."""
        self._ck(lib().tfhe_ctx_set_fft_check(self._h, 1 if on else 0))

    def set_latency_config(self, mode=4):
        """FFT path, small batches: 4 (default) / 3 a cluster of L CTAs per ciphertext, one gadget level each (batch <= SMs / L;
        4 also splits every CTA by key limb over twice the warps),
        2 all teams of a CTA on its one ciphertext (batch <= SMs), 1 one team with a deep key ring, 0 / False the throughput
        configuration.  Same bits."""
        self._ck(lib().tfhe_ctx_set_latency_config(self._h, 4 if mode is True else int(mode)))

    def set_fft_exchange(self, tensor_memory: bool = True):
        """FFT path, N = 512: exchange the register passes of the transforms through tensor memory (default) or shared memory.
        Keys uploaded while it is on carry the second key copy that kernel needs; same bits either way."""
        self._ck(lib().tfhe_ctx_set_fft_exchange(self._h, 1 if tensor_memory else 0))

    def fft_rounding_margin(self) -> float:
        """Largest distance to an integer of any value rounded by the FFT path since the last call (must be << 0.5)."""
        out = C.c_double()
        self._ck(lib().tfhe_fft_rounding_margin(self._h, C.byref(out)))
        return out.value

    def _ck(self, rc):
        if rc != 0:
            raise TfheError(rc, (lib().tfhe_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            for k in list(self._keys):
                k.free()
            lib().tfhe_ctx_destroy(self._h)
            self._h = None

    def _on_device(self, *xs):
        for x in xs:
            if x is not None and _is_torch(x) and x.is_cuda and x.device.index != self.device:   # CPU (e.g. pinned) tensors are host pointers
                raise ValueError(f"tensor on {x.device}: this context drives cuda:{self.device}")

    def set_stream(self, cuda_stream_ptr):
        self._ck(lib().tfhe_ctx_set_stream(self._h, cuda_stream_ptr))

    @property
    def launch_count(self) -> int:
        return int(lib().tfhe_ctx_launch_count(self._h))

    def upload_key(self, bsk, ksk) -> BootstrappingKey:
        bsk, ksk = _u32(bsk), _u32(ksk)
        h = C.c_void_p()
        self._ck(lib().tfhe_bk_upload(self._h, _ptr(bsk), _ptr(ksk), C.byref(h)))
        return BootstrappingKey(self, h)

    def upload_key_bmmp(self, bsk3, ksk) -> BootstrappingKey:
        """Key triples of the unrolled-by-two blind rotation; use the returned key with bootstrap / gate / blind_rotate."""
        bsk3, ksk = _u32(bsk3), _u32(ksk)
        h = C.c_void_p()
        self._ck(lib().tfhe_bk_upload_bmmp(self._h, _ptr(bsk3), _ptr(ksk), C.byref(h)))
        return BootstrappingKey(self, h)

    # -- bootstrapping.rs:58-120, batched
    def bootstrap(self, bk: BootstrappingKey, lwe_in, test_vectors, lut_idx=None, out=None):
        p = self.params
        lwe_in, tvs = _u32(lwe_in), _u32(test_vectors)
        B, T = _rows(lwe_in, p.n + 1, "lwe_in"), _rows(tvs, p.N, "test_vectors")
        idx = None if lut_idx is None else _u32(lut_idx)
        if idx is not None and _shape(idx) != (B,):
            raise ValueError(f"lut_idx: expected shape [{B}], got {list(_shape(idx))}")
        self._on_device(lwe_in, tvs, idx, out)
        if out is None:
            out = _like(lwe_in, (B, p.n + 1))
        else:
            _check_out(out, (B, p.n + 1), lwe_in)
        self._ck(lib().tfhe_bootstrap_batch(self._h, bk._h, _ptr(lwe_in), _ptr(tvs), T, _ptr(idx), B, _ptr(out)))
        return out

    # -- boolean.rs:9-53, batched
    def gate(self, bk: BootstrappingKey, op, ct0, ct1, out=None):
        p = self.params
        ct0, ct1 = _u32(ct0), _u32(ct1)
        B = _rows(ct0, p.n + 1, "ct0")
        if _rows(ct1, p.n + 1, "ct1") != B:
            raise ValueError("ct0 and ct1 must hold the same number of ciphertexts")
        self._on_device(ct0, ct1, out)
        if out is None:
            out = _like(ct0, (B, p.n + 1))
        else:
            _check_out(out, (B, p.n + 1), ct0)
        if isinstance(op, (int, np.integer)):
            self._ck(lib().tfhe_gate_batch(self._h, bk._h, int(op), _ptr(ct0), _ptr(ct1), B, _ptr(out)))
        else:
            ops = np.ascontiguousarray(op, dtype=np.uint8)
            assert len(ops) == B
            self._ck(lib().tfhe_gates_batch(self._h, bk._h, ops.ctypes.data, _ptr(ct0), _ptr(ct1), B, _ptr(out)))
        return out

    # -- notes/TFHE.md:365-400: key switch first, then blind rotation + sample extraction (input/output dimension kN)
    def bootstrap_ks_first(self, bk: BootstrappingKey, lwe_in, test_vectors, lut_idx=None, out=None):
        p = self.params
        lwe_in, tvs = _u32(lwe_in), _u32(test_vectors)
        B, T = _rows(lwe_in, p.k * p.N + 1, "lwe_in"), _rows(tvs, p.N, "test_vectors")
        idx = None if lut_idx is None else _u32(lut_idx)
        if idx is not None and _shape(idx) != (B,):
            raise ValueError(f"lut_idx: expected shape [{B}], got {list(_shape(idx))}")
        self._on_device(lwe_in, tvs, idx, out)
        if out is None:
            out = _like(lwe_in, (B, p.k * p.N + 1))
        else:
            _check_out(out, (B, p.k * p.N + 1), lwe_in)
        self._ck(lib().tfhe_bootstrap_batch_ks_first(self._h, bk._h, _ptr(lwe_in), _ptr(tvs), T, _ptr(idx), B, _ptr(out)))
        return out

    # -- notes/Boolean Gates.md:9-11: k-input gate, ct_in = sum 2^i * cts[i]; truth_table bit j = f(j)
    def gate_k(self, bk: BootstrappingKey, truth_table: int, cts, out=None):
        p = self.params
        cts = [_u32(c) for c in cts]
        B = _rows(cts[0], p.n + 1, "cts[0]")
        if any(_rows(c, p.n + 1, "cts") != B for c in cts):
            raise ValueError("all inputs of a k-input gate must hold the same number of ciphertexts")
        self._on_device(*cts, out)
        if out is None:
            out = _like(cts[0], (B, p.n + 1))
        else:
            _check_out(out, (B, p.n + 1), cts[0])
        ptrs = (C.c_void_p * len(cts))(*[_ptr(c) for c in cts])
        self._ck(lib().tfhe_gate_k_batch(self._h, bk._h, len(cts), truth_table, ptrs, B, _ptr(out)))
        return out

    def and_(self, bk, ct0, ct1):
        """boolean.rs:9-30"""
        return self.gate(bk, AND, ct0, ct1)

    def or_(self, bk, ct0, ct1):
        """boolean.rs:32-53"""
        return self.gate(bk, OR, ct0, ct1)

    # -- sub-operations
    def switch_modulus(self, values):
        values = _u32(values)
        out = _like(values, values.shape)
        self._ck(lib().tfhe_switch_modulus(self._h, _ptr(values), values.size if not _is_torch(values) else values.numel(), _ptr(out)))
        return out

    def decompose(self, values, which=0):
        values = _u32(values)
        lv = self.params.ks_levels if which else self.params.pbs_levels
        n = values.size if not _is_torch(values) else values.numel()
        out = _like(values, (n, lv))
        self._ck(lib().tfhe_decompose(self._h, which, _ptr(values), n, _ptr(out)))
        return out

    def glwe_mul_monomial(self, glwe, index):
        glwe = _u32(glwe)
        index = np.ascontiguousarray(index, dtype=np.int64)
        if _shape(glwe) != (len(index), self.params.k + 1, self.params.N):
            raise ValueError(f"glwe: expected shape [{len(index)}, {self.params.k + 1}, {self.params.N}], got {list(_shape(glwe))}")
        self._on_device(glwe)
        out = _like(glwe, glwe.shape)
        self._ck(lib().tfhe_glwe_mul_monomial(self._h, _ptr(glwe), index.ctypes.data, len(index), _ptr(out)))
        return out

    def negacyclic_mul(self, a_small, g):
        """utils.rs:155-160 poly_mul: a int32 [B, N] (small signed digits, or any word mod 2^32), g uint32 [B, N]."""
        a_small = np.ascontiguousarray(a_small, dtype=np.int32)
        g = np.ascontiguousarray(g, dtype=np.uint32)
        if g.ndim != 2 or g.shape[1] != self.params.N or a_small.shape != g.shape:
            raise ValueError(f"a_small, g: expected two [B, {self.params.N}] arrays, got {list(a_small.shape)} and {list(g.shape)}")
        out = np.empty(g.shape, dtype=np.uint32)
        self._ck(lib().tfhe_negacyclic_mul(self._h, a_small.ctypes.data, _ptr(g), g.shape[0], out.ctypes.data))
        return out

    def external_product(self, bk, ggsw_index, glwe):
        glwe = _u32(glwe)
        gi = np.ascontiguousarray(ggsw_index, dtype=np.uint32)
        if _shape(glwe) != (len(gi), self.params.k + 1, self.params.N):
            raise ValueError(f"glwe: expected shape [{len(gi)}, {self.params.k + 1}, {self.params.N}], got {list(_shape(glwe))}")
        self._on_device(glwe)
        out = _like(glwe, glwe.shape)
        self._ck(lib().tfhe_external_product(self._h, bk._h, gi.ctypes.data, _ptr(glwe), len(gi), _ptr(out)))
        return out

    def cmux(self, bk, ggsw_index, ct0, ct1):
        ct0, ct1 = _u32(ct0), _u32(ct1)
        gi = np.ascontiguousarray(ggsw_index, dtype=np.uint32)
        want = (len(gi), self.params.k + 1, self.params.N)
        if _shape(ct0) != want or _shape(ct1) != want:
            raise ValueError(f"ct0, ct1: expected shape {list(want)}, got {list(_shape(ct0))} and {list(_shape(ct1))}")
        self._on_device(ct0, ct1)
        out = _like(ct0, ct0.shape)
        self._ck(lib().tfhe_cmux(self._h, bk._h, gi.ctypes.data, _ptr(ct0), _ptr(ct1), len(gi), _ptr(out)))
        return out

    def blind_rotate(self, bk, lwe_in, test_vectors, lut_idx=None):
        p = self.params
        lwe_in, tvs = _u32(lwe_in), _u32(test_vectors)
        B, T = _rows(lwe_in, p.n + 1, "lwe_in"), _rows(tvs, p.N, "test_vectors")
        idx = None if lut_idx is None else _u32(lut_idx)
        if idx is not None and _shape(idx) != (B,):
            raise ValueError(f"lut_idx: expected shape [{B}], got {list(_shape(idx))}")
        self._on_device(lwe_in, tvs, idx)
        out = _like(lwe_in, (B, p.k + 1, p.N))
        self._ck(lib().tfhe_blind_rotate(self._h, bk._h, _ptr(lwe_in), _ptr(tvs), T, _ptr(idx), B, _ptr(out)))
        return out

    def sample_extract(self, glwe):
        p = self.params
        glwe = _u32(glwe)
        if _shape(glwe)[-2:] != (p.k + 1, p.N):
            raise ValueError(f"glwe: expected shape [B, {p.k + 1}, {p.N}], got {list(_shape(glwe))}")
        B = glwe.shape[0] if glwe.ndim == 3 else 1
        self._on_device(glwe)
        out = _like(glwe, (B, p.k * p.N + 1))
        self._ck(lib().tfhe_sample_extract(self._h, _ptr(glwe), B, _ptr(out)))
        return out

    def key_switch(self, bk, lwe_in):
        p = self.params
        lwe_in = _u32(lwe_in)
        B = _rows(lwe_in, p.k * p.N + 1, "lwe_in")
        self._on_device(lwe_in)
        out = _like(lwe_in, (B, p.n + 1))
        self._ck(lib().tfhe_key_switch(self._h, bk._h, _ptr(lwe_in), B, _ptr(out)))
        return out

    def gate_linear(self, ct0, ct1):
        ct0, ct1 = _u32(ct0), _u32(ct1)
        B = _rows(ct0, self.params.n + 1, "ct0")
        if _shape(ct1) != _shape(ct0):
            raise ValueError("ct0 and ct1 must have the same shape")
        self._on_device(ct0, ct1)
        out = _like(ct0, ct0.shape)
        self._ck(lib().tfhe_gate_linear(self._h, _ptr(ct0), _ptr(ct1), B, _ptr(out)))
        return out

    def measure_int_peak(self):
        out = (C.c_double * 8)()
        self._ck(lib().tfhe_measure_int_peak(self._h, C.byref(out)))
        return {"imad": out[0], "imad_hi": out[1], "imad_wide": out[2], "shoup_butterfly_shared_tw": out[3],
                "shoup_butterfly": out[4], "shoup_butterfly_imm_q": out[5],
                "butterfly_stream_3cta": out[6], "butterfly_stream_8cta": out[7]}

    def measure_fp64_peak(self):
        out = (C.c_double * 4)()
        self._ck(lib().tfhe_measure_fp64_peak(self._h, C.byref(out)))
        return {"dfma": out[0], "dfma_3reg": out[1], "fft_butterfly": out[2]}

    def last_timing(self):
        out = (C.c_double * 3)()
        self._ck(lib().tfhe_last_timing(self._h, C.byref(out)))
        return {"blind_rotate_ms": out[0], "key_switch_ms": out[1], "total_ms": out[2]}


# ------------------------------------------------------------------ one process, all GPUs of the box (include/tfhe_b200.h "multi-GPU")
class MultiGpuKey:
    """BootstrappingKey replicated on every device of a MultiGpuContext."""

    def __init__(self, m, handle):
        self.m, self._h = m, handle
        m._keys.add(self)

    def free(self):
        if self._h:
            lib().tfhe_mgpu_bk_free(self._h)
            self._h = None
            self.m._keys.discard(self)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MultiGpuContext:
    """tfhe_mgpu: batches are split into contiguous balanced ranges, one per GPU, keys replicated.

    numpy inputs: every GPU copies its own shard in and out; CUDA tensors (all on ONE of the devices): NCCL
    send/recv scatter, local bootstraps, NCCL gather into the output tensor on that device.  Same bits as Context.
    """

    def __init__(self, params: TfheParams, n_gpus: int, devices=None):
        self.params = params
        h = C.c_void_p()
        dev = None if devices is None else (C.c_int * n_gpus)(*devices)
        rc = lib().tfhe_mgpu_create(C.byref(params), n_gpus, dev, C.byref(h))
        if rc == TFHE_E_CUDA:
            raise TfheError(rc, "no usable CUDA device: the PBS path has no CPU fallback")
        _check(rc, "tfhe_mgpu_create")
        self._h = h
        self.n_gpus = n_gpus
        self.devices = list(range(n_gpus)) if devices is None else list(devices)
        self._keys = weakref.WeakSet()

    def _ck(self, rc):
        if rc != 0:
            raise TfheError(rc, (lib().tfhe_mgpu_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            for k in list(self._keys):
                k.free()
            lib().tfhe_mgpu_destroy(self._h)
            self._h = None

    def launch_count(self) -> int:
        return sum(int(lib().tfhe_ctx_launch_count(lib().tfhe_mgpu_ctx(self._h, i))) for i in range(self.n_gpus))

    def upload_key(self, bsk, ksk) -> MultiGpuKey:
        bsk, ksk = np.ascontiguousarray(bsk, dtype=np.uint32), np.ascontiguousarray(ksk, dtype=np.uint32)
        h = C.c_void_p()
        self._ck(lib().tfhe_mgpu_bk_upload(self._h, bsk.ctypes.data, ksk.ctypes.data, C.byref(h)))
        return MultiGpuKey(self, h)

    def upload_key_bmmp(self, bsk3, ksk) -> MultiGpuKey:
        bsk3, ksk = np.ascontiguousarray(bsk3, dtype=np.uint32), np.ascontiguousarray(ksk, dtype=np.uint32)
        h = C.c_void_p()
        self._ck(lib().tfhe_mgpu_bk_upload_bmmp(self._h, bsk3.ctypes.data, ksk.ctypes.data, C.byref(h)))
        return MultiGpuKey(self, h)

    def _same_place(self, *xs):
        dev = {(x.device.index if _is_torch(x) else None) for x in xs if x is not None}
        if len(dev) != 1:
            raise ValueError("inputs and output must be all numpy arrays or all CUDA tensors on one device")
        d = dev.pop()
        if d is not None and d not in self.devices:
            raise ValueError(f"tensors on cuda:{d}, which is not one of this context's devices {self.devices}")

    def bootstrap(self, bk: MultiGpuKey, lwe_in, test_vectors, lut_idx=None, out=None):
        p = self.params
        lwe_in = _u32(lwe_in)
        tvs = np.ascontiguousarray(test_vectors, dtype=np.uint32)
        B, T = _rows(lwe_in, p.n + 1, "lwe_in"), _rows(tvs, p.N, "test_vectors")
        idx = None if lut_idx is None else np.ascontiguousarray(lut_idx, dtype=np.uint32)
        if idx is not None and idx.shape != (B,):
            raise ValueError(f"lut_idx: expected shape [{B}], got {list(idx.shape)}")
        if out is None:
            out = _like(lwe_in, (B, p.n + 1))
        else:
            _check_out(out, (B, p.n + 1), lwe_in)
        self._same_place(lwe_in, out)
        self._ck(lib().tfhe_mgpu_bootstrap_batch(self._h, bk._h, _ptr(lwe_in), tvs.ctypes.data, T, None if idx is None else idx.ctypes.data, B, _ptr(out)))
        return out

    def gate(self, bk: MultiGpuKey, ops, ct0, ct1, out=None):
        p = self.params
        ct0, ct1 = _u32(ct0), _u32(ct1)
        B = _rows(ct0, p.n + 1, "ct0")
        if _rows(ct1, p.n + 1, "ct1") != B:
            raise ValueError("ct0 and ct1 must hold the same number of ciphertexts")
        ops = np.full(B, int(ops), dtype=np.uint8) if isinstance(ops, (int, np.integer)) else np.ascontiguousarray(ops, dtype=np.uint8)
        if ops.shape != (B,):
            raise ValueError(f"ops: expected shape [{B}]")
        if out is None:
            out = _like(ct0, (B, p.n + 1))
        else:
            _check_out(out, (B, p.n + 1), ct0)
        self._same_place(ct0, ct1, out)
        self._ck(lib().tfhe_mgpu_gates_batch(self._h, bk._h, ops.ctypes.data, _ptr(ct0), _ptr(ct1), B, _ptr(out)))
        return out

    def last_timing(self):
        out = (C.c_double * 4)()
        self._ck(lib().tfhe_mgpu_last_timing(self._h, C.byref(out)))
        return {"scatter_ms": out[0], "compute_ms": out[1], "gather_ms": out[2], "total_ms": out[3]}
