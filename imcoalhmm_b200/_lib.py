"""ctypes binding of libimcoalhmm_b200.so (include/imcoalhmm_b200.h).

There is no CPU fallback: if the shared library is missing the import fails loudly, and every
forward call fails loudly when no B200 is usable.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IMC_LIB_PATH") or os.path.join(_HERE, "libimcoalhmm_b200.so")   # IMC_LIB_PATH: experiment builds


class IMCError(RuntimeError):
    """Raised for every non-zero return code of the C ABI."""

    def __init__(self, code, message):
        super().__init__("imcoalhmm_b200 error %d: %s" % (code, message))
        self.code = code


c_i32p = ctypes.POINTER(ctypes.c_int32)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_vp = ctypes.c_void_p

_SIGNATURES = {
    # name: (restype, argtypes)
    "imc_last_error": (ctypes.c_char_p, []),
    "imc_version": (ctypes.c_int, []),
    "imc_init": (ctypes.c_int, [ctypes.c_int]),
    "imc_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "imc_seq_create": (ctypes.c_int, [c_i32p, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "imc_seq_create_u8": (ctypes.c_int, [c_u8p, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "imc_seq_from_file": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "imc_seq_from_pair": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int64, ctypes.POINTER(c_vp)]),
    "imc_seq_from_fasta": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(c_vp)]),
    "imc_seq_from_columns": (ctypes.c_int, [ctypes.POINTER(ctypes.c_char_p), ctypes.c_int, ctypes.c_int64, ctypes.POINTER(c_vp)]),
    "imc_seq_from_fasta_n": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), ctypes.c_int, ctypes.POINTER(c_vp)]),
    "imc_seq_from_alignment": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), ctypes.c_int,
                                              ctypes.POINTER(c_vp)]),
    "imc_seq_save": (ctypes.c_int, [c_vp, ctypes.c_char_p]),
    "imc_seq_load": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(c_vp)]),
    "imc_seq_length": (ctypes.c_int, [c_vp, c_i64p]),
    "imc_seq_nsym": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_int)]),
    "imc_seq_symbol_counts": (ctypes.c_int, [c_vp, c_i64p]),
    "imc_seq_symbols": (ctypes.c_int, [c_vp, c_u8p, ctypes.c_int64]),
    "imc_seq_destroy": (ctypes.c_int, [c_vp]),
    "imc_seqset_create": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, ctypes.POINTER(c_vp)]),
    "imc_seqset_create_parts": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "imc_seqset_destroy": (ctypes.c_int, [c_vp]),
    "imc_seqset_info": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_int), c_i64p, c_i64p]),
    "imc_seqset_zip_info": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                           c_i64p, ctypes.POINTER(ctypes.c_int)]),
    "imc_seqset_zip_pairs": (ctypes.c_int, [c_vp, c_u8p, ctypes.c_int]),
    "imc_seqset_zip_tokens": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_u8p, ctypes.c_int64, c_i64p]),
    "imc_seqset_run_info": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                           ctypes.POINTER(ctypes.c_int), c_i64p, ctypes.POINTER(ctypes.c_int)]),
    "imc_seqset_run_pairs": (ctypes.c_int, [c_vp, c_u8p, ctypes.c_int]),
    "imc_seqset_run_tokens": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32), ctypes.c_int64,
                                             c_i64p, ctypes.POINTER(ctypes.c_int)]),
    "imc_seqset_spectral_counts": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "imc_forward": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_f64p, c_f64p, c_f64p, c_f64p]),
    "imc_forward_batch": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_f64p, c_f64p, c_f64p,
                                         c_f64p]),
    "imc_forward_batch_dev": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp, c_vp,
                                             c_vp]),
    "imc_comm_unique_id": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "imc_comm_init": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_vp]),
    "imc_comm_destroy": (ctypes.c_int, []),
    "imc_comm_info": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "imc_model_create": (ctypes.c_int, [ctypes.c_int, c_i32p, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "imc_model_info": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "imc_model_destroy": (ctypes.c_int, [c_vp]),
    "imc_model_build_batch": (ctypes.c_int, [c_vp, ctypes.c_int, c_f64p, c_f64p, c_f64p, c_f64p, c_i32p]),
    "imc_model_build_batch_dev": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "imc_break_points": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, c_f64p]),
    "imc_model_break_points": (ctypes.c_int, [c_vp, ctypes.c_int, c_f64p, c_f64p]),
    "imc_loglik_batch": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_f64p, c_f64p, c_i32p]),
    "imc_loglik_batch_dev": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp]),
    "imc_statespace_describe": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                               ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), c_i32p, c_i32p,
                                               c_u8p]),
    "imc_measure_fp64_peak": (ctypes.c_int, [c_f64p, c_f64p]),
    "imc_set_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int64]),
    "imc_get_option": (ctypes.c_int, [ctypes.c_char_p, c_i64p]),
    "imc_kernel_launches": (ctypes.c_int64, []),
    "imc_mma_passes": (ctypes.c_int, [c_i64p]),
    "imc_seqset_align_info": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_i64p, c_f64p, c_i64p, c_i32p, c_i32p]),
    "imc_seqset_align_quad": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32), ctypes.c_int64,
                                             c_i64p, c_i32p]),
    "imc_last_forward_kernel": (ctypes.c_char_p, []),
}

EXPORTS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load the shared library (once).  Raises ImportError with the build hint when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not found: build it with `python imcoalhmm_b200/build.py` "
                              "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            if os.environ.get("IMC_LIB_PATH") and not hasattr(lib, name):
                continue              # an older experiment build (tools/zip_bench.py A/B runs): the call fails when it is made
            fn = getattr(lib, name)   # AttributeError here == header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise IMCError(rc, load().imc_last_error().decode("utf-8", "replace"))


def set_option(key, value):
    check(load().imc_set_option(key.encode(), int(value)))


def get_option(key):
    v = ctypes.c_int64()
    check(load().imc_get_option(key.encode(), ctypes.byref(v)))
    return v.value


def measure_fp64_peak():
    """(dfma_tflops, dmma_tflops) measured on the bound device."""
    a, b = ctypes.c_double(), ctypes.c_double()
    check(load().imc_measure_fp64_peak(ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


def comm_unique_id():
    """128 opaque bytes drawn by rank 0; hand them to every other rank (file, pipe, MPI, torch.distributed...)."""
    buf = ctypes.create_string_buffer(128)
    check(load().imc_comm_unique_id(buf, 128))
    return buf.raw


def comm_init(nranks, rank, unique_id=None):
    """After imc_init(device): from now on every forward / likelihood call returns the sum over ranks (one all-reduce)."""
    check(load().imc_comm_init(int(nranks), int(rank), ctypes.c_char_p(unique_id) if unique_id is not None else None))


def comm_destroy():
    check(load().imc_comm_destroy())


def comm_info():
    """dict(nranks, rank, fused): fused = the all-reduce runs inside the reduction kernel over peer memory (NVLink)."""
    n, r, f = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(load().imc_comm_info(ctypes.byref(n), ctypes.byref(r), ctypes.byref(f)))
    return {"nranks": n.value, "rank": r.value, "fused": bool(f.value)}


def mma_passes():
    v = ctypes.c_int64()
    check(load().imc_mma_passes(ctypes.byref(v)))
    return v.value


def kernel_launches():
    return int(load().imc_kernel_launches())


def last_forward_kernel():
    return load().imc_last_forward_kernel().decode()
