"""ctypes loader for ``libmp2gpu.so`` (the C ABI of ``include/mp2gpu.h``).

There is no CPU fallback: if the shared library is missing, cannot be loaded, or reports no CUDA
device, every entry point raises :class:`Mp2GpuError`.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MP2GPU_LIB") or os.path.join(HERE, "libmp2gpu.so")  # override: tuning variants

u64p = C.POINTER(C.c_uint64)
u64pp = C.POINTER(u64p)
size_p = C.POINTER(C.c_size_t)
u32p = C.POINTER(C.c_uint32)


class Mp2GpuError(RuntimeError):
    """An error string returned across the C ABI (where the Rust shim would panic / bail!)."""


# every exported symbol of include/mp2gpu.h: name -> (restype, argtypes)
_ERR = C.c_void_p  # const char* that we must free ourselves
SIGNATURES = {
    "mp2gpu_init": (_ERR, [C.c_int]),
    "mp2gpu_device_count": (_ERR, [C.POINTER(C.c_int)]),
    "mp2gpu_free_string": (None, [C.c_void_p]),
    "mp2gpu_version": (C.c_char_p, []),
    "mp2gpu_host_alloc": (_ERR, [C.POINTER(C.c_void_p), C.c_size_t]),
    "mp2gpu_host_free": (_ERR, [C.c_void_p]),
    "mp2gpu_commit_from_values": (_ERR, [u64pp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                         u64pp, u64p, u64p, u64p, C.POINTER(C.c_void_p)]),
    "mp2gpu_commit_from_coeffs": (_ERR, [u64pp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                         u64pp, u64p, u64p, u64p, C.POINTER(C.c_void_p)]),
    "mp2gpu_merkle_new": (_ERR, [u64p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32, u64p, u64p]),
    "mp2gpu_merkle_new_ragged": (_ERR, [u64pp, size_p, C.c_size_t, C.c_uint32, C.c_uint32, u64p, u64p]),
    "mp2gpu_merkle_prove": (_ERR, [u64p, C.c_size_t, C.c_uint32, C.c_size_t, u64p, size_p]),
    "mp2gpu_hash_no_pad_batch": (_ERR, [u64p, C.c_size_t, C.c_size_t, C.c_uint32, u64p]),
    "mp2gpu_two_to_one_batch": (_ERR, [u64p, u64p, C.c_size_t, C.c_uint32, u64p]),
    "mp2gpu_permute_batch": (_ERR, [u64p, C.c_size_t, C.c_uint32]),
    "mp2gpu_batch_fetch_rows": (_ERR, [C.c_void_p, u64p, C.c_size_t, u64p]),
    "mp2gpu_batch_prove": (_ERR, [C.c_void_p, C.c_size_t, u64p, size_p]),
    "mp2gpu_batch_open": (_ERR, [C.c_void_p, u64p, C.c_size_t, u64p, u64p]),
    "mp2gpu_fri_open_layer": (_ERR, [C.c_void_p, C.c_uint32, u64p, C.c_size_t, u64p, u64p]),
    "mp2gpu_batch_fetch": (_ERR, [C.c_void_p, u64pp, u64p, u64p, u64p]),
    "mp2gpu_batch_eval": (_ERR, [C.c_void_p, u64p, C.c_size_t, u64p]),
    "mp2gpu_batch_shape": (_ERR, [C.c_void_p, size_p, u32p, u32p, u32p, u32p]),
    "mp2gpu_batch_free": (None, [C.c_void_p]),
    "mp2gpu_fri_begin": (_ERR, [u64p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "mp2gpu_fri_begin_openings": (_ERR, [C.POINTER(C.c_void_p), C.c_size_t, u64p, u32p, C.c_size_t, u32p, u32p, u64p,
                                         C.c_uint32, C.c_uint32, u64p, C.POINTER(C.c_void_p)]),
    "mp2gpu_fri_commit_layer": (_ERR, [C.c_void_p, C.c_uint32, u64p]),
    "mp2gpu_fri_fold": (_ERR, [C.c_void_p, u64p]),
    "mp2gpu_fri_layer_shape": (_ERR, [C.c_void_p, C.c_uint32, size_p, size_p, size_p, size_p]),
    "mp2gpu_fri_fetch_layer": (_ERR, [C.c_void_p, C.c_uint32, u64p, u64p, u64p]),
    "mp2gpu_fri_finish": (_ERR, [C.c_void_p, u64p, size_p]),
    "mp2gpu_fri_free": (None, [C.c_void_p]),
    "mp2gpu_fri_proof_of_work": (_ERR, [u64p, C.c_uint32, C.c_uint32, C.c_uint32, u64p]),
    "mp2gpu_dev_intt": (_ERR, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_void_p]),
    "mp2gpu_dev_coset_lde": (_ERR, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32,
                                    C.c_uint32, C.c_uint32, C.c_size_t, C.c_void_p]),
    "mp2gpu_dev_coset_lde_peer": (_ERR, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_size_t, C.c_size_t,
                                         C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "mp2gpu_dev_merkle_colmajor": (_ERR, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mp2gpu_dev_merkle_colmajor_leaves": (_ERR, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32,
                                                 C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mp2gpu_dev_merkle_levels": (_ERR, [C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mp2gpu_comm_init": (_ERR, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "mp2gpu_comm_free": (None, [C.c_void_p]),
    "mp2gpu_commit_from_values_sharded": (_ERR, [C.c_void_p, u64pp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32,
                                                 C.c_uint32, C.c_int, u64pp, u64p, u64p, u64p]),
    "mp2gpu_dev_merkle_rowmajor": (_ERR, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "mp2gpu_dev_commit": (_ERR, [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mp2gpu_dev_canonicalize": (_ERR, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mp2gpu_quotient_polys": (_ERR, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, u64p, u64p, u64p, u64p, C.c_uint32,
                                     C.c_uint32, C.c_uint32, u64pp, u64p, u64p, u64p, C.POINTER(C.c_void_p)]),
    "mp2gpu_partial_products_and_zs": (_ERR, [C.c_void_p, C.c_void_p, C.c_void_p, u64p, u64p, C.c_uint32, C.c_uint32,
                                              C.c_uint32, u64pp, u64p, u64p, u64p, C.POINTER(C.c_void_p)]),
    "mp2gpu_prove": (_ERR, [C.c_void_p, C.c_void_p, C.c_void_p, u64p, u64pp, u64p, C.c_size_t, u64p,
                            C.POINTER(C.c_void_p), size_p]),
    "mp2gpu_free_bytes": (None, [C.c_void_p]),
    "mp2gpu_transcript_permute": (_ERR, [u64p, C.c_uint32]),
    "mp2gpu_transcript_observe": (_ERR, [u64p, u64p, u32p, u64p, C.c_size_t, C.c_uint32, u32p]),
    "mp2gpu_trim": (_ERR, []),
    "mp2gpu_sync": (_ERR, [C.c_void_p]),
    "mp2gpu_profile_enable": (_ERR, [C.c_int]),
    "mp2gpu_profile_report": (_ERR, [C.c_char_p, C.c_size_t]),
    "mp2gpu_debug_int_pipe_peak": (_ERR, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mp2gpu_debug_field_selftest": (_ERR, [u64p, C.c_size_t]),
    "mp2gpu_debug_dft": (_ERR, [u64p, C.c_uint32, C.c_size_t]),
    "mp2gpu_debug_field_probe": (_ERR, [C.POINTER(C.c_double)]),
    "mp2gpu_launch_count": (C.c_uint64, []),
}

_lib = None


def load() -> C.CDLL:
    """Loads the library (never builds it implicitly -- ``__graft_entry__.build()`` does that)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Mp2GpuError("libmp2gpu.so not built (%s missing): run `python -m mapreduce_plonky2_b200.build`; "
                              "there is no CPU fallback" % LIB_PATH)
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise Mp2GpuError("cannot load %s: %s (no CPU fallback)" % (LIB_PATH, e)) from e
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                if os.environ.get("MP2GPU_LIB"):  # an older tuning variant: its missing entry points just cannot be called
                    continue
                raise
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(err) -> None:
    """NULL = success; otherwise copy + free the C string and raise."""
    if err:
        msg = C.string_at(err).decode("utf-8", "replace")
        _lib.mp2gpu_free_string(err)
        raise Mp2GpuError(msg)


def call(name: str, *args) -> None:
    lib = load()
    check(getattr(lib, name)(*args))
