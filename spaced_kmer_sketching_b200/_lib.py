"""ctypes binding of libsks.so (C ABI: include/sks.h).

The library is built in-tree by `make -C spaced_kmer_sketching_b200/csrc` (or
`__graft_entry__.build()`).  There is no CPU fallback: if the shared object is missing the import
fails, and compute calls fail with SKS_ERR_CUDA when no B200 is visible.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsks.so")

SKS_OK, SKS_ERR_INVALID, SKS_ERR_CUDA, SKS_ERR_CAPACITY, SKS_ERR_MISMATCH, SKS_ERR_IO = range(6)
PRED_ALL, PRED_FMH = 0, 1
HASH_BOOST_171, HASH_BOOST_181 = 171, 181
REPR_AUTO, REPR_SORTED, REPR_BITSET, REPR_BITSET_ONCHIP = 0, 1, 2, 3
KERNEL_KINDS = 15


class SksPred(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nonce", C.c_int32), ("modulus", C.c_uint64), ("hash_variant", C.c_int32),
                ("reserved", C.c_int32)]


class SksPairResult(C.Structure):
    _fields_ = [("size_a", C.c_int64), ("size_b", C.c_int64), ("intersection", C.c_int64), ("ani_ab", C.c_double),
                ("ani_ba", C.c_double)]


class SksError(RuntimeError):
    """Raised for any non-zero status of the C ABI (the C++ wrappers raise std::runtime_error)."""

    def __init__(self, code: int, message: str):
        super().__init__("libsks error %d: %s" % (code, message))
        self.code = code


# name -> (restype, argtypes); every symbol include/sks.h declares
vp, u64, i64, ci = C.c_void_p, C.c_uint64, C.c_int64, C.c_int
u64p, i64p, u32p, i32p = C.POINTER(u64), C.POINTER(i64), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
PROTOTYPES = {
    "sks_version": (ci, []),
    "sks_last_error": (C.c_char_p, []),
    "sks_device_count": (ci, []),
    "sks_ctx_create": (ci, [ci, C.POINTER(vp)]),
    "sks_ctx_destroy": (None, [vp]),
    "sks_ctx_set_stream": (ci, [vp, vp]),
    "sks_ctx_sync": (ci, [vp]),
    "sks_timer_begin": (ci, [vp]),
    "sks_timer_end": (ci, [vp, C.POINTER(C.c_float)]),
    "sks_ctx_launch_count": (i64, [vp]),
    "sks_ctx_in_place_count": (i64, [vp]),
    "sks_ctx_streamed_count": (i64, [vp]),
    "sks_ctx_profile": (ci, [vp, ci]),
    "sks_ctx_kernel_stats": (ci, [vp, ci, i64p, C.POINTER(C.c_double)]),
    "sks_kernel_name": (C.c_char_p, [ci]),
    "sks_seed_to_mask": (ci, [C.c_char_p, u64p, C.POINTER(ci)]),
    "sks_mask_weight": (ci, [u64p]),
    "sks_contiguous_mask": (ci, [ci, u64p]),
    "sks_random_mask": (ci, [ci, ci, u64, u64p]),
    "sks_reverse_bitset": (None, [u64p, u64p]),
    "sks_boost_hash_bitset": (u64, [u64p, ci]),
    "sks_fmh_hash": (u64, [u64p, u64p, ci, ci, ci]),
    "sks_containment": (C.c_double, [ci, ci]),
    "sks_binomial_estimator": (C.c_double, [C.c_double, ci]),
    "sks_packed_words": (C.c_size_t, [u64]),
    "sks_pack_codes": (ci, [vp, u64, vp]),
    "sks_unpack_codes": (ci, [vp, u64, vp]),
    "sks_fasta_parse": (ci, [C.c_char_p, C.c_size_t, u64p, u64p, vp, vp]),
    "sks_fasta_parse_file": (ci, [C.c_char_p, u64p, u64p, C.POINTER(vp), C.POINTER(vp)]),
    "sks_free": (None, [vp]),
    "sks_batch_upload": (ci, [vp, ci, C.POINTER(vp), u64p, C.POINTER(vp), u64p, C.POINTER(vp)]),
    "sks_batch_from_fasta_text": (ci, [vp, ci, C.POINTER(C.c_char_p), u64p, C.POINTER(vp)]),
    "sks_batch_from_fasta_files": (ci, [vp, ci, C.POINTER(C.c_char_p), C.POINTER(vp)]),
    "sks_batch_segments": (ci, [vp, ci, u64p, vp]),
    "sks_batch_synth": (ci, [vp, ci, u64, u64p, u64p, u64p, C.POINTER(vp)]),
    "sks_batch_synth_at": (ci, [vp, ci, u64, u64p, u64p, u64p, u64p, C.POINTER(vp)]),
    "sks_batch_slice": (ci, [vp, vp, ci, u64, u64, ci, C.POINTER(vp)]),
    "sks_batch_n_genomes": (ci, [vp]),
    "sks_batch_n_bases": (u64, [vp, ci]),
    "sks_batch_download": (ci, [vp, vp, ci, vp]),
    "sks_batch_destroy": (None, [vp, vp]),
    "sks_sketch": (ci, [vp, vp, u64p, ci, C.POINTER(SksPred), ci, C.POINTER(vp)]),
    "sks_kmer_list": (ci, [vp, vp, ci, u64p, ci, C.POINTER(SksPred), u64p, vp, vp, u64]),
    "sks_set_repr": (ci, [vp]),
    "sks_set_window": (ci, [vp]),
    "sks_set_weight": (ci, [vp]),
    "sks_set_size": (ci, [vp, vp, i64p]),
    "sks_set_keys": (ci, [vp, vp, vp, u64]),
    "sks_set_device_keys": (ci, [vp, vp, C.POINTER(vp), i64p, C.POINTER(ci)]),
    "sks_set_from_device_keys": (ci, [vp, vp, i64, ci, u64p, ci, C.POINTER(vp)]),
    "sks_sets_from_device_keys": (ci, [vp, vp, i64, i64p, ci, u64p, ci, C.POINTER(vp)]),
    "sks_set_from_unsorted_device_keys": (ci, [vp, vp, i64, ci, u64p, ci, C.POINTER(vp)]),
    "sks_set_from_host_keys": (ci, [vp, vp, i64, u64p, ci, C.POINTER(vp)]),
    "sks_set_device_index": (ci, [vp]),
    "sks_set_clone_to": (ci, [vp, vp, C.POINTER(vp)]),
    "sks_set_destroy": (None, [vp, vp]),
    "sks_set_save": (ci, [vp, vp, C.POINTER(SksPred), C.c_char_p]),
    "sks_set_load": (ci, [vp, C.c_char_p, C.POINTER(vp), C.POINTER(SksPred)]),
    "sks_intersect": (ci, [vp, vp, vp, i64p]),
    "sks_intersect_pairs": (ci, [vp, C.POINTER(vp), i64, C.POINTER(vp), i64, vp]),
    "sks_intersect_all_pairs": (ci, [vp, C.POINTER(vp), i64, i64, i64, vp]),
    "sks_intersect_block": (ci, [vp, C.POINTER(vp), i64, i64, i64, i64, i64, vp]),
    "sks_intersect_rects": (ci, [vp, C.POINTER(vp), i64, vp, i64, vp]),
    "sks_all_vs_all": (ci, [vp, C.POINTER(vp), i64, i64, i64, vp, vp, vp]),
    "sks_ani_from_counts": (None, [vp, vp, i64, ci, vp]),
    "sks_shard_range": (None, [i64, ci, ci, i64p, i64p]),
    "sks_comm_unique_id": (ci, [vp]),
    "sks_comm_init_rank": (ci, [vp, vp, ci, ci, C.POINTER(vp)]),
    "sks_comm_init_all": (ci, [C.POINTER(vp), ci, C.POINTER(vp)]),
    "sks_comm_destroy": (None, [vp]),
    "sks_comm_rank": (ci, [vp]),
    "sks_comm_world": (ci, [vp]),
    "sks_comm_nccl_version": (ci, []),
    "sks_comm_allgather_sets": (ci, [vp, vp, C.POINTER(vp), i64, i64, C.POINTER(vp)]),
    "sks_all_vs_all_sharded": (ci, [vp, vp, C.POINTER(vp), i64, i64, vp, vp, vp]),
    "sks_all_vs_all_resident": (ci, [vp, vp, vp, i64, u64p, ci, C.POINTER(SksPred), vp, vp, vp]),
    "sks_all_vs_all_from_host": (ci, [vp, vp, ci, C.POINTER(vp), u64p, i64, u64p, ci, C.POINTER(SksPred), vp, vp, vp]),
    "sks_sketch_sequence_sharded": (ci, [vp, vp, vp, u64p, ci, C.POINTER(SksPred), ci, C.POINTER(vp), i64p]),
    "sks_pair_ani": (ci, [vp, vp, u64, vp, u64, u64p, ci, C.POINTER(SksPred), ci, C.POINTER(SksPairResult)]),
    "sks_pair_ani_resident": (ci, [vp, vp, u64p, ci, C.POINTER(SksPred), ci, C.POINTER(SksPairResult)]),
}

_lib = None


def load():
    """Loads libsks.so; raises if it has not been built (the product has no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libsks.so is missing at %s: run `make -C spaced_kmer_sketching_b200/csrc` "
                              "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError if include/sks.h and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


_nccl_preloaded = False


def preload_nccl() -> None:
    """libsks binds NCCL at run time (csrc/sks_comm.cu: SKS_NCCL_LIB, else a libnccl.so.2 already in the process, else
    the system's).  A Python process that imports torch AFTER its first communicator would then hold the system's NCCL
    where libtorch_cuda.so expects the newer one of the `nvidia-nccl` wheel and fail to import; so the wheel's library,
    when there is one, is loaded first."""
    global _nccl_preloaded
    if _nccl_preloaded or os.environ.get("SKS_NCCL_LIB"):
        return
    _nccl_preloaded = True
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for loc in (spec.submodule_search_locations if spec else []):
            path = os.path.join(loc, "lib", "libnccl.so.2")
            if os.path.exists(path):
                C.CDLL(path, mode=C.RTLD_GLOBAL)
                return
    except Exception:      # no wheel, or it does not load here: libsks finds its own
        pass


def check(status: int) -> None:
    if status != SKS_OK:
        raise SksError(status, load().sks_last_error().decode("utf-8", "replace"))
