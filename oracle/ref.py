"""TEST INFRASTRUCTURE ONLY -- ctypes view of oracle/_ref/libref.so.

libref.so is the UNMODIFIED reference (/root/reference/src/*.cpp) compiled against
oracle/shim by oracle/Makefile, plus oracle/ref_glue.cpp.  Only tests/, bench.py's CPU
baseline legs and __graft_entry__.smoke() may import this module; the product never does.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref.so")
CLI_PATH = os.path.join(_HERE, "_ref", "ref_cli")

_lib = None

ALL = 0
FMH = 1


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        u64p = C.POINTER(C.c_uint64)
        L.ref_set_boost_variant.argtypes = [C.c_int]
        L.ref_get_boost_variant.restype = C.c_int
        L.ref_random_mask.argtypes = [C.c_int, C.c_int, C.c_uint64, u64p]
        L.ref_contiguous_mask.argtypes = [C.c_int, u64p]
        L.ref_contiguous_mask.restype = C.c_int
        L.ref_reverse_bitset.argtypes = [u64p, u64p]
        L.ref_fmh.argtypes = [C.c_int, C.c_int, u64p, u64p]
        L.ref_fmh.restype = C.c_uint64
        L.ref_boost_hash_bitset.argtypes = [u64p]
        L.ref_boost_hash_bitset.restype = C.c_uint64
        L.ref_canonical_kmer.argtypes = [C.c_int, u64p, u64p, u64p, u64p]
        L.ref_strings_from_fasta.argtypes = [C.c_char_p]
        L.ref_strings_from_fasta.restype = C.c_void_p
        L.ref_strings_from_codes.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_strings_from_codes.restype = C.c_void_p
        L.ref_strings_from_raw.argtypes = [C.c_char_p, C.c_int64]
        L.ref_strings_from_raw.restype = C.c_void_p
        L.ref_strings_count.argtypes = [C.c_void_p]
        L.ref_strings_count.restype = C.c_int64
        L.ref_string_len.argtypes = [C.c_void_p, C.c_int64]
        L.ref_string_len.restype = C.c_int64
        L.ref_string_copy.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.ref_strings_free.argtypes = [C.c_void_p]
        L.ref_kmers.argtypes = [C.c_void_p, u64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_kmers.restype = C.c_void_p
        L.ref_kmers_count.argtypes = [C.c_void_p]
        L.ref_kmers_count.restype = C.c_int64
        L.ref_kmers_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_kmers_free.argtypes = [C.c_void_p]
        L.ref_set_from_kmers.argtypes = [C.c_void_p]
        L.ref_set_from_kmers.restype = C.c_void_p
        L.ref_set_from_fasta.argtypes = [C.c_char_p, u64p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_set_from_fasta.restype = C.c_void_p
        L.ref_sets_from_fasta_files.argtypes = [C.c_int, C.POINTER(C.c_char_p), u64p, C.c_int, C.c_int,
                                                C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.ref_set_size.argtypes = [C.c_void_p]
        L.ref_set_size.restype = C.c_int
        L.ref_set_keys.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_set_free.argtypes = [C.c_void_p]
        L.ref_intersection.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_intersection.restype = C.c_int
        L.ref_pairwise_intersections.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p),
                                                 C.c_int, C.c_int, C.c_void_p]
        L.ref_pairwise_intersections.restype = C.c_int
        L.ref_all_pairs.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.ref_ring_pairs.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.ref_containment.argtypes = [C.c_int, C.c_int]
        L.ref_containment.restype = C.c_double
        L.ref_binomial_estimator.argtypes = [C.c_double, C.c_int]
        L.ref_binomial_estimator.restype = C.c_double
        L.ref_init()
        _lib = L
    return _lib


def _w2(v: int):
    return (C.c_uint64 * 2)(v & 0xFFFFFFFFFFFFFFFF, (v >> 64) & 0xFFFFFFFFFFFFFFFF)


def _from_w2(a) -> int:
    return int(a[0]) | (int(a[1]) << 64)


def set_boost_variant(v: int) -> None:
    lib().ref_set_boost_variant(v)


def random_mask(window: int, k: int, seed: int = 0) -> int:
    out = (C.c_uint64 * 2)()
    lib().ref_random_mask(window, k, seed, out)
    return _from_w2(out)


def contiguous_mask(k: int) -> Optional[int]:
    out = (C.c_uint64 * 2)()
    if lib().ref_contiguous_mask(k, out) != 0:
        return None
    return _from_w2(out)


def reverse_bitset(v: int) -> int:
    out = (C.c_uint64 * 2)()
    lib().ref_reverse_bitset(_w2(v), out)
    return _from_w2(out)


def fmh(nonce: int, window: int, masked: int, mask: int) -> int:
    return int(lib().ref_fmh(nonce, window, _w2(masked), _w2(mask)))


def boost_hash_bitset(v: int) -> int:
    return int(lib().ref_boost_hash_bitset(_w2(v)))


def canonical_kmer(window: int, bits: int, mask: int) -> Tuple[int, int]:
    ob, om = (C.c_uint64 * 2)(), (C.c_uint64 * 2)()
    lib().ref_canonical_kmer(window, _w2(bits), _w2(mask), ob, om)
    return _from_w2(ob), _from_w2(om)


class Strings:
    """Owns a reference std::vector<acgt_string>."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def from_fasta(cls, path: str) -> "Strings":
        return cls(lib().ref_strings_from_fasta(path.encode()))

    @classmethod
    def from_codes(cls, segs: Sequence[np.ndarray]) -> "Strings":
        lens = np.array([len(s) for s in segs], dtype=np.int64)
        codes = np.concatenate([np.asarray(s, dtype=np.uint8) for s in segs]) if len(segs) else np.zeros(0, np.uint8)
        codes = np.ascontiguousarray(codes)
        return cls(lib().ref_strings_from_codes(codes.ctypes.data, lens.ctypes.data, len(segs)))

    @classmethod
    def from_raw(cls, raw: bytes) -> "Strings":
        return cls(lib().ref_strings_from_raw(raw, len(raw)))

    def segments(self) -> List[np.ndarray]:
        out = []
        for i in range(lib().ref_strings_count(self.h)):
            n = lib().ref_string_len(self.h, i)
            a = np.empty(n, dtype=np.uint8)
            if n:
                lib().ref_string_copy(self.h, i, a.ctypes.data)
            out.append(a)
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_strings_free(self.h)
            self.h = None


def _keys_to_int(a: np.ndarray) -> List[int]:
    return [int(a[i, 0]) | (int(a[i, 1]) << 64) for i in range(a.shape[0])]


def kmers(strings: Strings, mask: int, window: int, pred=ALL, nonce=1, modulus=200, legacy=False):
    """Ordered list (duplicates kept) -> (masked[n,2], kmer_bits[n,2]) uint64 arrays."""
    h = lib().ref_kmers(strings.h, _w2(mask), window, pred, nonce, modulus, int(legacy))
    n = lib().ref_kmers_count(h)
    masked = np.empty((n, 2), dtype=np.uint64)
    bits = np.empty((n, 2), dtype=np.uint64)
    if n:
        lib().ref_kmers_copy(h, masked.ctypes.data, bits.ctypes.data)
    lib().ref_kmers_free(h)
    return masked, bits


class KmerSet:
    def __init__(self, handle):
        self.h = handle

    @classmethod
    def from_strings(cls, strings: Strings, mask: int, window: int, pred=ALL, nonce=1, modulus=200):
        kh = lib().ref_kmers(strings.h, _w2(mask), window, pred, nonce, modulus, 0)
        s = cls(lib().ref_set_from_kmers(kh))
        lib().ref_kmers_free(kh)
        return s

    @classmethod
    def from_fasta(cls, path: str, mask: int, window: int, pred=ALL, nonce=1, modulus=200):
        return cls(lib().ref_set_from_fasta(path.encode(), _w2(mask), window, pred, nonce, modulus))

    def size(self) -> int:
        return lib().ref_set_size(self.h)

    def keys(self) -> np.ndarray:
        """Sorted masked_bits, shape [n, 2] (lo, hi)."""
        a = np.empty((self.size(), 2), dtype=np.uint64)
        if a.shape[0]:
            lib().ref_set_keys(self.h, a.ctypes.data)
        return a

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_set_free(self.h)
            self.h = None


def sets_from_fasta_files(paths: Sequence[str], mask: int, window: int, pred=ALL, nonce=1, modulus=200,
                          parallel=True) -> List[KmerSet]:
    n = len(paths)
    arr = (C.c_char_p * n)(*[p.encode() for p in paths])
    out = (C.c_void_p * n)()
    lib().ref_sets_from_fasta_files(n, arr, _w2(mask), window, pred, nonce, modulus, int(parallel), out)
    return [KmerSet(out[i]) for i in range(n)]


def intersection(a: KmerSet, b: KmerSet) -> int:
    return lib().ref_intersection(a.h, b.h)


def pairwise_intersections(a: Sequence[KmerSet], b: Sequence[KmerSet], parallel=True) -> Optional[np.ndarray]:
    na, nb = len(a), len(b)
    pa = (C.c_void_p * max(na, 1))(*[x.h for x in a])
    pb = (C.c_void_p * max(nb, 1))(*[x.h for x in b])
    out = np.zeros(max(na, 1), dtype=np.int32)
    rc = lib().ref_pairwise_intersections(pa, na, pb, nb, int(parallel), out.ctypes.data)
    return None if rc != 0 else out[:na]


def all_pairs(n: int):
    f = np.zeros(n * n, dtype=np.int32)
    s = np.zeros(n * n, dtype=np.int32)
    lib().ref_all_pairs(n, f.ctypes.data, s.ctypes.data)
    return f, s


def ring_pairs(n: int):
    f = np.zeros(n, dtype=np.int32)
    s = np.zeros(n, dtype=np.int32)
    lib().ref_ring_pairs(n, f.ctypes.data, s.ctypes.data)
    return f, s


def containment(i: int, size: int) -> float:
    return lib().ref_containment(i, size)


def binomial_estimator(c: float, k: int) -> float:
    return lib().ref_binomial_estimator(c, k)
