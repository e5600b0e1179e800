"""TEST INFRASTRUCTURE ONLY -- ctypes view of the plain-C oracle (oracle/oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  The product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

ALL = 0
FMH = 1

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, i64, u64, ci = C.c_void_p, C.c_int64, C.c_uint64, C.c_int
        L.orc_boost_hash_bitset.argtypes = [u64, u64, ci]
        L.orc_boost_hash_bitset.restype = u64
        L.orc_fmh.argtypes = [u64, u64, u64, u64, ci, ci, ci]
        L.orc_fmh.restype = u64
        L.orc_fasta_parse.argtypes = [C.c_char_p, i64, vp, C.POINTER(i64), vp, C.POINTER(i64)]
        L.orc_split_raw.argtypes = [C.c_char_p, i64, vp, C.POINTER(i64), vp, C.POINTER(i64)]
        L.orc_kmers.argtypes = [vp, vp, i64, u64, u64, ci, ci, ci, u64, ci, vp, vp, i64]
        L.orc_kmers.restype = i64
        L.orc_reverse_bitset.argtypes = [u64, u64, C.POINTER(u64)]
        L.orc_legacy_canonical.argtypes = [ci, u64, u64, u64, u64, C.POINTER(u64), C.POINTER(u64)]
        L.orc_contiguous_mask.argtypes = [ci, C.POINTER(u64)]
        L.orc_contiguous_mask.restype = ci
        L.orc_random_mask.argtypes = [ci, ci, u64, C.POINTER(u64)]
        L.orc_sort_unique.argtypes = [vp, i64]
        L.orc_sort_unique.restype = i64
        L.orc_intersection.argtypes = [vp, i64, vp, i64]
        L.orc_intersection.restype = i64
        L.orc_containment.argtypes = [ci, ci]
        L.orc_containment.restype = C.c_double
        L.orc_binomial_estimator.argtypes = [C.c_double, ci]
        L.orc_binomial_estimator.restype = C.c_double
        L.orc_all_pairs.argtypes = [ci, vp, vp]
        L.orc_ring_pairs.argtypes = [ci, vp, vp]
        L.orc_gen.argtypes = [i64, u64, vp]
        L.orc_mutate.argtypes = [vp, i64, u64, u64, vp]
        _lib = L
    return _lib


M64 = (1 << 64) - 1


def _lo(v: int) -> int:
    return v & M64


def _hi(v: int) -> int:
    return (v >> 64) & M64


# ---- masks ---------------------------------------------------------------------------------
def seed_to_mask(seed: str) -> Tuple[int, int]:
    """Seed string (README notation, left = first base of the window) -> (128-bit mask, window).

    s[i] == '1' sets bits 2(w-1-i) and 2(w-1-i)+1."""
    w = len(seed)
    m = 0
    for i, ch in enumerate(seed):
        if ch == "1":
            m |= 3 << (2 * (w - 1 - i))
        elif ch != "0":
            raise ValueError("seed strings hold only '0' and '1'")
    return m, w


def mask_to_seed(mask: int, window: int) -> str:
    return "".join("1" if (mask >> (2 * (window - 1 - i))) & 3 else "0" for i in range(window))


def mask_weight(mask: int) -> int:
    return bin(mask).count("1") // 2


def random_mask(window: int, k: int, seed: int = 0) -> int:
    out = (C.c_uint64 * 2)()
    lib().orc_random_mask(window, k, seed, out)
    return int(out[0]) | (int(out[1]) << 64)


def contiguous_mask(k: int) -> Optional[int]:
    out = (C.c_uint64 * 2)()
    if lib().orc_contiguous_mask(k, out) != 0:
        return None
    return int(out[0]) | (int(out[1]) << 64)


def reverse_bitset(v: int) -> int:
    out = (C.c_uint64 * 2)()
    lib().orc_reverse_bitset(_lo(v), _hi(v), out)
    return int(out[0]) | (int(out[1]) << 64)


def legacy_canonical(window: int, bits: int, mask: int) -> Tuple[int, int]:
    ob, om = (C.c_uint64 * 2)(), (C.c_uint64 * 2)()
    lib().orc_legacy_canonical(window, _lo(bits), _hi(bits), _lo(mask), _hi(mask), ob, om)
    return int(ob[0]) | (int(ob[1]) << 64), int(om[0]) | (int(om[1]) << 64)


# ---- hashes --------------------------------------------------------------------------------
def boost_hash_bitset(v: int, variant: int = 181) -> int:
    return int(lib().orc_boost_hash_bitset(_lo(v), _hi(v), variant))


def fmh(masked: int, mask: int, window: int, nonce: int = 1, variant: int = 181) -> int:
    return int(lib().orc_fmh(_lo(masked), _hi(masked), _lo(mask), _hi(mask), window, nonce, variant))


# ---- FASTA ---------------------------------------------------------------------------------
def fasta_parse(text: bytes) -> Tuple[np.ndarray, np.ndarray]:
    """File bytes -> (codes uint8[n], seg_len int64[s]) with the reference's record/split rules."""
    nc, ns = C.c_int64(), C.c_int64()
    lib().orc_fasta_parse(text, len(text), None, C.byref(nc), None, C.byref(ns))
    codes = np.empty(nc.value, dtype=np.uint8)
    segs = np.empty(ns.value, dtype=np.int64)
    lib().orc_fasta_parse(text, len(text), codes.ctypes.data, C.byref(nc), segs.ctypes.data, C.byref(ns))
    return codes, segs


def split_raw(raw: bytes) -> Tuple[np.ndarray, np.ndarray]:
    nc, ns = C.c_int64(), C.c_int64()
    lib().orc_split_raw(raw, len(raw), None, C.byref(nc), None, C.byref(ns))
    codes = np.empty(nc.value, dtype=np.uint8)
    segs = np.empty(ns.value, dtype=np.int64)
    lib().orc_split_raw(raw, len(raw), codes.ctypes.data, C.byref(nc), segs.ctypes.data, C.byref(ns))
    return codes, segs


# ---- the hot loop --------------------------------------------------------------------------
def kmers(codes: np.ndarray, seg_len: Sequence[int], mask: int, window: int, pred: int = ALL, nonce: int = 1,
          modulus: int = 200, variant: int = 181, want_bits: bool = False):
    """Ordered, duplicate-preserving canonical k-mer list: masked[n,2] (and kmer_bits[n,2])."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    segs = np.ascontiguousarray(seg_len, dtype=np.int64)
    args = (codes.ctypes.data, segs.ctypes.data, len(segs), _lo(mask), _hi(mask), window, pred, nonce, modulus,
            variant)
    n = lib().orc_kmers(*args, None, None, 0)
    masked = np.empty((n, 2), dtype=np.uint64)
    bits = np.empty((n, 2), dtype=np.uint64) if want_bits else None
    lib().orc_kmers(*args, masked.ctypes.data, bits.ctypes.data if want_bits else None, n)
    return (masked, bits) if want_bits else masked


def sort_unique(keys: np.ndarray) -> np.ndarray:
    k = np.ascontiguousarray(keys, dtype=np.uint64).copy().reshape(-1, 2)
    m = lib().orc_sort_unique(k.ctypes.data, k.shape[0])
    return k[:m].copy()


def sketch_set(codes, seg_len, mask, window, pred=ALL, nonce=1, modulus=200, variant=181) -> np.ndarray:
    """kmer_set as ascending distinct masked_bits, shape [m, 2] (lo, hi)."""
    return sort_unique(kmers(codes, seg_len, mask, window, pred, nonce, modulus, variant))


def intersection(a: np.ndarray, b: np.ndarray) -> int:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    return int(lib().orc_intersection(a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0]))


def containment(i: int, size: int) -> float:
    return lib().orc_containment(i, size)


def binomial_estimator(c: float, k: int) -> float:
    return lib().orc_binomial_estimator(c, k)


def ani(intersection_count: int, first_set_size: int, weight: int) -> float:
    """src/kmer-sketching.cpp:196-200: containment on the FIRST set, then ^(1/weight)."""
    return binomial_estimator(containment(intersection_count, first_set_size), weight)


def all_pairs(n: int):
    f = np.zeros(n * n, dtype=np.int32)
    s = np.zeros(n * n, dtype=np.int32)
    lib().orc_all_pairs(n, f.ctypes.data, s.ctypes.data)
    return f, s


def ring_pairs(n: int):
    f = np.zeros(n, dtype=np.int32)
    s = np.zeros(n, dtype=np.int32)
    lib().orc_ring_pairs(n, f.ctypes.data, s.ctypes.data)
    return f, s


# ---- synthetic genomes -------------------------------------------------
def gen(L: int, seed: int) -> np.ndarray:
    out = np.empty(L, dtype=np.uint8)
    lib().orc_gen(L, seed, out.ctypes.data)
    return out


def mutate(seq: np.ndarray, seed: int, D: int) -> np.ndarray:
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    out = np.empty_like(seq)
    lib().orc_mutate(seq.ctypes.data, len(seq), seed, D, out.ctypes.data)
    return out


def codes_to_text(codes: np.ndarray) -> bytes:
    return np.frombuffer(b"ACGT", dtype=np.uint8)[np.asarray(codes, dtype=np.uint8)].tobytes()


def write_fasta(path: str, codes: np.ndarray, name: str = "seq", width: int = 80) -> None:
    """Single-record FASTA, `width` columns, LF line ends."""
    text = codes_to_text(codes)
    with open(path, "wb") as f:
        f.write(b">" + name.encode() + b"\n")
        for i in range(0, len(text), width):
            f.write(text[i:i + width] + b"\n")
