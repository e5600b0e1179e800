"""Python host mirror of the reference's sketch-and-compare API on top of libsks.so.

The product's host side is C++ (include/kmer.hpp ... over the C ABI in include/sks.h); this module
is the same surface for tests, bench.py and multi-GPU drivers, with the reference's names:

    generate_random_spaced_seed_mask / contiguous_kmer        src/kmer_bitset.cpp:51-56,132-152
    kmer_set_from_fasta_file / kmer_sets_from_fasta_files     src/kmer_set.cpp:54-133
    nucleotide_string_list_to_kmers                           src/kmer_sliding.cpp:224-238
    kmer_set_intersection / compute_pairwise_...              src/kmer_set.cpp:23-41,143-184
    containment / binomial_estimator                          src/ani_estimation.cpp:24-42
    generate_all_pairs_from_vector / generate_pairwise_...    src/generators.hpp:20-58

All compute runs in the CUDA kernels of libsks.so; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import (HASH_BOOST_171, HASH_BOOST_181, PRED_ALL, PRED_FMH, REPR_AUTO, REPR_BITSET, REPR_BITSET_ONCHIP, REPR_SORTED, SksError,
                   SksPairResult, SksPred, check)

M64 = (1 << 64) - 1


def _w2(v: int):
    return (C.c_uint64 * 2)(v & M64, (v >> 64) & M64)


def _from_w2(a) -> int:
    return int(a[0]) | (int(a[1]) << 64)


# ---- predicates -------------------------------------------------------------------------------
class Predicate:
    """The recognised forms of the reference's `std::function<bool(const kmer)>` sketching condition."""

    def __init__(self, kind: int, nonce: int = 1, modulus: int = 200, hash_variant: int = HASH_BOOST_181):
        self.kind, self.nonce, self.modulus, self.hash_variant = kind, nonce, modulus, hash_variant

    def c(self) -> SksPred:
        return SksPred(self.kind, self.nonce, self.modulus, self.hash_variant, 0)


def all_kmers() -> Predicate:
    return Predicate(PRED_ALL)


def frac_min_hash(nonce: int = 1, modulus: int = 200, hash_variant: int = HASH_BOOST_181) -> Predicate:
    """`frac_min_hash(nonce)(k) % modulus == 0` (src/kmer-sketching.cpp:29-34)."""
    return Predicate(PRED_FMH, nonce, modulus, hash_variant)


# ---- masks and host helpers -------------------------------------------------------------------
def seed_to_mask(seed: str) -> Tuple[int, int]:
    out, w = (C.c_uint64 * 2)(), C.c_int()
    check(_lib.load().sks_seed_to_mask(seed.encode(), out, C.byref(w)))
    return _from_w2(out), w.value


def mask_weight(mask: int) -> int:
    return _lib.load().sks_mask_weight(_w2(mask))


def contiguous_kmer(k: int) -> int:
    out = (C.c_uint64 * 2)()
    check(_lib.load().sks_contiguous_mask(k, out))
    return _from_w2(out)


def generate_random_spaced_seed_mask(window_size: int, kmer_size: int, random_seed: int = 0) -> int:
    out = (C.c_uint64 * 2)()
    check(_lib.load().sks_random_mask(window_size, kmer_size, random_seed, out))
    return _from_w2(out)


def reverse_kmer_bitset(v: int) -> int:
    out = (C.c_uint64 * 2)()
    _lib.load().sks_reverse_bitset(_w2(v), out)
    return _from_w2(out)


def fmh_hash(masked: int, mask: int, window: int, nonce: int = 1, hash_variant: int = HASH_BOOST_181) -> int:
    return int(_lib.load().sks_fmh_hash(_w2(masked), _w2(mask), window, nonce, hash_variant))


def containment(intersection: int, set_size: int) -> float:
    return _lib.load().sks_containment(intersection, set_size)


def binomial_estimator(c: float, kmer_num_ones: int) -> float:
    return _lib.load().sks_binomial_estimator(c, kmer_num_ones)


def ani_from_counts(intersections: np.ndarray, first_sizes: np.ndarray, weight: int) -> np.ndarray:
    a = np.ascontiguousarray(intersections, dtype=np.int32).ravel()
    b = np.ascontiguousarray(first_sizes, dtype=np.int32).ravel()
    out = np.empty(a.shape[0], dtype=np.float64)
    _lib.load().sks_ani_from_counts(a.ctypes.data, b.ctypes.data, a.shape[0], weight, out.ctypes.data)
    return out


def generate_all_pairs_from_vector(v: Sequence):
    """All n^2 ordered pairs, row-major (src/generators.hpp:44-58)."""
    return [a for a in v for _ in v], [b for _ in v for b in v]


def generate_pairwise_from_vector(v: Sequence):
    """Ring pairs (i, i+1 mod n) (src/generators.hpp:20-33)."""
    n = len(v)
    return list(v), [v[(i + 1) % n] for i in range(n)]


def pack_codes(codes: np.ndarray) -> np.ndarray:
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    L = _lib.load()
    out = np.zeros(L.sks_packed_words(len(codes)), dtype=np.uint32)
    check(L.sks_pack_codes(codes.ctypes.data, len(codes), out.ctypes.data))
    return out


def unpack_codes(words: np.ndarray, n_bases: int) -> np.ndarray:
    words = np.ascontiguousarray(words, dtype=np.uint32)
    out = np.empty(n_bases, dtype=np.uint8)
    check(_lib.load().sks_unpack_codes(words.ctypes.data, n_bases, out.ctypes.data))
    return out


def fasta_parse(text: bytes) -> Tuple[np.ndarray, int, np.ndarray]:
    """FASTA bytes -> (packed words, n_bases, segment lengths) with the reference's rules."""
    L = _lib.load()
    nb, ns = C.c_uint64(), C.c_uint64()
    check(L.sks_fasta_parse(text, len(text), C.byref(nb), C.byref(ns), None, None))
    words = np.zeros(L.sks_packed_words(nb.value), dtype=np.uint32)
    segs = np.zeros(ns.value, dtype=np.uint64)
    check(L.sks_fasta_parse(text, len(text), C.byref(nb), C.byref(ns), words.ctypes.data, segs.ctypes.data))
    return words, nb.value, segs


def fasta_parse_file(path: str) -> Tuple[np.ndarray, int, np.ndarray]:
    L = _lib.load()
    nb, ns, pw, ps = C.c_uint64(), C.c_uint64(), C.c_void_p(), C.c_void_p()
    check(L.sks_fasta_parse_file(path.encode(), C.byref(nb), C.byref(ns), C.byref(pw), C.byref(ps)))
    try:
        nw = L.sks_packed_words(nb.value)
        words = np.ctypeslib.as_array(C.cast(pw, C.POINTER(C.c_uint32)), shape=(max(nw, 1),))[:nw].copy()
        segs = np.ctypeslib.as_array(C.cast(ps, C.POINTER(C.c_uint64)), shape=(max(ns.value, 1),))[:ns.value].copy()
    finally:
        L.sks_free(pw)
        L.sks_free(ps)
    return words, nb.value, segs


# ---- device objects ---------------------------------------------------------------------------
class Context:
    """One per (process, GPU): stream, scratch, event timers."""

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        h = C.c_void_p()
        check(self._L.sks_ctx_create(device, C.byref(h)))
        self.h, self.device = h, device

    def close(self):
        if self.h:
            self._L.sks_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        check(self._L.sks_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def sync(self):
        check(self._L.sks_ctx_sync(self.h))

    def timer_begin(self):
        check(self._L.sks_timer_begin(self.h))

    def timer_end(self) -> float:
        ms = C.c_float()
        check(self._L.sks_timer_end(self.h, C.byref(ms)))
        return ms.value

    @property
    def launches(self) -> int:
        return int(self._L.sks_ctx_launch_count(self.h))

    @property
    def in_place_calls(self) -> int:
        """pair_ani calls that read the genomes in place from pinned host buffers (no host-to-device copy)."""
        return int(self._L.sks_ctx_in_place_count(self.h))

    @property
    def streamed_calls(self) -> int:
        """all_vs_all_from_host calls whose pinned host genomes were copied chunk by chunk under the sketch kernel."""
        return int(self._L.sks_ctx_streamed_count(self.h))

    def profile(self, enable: bool):
        check(self._L.sks_ctx_profile(self.h, int(enable)))

    def kernel_stats(self) -> dict:
        """{kernel name: (launches, total ms)} since the last query (synchronises)."""
        out = {}
        for kind in range(_lib.KERNEL_KINDS):
            n, ms = C.c_int64(), C.c_double()
            check(self._L.sks_ctx_kernel_stats(self.h, kind, C.byref(n), C.byref(ms)))
            if n.value:
                out[self._L.sks_kernel_name(kind).decode()] = (n.value, ms.value)
        return out

    # -- batches
    def upload(self, genomes: Sequence[Tuple[np.ndarray, int, Optional[np.ndarray]]]) -> "Batch":
        """genomes: (packed uint32 words, n_bases, segment lengths or None) per genome, HOST arrays."""
        n = len(genomes)
        words = [np.ascontiguousarray(g[0], dtype=np.uint32) for g in genomes]
        segs = [None if g[2] is None else np.ascontiguousarray(g[2], dtype=np.uint64) for g in genomes]
        pw = (C.c_void_p * max(n, 1))(*[w.ctypes.data for w in words])
        nb = (C.c_uint64 * max(n, 1))(*[int(g[1]) for g in genomes])
        ps = (C.c_void_p * max(n, 1))(*[None if s is None or len(s) == 0 else s.ctypes.data for s in segs])
        ns = (C.c_uint64 * max(n, 1))(*[0 if s is None else len(s) for s in segs])
        h = C.c_void_p()
        check(self._L.sks_batch_upload(self.h, n, pw, nb, ps, ns, C.byref(h)))
        return Batch(self, h)

    def upload_codes(self, seqs: Sequence[np.ndarray], seg_lens: Optional[Sequence] = None) -> "Batch":
        return self.upload([(pack_codes(s), len(s), None if seg_lens is None else seg_lens[i])
                            for i, s in enumerate(seqs)])

    def batch_from_fasta_text(self, texts: Sequence[bytes]) -> "Batch":
        """Device-side FASTA ingest: raw file bytes are uploaded and parsed / packed by kernels."""
        n = len(texts)
        ptr = (C.c_char_p * max(n, 1))(*texts)
        ln = (C.c_uint64 * max(n, 1))(*[len(t) for t in texts])
        h = C.c_void_p()
        check(self._L.sks_batch_from_fasta_text(self.h, n, ptr, ln, C.byref(h)))
        return Batch(self, h)

    def batch_from_fasta_files(self, paths: Sequence[str]) -> "Batch":
        n = len(paths)
        ptr = (C.c_char_p * max(n, 1))(*[p.encode() for p in paths])
        h = C.c_void_p()
        check(self._L.sks_batch_from_fasta_files(self.h, n, ptr, C.byref(h)))
        return Batch(self, h)

    def synth(self, n_bases: int, gen_seed: Sequence[int], mut_seed: Sequence[int], mut_D: Sequence[int]) -> "Batch":
        n = len(gen_seed)
        a = (C.c_uint64 * max(n, 1))(*gen_seed)
        b = (C.c_uint64 * max(n, 1))(*mut_seed)
        d = (C.c_uint64 * max(n, 1))(*mut_D)
        h = C.c_void_p()
        check(self._L.sks_batch_synth(self.h, n, n_bases, a, b, d, C.byref(h)))
        return Batch(self, h)

    def synth_at(self, n_bases: int, first_base: Sequence[int], gen_seed: Sequence[int], mut_seed: Sequence[int],
                 mut_D: Sequence[int]) -> "Batch":
        n = len(gen_seed)
        f = (C.c_uint64 * max(n, 1))(*first_base)
        a = (C.c_uint64 * max(n, 1))(*gen_seed)
        b = (C.c_uint64 * max(n, 1))(*mut_seed)
        d = (C.c_uint64 * max(n, 1))(*mut_D)
        h = C.c_void_p()
        check(self._L.sks_batch_synth_at(self.h, n, n_bases, f, a, b, d, C.byref(h)))
        return Batch(self, h)

    # -- sketching
    def sketch(self, batch: "Batch", mask: int, window: int, pred: Predicate, repr_: int = REPR_AUTO) -> List["KmerSet"]:
        n = batch.n_genomes
        out = (C.c_void_p * max(n, 1))()
        p = pred.c()
        check(self._L.sks_sketch(self.h, batch.h, _w2(mask), window, C.byref(p), repr_, out))
        return [KmerSet(self, C.c_void_p(out[i])) for i in range(n)]

    def kmer_list(self, batch: "Batch", genome: int, mask: int, window: int, pred: Predicate):
        """Ordered list with duplicates: (masked[n,2], kmer_bits[n,2]) uint64 (lo, hi)."""
        p = pred.c()
        n = C.c_uint64()
        m2 = _w2(mask)
        check(self._L.sks_kmer_list(self.h, batch.h, genome, m2, window, C.byref(p), C.byref(n), None, None, 0))
        masked = np.zeros((n.value, 2), dtype=np.uint64)
        bits = np.zeros((n.value, 2), dtype=np.uint64)
        if n.value:
            check(self._L.sks_kmer_list(self.h, batch.h, genome, m2, window, C.byref(p), C.byref(n), masked.ctypes.data,
                                        bits.ctypes.data, n.value))
        return masked, bits

    def set_from_device_keys(self, dptr: int, n_keys: int, words_per_key: int, mask: int, window: int,
                             sorted_unique: bool = True) -> "KmerSet":
        h = C.c_void_p()
        fn = self._L.sks_set_from_device_keys if sorted_unique else self._L.sks_set_from_unsorted_device_keys
        check(fn(self.h, C.c_void_p(dptr), n_keys, words_per_key, _w2(mask), window, C.byref(h)))
        return KmerSet(self, h)

    def sets_from_device_keys(self, dptr: int, counts: Sequence[int], words_per_key: int, mask: int,
                              window: int) -> List["KmerSet"]:
        """Sets whose sorted keys lie back to back at `dptr` (one device copy, shared buffer)."""
        n = len(counts)
        cnt = (C.c_int64 * max(n, 1))(*[int(c) for c in counts])
        out = (C.c_void_p * max(n, 1))()
        check(self._L.sks_sets_from_device_keys(self.h, C.c_void_p(dptr), n, cnt, words_per_key, _w2(mask), window, out))
        return [KmerSet(self, C.c_void_p(out[i])) for i in range(n)]

    def load_set(self, path: str) -> Tuple["KmerSet", Predicate]:
        """Reads a sketch file written by KmerSet.save."""
        h, p = C.c_void_p(), SksPred()
        check(self._L.sks_set_load(self.h, path.encode(), C.byref(h), C.byref(p)))
        return KmerSet(self, h), Predicate(p.kind, p.nonce, p.modulus, p.hash_variant)

    # -- comparison
    def intersect(self, a: "KmerSet", b: "KmerSet") -> int:
        out = C.c_int64()
        check(self._L.sks_intersect(self.h, a.h, b.h, C.byref(out)))
        return out.value

    def intersect_pairs(self, a: Sequence["KmerSet"], b: Sequence["KmerSet"]) -> np.ndarray:
        pa = (C.c_void_p * max(len(a), 1))(*[s.h for s in a])
        pb = (C.c_void_p * max(len(b), 1))(*[s.h for s in b])
        out = np.zeros(max(len(a), 1), dtype=np.int32)
        check(self._L.sks_intersect_pairs(self.h, pa, len(a), pb, len(b), out.ctypes.data))
        return out[:len(a)]

    def intersect_all_pairs(self, sets: Sequence["KmerSet"], row_begin: int = 0, row_end: Optional[int] = None,
                            out: Optional[np.ndarray] = None) -> np.ndarray:
        n = len(sets)
        row_end = n if row_end is None else row_end
        ps = (C.c_void_p * max(n, 1))(*[s.h for s in sets])
        if out is None:
            out = np.zeros((n, n), dtype=np.int32)
        check(self._L.sks_intersect_all_pairs(self.h, ps, n, row_begin, row_end, out.ctypes.data))
        return out

    def all_vs_all(self, sets: Sequence["KmerSet"], row_begin: int = 0, row_end: Optional[int] = None, want_ani: bool = True):
        """Rows [row_begin, row_end) of the all-pairs matrix in one device-resident pass: (counts[rows, n] int32,
        sizes[n] int32, ani[rows, n] float64 or None).  The comparison phase of the reference driver
        (src/kmer-sketching.cpp:185-200)."""
        n = len(sets)
        row_end = n if row_end is None else row_end
        rows = max(row_end - row_begin, 0)
        ps = (C.c_void_p * max(n, 1))(*[s.h for s in sets])
        counts = np.zeros((rows, n), dtype=np.int32)
        sizes = np.zeros(n, dtype=np.int32)
        ani = np.zeros((rows, n), dtype=np.float64) if want_ani else None
        check(self._L.sks_all_vs_all(self.h, ps, n, row_begin, row_end, counts.ctypes.data, sizes.ctypes.data,
                                     ani.ctypes.data if want_ani else None))
        return counts, sizes, ani

    # -- several GPUs (comm None: one rank)
    def allgather_sets(self, comm: Optional["Comm"], local_sets: Sequence["KmerSet"], n_total: int) -> List["KmerSet"]:
        ps = (C.c_void_p * max(len(local_sets), 1))(*[s.h for s in local_sets])
        out = (C.c_void_p * max(n_total, 1))()
        check(self._L.sks_comm_allgather_sets(self.h, comm.h if comm else None, ps, len(local_sets), n_total, out))
        return [KmerSet(self, C.c_void_p(out[i])) for i in range(n_total)]

    def all_vs_all_sharded(self, comm: Optional["Comm"], local_sets: Sequence["KmerSet"], n_total: int, want_ani: bool = True,
                           out: Optional[tuple] = None):
        """This rank's block rows of the all-pairs matrix: (counts[rows, n] int32, sizes[n] int32, ani[rows, n] float64).
        `out`: (counts, sizes, ani) arrays to fill (e.g. views of pinned memory)."""
        rank, world = (comm.rank, comm.world) if comm else (0, 1)
        b, e = shard_range(n_total, rank, world)
        if out is None:
            out = (np.zeros((e - b, n_total), dtype=np.int32), np.zeros(n_total, dtype=np.int32),
                   np.zeros((e - b, n_total), dtype=np.float64) if want_ani else None)
        counts, sizes, ani = out
        ps = (C.c_void_p * max(len(local_sets), 1))(*[s.h for s in local_sets])
        check(self._L.sks_all_vs_all_sharded(self.h, comm.h if comm else None, ps, len(local_sets), n_total,
                                             counts.ctypes.data, sizes.ctypes.data, ani.ctypes.data if ani is not None else None))
        return counts, sizes, ani

    def all_vs_all_resident(self, comm: Optional["Comm"], batch: Optional["Batch"], n_total: int, mask: int, window: int,
                            pred: Predicate, out: tuple):
        """Sketch + exchange + comparison for the rank's resident genomes; `out` = (counts, sizes, ani) arrays to fill."""
        counts, sizes, ani = out
        p = pred.c()
        check(self._L.sks_all_vs_all_resident(self.h, comm.h if comm else None, batch.h if batch is not None else None, n_total,
                                              _w2(mask), window, C.byref(p), counts.ctypes.data, sizes.ctypes.data,
                                              ani.ctypes.data if ani is not None else None))
        return counts, sizes, ani

    def all_vs_all_from_host(self, comm: Optional["Comm"], ptrs: Sequence[int], n_bases: Sequence[int], n_total: int, mask: int,
                             window: int, pred: Predicate, out: tuple):
        """HOST packed genomes (raw pointers, e.g. into pinned memory) -> this rank's rows; `out` = (counts, sizes, ani)."""
        counts, sizes, ani = out
        n = len(ptrs)
        pp = (C.c_void_p * max(n, 1))(*ptrs)
        nb = (C.c_uint64 * max(n, 1))(*n_bases)
        p = pred.c()
        check(self._L.sks_all_vs_all_from_host(self.h, comm.h if comm else None, n, pp, nb, n_total, _w2(mask), window, C.byref(p),
                                               counts.ctypes.data, sizes.ctypes.data, ani.ctypes.data if ani is not None else None))
        return counts, sizes, ani

    def sketch_sequence_sharded(self, comm: Optional["Comm"], slice_batch: "Batch", mask: int, window: int, pred: Predicate,
                                gather: bool = True) -> Tuple["KmerSet", int]:
        """One long sequence split by position: (this rank's key range of the global set, or the whole set when
        `gather`; size of the global set)."""
        h, n, p = C.c_void_p(), C.c_int64(), pred.c()
        check(self._L.sks_sketch_sequence_sharded(self.h, comm.h if comm else None, slice_batch.h, _w2(mask), window, C.byref(p),
                                                  int(gather), C.byref(h), C.byref(n)))
        return KmerSet(self, h), n.value

    def intersect_block(self, sets: Sequence["KmerSet"], rows: Tuple[int, int], cols: Tuple[int, int],
                        out: np.ndarray) -> np.ndarray:
        """out[i, j] = |sets[i] n sets[j]| for i in rows, j in cols (half-open ranges); other entries untouched."""
        n = len(sets)
        ps = (C.c_void_p * max(n, 1))(*[s.h for s in sets])
        check(self._L.sks_intersect_block(self.h, ps, n, rows[0], rows[1], cols[0], cols[1], out.ctypes.data))
        return out

    def intersect_rects(self, sets: Sequence["KmerSet"], rects: Sequence[Tuple[Tuple[int, int], Tuple[int, int]]],
                        out: np.ndarray) -> np.ndarray:
        """Several (rows, cols) rectangles of the all-pairs matrix in one pass (one pair table, one launch)."""
        n = len(sets)
        ps = (C.c_void_p * max(n, 1))(*[s.h for s in sets])
        flat = np.array([[r[0][0], r[0][1], r[1][0], r[1][1]] for r in rects], dtype=np.int64).reshape(-1)
        check(self._L.sks_intersect_rects(self.h, ps, n, flat.ctypes.data, len(rects), out.ctypes.data))
        return out

    def pair_ani(self, packed_a: np.ndarray, n_a: int, packed_b: np.ndarray, n_b: int, mask: int, window: int,
                 pred: Predicate, repr_: int = REPR_AUTO) -> SksPairResult:
        r, p = SksPairResult(), pred.c()
        check(self._L.sks_pair_ani(self.h, packed_a.ctypes.data, n_a, packed_b.ctypes.data, n_b, _w2(mask), window,
                                   C.byref(p), repr_, C.byref(r)))
        return r

    def pair_ani_ptr(self, ptr_a: int, n_a: int, ptr_b: int, n_b: int, mask: int, window: int, pred: Predicate,
                     repr_: int = REPR_AUTO) -> SksPairResult:
        """Same with raw host pointers (e.g. pinned torch tensors)."""
        r, p = SksPairResult(), pred.c()
        check(self._L.sks_pair_ani(self.h, C.c_void_p(ptr_a), n_a, C.c_void_p(ptr_b), n_b, _w2(mask), window,
                                   C.byref(p), repr_, C.byref(r)))
        return r

    def pair_ani_resident(self, batch: "Batch", mask: int, window: int, pred: Predicate,
                          repr_: int = REPR_AUTO) -> SksPairResult:
        r, p = SksPairResult(), pred.c()
        check(self._L.sks_pair_ani_resident(self.h, batch.h, _w2(mask), window, C.byref(p), repr_, C.byref(r)))
        return r


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the genomes (and pair-matrix rows) of `rank`: contiguous blocks of ceil(n / world)."""
    b, e = C.c_int64(), C.c_int64()
    _lib.load().sks_shard_range(n, rank, world, C.byref(b), C.byref(e))
    return b.value, e.value


COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """The 128-byte NCCL id rank 0 makes; the launcher carries it to the other ranks (multi_gpu.init_comm)."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _lib.preload_nccl()
    check(_lib.load().sks_comm_unique_id(buf))
    return buf.raw


class Comm:
    """One rank's communicator (C ABI: sks_comm_*).  `None` everywhere a Comm is expected means a single rank."""

    def __init__(self, ctx: "Context", comm_id: bytes, rank: int, world: int):
        _lib.preload_nccl()
        self._L = _lib.load()
        self.ctx, self.rank, self.world = ctx, rank, world
        h = C.c_void_p()
        check(self._L.sks_comm_init_rank(ctx.h, C.create_string_buffer(comm_id, COMM_ID_BYTES), rank, world, C.byref(h)))
        self.h = h

    @classmethod
    def init_all(cls, ctxs: Sequence["Context"]) -> List["Comm"]:
        """One process, one context per GPU: communicators for all of them (sks_comm_init_all); the sharded calls are
        then made from one thread per context."""
        _lib.preload_nccl()
        L = _lib.load()
        n = len(ctxs)
        hs = (C.c_void_p * n)(*[c.h for c in ctxs])
        out = (C.c_void_p * n)()
        check(L.sks_comm_init_all(hs, n, out))
        comms = []
        for r in range(n):
            c = cls.__new__(cls)
            c._L, c.ctx, c.rank, c.world, c.h = L, ctxs[r], r, n, C.c_void_p(out[r])
            comms.append(c)
        return comms

    def close(self):
        if getattr(self, "h", None):
            self._L.sks_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """Genomes resident in HBM: 2-bit packed bases plus segment tables."""

    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h

    @property
    def n_genomes(self) -> int:
        return self.ctx._L.sks_batch_n_genomes(self.h)

    def n_bases(self, genome: int) -> int:
        return int(self.ctx._L.sks_batch_n_bases(self.h, genome))

    def segments(self, genome: int) -> np.ndarray:
        """Lengths of the ACGT runs of one genome."""
        n = C.c_uint64()
        check(self.ctx._L.sks_batch_segments(self.h, genome, C.byref(n), None))
        out = np.zeros(max(n.value, 1), dtype=np.uint64)
        check(self.ctx._L.sks_batch_segments(self.h, genome, C.byref(n), out.ctypes.data))
        return out[:n.value]

    def download(self, genome: int) -> np.ndarray:
        out = np.zeros(self.ctx._L.sks_packed_words(self.n_bases(genome)), dtype=np.uint32)
        check(self.ctx._L.sks_batch_download(self.ctx.h, self.h, genome, out.ctypes.data))
        return out

    def slice(self, genome: int, first_base: int, n_starts: int, window: int) -> "Batch":
        h = C.c_void_p()
        check(self.ctx._L.sks_batch_slice(self.ctx.h, self.h, genome, first_base, n_starts, window, C.byref(h)))
        return Batch(self.ctx, h)

    def close(self):
        if self.h and self.ctx.h:
            self.ctx._L.sks_batch_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class KmerSet:
    """Device-resident `kmer_set` (src/kmer.hpp:160-190)."""

    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h

    @property
    def repr(self) -> int:
        return self.ctx._L.sks_set_repr(self.h)

    @property
    def weight(self) -> int:
        return self.ctx._L.sks_set_weight(self.h)

    @property
    def window(self) -> int:
        return self.ctx._L.sks_set_window(self.h)

    def kmer_set_size(self) -> int:
        out = C.c_int64()
        check(self.ctx._L.sks_set_size(self.ctx.h, self.h, C.byref(out)))
        return out.value

    def keys(self) -> np.ndarray:
        """Ascending distinct masked_bits as uint64 [n, 2] = (lo, hi)."""
        n = self.kmer_set_size()
        out = np.zeros((n, 2), dtype=np.uint64)
        check(self.ctx._L.sks_set_keys(self.ctx.h, self.h, out.ctypes.data, n))
        return out

    def save(self, path: str, pred: Optional[Predicate] = None) -> None:
        """Writes the set as a sketch file (include/sks.h: sks_set_save)."""
        p = pred.c() if pred is not None else None
        check(self.ctx._L.sks_set_save(self.ctx.h, self.h, C.byref(p) if p is not None else None, path.encode()))

    def device_keys(self) -> Tuple[int, int, int]:
        p, n, kw = C.c_void_p(), C.c_int64(), C.c_int()
        check(self.ctx._L.sks_set_device_keys(self.ctx.h, self.h, C.byref(p), C.byref(n), C.byref(kw)))
        return (p.value or 0), n.value, kw.value

    def close(self):
        if self.h and self.ctx.h:
            self.ctx._L.sks_set_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- the reference's free functions -------------------------------------------------------------
def kmer_sets_from_fasta_files(ctx: Context, fasta_filenames: Sequence[str], mask: int, window_length: int,
                               sketching_cond: Predicate, repr_: int = REPR_AUTO) -> List[KmerSet]:
    """(parallel_)kmer_sets_from_fasta_files, src/kmer_set.cpp:81-133: one batched launch."""
    batch = ctx.batch_from_fasta_files(list(fasta_filenames))   # raw bytes up, parse + split + pack on the device
    try:
        return ctx.sketch(batch, mask, window_length, sketching_cond, repr_)
    finally:
        batch.close()


def kmer_set_from_fasta_file(ctx: Context, fasta_filename: str, mask: int, window_length: int,
                             sketching_cond: Predicate, repr_: int = REPR_AUTO) -> KmerSet:
    return kmer_sets_from_fasta_files(ctx, [fasta_filename], mask, window_length, sketching_cond, repr_)[0]


def kmer_set_intersection(ctx: Context, a: KmerSet, b: KmerSet) -> int:
    return ctx.intersect(a, b)


def compute_pairwise_kmer_set_intersections(ctx: Context, v1: Sequence[KmerSet], v2: Sequence[KmerSet]) -> np.ndarray:
    """src/kmer_set.cpp:143-184; a length mismatch raises (the reference throws std::runtime_error)."""
    return ctx.intersect_pairs(v1, v2)
