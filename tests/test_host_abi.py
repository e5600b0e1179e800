"""CPU-only checks of libsks.so: the library loads, exports every symbol include/sks.h declares, and
its host-side helpers (masks, Boost hash, FASTA ingest, packing, ANI) agree with the oracle and the
golden fixtures.  No compute entry point is called here (there is no GPU in this suite)."""
import os
import re

import numpy as np
import pytest

import spaced_kmer_sketching_b200 as sks
from oracle import port

from conftest import ROOT, load_golden


def ival(h):
    return int(h, 16)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sks.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sks_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 40
    lib = sks.load()
    for name in sorted(declared):
        assert hasattr(lib, name), "libsks.so does not export " + name
    assert declared == set(sks.PROTOTYPES), declared ^ set(sks.PROTOTYPES)
    assert lib.sks_version() == 100


def test_no_cpu_fallback_without_device():
    if sks.load().sks_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(sks.SksError) as e:
        sks.Context(0)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_seed_masks():
    assert sks.seed_to_mask("11001011") == (0xF0CF, 8)
    assert sks.mask_weight(0xF0CF) == 5
    for s in ("", "1" * 65, "10x1"):
        with pytest.raises(sks.SksError):
            sks.seed_to_mask(s)
    g = load_golden("masks.json")
    for k, h in g["contiguous"].items():
        assert sks.contiguous_kmer(int(k)) == ival(h)
    with pytest.raises(sks.SksError):  # src/kmer_bitset.cpp:53-54
        sks.contiguous_kmer(65)
    for w, k, seed, h in g["random"]:
        assert sks.generate_random_spaced_seed_mask(w, k, seed) == ival(h)
    for w, k, h in g["driver_sweep"]:
        assert sks.generate_random_spaced_seed_mask(w, k) == ival(h)
    for v, r in g["reverse"]:
        assert sks.reverse_kmer_bitset(ival(v)) == ival(r)


def test_fmh_hash_golden():
    g = load_golden("hash.json")
    for variant, nonce, w, masked, mask, h in g["fmh"]:
        assert sks.fmh_hash(ival(masked), ival(mask), w, nonce, variant) == ival(h)


def test_fasta_parser_golden():
    g = load_golden("fasta.json")
    for name, case in g.items():
        text = case["text"].encode("latin1")
        words, nb, segs = sks.fasta_parse(text)
        codes = sks.unpack_codes(words, nb)
        got, o = [], 0
        for L in segs:
            got.append("".join(str(int(c)) for c in codes[o:o + int(L)]))
            o += int(L)
        assert got == [s for s in case["segments"] if s], name
        oc, os_ = port.fasta_parse(text)
        assert np.array_equal(oc, codes)


def test_fasta_parse_file(tmp_path):
    p = tmp_path / "x.fna"
    p.write_bytes(b">r1\nACGTAC\nGTNNAC\n\nGGGG\n>r2\nacgtRYac\n")
    words, nb, segs = sks.fasta_parse_file(str(p))
    assert nb == 20 and list(segs) == [8, 2, 4, 4, 2]
    with pytest.raises(sks.SksError) as e:
        sks.fasta_parse_file(str(tmp_path / "missing.fna"))
    assert e.value.code == 5


def test_pack_roundtrip():
    rng = np.random.default_rng(0)
    for n in (0, 1, 15, 16, 17, 1000, 4099):
        codes = rng.integers(0, 4, n, dtype=np.uint8)
        words = sks.pack_codes(codes)
        assert len(words) == (n + 15) // 16
        assert np.array_equal(sks.unpack_codes(words, n), codes)
        for i in (0, n // 2, n - 1):
            if n:
                assert (int(words[i // 16]) >> (2 * (i % 16))) & 3 == codes[i]


def test_ani_math():
    assert sks.containment(0, 0) == 0.0 and sks.containment(0, 7) == 0.0
    assert sks.binomial_estimator(0.0, 5) == 0.0 and sks.binomial_estimator(-0.5, 5) == 0.0
    assert sks.binomial_estimator(sks.containment(84848, 99975), 16) == port.ani(84848, 99975, 16)
    out = sks.ani_from_counts(np.array([544, 0, 4244791]), np.array([544, 5, 4994572]), 16)
    assert list(out) == [port.ani(544, 544, 16), 0.0, port.ani(4244791, 4994572, 16)]


def test_pair_generators():
    assert sks.generate_all_pairs_from_vector([0, 1, 2]) == ([0, 0, 0, 1, 1, 1, 2, 2, 2], [0, 1, 2, 0, 1, 2, 0, 1, 2])
    assert sks.generate_pairwise_from_vector([0, 1, 2, 3]) == ([0, 1, 2, 3], [1, 2, 3, 0])
