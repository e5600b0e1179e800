"""Pins the CPU oracle (oracle/oracle.c) against the fixtures in tests/golden/, which were produced
by the UNMODIFIED reference sources (tests/golden/make_golden.py), and against the survey's
known answers.  CPU only."""
import hashlib
import os

import numpy as np
import pytest

from oracle import port

from conftest import load_golden

CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def ival(h):
    return int(h, 16)


def to_int(a):
    return [int(x[0]) | (int(x[1]) << 64) for x in a]


def key_digest(keys):
    return hashlib.sha256(np.ascontiguousarray(keys, dtype="<u8").tobytes()).hexdigest()


def test_readme_example():
    # README.md:35-41: seed 11001011 over ACGTACGT picks A C . . A . G T -> ACAGT
    mask, w = port.seed_to_mask("11001011")
    assert (mask, w) == (0xF0CF, 8) and port.mask_weight(mask) == 5
    codes = np.array([CODE[c] for c in "ACGTACGT"], dtype=np.uint8)
    (m,) = to_int(port.kmers(codes, [8], mask, w))
    picked = "".join("ACGT"[(m >> (2 * (w - 1 - i))) & 3] for i in range(w) if "11001011"[i] == "1")
    assert picked == "ACAGT"


def test_kat1_table():
    mask, w = port.seed_to_mask("11001011")
    codes = np.array([CODE[c] for c in "AAACGTACGTTT"], dtype=np.uint8)
    masked, bits = port.kmers(codes, [12], mask, w, want_bits=True)
    assert to_int(masked) == [0x81, 0xC6, 0x100B, 0xC6, 0x81]
    assert [port.fmh(m, mask, w, 1, 171) % 200 for m in to_int(masked)] == [34, 41, 79, 41, 34]
    assert [port.fmh(m, mask, w, 1, 181) % 200 for m in to_int(masked)] == [106, 50, 56, 50, 106]
    assert port.boost_hash_bitset(mask, 171) == 0x25BA93A0BC4B4856
    assert port.boost_hash_bitset(mask, 181) == 0xB0E7CEA2795E7D39
    assert port.boost_hash_bitset(0, 171) == 0x6BA3C4643770594A
    assert port.boost_hash_bitset(0, 181) == 0xFD13631E6FAECB1A


def test_kat2_seed_strings():
    table = {(8, 5): "00101111", (20, 10): "00110010111001010011", (22, 12): "0001110010111001011011",
             (24, 16): "011101110010111110011011", (26, 16): "01011101110010111110010011",
             (31, 21): "0011111011010111111011001011101", (38, 28): "11111011101101111111101100101101110110",
             (50, 40): "01111101111111010111111101101111111111101100111011"}
    for (w, k), s in table.items():
        assert port.mask_to_seed(port.random_mask(w, k, 0), w) == s
    assert port.mask_to_seed(port.random_mask(24, 16, 1), 24) == "110101101111001011111001"
    for k in (1, 10, 33, 64):
        assert port.mask_to_seed(port.random_mask(k, k, 0), k) == "1" * k


def test_masks_golden():
    g = load_golden("masks.json")
    for k, h in g["contiguous"].items():
        assert port.contiguous_mask(int(k)) == ival(h)
    assert g["contiguous_65_throws"] and port.contiguous_mask(65) is None
    for w, k, seed, h in g["random"]:
        assert port.random_mask(w, k, seed) == ival(h), (w, k, seed)
    for w, k, h in g["driver_sweep"]:
        assert port.random_mask(w, k, 0) == ival(h)
    for v, r in g["reverse"]:
        assert port.reverse_bitset(ival(v)) == ival(r)


def test_hash_golden():
    g = load_golden("hash.json")
    for variant, v, h in g["bitset_hash"]:
        assert port.boost_hash_bitset(ival(v), variant) == ival(h)
    for variant, nonce, w, masked, mask, h in g["fmh"]:
        assert port.fmh(ival(masked), ival(mask), w, nonce, variant) == ival(h)


def test_fasta_golden():
    g = load_golden("fasta.json")
    for name, case in g.items():
        codes, segs = port.fasta_parse(case["text"].encode("latin1"))
        got, o = [], 0
        for L in segs:
            got.append("".join(str(int(c)) for c in codes[o:o + L]))
            o += L
        assert got == case["segments"], name
    # the four survey fixtures, as segment lengths
    lens = lambda t: list(port.fasta_parse(t)[1])
    assert lens(b">r1\nACGTAC\nGTNNAC\n\nGGGG\n>r2\nacgtRYac\n") == [8, 2, 4, 4, 2]
    assert lens(b">r1\r\nACGT\r\nACGT\r\n") == [4, 4]
    assert lens(b">\nACGT\n>ok\nAC GT\nAAAA\n>ok2\nTTTT\n") == [4]
    assert lens(b"ACGT\n>x\n\nCCCC\n") == [4]


def test_kmer_lists_golden():
    g = load_golden("kmer_lists.json")
    seqs = {k: np.array([CODE[c] for c in v], dtype=np.uint8) for k, v in g["sequences"].items()}
    for case in g["cases"]:
        mask, w = port.seed_to_mask(case["seed"])
        codes = seqs[case["seq"]]
        masked, bits = port.kmers(codes, case["segs"], mask, w, want_bits=True)
        assert to_int(masked) == [ival(x) for x in case["masked"]], (case["seq"], case["seed"], case["segs"])
        assert to_int(bits) == [ival(x) for x in case["bits"]], (case["seq"], case["seed"], case["segs"])
        # the legacy route (src/kmers.cpp:16-35) gives the same canonical k-mers
        for b, m in zip(case["bits"][:8], case["masked"][:8]):
            lb, lm = port.legacy_canonical(w, ival(b), mask)
            assert lm == ival(m)
        for variant in (171, 181):
            fm = port.kmers(codes, case["segs"], mask, w, port.FMH, 1, 4, variant)
            assert to_int(fm) == [ival(x) for x in case["fmh4_%d" % variant]]


@pytest.mark.parametrize("name", ["sets_100k.json", "sets_5m.json"])
def test_sets_golden(name):
    if not os.path.exists(os.path.join(os.path.dirname(__file__), "golden", name)):
        pytest.skip(name + " not generated")
    g = load_golden(name)
    A = port.gen(g["L"], g["gen_seed"])
    B = port.mutate(A, g["mut_seed"], g["D"])
    assert port.codes_to_text(A[:32]).decode() == g["prefix"]
    assert int((A != B).sum()) == g["hamming"]
    for case in g["cases"]:
        if g["L"] > 10 ** 6 and case["pred"] == "ALL" and len(case["seed"]) > 8:
            continue  # 5M-key qsort per set: covered on the GPU side; keep the CPU suite short
        mask, w = port.seed_to_mask(case["seed"])
        pred = port.ALL if case["pred"] == "ALL" else port.FMH
        sa = port.sketch_set(A, [len(A)], mask, w, pred, case["nonce"], case["modulus"], case["variant"])
        sb = port.sketch_set(B, [len(B)], mask, w, pred, case["nonce"], case["modulus"], case["variant"])
        I = port.intersection(sa, sb)
        assert (len(sa), len(sb), I) == (case["size_a"], case["size_b"], case["intersection"]), case["seed"]
        assert key_digest(sa) == case["digest_a"] and key_digest(sb) == case["digest_b"]
        wt = port.mask_weight(mask)
        assert port.ani(I, len(sa), wt) == float(case["ani_ab"])
        assert port.ani(I, len(sb), wt) == float(case["ani_ba"])


def test_multi_golden():
    g = load_golden("multi.json")
    base = port.gen(g["L"], g["base_seed"])
    genomes = [base if D == 0 else port.mutate(base, 2000 + i, D) for i, D in enumerate(g["Ds"])]
    for case in g["cases"]:
        mask, w = port.seed_to_mask(case["seed"])
        pred = port.ALL if case["pred"] == "ALL" else port.FMH
        sets = [port.sketch_set(x, [len(x)], mask, w, pred, case["nonce"], case["modulus"], case["variant"])
                for x in genomes]
        assert [len(s) for s in sets] == case["sizes"]
        assert [key_digest(s) for s in sets] == case["digests"]
        f, s = port.all_pairs(len(sets))
        ints = [port.intersection(sets[i], sets[j]) for i, j in zip(f, s)]
        assert ints == case["intersections"]
        wt = port.mask_weight(mask)
        assert [port.ani(I, len(sets[i]), wt) for I, i in zip(ints, f)] == [float(x) for x in case["ani"]]


def test_pair_generators():
    f, s = port.all_pairs(3)
    assert list(f) == [0, 0, 0, 1, 1, 1, 2, 2, 2] and list(s) == [0, 1, 2, 0, 1, 2, 0, 1, 2]
    f, s = port.ring_pairs(4)
    assert list(f) == [0, 1, 2, 3] and list(s) == [1, 2, 3, 0]


def test_strand_symmetry():
    rng = np.random.default_rng(3)
    codes = rng.integers(0, 4, 3000, dtype=np.uint8)
    rc = (3 - codes[::-1]).astype(np.uint8)
    for seed in ("11001011", "011101110010111110011011", "1" * 33):
        mask, w = port.seed_to_mask(seed)
        assert np.array_equal(port.sketch_set(codes, [3000], mask, w), port.sketch_set(rc, [3000], mask, w))


def test_ani_edge_cases():
    assert port.containment(0, 0) == 0.0 and port.containment(0, 10) == 0.0
    assert port.binomial_estimator(0.0, 5) == 0.0 and port.binomial_estimator(-1.0, 5) == 0.0
    assert port.ani(544, 544, 5) == 1.0
    assert abs(port.ani(84848, 99975, 16) - 0.989798718772253) < 1e-15


def test_against_reference_build_random():
    """When oracle/_ref is present (it is git-ignored), fuzz the oracle against it."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(17)
    for trial in range(40):
        w = int(rng.integers(1, 65))
        k = int(rng.integers(1, w + 1))
        mask = ref.random_mask(w, k, trial)
        nseg = int(rng.integers(1, 5))
        segs = [rng.integers(0, 4, int(rng.integers(0, 400)), dtype=np.uint8) for _ in range(nseg)]
        segs = [s for s in segs if len(s)] or [np.zeros(3, np.uint8)]
        codes = np.concatenate(segs)
        lens = [len(s) for s in segs]
        variant = 171 if trial % 2 else 181
        ref.set_boost_variant(variant)
        nonce = int(rng.integers(-3, 5))
        modulus = int(rng.integers(1, 9))
        S = ref.Strings.from_codes(segs)
        for pred in (ref.ALL, ref.FMH):
            rm, rb = ref.kmers(S, mask, w, pred, nonce, modulus)
            om, ob = port.kmers(codes, lens, mask, w, pred, nonce, modulus, variant, want_bits=True)
            assert np.array_equal(rm, om) and np.array_equal(rb, ob), (trial, w, k, pred)
    ref.set_boost_variant(181)
