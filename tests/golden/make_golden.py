#!/usr/bin/env python
"""Regenerates tests/golden/*.json by running the UNMODIFIED reference (oracle/_ref/libref.so,
built from /root/reference/src by oracle/Makefile against oracle/shim).

Run from the repo root in the development container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py [--big]

The fixtures are what the reference's own code returns; the C oracle (oracle/oracle.c) and the CUDA
path are both checked against them.  FracMinHash entries are produced for both Boost hash_combine
variants the shim restates (171 = Boost 1.71..1.80, 181 = Boost >= 1.81).
--big also regenerates the config-sized KAT-4 entries (5 Mbp, ~2 min of CPU).
"""
import hashlib
import json
import os
import random
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import port, ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def hx(v):
    return "%x" % int(v)


def key_digest(keys: np.ndarray) -> str:
    """sha256 over the ascending (lo, hi) little-endian u64 pairs of a set."""
    return hashlib.sha256(np.ascontiguousarray(keys, dtype="<u8").tobytes()).hexdigest()


def golden_masks():
    out = {"contiguous": {}, "random": []}
    for k in range(0, 65):
        out["contiguous"][str(k)] = hx(ref.contiguous_mask(k))
    out["contiguous_65_throws"] = ref.contiguous_mask(65) is None
    for w in range(1, 65):
        for k in sorted(set([0, 1, w // 3, w // 2, (2 * w) // 3, w - 1, w])):
            if 0 <= k <= w:
                for seed in (0, 1, 12345):
                    out["random"].append([w, k, seed, hx(ref.random_mask(w, k, seed))])
    # the driver's sweep (src/kmer-sketching.cpp:219-239)
    out["driver_sweep"] = [[k, k, hx(ref.random_mask(k, k, 0))] for k in range(10, 41)] + \
                          [[k + 10, k, hx(ref.random_mask(k + 10, k, 0))] for k in range(10, 41)]
    rnd = random.Random(7)
    out["reverse"] = []
    for _ in range(64):
        v = rnd.getrandbits(128)
        out["reverse"].append([hx(v), hx(ref.reverse_bitset(v))])
    return out


def golden_hash():
    rnd = random.Random(11)
    out = {"bitset_hash": [], "fmh": []}
    vals = [0, 1, (1 << 128) - 1, 0xF0CF, 1 << 64, (1 << 64) - 1] + [rnd.getrandbits(128) for _ in range(40)] + \
           [rnd.getrandbits(48) for _ in range(20)]
    for variant in (171, 181):
        ref.set_boost_variant(variant)
        for v in vals:
            out["bitset_hash"].append([variant, hx(v), hx(ref.boost_hash_bitset(v))])
        for _ in range(60):
            w = rnd.randint(1, 64)
            mask = ref.random_mask(w, rnd.randint(1, w), rnd.randint(0, 99))
            masked = rnd.getrandbits(2 * w) & mask
            nonce = rnd.choice([0, 1, 2, 200, -1, -7, 2**31 - 1])
            out["fmh"].append([variant, nonce, w, hx(masked), hx(mask), hx(ref.fmh(nonce, w, masked, mask))])
    ref.set_boost_variant(181)
    return out


FASTA_CASES = {
    # SURVEY.md 3.6 fixtures + extra quirks
    "multi_n_blank": b">r1\nACGTAC\nGTNNAC\n\nGGGG\n>r2\nacgtRYac\n",
    "crlf": b">r1\r\nACGT\r\nACGT\r\n",
    "empty_header_and_space": b">\nACGT\n>ok\nAC GT\nAAAA\n>ok2\nTTTT\n",
    "seq_before_header": b"ACGT\n>x\n\nCCCC\n",
    "no_trailing_newline": b">a\nACGTACGTAC",
    "only_header": b">a\n",
    "empty_file": b"",
    "blank_lines_only": b"\n\n\n",
    "lowercase_iupac": b">a desc here\nacgtnNacgtRYKMacgt\nACGT\n>b\nNNNN\n>c\nA\n",
    "header_with_space_ok": b">name with spaces\nACGTACGT\nTTTT\n",
    "gt_inside": b">a\nAC>GT\nACGT\n",
    "tab_in_seq": b">a\nAC\tGT\nACGT\n",
    "blank_then_more": b">a\nAAAA\n\nCCCC\n\n\nGGGG\n>b\nTTTT\n",
}


def golden_fasta():
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for name, text in FASTA_CASES.items():
            p = os.path.join(td, name + ".fa")
            with open(p, "wb") as f:
                f.write(text)
            segs = ref.Strings.from_fasta(p).segments()
            out[name] = {"text": text.decode("latin1"),
                         "segments": ["".join(str(int(c)) for c in s) for s in segs]}
    return out


def small_sequences():
    rnd = random.Random(5)
    seqs = {
        "readme": "AAACGTACGTTT",
        "homopolymer_a": "A" * 40,
        "homopolymer_t": "T" * 40,
        "palindrome": "ACGTACGTACGTACGTACGTACGTACGTACGT",
        "at_repeat": "AT" * 40,
        "random_150": "".join(rnd.choice("ACGT") for _ in range(150)),
        "random_70": "".join(rnd.choice("ACGT") for _ in range(70)),
    }
    return seqs


def golden_kmer_lists():
    """Ordered duplicate-preserving lists incl. kmer_bits on small inputs."""
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    seeds = ["11001011", "11111111", "1", "101", "011101110010111110011011",
             "0011111011010111111011001011101", "1" * 32, "1" * 33, "10" * 20, "1" * 64,
             "1" + "0" * 62 + "1", "0" * 10 + "1" * 5 + "0" * 3]
    out = []
    seqs = small_sequences()
    for sname, s in seqs.items():
        codes = np.array([code[c] for c in s], dtype=np.uint8)
        for seed in seeds:
            mask, w = port.seed_to_mask(seed)
            for segs in ([len(codes)], None):
                if segs is None:  # a ragged split: lengths w-1, w, w+1, rest (when it fits)
                    segs = [x for x in (w - 1, w, w + 1) if x > 0]
                    if sum(segs) >= len(codes):
                        continue
                    segs = segs + [len(codes) - sum(segs)]
                pieces, o = [], 0
                for L in segs:
                    pieces.append(codes[o:o + L])
                    o += L
                S = ref.Strings.from_codes(pieces)
                masked, bits = ref.kmers(S, mask, w, ref.ALL)
                lm, lb = ref.kmers(S, mask, w, ref.ALL, legacy=True)
                assert np.array_equal(masked, lm) and np.array_equal(bits, lb), "legacy path differs"
                entry = {"seq": sname, "seed": seed, "segs": segs,
                         "masked": [hx(int(a) | (int(b) << 64)) for a, b in masked],
                         "bits": [hx(int(a) | (int(b) << 64)) for a, b in bits]}
                for variant in (171, 181):
                    ref.set_boost_variant(variant)
                    fm, _ = ref.kmers(S, mask, w, ref.FMH, nonce=1, modulus=4)
                    entry["fmh4_%d" % variant] = [hx(int(a) | (int(b) << 64)) for a, b in fm]
                ref.set_boost_variant(181)
                out.append(entry)
    return {"sequences": seqs, "cases": out}


def set_stats(A, B, mask, w, pred, variant, nonce=1, modulus=200, via_fasta_dir=None):
    ref.set_boost_variant(variant)
    if via_fasta_dir:
        pa, pb = os.path.join(via_fasta_dir, "A.fna"), os.path.join(via_fasta_dir, "B.fna")
        sa, sb = ref.sets_from_fasta_files([pa, pb], mask, w, pred, nonce, modulus, parallel=True)
    else:
        sa = ref.KmerSet.from_strings(ref.Strings.from_codes([A]), mask, w, pred, nonce, modulus)
        sb = ref.KmerSet.from_strings(ref.Strings.from_codes([B]), mask, w, pred, nonce, modulus)
    I = ref.intersection(sa, sb)
    wt = port.mask_weight(mask)
    r = {"size_a": sa.size(), "size_b": sb.size(), "intersection": I,
         "ani_ab": repr(ref.binomial_estimator(ref.containment(I, sa.size()), wt)),
         "ani_ba": repr(ref.binomial_estimator(ref.containment(I, sb.size()), wt)),
         "digest_a": key_digest(sa.keys()), "digest_b": key_digest(sb.keys())}
    ref.set_boost_variant(181)
    return r


def golden_sets(L, tag):
    """KAT-3 / KAT-4: A = gen(L, 42), B = mutate(A, 43, 100)."""
    A = port.gen(L, 42)
    B = port.mutate(A, 43, 100)
    out = {"L": L, "gen_seed": 42, "mut_seed": 43, "D": 100, "hamming": int((A != B).sum()),
           "prefix": port.codes_to_text(A[:32]).decode(), "cases": []}
    seeds = ["11001011", "011101110010111110011011", "0011111011010111111011001011101"]
    if L <= 200000:
        seeds += ["1" * 40, "01111101111111010111111101101111111111101100111011", "1" * 64]
    with tempfile.TemporaryDirectory() as td:
        port.write_fasta(os.path.join(td, "A.fna"), A, "A")
        port.write_fasta(os.path.join(td, "B.fna"), B, "B")
        for seed in seeds:
            mask, w = port.seed_to_mask(seed)
            for pred, variant in ((ref.ALL, 181), (ref.FMH, 171), (ref.FMH, 181)):
                r = set_stats(A, B, mask, w, pred, variant, via_fasta_dir=td)
                r.update({"seed": seed, "pred": "ALL" if pred == ref.ALL else "FMH", "nonce": 1, "modulus": 200,
                          "variant": variant})
                out["cases"].append(r)
                print(tag, seed, r["pred"], variant, r["size_a"], r["size_b"], r["intersection"], r["ani_ab"],
                      flush=True)
    return out


def golden_multi():
    """A small all-vs-all: 6 genomes of 20 kbp at graded mutation rates; full n x n matrices in
    generate_all_pairs_from_vector order, through the reference's pairwise function."""
    base = port.gen(20000, 1000)
    Ds = [0, 1000, 200, 100, 50, 20]
    genomes = [base if D == 0 else port.mutate(base, 2000 + g, D) for g, D in enumerate(Ds)]
    out = {"L": 20000, "base_seed": 1000, "Ds": Ds, "cases": []}
    for seed, pred, modulus in (("0011111011010111111011001011101", ref.FMH, 20),
                                ("011101110010111110011011", ref.ALL, 1),
                                ("11001011", ref.FMH, 3)):
        mask, w = port.seed_to_mask(seed)
        wt = port.mask_weight(mask)
        for variant in (171, 181):
            ref.set_boost_variant(variant)
            sets = [ref.KmerSet.from_strings(ref.Strings.from_codes([g]), mask, w, pred, 1, modulus) for g in genomes]
            f, s = ref.all_pairs(len(sets))
            ints = ref.pairwise_intersections([sets[i] for i in f], [sets[j] for j in s])
            anis = [repr(ref.binomial_estimator(ref.containment(int(I), sets[i].size()), wt)) for I, i in zip(ints, f)]
            out["cases"].append({"seed": seed, "pred": "ALL" if pred == ref.ALL else "FMH", "nonce": 1,
                                 "modulus": modulus, "variant": variant, "sizes": [x.size() for x in sets],
                                 "intersections": [int(x) for x in ints], "ani": anis,
                                 "digests": [key_digest(x.keys()) for x in sets]})
    ref.set_boost_variant(181)
    mism = ref.pairwise_intersections(sets[:2], sets[:3])
    out["length_mismatch_throws"] = mism is None
    return out


def dump(name, obj):
    with open(os.path.join(OUT, name), "w") as f:
        json.dump(obj, f, indent=0, separators=(",", ":"))
        f.write("\n")
    print("wrote", name, os.path.getsize(os.path.join(OUT, name)), "bytes")


if __name__ == "__main__":
    dump("masks.json", golden_masks())
    dump("hash.json", golden_hash())
    dump("fasta.json", golden_fasta())
    dump("kmer_lists.json", golden_kmer_lists())
    dump("multi.json", golden_multi())
    dump("sets_100k.json", golden_sets(100000, "KAT-3"))
    if "--big" in sys.argv:
        dump("sets_5m.json", golden_sets(5000000, "KAT-4"))
