"""GPU parity tests: the CUDA path (through the C ABI of libsks.so) against the CPU oracle on the same
seeded inputs, against the golden fixtures produced by the unmodified reference, and -- at the
BASELINE.json sizes -- through size-independent properties.  Integer results are compared bit-exactly;
ANI within 1e-12 (it is computed on the host in double from identical integer counts)."""
import hashlib

import numpy as np
import pytest

import spaced_kmer_sketching_b200 as sks
from oracle import port

from conftest import load_golden

pytestmark = pytest.mark.gpu

CODE = {"A": 0, "C": 1, "G": 2, "T": 3}
C2_SEED = "011101110010111110011011"          # (24,16,seed 0)
C3_SEED = "0011111011010111111011001011101"   # (31,21,seed 0)


def ival(h):
    return int(h, 16)


def to_int(a):
    return [int(x[0]) | (int(x[1]) << 64) for x in a]


def key_digest(keys):
    return hashlib.sha256(np.ascontiguousarray(keys, dtype="<u8").tobytes()).hexdigest()


@pytest.fixture(scope="module")
def ctx():
    c = sks.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_bucket():
    """A context that builds every bitset above one slice through the bucketed (slice-assembly) path."""
    import os
    old = os.environ.get("SKS_BUCKET_MIN_BITS")
    os.environ["SKS_BUCKET_MIN_BITS"] = "20"
    c = sks.Context(0)
    if old is None:
        del os.environ["SKS_BUCKET_MIN_BITS"]
    else:
        os.environ["SKS_BUCKET_MIN_BITS"] = old
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_exact():
    """Bucketed build through the exact counting partition only (the fused fast path switched off)."""
    import os
    os.environ["SKS_BUCKET_MIN_BITS"] = "20"
    os.environ["SKS_EXACT_PARTITION"] = "1"
    c = sks.Context(0)
    del os.environ["SKS_BUCKET_MIN_BITS"], os.environ["SKS_EXACT_PARTITION"]
    yield c
    c.close()


def opred(pred):
    if pred.kind == sks.PRED_ALL:
        return (port.ALL,)
    return (port.FMH, pred.nonce, pred.modulus, pred.hash_variant)


# ---- ordered lists: the most direct view of the hot loop --------------------------------------
def test_kmer_lists_golden(ctx):
    g = load_golden("kmer_lists.json")
    seqs = {k: np.array([CODE[c] for c in v], dtype=np.uint8) for k, v in g["sequences"].items()}
    for case in g["cases"]:
        mask, w = sks.seed_to_mask(case["seed"])
        batch = ctx.upload_codes([seqs[case["seq"]]], [np.array(case["segs"], dtype=np.uint64)])
        masked, bits = ctx.kmer_list(batch, 0, mask, w, sks.all_kmers())
        tag = (case["seq"], case["seed"], case["segs"])
        assert to_int(masked) == [ival(x) for x in case["masked"]], tag
        assert to_int(bits) == [ival(x) for x in case["bits"]], tag
        for variant in (171, 181):
            fm, _ = ctx.kmer_list(batch, 0, mask, w, sks.frac_min_hash(1, 4, variant))
            assert to_int(fm) == [ival(x) for x in case["fmh4_%d" % variant]], tag + (variant,)
        batch.close()


def test_kat1_readme(ctx):
    mask, w = sks.seed_to_mask("11001011")
    codes = np.array([CODE[c] for c in "AAACGTACGTTT"], dtype=np.uint8)
    batch = ctx.upload_codes([codes])
    masked, bits = ctx.kmer_list(batch, 0, mask, w, sks.all_kmers())
    assert to_int(masked) == [0x81, 0xC6, 0x100B, 0xC6, 0x81]   # SURVEY.md 4.2 KAT-1 (tie -> rc at i=2)
    (s,) = ctx.sketch(batch, mask, w, sks.all_kmers(), sks.REPR_SORTED)
    assert to_int(s.keys()) == [0x81, 0xC6, 0x100B]
    (s2,) = ctx.sketch(batch, mask, w, sks.all_kmers(), sks.REPR_BITSET)
    assert to_int(s2.keys()) == [0x81, 0xC6, 0x100B] and s2.kmer_set_size() == 3


def random_case(rng, trial):
    w = int(rng.integers(1, 65))
    k = int(rng.integers(1, w + 1))
    mask = sks.generate_random_spaced_seed_mask(w, k, trial)
    nseg = int(rng.integers(1, 6))
    lens = []
    for _ in range(nseg):
        kind = int(rng.integers(0, 6))
        lens.append([w - 1, w, w + 1, int(rng.integers(1, 300)), int(rng.integers(300, 9000)),
                     int(rng.integers(8000, 40000))][kind])
    lens = [L for L in lens if L > 0] or [w]
    codes = rng.integers(0, 4, sum(lens), dtype=np.uint8)
    if trial % 5 == 0:   # low-complexity stretches: duplicates and strand ties
        codes[: len(codes) // 2] = np.tile(np.array([0, 3], dtype=np.uint8), len(codes))[: len(codes) // 2]
    return w, k, mask, codes, lens


@pytest.mark.parametrize("block", range(4))
def test_fuzz_lists_and_sets_vs_oracle(ctx, ctx_bucket, block):
    rng = np.random.default_rng(100 + block)
    for t in range(12):
        trial = block * 12 + t
        w, k, mask, codes, lens = random_case(rng, trial)
        variant = 171 if trial % 2 else 181
        preds = [sks.all_kmers(), sks.frac_min_hash(int(rng.integers(-3, 5)), int(rng.integers(1, 12)), variant)]
        batch = ctx.upload_codes([codes], [np.array(lens, dtype=np.uint64)])
        for pred in preds:
            om, ob = port.kmers(codes, lens, mask, w, *opred(pred), want_bits=True)
            gm, gb = ctx.kmer_list(batch, 0, mask, w, pred)
            tag = (trial, w, k, lens, pred.kind)
            assert np.array_equal(gm, om), tag
            assert np.array_equal(gb, ob), tag
            oset = port.sort_unique(om)
            (s,) = ctx.sketch(batch, mask, w, pred, sks.REPR_SORTED)
            assert s.kmer_set_size() == len(oset) and np.array_equal(s.keys(), oset), tag
            if k <= 13:
                (b,) = ctx.sketch(batch, mask, w, pred, sks.REPR_BITSET)
                assert b.kmer_set_size() == len(oset) and np.array_equal(b.keys(), oset), tag
                if k >= 10:   # same through the bucketed build (atomic path above when 2k < 26)
                    b2 = ctx_bucket.upload_codes([codes], [np.array(lens, dtype=np.uint64)])
                    (b,) = ctx_bucket.sketch(b2, mask, w, pred, sks.REPR_BITSET)
                    assert b.kmer_set_size() == len(oset) and np.array_equal(b.keys(), oset), tag
                    b2.close()
        batch.close()


def test_bitset_paths_agree_multi_genome(ctx, ctx_bucket, ctx_exact):
    """Atomic-insert and bucketed bitset builds give identical bitsets (sizes, members, intersections)."""
    rng = np.random.default_rng(11)
    genomes = [rng.integers(0, 4, n, dtype=np.uint8) for n in (70000, 5, 8192, 33333)]
    genomes.append(genomes[0].copy())
    genomes[4][::97] = (genomes[4][::97] + 1) & 3
    # skewed index distributions: one slice / one coarse bucket holds (nearly) every index, which drives
    # the bucketed build through its span-halving and direct-scan paths
    genomes.append(np.zeros(60000, dtype=np.uint8))                                  # poly-A
    genomes.append((rng.integers(0, 2, 90000) * 3).astype(np.uint8))                 # A/T only
    genomes.append(np.where(rng.random(120000) < 0.9, 0, rng.integers(0, 4, 120000)).astype(np.uint8))
    segl = [None, None, np.array([4000, 4192], dtype=np.uint64), None, None, None, None, None]
    for seed, pred in (("1101100111011", sks.all_kmers()), ("110110011101101", sks.frac_min_hash(1, 3))):
        mask, w = sks.seed_to_mask(seed)
        osets = [port.sketch_set(g, [len(g)] if s is None else list(s), mask, w, *opred(pred))
                 for g, s in zip(genomes, segl)]
        for c in (ctx, ctx_bucket, ctx_exact):
            sets = c.sketch(c.upload_codes(genomes, segl), mask, w, pred, sks.REPR_BITSET)
            assert [s.kmer_set_size() for s in sets] == [len(o) for o in osets]
            for s, o in zip(sets, osets):
                assert np.array_equal(s.keys(), o)
            assert c.intersect(sets[0], sets[4]) == port.intersection(osets[0], osets[4])


def test_multi_genome_batch_and_intersections(ctx):
    rng = np.random.default_rng(7)
    base = rng.integers(0, 4, 30000, dtype=np.uint8)
    genomes, seglens = [], []
    for i in range(7):
        g = base.copy()
        flips = rng.random(len(g)) < 0.01 * i
        g[flips] = (g[flips] + 1) & 3
        n = [30000, 8192, 8193, 16384 + 23, 5, 29999, 12000][i]
        genomes.append(g[:n])
        seglens.append(np.array([n // 3, n - n // 3], dtype=np.uint64) if i % 2 else None)
    batch = ctx.upload_codes(genomes, seglens)
    for seed, pred, reprs in ((C2_SEED, sks.all_kmers(), (sks.REPR_SORTED,)),
                              ("110101101", sks.all_kmers(), (sks.REPR_SORTED, sks.REPR_BITSET)),
                              (C3_SEED, sks.frac_min_hash(1, 20), (sks.REPR_SORTED,)),
                              ("1" * 40, sks.frac_min_hash(2, 7, 171), (sks.REPR_SORTED,))):
        mask, w = sks.seed_to_mask(seed)
        osets = [port.sketch_set(g, [len(g)] if s is None else list(s), mask, w, *opred(pred))
                 for g, s in zip(genomes, seglens)]
        want = np.array([[port.intersection(a, b) for b in osets] for a in osets], dtype=np.int32)
        for r in reprs:
            sets = ctx.sketch(batch, mask, w, pred, r)
            for s, o in zip(sets, osets):
                assert np.array_equal(s.keys(), o)
            got = ctx.intersect_all_pairs(sets)
            assert np.array_equal(got, want)
            f, s2 = sks.generate_all_pairs_from_vector(list(range(len(sets))))
            flat = ctx.intersect_pairs([sets[i] for i in f], [sets[j] for j in s2])
            assert np.array_equal(flat, want.ravel())
            # row tiling (rank tiling of the pair matrix) fills exactly its rows
            part = np.full_like(want, -1)
            ctx.intersect_all_pairs(sets, 2, 5, part)
            assert np.array_equal(part[2:5], want[2:5]) and (part[:2] == -1).all() and (part[5:] == -1).all()
    with pytest.raises(sks.SksError) as e:   # src/kmer_set.cpp:147-150,174-177
        ctx.intersect_pairs(sets[:2], sets[:3])
    assert e.value.code == 4 and "different lengths" in str(e.value)


def test_multi_golden(ctx):
    g = load_golden("multi.json")
    n = len(g["Ds"])
    batch = ctx.synth(g["L"], [g["base_seed"]] * n, [2000 + i for i in range(n)], g["Ds"])
    for case in g["cases"]:
        mask, w = sks.seed_to_mask(case["seed"])
        pred = sks.all_kmers() if case["pred"] == "ALL" else sks.frac_min_hash(case["nonce"], case["modulus"], case["variant"])
        sets = ctx.sketch(batch, mask, w, pred)
        assert [s.kmer_set_size() for s in sets] == case["sizes"]
        assert [key_digest(s.keys()) for s in sets] == case["digests"]
        ints = ctx.intersect_all_pairs(sets)
        assert ints.ravel().tolist() == case["intersections"]
        sizes = np.repeat(np.array(case["sizes"], dtype=np.int32), n)
        ani = sks.ani_from_counts(ints.ravel(), sizes, sks.mask_weight(mask))
        assert np.max(np.abs(ani - np.array([float(x) for x in case["ani"]]))) <= 1e-12
        # the multi-GPU tiling on one device: every rank's block rows (sks_all_vs_all_sharded with a single rank is the
        # same call on rows [begin, end)), stacked in rank order; and the rectangles of the pairwise kernels
        for world in (1, 2, 3, 4, 6):
            parts = [ctx.all_vs_all(sets, *sks.shard_range(n, r, world)) for r in range(world)]
            assert np.array_equal(np.concatenate([p[0] for p in parts]), ints), world
            assert np.max(np.abs(np.concatenate([p[2] for p in parts]).ravel() - np.array([float(x) for x in case["ani"]]))) <= 1e-12
            full = np.full((n, n), -1, dtype=np.int32)
            for r in range(world):
                rows = sks.shard_range(n, r, world)
                ctx.intersect_block(sets, rows, (0, n), full)
            assert np.array_equal(full, ints), world


def test_all_pairs_row_resident_and_merge_paths(ctx):
    """All-vs-all through the row-resident intersection kernel (row set in shared memory, bucket index on the top
    key bits), incl. 16-byte keys, a row too large to be resident (merge kernel), empty sets, tiny key ranges."""
    rng = np.random.default_rng(21)
    base = rng.integers(0, 4, 20000, dtype=np.uint8)
    genomes = [base]
    for d in (500, 100, 40, 15, 7, 3):
        g = base.copy()
        idx = rng.integers(0, len(g), len(g) // d)
        g[idx] = (g[idx] + rng.integers(1, 4, len(idx))) & 3
        genomes.append(g)
    genomes.append(rng.integers(0, 4, 45000, dtype=np.uint8))      # more keys than the resident kernel holds
    genomes.append(rng.integers(0, 4, 12, dtype=np.uint8))         # shorter than most windows: empty set
    genomes.append(np.zeros(5000, dtype=np.uint8))                 # poly-A: one key
    batch = ctx.upload_codes(genomes)
    n = len(genomes)
    for seed, pred in (("0011111011010111111011001011101", sks.all_kmers()),
                       ("0011111011010111111011001011101", sks.frac_min_hash(1, 3)),
                       ("1110110111011011101101110110111011011101", sks.all_kmers()),     # window 40: 16-byte keys
                       ("11001011", sks.all_kmers()),                                      # 16-bit keys
                       ("1", sks.all_kmers())):
        mask, w = sks.seed_to_mask(seed)
        sets = ctx.sketch(batch, mask, w, pred, sks.REPR_SORTED)
        osets = [port.sketch_set(g, [len(g)], mask, w, *opred(pred)) for g in genomes]
        want = np.array([[port.intersection(a, b) for b in osets] for a in osets], dtype=np.int32)
        got = ctx.intersect_all_pairs(sets)            # dictionary route where the keys allow it
        assert np.array_equal(got, want), seed
        pairwise = ctx.intersect_block(sets, (0, n), (0, n), np.full((n, n), -1, dtype=np.int32))   # pairwise kernels
        assert np.array_equal(pairwise, want), seed
        # rectangles, as the multi-GPU tiling issues them
        part = np.full((n, n), -1, dtype=np.int32)
        ctx.intersect_block(sets, (2, 7), (0, n), part)
        ctx.intersect_block(sets, (7, n), (1, 6), part)
        assert np.array_equal(part[2:7], want[2:7]) and np.array_equal(part[7:, 1:6], want[7:, 1:6])
        assert (part[:2] == -1).all() and (part[7:, 6:] == -1).all()
        for x in sets:
            x.close()
    batch.close()


def test_fuzz_all_pairs_random_masks(ctx):
    """Random spaced seeds of every span (bucket plans of 1..4 mask runs, masks below bit 32, 16-byte keys), random
    FracMinHash moduli: the all-pairs counts equal numpy intersections of the key arrays, which equal the oracle's
    sets for a sample of genomes."""
    rng = np.random.default_rng(77)
    base = rng.integers(0, 4, 60000, dtype=np.uint8)
    genomes = [base]
    for d in (300, 60, 25, 9, 4):
        g = base.copy()
        idx = rng.integers(0, len(g), len(g) // d)
        g[idx] = (g[idx] + rng.integers(1, 4, len(idx))) & 3
        genomes.append(g)
    genomes.append(rng.integers(0, 4, 20000, dtype=np.uint8))
    batch = ctx.upload_codes(genomes)
    n = len(genomes)
    for trial in range(28):
        w = int(rng.integers(6, 65))
        k = int(rng.integers(max(3, w // 3), w + 1))
        mask = sks.generate_random_spaced_seed_mask(w, k, int(rng.integers(0, 1000)))
        pred = sks.all_kmers() if trial % 4 == 0 else sks.frac_min_hash(int(rng.integers(0, 5)), int(rng.integers(2, 40)))
        sets = ctx.sketch(batch, mask, w, pred, sks.REPR_SORTED)
        keys = [s.keys() for s in sets]
        as_set = [set(map(tuple, kk.tolist())) for kk in keys]
        want = np.array([[len(a & b) for b in as_set] for a in as_set], dtype=np.int32)
        got = ctx.intersect_all_pairs(sets)
        assert np.array_equal(got, want), (trial, w, k, hex(mask))
        pairwise = ctx.intersect_block(sets, (0, n), (0, n), np.full((n, n), -1, dtype=np.int32))
        assert np.array_equal(pairwise, want), (trial, w, k, hex(mask))
        r0, r1 = sorted(int(x) for x in rng.integers(0, n + 1, 2))
        cnt, sizes, ani = ctx.all_vs_all(sets, r0, r1)
        assert np.array_equal(cnt, want[r0:r1]) and sizes.tolist() == [len(a) for a in as_set]
        wani = sks.ani_from_counts(want[r0:r1].ravel(), np.repeat(sizes[r0:r1], n), sks.mask_weight(mask))
        assert ani.size == 0 or np.max(np.abs(ani.ravel() - wani)) <= 1e-12
        g = int(rng.integers(0, n))
        assert np.array_equal(keys[g], port.sketch_set(genomes[g], [len(genomes[g])], mask, w, *opred(pred))), (trial, g)
        # single pairs and the ring (src/generators.hpp:20-31) go through the same tables
        a, b = int(rng.integers(0, n)), int(rng.integers(0, n))
        assert ctx.intersect(sets[a], sets[b]) == want[a, b]
        ring = ctx.intersect_pairs(sets, sets[1:] + sets[:1])
        assert ring.tolist() == [int(want[i, (i + 1) % n]) for i in range(n)]
        for x in sets:
            x.close()
    batch.close()


def test_all_vs_all_dictionary_route(ctx):
    """sks_all_vs_all on sets that exercise every shape of the dictionary route (csrc/sks_allpairs.cu): keys shared by
    most sets (class A ids, dense bitmaps, several id ranges), keys shared by exactly two sets (class B), private
    keys (dropped), empty sets, the all-zero key, row ranges -- against the pairwise kernels, numpy and the host ANI
    (src/kmer-sketching.cpp:185-200)."""
    rng = np.random.default_rng(5)
    base = rng.integers(0, 4, 150_000, dtype=np.uint8)
    other = rng.integers(0, 4, 90_000, dtype=np.uint8)
    genomes = []
    for g in range(50):
        src = base if g % 5 else other
        d = [0, 2000, 300, 50, 12, 5, 3][g % 7]
        x = src.copy()
        if d:
            idx = rng.integers(0, len(x), len(x) // d)
            x[idx] = (x[idx] + rng.integers(1, 4, len(idx))) & 3
        genomes.append(x)
    genomes[7] = rng.integers(0, 4, 10, dtype=np.uint8)       # empty set
    genomes[11] = np.zeros(3000, dtype=np.uint8)              # poly-A: the all-zero key only
    genomes[13] = np.concatenate([np.zeros(500, dtype=np.uint8), base[:20000]])   # shares the all-zero key with 11
    batch = ctx.upload_codes(genomes)
    n = len(genomes)
    for seed, pred in ((C3_SEED, sks.all_kmers()), (C3_SEED, sks.frac_min_hash(1, 7)),
                       ("1110110111011011101101110110111011011101", sks.frac_min_hash(1, 3))):   # 16-byte keys, compacted
        mask, w = sks.seed_to_mask(seed)
        sets = ctx.sketch(batch, mask, w, pred, sks.REPR_SORTED)
        want = ctx.intersect_block(sets, (0, n), (0, n), np.full((n, n), -1, dtype=np.int32))    # pairwise kernels
        keys = [s.keys() for s in sets]
        for i, j in ((0, 1), (3, 44), (11, 13), (5, 10), (7, 2)):
            a, b = set(map(tuple, keys[i].tolist())), set(map(tuple, keys[j].tolist()))
            assert want[i, j] == len(a & b)
        for rows in ((0, n), (0, 1), (17, 40), (n - 1, n)):
            cnt, sizes, ani = ctx.all_vs_all(sets, rows[0], rows[1])
            assert np.array_equal(cnt, want[rows[0]:rows[1]]), (seed, rows)
            assert sizes.tolist() == [len(k) for k in keys]
            wani = sks.ani_from_counts(want[rows[0]:rows[1]].ravel(), np.repeat(sizes[rows[0]:rows[1]], n), sks.mask_weight(mask))
            assert np.max(np.abs(ani.ravel() - wani)) <= 1e-12
        assert np.array_equal(ctx.intersect_all_pairs(sets), want)
        for x in sets:
            x.close()
    batch.close()


def test_imported_keys_are_validated(ctx, tmp_path):
    """Foreign keys (device buffers, host lists, sketch files) are checked before a set is built from them: subsets of
    the mask, ascending and distinct where that is claimed, a key width that fits the window; a sketch file whose
    header lies about its length is refused without a huge allocation."""
    import torch
    mask, w = sks.seed_to_mask(C3_SEED)
    good = np.sort(np.array([0, mask & 0x3, mask & 0xFFFF0000, mask], dtype=np.uint64))
    good = np.unique(good)
    def dev(a):
        return torch.from_numpy(a.view(np.int64)).cuda()
    t = dev(good)
    s = ctx.set_from_device_keys(t.data_ptr(), len(good), 1, mask, w)
    assert s.keys()[:, 0].tolist() == good.tolist()
    s.close()
    for bad, what in ((np.array([1, 2, 4 | (1 << 63)], dtype=np.uint64), "outside the mask"),     # bit 63 is not in a 62-bit mask
                      (good[::-1].copy(), "ascending"),
                      (np.concatenate([good[:2], good[1:]]), "ascending")):
        t = dev(bad)
        with pytest.raises(sks.SksError) as e:
            ctx.set_from_device_keys(t.data_ptr(), len(bad), 1, mask, w)
        assert what in str(e.value), (what, str(e.value))
    t = dev(np.array([5, 1, 1 << 63], dtype=np.uint64))
    with pytest.raises(sks.SksError) as e:   # unsorted import: only the mask is checked
        ctx.set_from_device_keys(t.data_ptr(), 3, 1, mask, w, sorted_unique=False)
    assert "outside the mask" in str(e.value)
    t = dev(good)
    with pytest.raises(sks.SksError):        # window 31 takes one-word keys
        ctx.set_from_device_keys(t.data_ptr(), len(good) // 2, 2, mask, w)
    # several sets back to back: ascending inside every set only
    two = np.concatenate([good, good[:2]])
    t = dev(two)
    a, b = ctx.sets_from_device_keys(t.data_ptr(), [len(good), 2], 1, mask, w)
    assert a.kmer_set_size() == len(good) and ctx.intersect(a, b) == 2
    with pytest.raises(sks.SksError):
        ctx.sets_from_device_keys(t.data_ptr(), [len(good) + 1, 1], 1, mask, w)
    # sketch files
    path = str(tmp_path / "s.sks")
    a.save(path, sks.frac_min_hash(1, 200))
    raw = bytearray(open(path, "rb").read())
    lying = bytearray(raw)
    lying[56:64] = (1 << 39).to_bytes(8, "little")          # n_keys: 2^39 keys in a 100-byte file
    open(path, "wb").write(lying)
    with pytest.raises(sks.SksError) as e:
        ctx.load_set(path)
    assert "truncated" in str(e.value)
    off_mask = bytearray(raw)
    off_mask[64 + 7] |= 0x80                                 # first key gets bit 63
    open(path, "wb").write(off_mask)
    with pytest.raises(sks.SksError) as e:
        ctx.load_set(path)
    assert "outside the mask" in str(e.value) or "ascending" in str(e.value)
    open(path, "wb").write(raw)
    back, _ = ctx.load_set(path)
    assert back.keys()[:, 0].tolist() == good.tolist()
    for x in (a, b, back):
        x.close()


def test_auto_representation_follows_cost_and_list_capacity(ctx):
    """SKS_REPR_AUTO: a 4^16-bit presence bitset for a 5 Mbp genome, sorted keys for a 100 kbp one (512 MiB per tiny
    genome would be absurd); the ordered k-mer list refuses genomes of 2^31 bases and more instead of aliasing the
    strand bit."""
    mask, w = sks.seed_to_mask(C2_SEED)
    small = ctx.synth(100_000, [1, 2, 3], [0, 0, 0], [0, 0, 0])
    sets = ctx.sketch(small, mask, w, sks.all_kmers())
    assert [s.repr for s in sets] == [sks.REPR_SORTED] * 3
    big = ctx.synth(5_000_000, [42], [0], [0])
    (b,) = ctx.sketch(big, mask, w, sks.all_kmers())
    assert b.repr == sks.REPR_BITSET and b.kmer_set_size() == 4994572      # KAT-4 |A|
    (t,) = ctx.sketch(small, sks.seed_to_mask("11001011")[0], 8, sks.all_kmers())[:1]
    assert t.repr == sks.REPR_BITSET                                       # 4^5 bits: always
    for x in sets + [b, t]:
        x.close()
    huge = ctx.synth(1 << 31, [9], [0], [0])
    with pytest.raises(sks.SksError) as e:
        ctx.kmer_list(huge, 0, mask, w, sks.frac_min_hash(1, 200))
    assert e.value.code == 3 and "2^31" in str(e.value)
    huge.close()
    small.close()
    big.close()


def test_bucket_sort_overflow_among_many_regions(ctx):
    """One region whose keys are all equal (poly-A) overflows its bucket while the buckets of 60 other regions are being
    sorted: the bucket sort has to give up as a whole and hand over to the library sort.  (Found by
    tools/stress_allpairs.py: threads used to leave a CTA one by one when the overflow flag went up, and the rest wrote
    through garbage prefix sums.)"""
    rng = np.random.default_rng(48)
    base = rng.integers(0, 4, 32838, dtype=np.uint8)
    genomes = [base.copy() for _ in range(60)]
    for g in genomes[1:]:
        idx = rng.integers(0, len(g), 300)
        g[idx] = (g[idx] + rng.integers(1, 4, len(idx))) & 3
    genomes.insert(35, np.zeros(1671, dtype=np.uint8))
    genomes[0] = genomes[0][:14]
    mask, w = 0xfcfffffffcf3fff3fff3cfffff3fc, 58          # 16-byte keys
    batch = ctx.upload_codes(genomes)
    want = {i: port.sketch_set(genomes[i], [len(genomes[i])], mask, w) for i in (0, 1, 35, 60)}
    for rep in range(6):
        sets = ctx.sketch(batch, mask, w, sks.all_kmers(), sks.REPR_SORTED)
        assert sets[35].kmer_set_size() == 1 and sets[0].kmer_set_size() == 0
        for i, o in want.items():
            assert np.array_equal(sets[i].keys(), o), (rep, i)
        for x in sets:
            x.close()
    batch.close()


def test_synth_matches_oracle_generator(ctx):
    batch = ctx.synth(100_003, [42, 42, 9], [0, 43, 5], [0, 100, 3])
    A = port.gen(100_003, 42)
    assert np.array_equal(sks.unpack_codes(batch.download(0), 100_003), A)
    assert np.array_equal(sks.unpack_codes(batch.download(1), 100_003), port.mutate(A, 43, 100))
    assert np.array_equal(sks.unpack_codes(batch.download(2), 100_003), port.mutate(port.gen(100_003, 9), 5, 3))


@pytest.mark.parametrize("name", ["sets_100k.json", "sets_5m.json"])
def test_sets_golden(ctx, name):
    """KAT-3 / KAT-4: counts and key digests produced by the unmodified reference."""
    g = load_golden(name)
    batch = ctx.synth(g["L"], [g["gen_seed"]] * 2, [0, g["mut_seed"]], [0, g["D"]])
    for case in g["cases"]:
        mask, w = sks.seed_to_mask(case["seed"])
        pred = sks.all_kmers() if case["pred"] == "ALL" else sks.frac_min_hash(case["nonce"], case["modulus"], case["variant"])
        weight = sks.mask_weight(mask)
        reprs = [sks.REPR_SORTED] + ([sks.REPR_BITSET] if weight <= 16 and pred.kind == sks.PRED_ALL else [])
        for r in reprs:
            sa, sb = ctx.sketch(batch, mask, w, pred, r)
            inter = ctx.intersect(sa, sb)
            tag = (case["seed"], case["pred"], case["variant"], r)
            assert (sa.kmer_set_size(), sb.kmer_set_size(), inter) == (case["size_a"], case["size_b"], case["intersection"]), tag
            if r == sks.REPR_SORTED or g["L"] <= 10 ** 6:
                assert key_digest(sa.keys()) == case["digest_a"] and key_digest(sb.keys()) == case["digest_b"], tag
            ab = sks.binomial_estimator(sks.containment(inter, sa.kmer_set_size()), weight)
            ba = sks.binomial_estimator(sks.containment(inter, sb.kmer_set_size()), weight)
            assert abs(ab - float(case["ani_ab"])) <= 1e-12 and abs(ba - float(case["ani_ba"])) <= 1e-12, tag
            sa.close()
            sb.close()
    # the one-call pipelines (resident and host-buffer) on C2
    case = [c for c in g["cases"] if c["seed"] == C2_SEED and c["pred"] == "ALL"][0]
    mask, w = sks.seed_to_mask(C2_SEED)
    r = ctx.pair_ani_resident(batch, mask, w, sks.all_kmers(), sks.REPR_BITSET)
    assert (r.size_a, r.size_b, r.intersection) == (case["size_a"], case["size_b"], case["intersection"])
    assert abs(r.ani_ab - float(case["ani_ab"])) <= 1e-12 and abs(r.ani_ba - float(case["ani_ba"])) <= 1e-12
    r3 = ctx.pair_ani_resident(batch, mask, w, sks.all_kmers(), sks.REPR_BITSET_ONCHIP)
    assert (r3.size_a, r3.size_b, r3.intersection, r3.ani_ab, r3.ani_ba) == (r.size_a, r.size_b, r.intersection, r.ani_ab, r.ani_ba)
    wa, wb = batch.download(0), batch.download(1)
    r2 = ctx.pair_ani(wa, g["L"], wb, g["L"], mask, w, sks.all_kmers())
    assert (r2.size_a, r2.size_b, r2.intersection, r2.ani_ab) == (r.size_a, r.size_b, r.intersection, r.ani_ab)


def test_pair_pipeline_fused_build_and_onchip(ctx, ctx_bucket, ctx_exact):
    """sks_pair_ani*: the fused slice-assembly + AND/popcount kernel (bitsets stored, and kept on chip) gives
    the counts of the separate build + intersection, on random, segmented, skewed and tiny genomes."""
    rng = np.random.default_rng(5)
    polya = np.zeros(60000, dtype=np.uint8)
    polya[rng.integers(0, 60000, 300)] = rng.integers(0, 4, 300)
    at_rich = rng.choice(np.array([0, 3], dtype=np.uint8), 50000, p=[0.5, 0.5])
    cases = [
        (rng.integers(0, 4, 80000, dtype=np.uint8), None),
        (polya, None),                                   # one bucket holds nearly everything (region overflow)
        (at_rich, None),
        (rng.integers(0, 4, 30011, dtype=np.uint8), [11, 5000, 3, 25, 24972]),
        (rng.integers(0, 4, 9, dtype=np.uint8), None),   # shorter than the window: empty set
    ]
    for seed in ("011101110010111110011011", "1101100111011", "1111111111111111"):
        mask, w = sks.seed_to_mask(seed)
        weight = sks.mask_weight(mask)
        for ia in range(len(cases)):
            ib = (ia + 1) % len(cases)
            A, sA = cases[ia]
            Bm = A.copy() if ia % 2 == 0 else cases[ib][0]
            sB = sA if ia % 2 == 0 else cases[ib][1]
            if ia % 2 == 0 and len(Bm) > 100:
                idx = rng.integers(0, len(Bm), len(Bm) // 50)
                Bm[idx] = (Bm[idx] + 1) & 3
            oa = port.sketch_set(A, [len(A)] if sA is None else list(sA), mask, w, port.ALL)
            ob = port.sketch_set(Bm, [len(Bm)] if sB is None else list(sB), mask, w, port.ALL)
            want = (len(oa), len(ob), port.intersection(oa, ob))
            for c in (ctx, ctx_bucket, ctx_exact):
                batch = c.upload_codes([A, Bm], [sA, sB])
                for r in (sks.REPR_BITSET, sks.REPR_BITSET_ONCHIP, sks.REPR_SORTED, sks.REPR_AUTO):
                    res = c.pair_ani_resident(batch, mask, w, sks.all_kmers(), r)
                    assert (res.size_a, res.size_b, res.intersection) == want, (seed, ia, r)
                    assert abs(res.ani_ab - port.ani(want[2], want[0], weight)) <= 1e-12
                    assert abs(res.ani_ba - port.ani(want[2], want[1], weight)) <= 1e-12
                batch.close()
    # the on-chip representation has no set to hand out
    batch = ctx.upload_codes([cases[0][0]])
    mask, w = sks.seed_to_mask("1101100111011")
    with pytest.raises(sks.SksError):
        ctx.sketch(batch, mask, w, sks.all_kmers(), sks.REPR_BITSET_ONCHIP)
    batch.close()


def test_pair_ani_reads_pinned_host_buffers_in_place(ctx):
    """sks_pair_ani on pinned, 16-byte aligned host buffers: the sketch kernel's bulk copies read the genomes in place
    (no host-to-device copy).  Lengths around every boundary of the tile fetch (first tile without history words,
    last 1-3 words by plain loads, genomes shorter than one bulk copy), every representation, dense and FracMinHash
    predicates; pageable and misaligned buffers take the copy and give the same counts."""
    import torch
    rng = np.random.default_rng(17)
    lengths = [1, 15, 16, 17, 47, 63, 64, 65, 100, 8191, 8192, 8193, 8255, 8256, 16384 + 37, 3 * 8192 - 1, 70001, 262147]

    def pinned(words):
        t = torch.empty(len(words) + 1, dtype=torch.int32).pin_memory()
        t.numpy().view(np.uint32)[: len(words)] = words
        assert t.data_ptr() % 16 == 0
        return t

    cases = [("1101100111011", sks.all_kmers(), port.ALL, ()),                       # weight 9: shared-memory bitset
             ("011101110010111110011011", sks.all_kmers(), port.ALL, ()),            # weight 16: bucketed build
             ("0011111011010111111011001011101", sks.frac_min_hash(1, 8), port.FMH, (1, 8, 181))]
    n_in_place = ctx.in_place_calls
    for li, L in enumerate(lengths):
        A = rng.integers(0, 4, L, dtype=np.uint8)
        Bm = A.copy()
        idx = rng.integers(0, L, max(L // 40, 1))
        Bm[idx] = (Bm[idx] + 1) & 3
        if li % 3 == 0:
            Bm = Bm[: max(L - 5, 1)]
        wa, wb = sks.pack_codes(A), sks.pack_codes(Bm)
        ta, tb = pinned(wa), pinned(wb)
        for seed, pred, kind, args in cases:
            mask, w = sks.seed_to_mask(seed)
            oa = port.sketch_set(A, [len(A)], mask, w, kind, *args)
            ob = port.sketch_set(Bm, [len(Bm)], mask, w, kind, *args)
            want = (len(oa), len(ob), port.intersection(oa, ob))
            reprs = (sks.REPR_SORTED,) if kind == port.FMH else (sks.REPR_BITSET, sks.REPR_BITSET_ONCHIP, sks.REPR_SORTED)
            for r in reprs:
                res = ctx.pair_ani_ptr(ta.data_ptr(), len(A), tb.data_ptr(), len(Bm), mask, w, pred, r)
                n_in_place += 1
                assert ctx.in_place_calls == n_in_place, "the pinned buffers were copied instead of read in place"
                assert (res.size_a, res.size_b, res.intersection) == want, (L, seed, r)
                # pageable memory, and a pinned buffer at a misaligned address: copied, same counts
                res = ctx.pair_ani(wa, len(A), wb, len(Bm), mask, w, pred, r)
                assert (res.size_a, res.size_b, res.intersection) == want, (L, seed, r, "pageable")
            if len(wa) > 1:
                off = pinned(np.concatenate([wa[:1], wa]))
                res = ctx.pair_ani_ptr(off.data_ptr() + 4, len(A), tb.data_ptr(), len(Bm), mask, w, pred, reprs[-1])
                assert (res.size_a, res.size_b, res.intersection) == want, (L, seed, "misaligned")
                assert ctx.in_place_calls == n_in_place


def test_sort_unique_of_raw_device_keys(ctx):
    """sks_set_from_unsorted_device_keys (the union step of the position-sharded sketch): random keys under a spaced
    mask, sizes from one key to a few hundred thousand, uniform and heavily skewed over the buckets, with duplicates --
    against numpy's sort + unique.  Exercises the bucket sort's small and large networks and its fall-backs to the library sort."""
    import torch
    rng = np.random.default_rng(23)
    mask, w = sks.seed_to_mask("0011111011010111111011001011101")
    mbits = [b for b in range(64) if (mask >> b) & 1]

    def spread(vals):   # deposit the low bits of vals at the mask's set positions (PDEP)
        out = np.zeros(len(vals), dtype=np.uint64)
        for k, b in enumerate(mbits):
            out |= ((vals >> np.uint64(k)) & np.uint64(1)) << np.uint64(b)
        return out

    for n in (1, 2, 31, 32, 33, 300, 1000, 5000, 40_000, 300_000):
        for kind in ("uniform", "skewed", "duplicates"):
            raw = rng.integers(0, 1 << len(mbits), n, dtype=np.uint64)
            if kind == "skewed":        # most keys share their top bits: one bucket far above the average
                raw[: n - n // 8] &= np.uint64((1 << (len(mbits) - 9)) - 1)
            if kind == "duplicates":
                raw = raw[rng.integers(0, max(n // 3, 1), n)]
            keys = spread(raw)
            want = np.unique(keys)
            d = torch.from_numpy(keys.view(np.int64)).cuda()
            s = ctx.set_from_device_keys(d.data_ptr(), n, 1, mask, w, sorted_unique=False)
            got = s.keys()
            assert s.kmer_set_size() == len(want), (n, kind)
            assert np.array_equal(got[:, 0], want) and not got[:, 1].any(), (n, kind)
            s.close()


def test_fasta_files_to_sets(ctx, tmp_path):
    A = port.gen(50_000, 5)
    B = port.mutate(A, 6, 50)
    port.write_fasta(str(tmp_path / "a.fna"), A)
    text = port.codes_to_text(B)
    # N runs, lower case, CRLF, blank line, second record: the section 3.6 quirks
    messy = b">b one\r\n" + text[:1000] + b"\r\n" + text[1000:3000].lower() + b"\n\n" + text[3000:20000] + \
            b"NNNN" + text[20000:] + b"\n>b2\n" + text[:777] + b"\n"
    (tmp_path / "b.fna").write_bytes(messy)
    files = [str(tmp_path / "a.fna"), str(tmp_path / "b.fna")]
    for seed, pred in ((C2_SEED, sks.all_kmers()), (C3_SEED, sks.frac_min_hash(1, 10))):
        mask, w = sks.seed_to_mask(seed)
        sets = sks.kmer_sets_from_fasta_files(ctx, files, mask, w, pred, sks.REPR_SORTED)
        osets = []
        for f in files:
            codes, segs = port.fasta_parse(open(f, "rb").read())
            osets.append(port.sketch_set(codes, list(segs), mask, w, *opred(pred)))
        for s, o in zip(sets, osets):
            assert np.array_equal(s.keys(), o)
        assert sks.kmer_set_intersection(ctx, sets[0], sets[1]) == port.intersection(osets[0], osets[1])
    one = sks.kmer_set_from_fasta_file(ctx, files[1], mask, w, pred, sks.REPR_SORTED)
    assert np.array_equal(one.keys(), osets[1])


# ---- properties at the BASELINE.json sizes -----------------------------------------------------
def test_properties_full_size(ctx):
    L = 5_000_000
    batch = ctx.synth(L, [42, 42], [0, 43], [0, 100])
    mask, w = sks.seed_to_mask(C2_SEED)
    sa, sb = ctx.sketch(batch, mask, w, sks.all_kmers(), sks.REPR_SORTED)
    ba, bb = ctx.sketch(batch, mask, w, sks.all_kmers(), sks.REPR_BITSET)
    # representation equivalence: bitset popcount == sorted-unique count, same members
    assert ba.kmer_set_size() == sa.kmer_set_size() and bb.kmer_set_size() == sb.kmer_set_size()
    assert ctx.intersect(ba, bb) == ctx.intersect(sa, sb)
    ka = sa.keys()
    assert np.array_equal(ba.keys(), ka)
    # sortedness, distinctness, idempotence, symmetry
    assert (ka[1:, 0] > ka[:-1, 0]).all() and (ka[:, 1] == 0).all()
    assert ctx.intersect(sa, sa) == sa.kmer_set_size()
    assert ctx.intersect(sa, sb) == ctx.intersect(sb, sa)
    # strand independence: the reverse complement gives the same set (src/kmer_sliding.cpp:159-175)
    A = sks.unpack_codes(batch.download(0), L)
    rc = (3 - A[::-1]).astype(np.uint8)
    (sr,) = ctx.sketch(ctx.upload_codes([rc]), mask, w, sks.all_kmers(), sks.REPR_SORTED)
    assert np.array_equal(sr.keys(), ka)
    # position sharding with a (w-1)-base halo: union of slices == whole (C3-style long sequence)
    mask3, w3 = sks.seed_to_mask(C3_SEED)
    pred = sks.frac_min_hash(1, 200)
    (whole,) = ctx.sketch(ctx.upload_codes([A]), mask3, w3, pred)
    parts = []
    n_starts = L - w3 + 1
    cuts = [0, 1_250_000 - 1_250_000 % 16, 2_500_000, 3_750_016, n_starts]
    full = ctx.upload_codes([A])
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        (p,) = ctx.sketch(full.slice(0, lo, hi - lo, w3), mask3, w3, pred)
        parts.append(p.keys())
    union = np.unique(np.concatenate(parts)[:, 0])
    assert np.array_equal(union, whole.keys()[:, 0])


def test_c3_chromosome_scale(ctx):
    """250 Mbp, weight-21 span-31 seed, FMH(200): sampled against the oracle + global properties."""
    L = 250_000_000
    batch = ctx.synth(L, [7], [0], [0])
    mask, w = sks.seed_to_mask(C3_SEED)
    pred = sks.frac_min_hash(1, 200)
    (s,) = ctx.sketch(batch, mask, w, pred)
    keys = s.keys()
    n = len(keys)
    assert abs(n - (L - w + 1) / 200) < 6 * ((L / 200) ** 0.5) + 0.001 * n   # ~1/200 of the windows survive
    assert (keys[1:, 0] > keys[:-1, 0]).all()
    # every member passes the predicate (host restatement of frac_min_hash in libsks itself is not the
    # checker: use the oracle's hash)
    for i in np.linspace(0, n - 1, 2000).astype(int):
        assert port.fmh(int(keys[i, 0]), mask, w, 1, 181) % 200 == 0
    # a 1 Mbp window of the chromosome, sketched by the oracle, must be a subset of the whole sketch
    lo = 123_456_784
    sub = sks.unpack_codes(batch.slice(0, lo, 1_000_000, w).download(0), 1_000_000 + w - 1)
    assert np.array_equal(sub[:64], port.gen(lo + 64, 7)[lo:])
    osub = port.sketch_set(sub, [len(sub)], mask, w, port.FMH, 1, 200, 181)
    assert len(osub) > 4000 and np.isin(osub[:, 0], keys[:, 0]).all()
    # full size, exactly: the 8-way position split of the whole chromosome -- the slices a run on 8 GPUs
    # sketches, a (w-1)-base halo each -- unites to the whole-sequence sketch ...
    from spaced_kmer_sketching_b200 import multi_gpu
    parts, cuts = [], []
    for r in range(8):
        first, count = multi_gpu.position_shard(L, w, r, 8)
        cuts.append(first)
        (p,) = ctx.sketch(batch.slice(0, first, count, w), mask, w, pred)
        parts.append(p.keys()[:, 0])
        p.close()
    assert np.array_equal(np.unique(np.concatenate(parts)), keys[:, 0])
    # ... 25 Mbp contiguous across the first cut are sketched identically by the oracle (every k-mer, not a sample) ...
    lo25 = cuts[1] - 12_500_000
    sl = batch.slice(0, lo25, 25_000_000, w)
    codes = sks.unpack_codes(sl.download(0), 25_000_000 + w - 1)
    (g25,) = ctx.sketch(sl, mask, w, pred)
    o25 = port.sketch_set(codes, [len(codes)], mask, w, port.FMH, 1, 200, 181)
    assert np.array_equal(g25.keys(), o25) and np.isin(o25[:, 0], keys[:, 0]).all()
    g25.close()
    # ... and so are the 400 kbp around every other cut point (windows that straddle two ranks' slices)
    for c in cuts[2:]:
        sl = batch.slice(0, c - 200_000, 400_000, w)
        codes = sks.unpack_codes(sl.download(0), 400_000 + w - 1)
        (gc,) = ctx.sketch(sl, mask, w, pred)
        oc = port.sketch_set(codes, [len(codes)], mask, w, port.FMH, 1, 200, 181)
        assert np.array_equal(gc.keys(), oc) and np.isin(oc[:, 0], keys[:, 0]).all(), c
        gc.close()


def test_c4_all_vs_all_at_genome_size(ctx):
    """BASELINE.json configs[3] in small: 24 graded mutants of one 5 Mbp genome, weight-21 span-31 seed, FMH(200):
    the n^2 intersection counts (dictionary route, row-resident and merge kernels) against numpy set intersections of
    the key arrays, two sketches against the oracle, and the rank tiling of the multi-GPU path."""
    L, n = 5_000_000, 24
    Ds = [[0, 1000, 200, 100, 50, 20][g % 6] for g in range(n)]
    batch = ctx.synth(L, [1000] * n, [2000 + g for g in range(n)], Ds)
    mask, w = sks.seed_to_mask(C3_SEED)
    pred = sks.frac_min_hash(1, 200)
    sets = ctx.sketch(batch, mask, w, pred)
    keys = [s.keys()[:, 0] for s in sets]
    assert all(abs(len(k) - L / 200) < 1500 and (k[1:] > k[:-1]).all() for k in keys)
    base = port.gen(L, 1000)
    for g in (0, 5):
        codes = base if Ds[g] == 0 else port.mutate(base, 2000 + g, Ds[g])
        assert np.array_equal(keys[g], port.sketch_set(codes, [L], mask, w, port.FMH, 1, 200, 181)[:, 0]), g
    want = np.array([[len(np.intersect1d(a, b, assume_unique=True)) for b in keys] for a in keys], dtype=np.int32)
    assert np.array_equal(ctx.intersect_all_pairs(sets), want)
    full = np.full((n, n), -1, dtype=np.int32)
    for r in range(4):   # the pairwise kernels, rank by rank
        ctx.intersect_block(sets, sks.shard_range(n, r, 4), (0, n), full)
    assert np.array_equal(full, want)
    rows = [ctx.all_vs_all(sets, *sks.shard_range(n, r, 4)) for r in range(4)]   # the dictionary route, rank by rank
    assert np.array_equal(np.concatenate([p[0] for p in rows]), want)
    # containment on the FIRST set of the ordered pair (src/kmer-sketching.cpp:198): the matrix is not symmetric
    sizes = np.array([len(k) for k in keys], dtype=np.int32)
    ani = sks.ani_from_counts(want.ravel(), np.repeat(sizes, n), sks.mask_weight(mask)).reshape(n, n)
    for g in range(1, n):
        if Ds[g]:
            assert abs(ani[0, g] - (1.0 - 1.0 / Ds[g])) < 0.002, (g, ani[0, g])
    for x in sets:
        x.close()
    batch.close()


def test_c5_multi_seed_sweep_ani_error(ctx):
    """BASELINE.json configs[4] in small: several (k+10, k) random seed masks over graded mutants of one base
    genome; the sketches match the oracle and the estimated ANI tracks the true substitution rate."""
    L, Ds = 1_000_000, [0, 1000, 200, 100, 50, 20, 1000, 200, 100, 50, 20, 500]
    n = len(Ds)
    batch = ctx.synth(L, [1000] * n, [2000 + g for g in range(n)], Ds)
    base = port.gen(L, 1000)
    for k in (12, 16, 20, 24, 28):
        w = k + 10
        mask = sks.generate_random_spaced_seed_mask(w, k)
        assert mask == port.random_mask(w, k, 0)
        pred = sks.frac_min_hash(1, 200)
        sets = ctx.sketch(batch, mask, w, pred)
        counts = ctx.intersect_all_pairs(sets)
        sizes = np.array([s.kmer_set_size() for s in sets], dtype=np.int32)
        ani = sks.ani_from_counts(counts.ravel(), np.repeat(sizes, n), k).reshape(n, n)
        for g in (0, 3, 5):   # oracle parity on three of the genomes
            codes = base if Ds[g] == 0 else port.mutate(base, 2000 + g, Ds[g])
            assert np.array_equal(sets[g].keys(), port.sketch_set(codes, [L], mask, w, port.FMH, 1, 200, 181)), (k, g)
        assert np.array_equal(counts, counts.T) and (np.diag(counts) == sizes).all()
        for g in range(1, n):   # ANI(base, mutant) vs 1 - 1/D (substitution-only, SURVEY.md 4.2 generator)
            true = 1.0 - 1.0 / Ds[g]
            # weight 12 on 1 Mbp has ~6 % chance matches (4^12 = 16.7 M k-mers), which biases the estimate upwards
            assert abs(ani[0, g] - true) < (0.008 if k == 12 else 0.004), (k, g, ani[0, g], true)


def test_output_capacity_retry_on_repetitive_genome(ctx):
    """FMH output regions are sized from the expected survivor rate; a homopolymer whose single k-mer passes
    the filter keeps EVERY window and must go through the exact-capacity retry."""
    codes = np.zeros(200_000, dtype=np.uint8)
    mask, w = sks.seed_to_mask("11001011")
    nonce = next(n for n in range(1, 200) if port.fmh(0, mask, w, n, 181) % 3 == 0)   # poly-A -> canonical key 0
    pred = sks.frac_min_hash(nonce, 3)
    batch = ctx.upload_codes([codes, port.gen(50_000, 3)])
    masked, bits = ctx.kmer_list(batch, 0, mask, w, pred)
    om, ob = port.kmers(codes, [len(codes)], mask, w, port.FMH, nonce, 3, 181, want_bits=True)
    assert len(masked) == len(codes) - w + 1 and np.array_equal(masked, om) and np.array_equal(bits, ob)
    sets = ctx.sketch(batch, mask, w, pred, sks.REPR_SORTED)
    assert sets[0].kmer_set_size() == 1 and sets[0].keys().tolist() == [[0, 0]]
    assert np.array_equal(sets[1].keys(), port.sketch_set(port.gen(50_000, 3), [50_000], mask, w, port.FMH, nonce, 3, 181))


def test_sketch_files_roundtrip(ctx, tmp_path):
    A = port.gen(80_000, 21)
    batch = ctx.upload_codes([A, port.mutate(A, 22, 30)])
    for seed, pred, r in ((C3_SEED, sks.frac_min_hash(1, 20), sks.REPR_SORTED), ("1" * 40, sks.frac_min_hash(3, 9, 171), sks.REPR_SORTED),
                          ("110101101", sks.all_kmers(), sks.REPR_BITSET)):
        mask, w = sks.seed_to_mask(seed)
        sa, sb = ctx.sketch(batch, mask, w, pred, r)
        pa = str(tmp_path / "a.sks")
        sa.save(pa, pred)
        la, lp = ctx.load_set(pa)
        assert (lp.kind, lp.nonce, lp.modulus) == (pred.kind, pred.nonce, pred.modulus)
        assert la.window == w and la.weight == sks.mask_weight(mask) and np.array_equal(la.keys(), sa.keys())
        if r == sks.REPR_SORTED:
            assert ctx.intersect(la, sb) == ctx.intersect(sa, sb)
    (tmp_path / "bad.sks").write_bytes(b"not a sketch")
    with pytest.raises(sks.SksError) as e:
        ctx.load_set(str(tmp_path / "bad.sks"))
    assert e.value.code == 5


def _check_device_fasta(ctx, texts):
    batch = ctx.batch_from_fasta_text(texts)
    assert batch.n_genomes == len(texts)
    for g, t in enumerate(texts):
        words, nb, segs = sks.fasta_parse(t)                    # host parser of libsks
        oc, osegs = port.fasta_parse(t)                         # the oracle
        assert nb == len(oc) and list(segs) == list(osegs)
        assert batch.n_bases(g) == nb, (g, t[:80])
        assert list(batch.segments(g)) == list(osegs), (g, t[:80])
        assert np.array_equal(sks.unpack_codes(batch.download(g), nb), oc), (g, t[:80])
    return batch


def test_device_fasta_parser_matches_reference_rules(ctx):
    """Device-side FASTA ingest against the oracle on every record / split quirk."""
    g = load_golden("fasta.json")
    texts = [case["text"].encode("latin1") for case in g.values()]
    texts += [b">r1\nACGTAC\nGTNNAC\n\nGGGG\n>r2\nacgtRYac\n", b">r1\r\nACGT\r\nACGT\r\n", b">\nACGT\n>ok\nAC GT\nAAAA\n>ok2\nTTTT\n",
              b"ACGT\n>x\n\nCCCC\n", b"", b"\n", b">", b">x", b">x\nACGT", b"AC GT\n>y\nAC\n\n\nGT\nA C\nTT\n>z\nGG", b">a b c\nACGT\n",
              b">x\nAAAA\nCC CC\n", b">x\nAAAA\n\nCC CC\nGG\n>y\nTT\n"]
    _check_device_fasta(ctx, texts)            # all in one batch: file boundaries must not leak state
    for t in texts:
        _check_device_fasta(ctx, [t])          # and alone
    rng = np.random.default_rng(5)
    pieces = [b"ACGT", b"acgt", b"N", b"NNNN", b" ", b"\n", b"\n\n", b"\r\n", b">", b">name\n", b">n m\n", b"R", b"GATTACA", b"\n>\n"]
    for trial in range(60):
        n_files = int(rng.integers(1, 5))
        files = []
        for _ in range(n_files):
            k = int(rng.integers(0, 60))
            files.append(b"".join(pieces[int(i)] if rng.random() < 0.5 else port.codes_to_text(rng.integers(0, 4, int(rng.integers(1, 90)), dtype=np.uint8))
                                  for i in rng.integers(0, len(pieces), k)))
        _check_device_fasta(ctx, files)


def test_device_fasta_parser_genome_sized(ctx, tmp_path):
    A = port.gen(5_000_000, 42)
    text = port.codes_to_text(A)
    fa = tmp_path / "a.fna"
    port.write_fasta(str(fa), A)
    single_line = b">one line\n" + text + b"\n"
    messy = b">m\n" + text[:1_000_000] + b"\r\n" + text[1_000_000:2_000_000].lower() + b"NNNNNNNNNN" + text[2_000_000:] + b"\n"
    batch = _check_device_fasta(ctx, [fa.read_bytes(), single_line, messy])
    fb = ctx.batch_from_fasta_files([str(fa)])
    assert fb.n_bases(0) == 5_000_000 and np.array_equal(fb.download(0), batch.download(0))
    mask, w = sks.seed_to_mask(C3_SEED)
    sets = ctx.sketch(batch, mask, w, sks.frac_min_hash(1, 200))
    want = port.sketch_set(A, [len(A)], mask, w, port.FMH, 1, 200, 181)
    assert np.array_equal(sets[0].keys(), want) and np.array_equal(sets[1].keys(), want)
    with pytest.raises(sks.SksError) as e:
        ctx.batch_from_fasta_files([str(tmp_path / "missing.fna")])
    assert e.value.code == 5


def test_alternative_routes_in_a_subprocess():
    """The merge-only intersection and the library-sort-only routes (process-wide switches, read once) against the
    same oracle results: the sorted-set tests again, in a child process with the switches set."""
    import os
    import subprocess
    import sys
    if os.environ.get("SKS_ROW_INTERSECT") in ("0", "2"):
        pytest.skip("already inside the child process")
    env = dict(os.environ, SKS_ROW_INTERSECT="0", SKS_BUCKET_SORT="0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k",
                        "fuzz_all_pairs or all_pairs_row or multi_golden or fuzz_lists or sets_golden or sort_unique"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    # the all-vs-all dictionary with the key space entered in three shares one after the other -- what three ranks do
    # before their reduce-scatter (csrc/sks_comm.cu) -- and, in another child, without the dictionary at all
    for extra in ({"SKS_DICT_PARTS": "3"}, {"SKS_DICT_INTERSECT": "0"}):
        env = dict(os.environ, SKS_ROW_INTERSECT="1", **extra)
        env["SKS_ROW_INTERSECT"] = "0" if "SKS_DICT_INTERSECT" in extra else "2"   # "2": enabled, and marks the child
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k",
                            "dictionary or fuzz_all_pairs or all_pairs_row or multi_golden"],
                           env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, (extra, r.stdout[-2000:] + r.stderr[-2000:])
