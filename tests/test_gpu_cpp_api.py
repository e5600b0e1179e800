"""GPU tests of the drop-in C++ API (include/kmer.hpp ... -> libsks_cpp.so -> C ABI -> CUDA kernels).

1. tests/cpp/test_cpp_api.cpp uses the API the way the reference's callers do and prints its results;
   they are compared with the CPU oracle here.
2. The reference's own, unmodified main() compiled against THIS repository's headers
   (oracle/_ref/dropin_cli, built by oracle/Makefile in the build container) and this repository's own
   driver (sks_cli) must write the same CSV, byte for byte, as the reference binary (oracle/_ref/ref_cli).
"""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import port

from conftest import ROOT

pytestmark = pytest.mark.gpu

PKG = os.path.join(ROOT, "spaced_kmer_sketching_b200")


def ival(h):
    return int(h, 16)


def write_inputs(d, n=60_000):
    A = port.gen(n, 77)
    B = port.mutate(A, 78, 40)
    fa, fb = os.path.join(d, "a.fna"), os.path.join(d, "b.fna")
    port.write_fasta(fa, A, "a")
    text = port.codes_to_text(B)
    with open(fb, "wb") as f:   # N run, lower case, blank line, second record: section 3.6 quirks
        f.write(b">b\n" + text[:10_000] + b"\n" + text[10_000:20_000].lower() + b"NNNNN" + text[20_000:45_000] +
                b"\n\n" + text[45_000:] + b"\n>b2\n" + text[:500] + b"\n")
    return fa, fb


def oracle_sets(files, mask, w, *pred):
    out = []
    for f in files:
        codes, segs = port.fasta_parse(open(f, "rb").read())
        out.append(port.sketch_set(codes, list(segs), mask, w, *pred))
    return out


def test_cpp_api_against_oracle(tmp_path):
    exe = os.path.join(PKG, "test_cpp_api")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    fa, fb = write_inputs(str(tmp_path))
    # (NCCL writes its version banner to stdout when NCCL_DEBUG is set: keep it out of the JSON)
    env = {k: v for k, v in os.environ.items() if k != "NCCL_DEBUG"}
    r = json.loads(subprocess.check_output([exe, fa, fb], timeout=300, env=env))
    assert ival(r["mask_24_16"]) == port.random_mask(24, 16, 0)
    assert ival(r["mask_31_21_s3"]) == port.random_mask(31, 21, 3)
    assert ival(r["contig_33"]) == port.contiguous_mask(33) and r["contig_65_throws"]
    assert ival(r["seed_mask"]) == 0xF0CF and r["seed_weight"] == 5 and r["seed_mask_printed_len"] == 128
    # ordered list incl. kmer_bits on two segments (the empty string in between yields nothing)
    codes = np.array([0, 0, 0, 1, 2, 3, 0, 1, 2, 3, 3, 3, 3, 2, 1, 0, 0, 1, 2, 3, 1], dtype=np.uint8)
    om, ob = port.kmers(codes, [12, 9], 0xF0CF, 8, want_bits=True)
    to_int = lambda a: [int(x[0]) | (int(x[1]) << 64) for x in a]
    assert [ival(x) for x in r["list_masked"]] == to_int(om)
    assert [ival(x) for x in r["list_bits"]] == to_int(ob)
    assert r["legacy_ok"]

    mask, w = port.random_mask(24, 16, 0), 24
    expect = {
        "all": ("device:all", oracle_sets([fa, fb], mask, w)),
        "fmh_struct": ("device:fmh", oracle_sets([fa, fb], mask, w, port.FMH, 1, 50, 181)),
        # opaque callables: on the host by default, on the device once probing is switched on (opt-in)
        "driver": ("host", oracle_sets([fa, fb], mask, w, port.FMH, 1, 200, 181)),
        "lambda_fmh7": ("host", oracle_sets([fa, fb], mask, w, port.FMH, -2, 7, 181)),
        "driver_probed": ("device:fmh", oracle_sets([fa, fb], mask, w, port.FMH, 1, 200, 181)),
        "lambda_fmh7_probed": ("device:fmh", oracle_sets([fa, fb], mask, w, port.FMH, -2, 7, 181)),
    }
    # a counting callable sees exactly one call per real k-mer unless probing was asked for
    ca0, sa0 = port.fasta_parse(open(fa, "rb").read())
    n_windows = sum(max(int(x) - w + 1, 0) for x in sa0)
    cnt = r["counting"]
    assert cnt["path_default"] == "host" and cnt["calls_default"] == n_windows
    assert cnt["path_probe"] == "device:fmh" and cnt["calls_probe"] > 0 and cnt["calls_probe"] != n_windows
    assert cnt["size_default"] == cnt["size_probe"] == len(expect["driver"][1][0]) and cnt["same"]
    par = []
    for s in oracle_sets([fa, fb], mask, w):   # opaque condition: even popcount of masked_bits, run on the host
        keep = np.array([bin(int(k[0])).count("1") % 2 == 0 for k in s], dtype=bool)
        par.append(s[keep])
    expect["parity"] = ("host", par)
    for name, (path, sets) in expect.items():
        got = r["cases"][name]
        assert got["path"] == path, name
        assert got["sizes"] == [len(sets[0]), len(sets[1])], name
        want = [port.intersection(sets[i], sets[j]) for i in (0, 1) for j in (0, 1)]
        assert got["inter"] == want, name
        ani = [port.ani(want[2 * i + j], len(sets[i]), 16) for i in (0, 1) for j in (0, 1)]
        assert max(abs(a - b) for a, b in zip(got["ani"], ani)) <= 1e-12, name

    # eight worker threads, each with its own implicit context, then intersections from other threads
    th = r["threads"]
    for base, name in ((0, "all"), (4, "fmh_struct")):
        sets = expect[name][1]
        assert th["sizes"][base:base + 4] == [len(sets[0]), len(sets[1])] * 2, name
        assert th["inter"][base:base + 4] == [port.intersection(sets[0], sets[1])] * 4, name

    # several GPUs in one process: equal to the single-device results, which equal the oracle's
    mu = r["multi"]
    fsets = expect["fmh_struct"][1]
    five = [fsets[0], fsets[1], fsets[1], fsets[0], fsets[1]]
    assert mu["sizes_1"] == [len(x) for x in five]
    assert mu["all_1"] == [port.intersection(a, b) for a in five for b in five]
    assert mu["ring_1"] == [port.intersection(five[i], five[(i + 1) % 5]) for i in range(5)]
    if mu["devices"] >= 2:
        assert (mu["all_n"], mu["ring_n"], mu["sizes_n"]) == (mu["all_1"], mu["ring_1"], mu["sizes_1"])

    m9, _ = port.seed_to_mask("110101101")
    ca, sa = port.fasta_parse(open(fa, "rb").read())
    lst = port.kmers(ca, list(sa), m9, 9)
    uniq = port.sort_unique(lst)
    assert r["host_list_len"] == len(lst) and r["n_strings"] == len(sa)
    assert r["host_set_size"] == r["dev_set_size"] == r["walked"] == r["found"] == r["host_dev_inter"] == len(uniq)
    assert r["different_mask_inter"] == 0 and r["empty_inter"] == 0 and r["empty_size"] == 0
    assert r["mismatch_throws"] and r["containment_0"] == 0 and r["estimator_neg"] == 0


def test_reference_main_on_our_headers_writes_the_same_csv(tmp_path):
    ref_cli = os.path.join(ROOT, "oracle", "_ref", "ref_cli")
    dropin = os.path.join(ROOT, "oracle", "_ref", "dropin_cli")
    ours = os.path.join(PKG, "sks_cli")
    if not (os.path.exists(ref_cli) and os.path.exists(dropin)):
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    d = str(tmp_path)
    fa, fb = write_inputs(d, 40_000)
    fc = os.path.join(d, "c.fna")
    port.write_fasta(fc, port.mutate(port.gen(40_000, 77), 79, 15), "c")
    outs = {}
    # the reference's main() passes a free function as the sketching condition: once on the default route (evaluated on
    # the host) and once with probing switched on (recognised, run on the device)
    for name, exe, env in (("ref", ref_cli, {}), ("dropin", dropin, {"SKS_PREDICATE_PROBE": "1"}), ("dropin_host", dropin, {}),
                           ("ours", ours, {})):
        csv = os.path.join(d, name + ".csv")
        log = subprocess.check_output([exe, csv, fa, fb, fc], timeout=900, env=dict(os.environ, **env)).decode()
        assert log.count("Time taken for sketching") == 62 and log.count("Time taken for comparison") == 62
        outs[name] = open(csv).read()
    assert outs["ref"].count("\n") == 1 + 62 * 9
    assert outs["dropin"] == outs["ref"]
    assert outs["dropin_host"] == outs["ref"]
    assert outs["ours"] == outs["ref"]
