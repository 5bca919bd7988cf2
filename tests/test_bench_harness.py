"""bench.py's reference handling: the unmodified reference races in its own binning phase
(src/histogram.cc:198) — crashes and silently lost bins must be retried, never trusted and never
fatal to the bench line.  CPU only; the reference binary is replaced by a scripted stand-in."""
import os
import subprocess
import sys
import types

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _write_hist(path, bins):
    with open(path, "w") as f:
        f.write("# Title:x\n###\n")
        for i, b in enumerate(bins):
            f.write(f"{i + 1} {b}\n")


def _fake_ref(script, good_bins):
    """run_ref stand-in: `script` is a list of outcomes, one per call: 'crash', 'drop', 'ok'."""
    calls = []

    def run_ref(args, cwd=None, timeout=0):
        what = script[min(len(calls), len(script) - 1)]
        calls.append(list(args))
        out = args[args.index("-o") + 1]
        if what == "crash":
            return types.SimpleNamespace(returncode=-6, stdout="double free or corruption (fasttop)\n")
        bins = list(good_bins)
        if what == "drop":
            bins[0] -= 7  # one thread's bins lost
        _write_hist(out, bins)
        return types.SimpleNamespace(returncode=0, stdout="Counting kmers ... Time taken: 1.5s\nBining kmers ... Time taken: 0.5s\n")
    return run_ref, calls


def test_reference_retry_on_crash_and_on_lost_bins(tmp_path, monkeypatch):
    good = [10, 5, 0, 0]  # 10*1 + 5*2 = 20 instances
    fake, calls = _fake_ref(["crash", "drop", "crash", "ok"], good)
    monkeypatch.setattr(O, "run_ref", fake)
    r = bench.run_reference_hist("x.fq", 32, 1 << 20, str(tmp_path), 20)
    assert r["attempts"] == 4 and r["threads"] == 32 and len(r["rejected"]) == 3
    assert list(r["bins"]) == good and r["phases"] == {"count": 1.5, "bin": 0.5}
    assert "-t32" in calls[0] and "-H1048576" in calls[0]


def test_reference_halves_threads_then_gives_up(tmp_path, monkeypatch):
    fake, calls = _fake_ref(["crash"], [1])
    monkeypatch.setattr(O, "run_ref", fake)
    with pytest.raises(bench.ReferenceFailed):
        bench.run_reference_hist("x.fq", 4, 0, str(tmp_path), 1, max_attempts=2)
    assert [c[2] for c in calls] == ["-t4", "-t4", "-t2", "-t2", "-t1", "-t1"]
    assert not any(a.startswith("-H") for a in calls[0])  # slots == 0: the reference's default -H


def test_synthref_matches_product_generator_and_instance_count():
    """oracle/synth_ref.c restates kat_b200/csrc/synth.cuh: same bytes, so the reference arm can
    build its sample without loading libkatgpu."""
    from kat_b200 import binding as B
    for first, n, L in ((0, 3000, 150), (12345678901, 257, 150), (5, 100, 31)):
        a = O.synth_reads(1_000_000, 7, 9, first, n, L, 0.01, 0.002, threads=3)
        b = B.synth_reads_host(1_000_000, 7, 9, first, n, L, 0.01, 0.002)
        assert np.array_equal(a, b)
    s = O.synth_reads(200_000_000, 1, 2, 0, 2000, 150, 0.005, 0.0001)
    assert O.count_instances(s, 27) == O.Table.from_chunks([s.tobytes()], 27, True).total
    assert O.count_instances(np.frombuffer(b"ACGTNacgtaX", dtype=np.uint8), 3) == 2 + 3


def test_write_fastq_roundtrip(tmp_path):
    s = O.synth_reads(100_000, 1, 2, 0, 50, 40, 0.0, 0.0)
    p = tmp_path / "s.fq"
    O.write_fastq(s, 50, 40, str(p))
    data = p.read_bytes()
    stream, st = O.parse_file(data, 5)
    assert st == 0 and stream == s.tobytes()[:-1]  # the parser puts 'N' between records, none after the last


def test_reference_arm_never_loads_the_product_library(tmp_path):
    """`bench.py --impl reference` must not map libkatgpu.so."""
    if not O.have_ref():
        pytest.skip("oracle/_ref/kat not built")
    code = ("import sys, runpy; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0','--cpu-sample-reads','20000'];"
            "\ntry:\n    runpy.run_path(%r, run_name='__main__')\nexcept SystemExit:\n    pass\n"
            "print('MAPPED', any('libkatgpu' in l for l in open('/proc/self/maps')), file=sys.stderr)") % os.path.join(ROOT, "bench.py")
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert "MAPPED False" in r.stderr, r.stderr[-500:]
    import json
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and "20000 reads" in line["config"]["workload"]
    assert line["cpu_baseline"]["attempts"] >= 1 and line["e2e"]["h2d_bytes_per_step"] == 0
