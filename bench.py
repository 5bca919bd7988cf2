#!/usr/bin/env python
"""bench.py — canonical k-mers/sec (k=27) into `kat hist` on B200.

    python bench.py --gpus N --steps K --warmup W           # libkatgpu (CUDA)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path

Workload (BASELINE.json configs[1]): `kat hist -m27` on
10 Gbp of synthetic 150 bp reads (66,666,667 reads sampled from a 200 Mbp
genome, 0.5 % substitutions, 0.01 % N; seeds 1/2).  One step = one whole job:
allocate + clear the table, extract/canonicalise/insert every k-mer of the
resident base stream, walk the table into the histogram bins.

`value`  : k-mer instances / s with the base stream already resident in HBM.
`e2e`    : same job through katgpu_count_feed with the base stream in pinned
           HOST memory (H2D inside the timed region) and the bins read back.
`roofline`: K1 = extract + canonicalise + insert (bin_bases + bin_keys + tile_insert kernels on the
           default path) — algorithmic bytes / CUDA-event time of those kernels; `traffic` = their
           DRAM bytes from the committed ncu capture.
`cpu_baseline`: the unmodified reference binary (oracle/_ref/kat, built by
           oracle/ref.mk) timed on this box's host cores on a bounded sample of
           the same reads, its histogram compared bit-for-bit with the GPU's.
`cli_file`: the same sample FILE through the product's own `kat hist` binary (process start to
           histogram file written) next to the reference's wall time, output files compared byte for byte.

At N > 1 (torchrun, one rank per GPU) reads are sharded by record, every k-mer
is routed to its owner rank (hash partition), exchanged, inserted by the owner,
and the per-rank bins are summed with one all-reduce; per-GPU
work is fixed (weak scaling: 10 Gbp of reads and 200 Mbp of genome per GPU).
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

K = 27
READ_LEN = 150
GENOME_PER_GPU = 200_000_000
READS_PER_GPU = 66_666_667
SEED_G, SEED_R = 1, 2
SUB_RATE, N_RATE = 0.005, 0.0001
KMERS_PER_READ = READ_LEN - K + 1
# algorithmic bytes per k-mer instance of the extract+insert kernel:
# L/(L-k+1) bytes of base stream + one table read-modify-write (8 B key + 4 B count read + 4 B written)
B_ALG = READ_LEN / KMERS_PER_READ + 16.0
# measured DRAM bytes per k-mer instance of K1 = partition_kernel + sweep_kernel, from the ncu --set full
# captures summarised in profiles/r1_v2_partition_sweep_ncu.txt (dram__bytes_read.sum + dram__bytes_write.sum)
TRAFFIC_B_PER_KMER_PARTITIONED = 26.0
# ... and of K1 = bin_bases_kernel + bin_keys_kernel + tile_insert_kernel (profiles/r2_k1_walks_route_ncu.txt, 4.09 G k-mers):
# 9.2 (stream read once + 8 B key written) + 15.5 (key read + written) + 15.7 (key read, 64 KB tiles read + written)
TRAFFIC_B_PER_KMER_TILED = 40.4


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  The sampler process
    is started early (NVML start-up contends with the CUDA driver for a moment) and only the
    samples that arrive between mark_begin() and mark_end() are used."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.t0, self.t1 = index, None, [], None, None

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits", "-lms", "100"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

        def pump():
            for ln in self.proc.stdout:
                self.lines.append((time.perf_counter(), ln))
        self.t = threading.Thread(target=pump, daemon=True)
        self.t.start()

    def mark_begin(self):
        self.t0, self.t1 = time.perf_counter(), None

    def mark_end(self):
        self.t1 = time.perf_counter()

    def summary(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in list(self.lines):
            if self.t0 is None or ts < self.t0 or (self.t1 is not None and ts > self.t1 + 0.11):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


class ReferenceFailed(RuntimeError):
    pass


def run_reference_hist(fastq, threads, slots, workdir, expect_instances, max_attempts=6):
    """`kat hist -m27 -t<threads> [-H<slots>]` with the UNMODIFIED reference binary (oracle/_ref/kat).

    The reference has a data race of its own in the binning phase — T threads `push_back` into one
    std::vector without a lock (src/histogram.cc:198) — which at 16-32 threads regularly ends in
    `double free or corruption` and can also silently lose one thread's bins.  Neither is a result:
    a run counts only if it exits 0 AND its histogram accounts for every k-mer instance of the
    sample (sum(label*bins) == instances when the open last bin is empty).  Anything else is
    retried (same binary, same options); after `max_attempts` failures the thread count is halved.
    Returns dict(wall, phases, hist, attempts, threads, rejected)."""
    import numpy as np
    import oracle as O
    out = os.path.join(workdir, "ref.hist")
    attempts, rejected, t = 0, [], max(1, threads)
    while True:
        for _ in range(max_attempts):
            attempts += 1
            if os.path.exists(out):
                os.unlink(out)
            t0 = time.perf_counter()
            args = ["hist", f"-m{K}", f"-t{t}"] + ([f"-H{slots}"] if slots else []) + ["-o", out, fastq]
            try:
                r = O.run_ref(args, cwd=workdir, timeout=1800)
            except subprocess.TimeoutExpired:
                rejected.append(f"-t{t}: timeout")
                continue
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                rejected.append(f"-t{t}: rc {r.returncode} ({(r.stdout or '').strip().splitlines()[-1][:80] if (r.stdout or '').strip() else 'no output'})")
                continue
            try:
                lab, bins = O.read_hist(out)
            except Exception as ex:
                rejected.append(f"-t{t}: unreadable output ({ex})")
                continue
            if len(bins) == 0 or (bins[-1] == 0 and int((lab * bins).sum()) != expect_instances):
                rejected.append(f"-t{t}: histogram accounts for {int((lab * bins).sum()) if len(bins) else 0} of {expect_instances} k-mer instances")
                continue
            phases = {}
            for name, pat in (("count", r"Counting kmers.*?Time taken: ([0-9.]+)s"), ("bin", r"Bining kmers.*?Time taken: ([0-9.]+)s")):
                m = re.search(pat, r.stdout, re.S)
                if m:
                    phases[name] = float(m.group(1))
            return {"wall": wall, "phases": phases, "hist": out, "bins": bins, "attempts": attempts, "threads": t, "rejected": rejected}
        if t == 1:
            raise ReferenceFailed("reference kat hist failed %d times: %s" % (attempts, "; ".join(rejected[-3:])))
        t = max(1, t // 2)


def reference_sample(sample_reads, workdir):
    """The first `sample_reads` reads of rank 0's shard of the workload, generated on the CPU by
    oracle/synth_ref.c (no product library involved), written as 4-line FASTQ.
    Returns (fastq path, k-mer instances in it, base stream)."""
    import oracle as O
    stream = O.synth_reads(GENOME_PER_GPU, SEED_G, SEED_R, 0, sample_reads, READ_LEN, SUB_RATE, N_RATE)
    fq = os.path.join(workdir, "sample.fastq")
    O.write_fastq(stream, sample_reads, READ_LEN, fq)
    return fq, O.count_instances(stream, K), stream


def sample_slots(sample_reads):
    return 1 << max(20, (sample_reads * KMERS_PER_READ - 1).bit_length())  # no doubling for this sample


def run_product_cli_hist(fastq, slots, workdir):
    """The product's own `kat hist` binary (C++ front-end over libkatgpu) on the same FASTQ file
    with the same options: file open -> (device) parse -> count -> bin -> output file written.
    Returns (wall seconds, output path, phase dict parsed from its own output and KATGPU_TRACE lines)."""
    kat = os.path.join(ROOT, "kat_b200", "bin", "kat")
    out = os.path.join(workdir, "ours.hist")
    t0 = time.perf_counter()
    r = subprocess.run([kat, "hist", f"-m{K}", f"-H{slots}", "-o", out, fastq], cwd=workdir, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=1800, env=dict(os.environ, KATGPU_TRACE="1"))
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("kat_b200/bin/kat failed: " + r.stdout[-400:])
    ph = {}
    m = re.search(r"katgpu_init\s+([0-9.]+) ms", r.stdout)
    if m:
        ph["katgpu_init_s"] = float(m.group(1)) / 1e3  # CUDA context creation + module load of a fresh process
    m = re.search(r"Counting kmers.*?Time taken: ([0-9.]+)s", r.stdout, re.S)
    if m:
        ph["count_s"] = float(m.group(1))
    m = re.search(r"Total runtime: ([0-9.]+)s", r.stdout)
    if m:
        ph["total_runtime_s"] = float(m.group(1))
    return wall, out, ph


def bind_to_gpu_numa_node(local_rank):
    """One rank per GPU on a multi-socket host: run this process (and so allocate its pinned feed buffers)
    on the NUMA node its GPU hangs off — what `numactl --cpunodebind` would do.  Returns a description
    for the bench line; any failure leaves the affinity alone."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             stdout=subprocess.PIPE, text=True, timeout=20).stdout.strip().lower()
        bus = out[-12:] if len(out) >= 12 else out  # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return "no NUMA information for the GPU"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"rank bound to NUMA node {node} ({len(cpus)} cpus)"
        return f"NUMA node {node} has no allowed cpus"
    except Exception as ex:
        return "affinity unchanged: " + str(ex)[:80]


def multi_gpu_parity(ctx, dist, rank, world, local_rank, reads_per_rank=150_000):
    """A bounded job through the same multi-GPU engine, checked against the ORACLE (not just
    sum invariants): every rank feeds `reads_per_rank` reads of its shard of a small read set, rank 0
    compares the all-reduced hist and gcp with the oracle's single table over all the reads."""
    import numpy as np
    import torch
    from kat_b200 import binding as B
    out = {"reads": reads_per_rank * world}
    try:
        mine = B.synth_reads_host(3_000_000, 11, 12, rank * reads_per_rank, reads_per_rank, READ_LEN, SUB_RATE, N_RATE)
        d = ctx.device_alloc(mine.size + 64)
        ctx.h2d(d, mine)
        t = ctx.count_begin(K, True, 1 << 24, True)
        t.feed_device(d, mine.size).end()
        _, gb = t.hist(1, 10000, 1)
        gm = t.gcp()
        nt, nd = t.n_total, t.n_distinct
        t.free()
        ctx.device_free(d)
        ok = torch.zeros(1, dtype=torch.int64, device="cuda")
        if rank == 0:
            import oracle as O
            full = O.synth_reads(3_000_000, 11, 12, 0, reads_per_rank * world, READ_LEN, SUB_RATE, N_RATE)
            ot = O.Table.from_chunks([full], K, True)
            same = (nt, nd) == (ot.total, ot.distinct) and np.array_equal(gb, ot.hist()[1]) and np.array_equal(gm, ot.gcp())
            out.update({"kmer_instances": nt, "distinct": nd, "hist_and_gcp_equal_oracle": bool(same)})
            ok[0] = 1 if same else 0
        dist.broadcast(ok, src=0)
        assert int(ok.item()) == 1, "multi-GPU result differs from the oracle"
    except AssertionError:
        raise
    except Exception as ex:
        out["error"] = (type(ex).__name__ + ": " + str(ex))[:300]
    return out


def cpu_baseline_leg(ctx, d_bases, n_reads, threads, args, line):
    """The unmodified reference timed on this box's host cores on a bounded sample of the workload
    (the first reads of the resident stream), its histogram compared bit for bit with the GPU's; the
    same FILE through the product's `kat hist` binary (cli_file); a second reference run with the
    default -H (1e8: the table doubles on the way).  Never raises: a failure is
    reported inside the object, the bench line is kept."""
    import numpy as np
    import oracle as O
    out = {"value": None, "unit": "k-mers/s", "cores": threads, "kind": "reference"}
    if not O.have_ref():
        out["sample"] = "oracle/_ref/kat not present"
        return out
    sample = min(args.cpu_sample_reads, n_reads)
    work = tempfile.mkdtemp(prefix="katref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        stream = np.empty(sample * (READ_LEN + 1), dtype=np.uint8)
        ctx.d2h(stream, d_bases)
        fq = os.path.join(work, "sample.fastq")
        O.write_fastq(stream, sample, READ_LEN, fq)
        n_inst = O.count_instances(stream, K)
        slots = sample_slots(sample)
        out["sample"] = f"first {sample} reads ({n_inst} k-mer instances) of the workload written as 4-line FASTQ in /dev/shm"
        # the GPU's histogram of the same sample (through the host-feed path)
        t = ctx.count_begin(K, True, slots, True)
        t.feed(stream).end()
        _, gbins = t.hist(1, 10000, 1)
        assert t.n_total == n_inst, "GPU k-mer instances differ from the CPU count of the sample"
        t.free()
        r = run_reference_hist(fq, threads, slots, work, n_inst)
        out.update({"value": n_inst / r["wall"], "cores": r["threads"], "host_cores": threads,
                    "sample": out["sample"] + f"; unmodified reference `kat hist -m27 -t{r['threads']} -H{slots}` wall {r['wall']:.2f}s (phases {r['phases']})",
                    "attempts": r["attempts"], "rejected_runs": r["rejected"][:6],
                    "parity_on_sample": bool(np.array_equal(gbins, r["bins"]))})
        ref_hist_bytes = open(r["hist"], "rb").read()
        # R2 of SURVEY.md §8(d): the same FILE through the product's CLI, output FILE compared byte for byte.
        # Measured as a user would run it: no other process holds the GPU (this process gives its memory back first).
        try:
            ctx.release_cached_memory()
            run_product_cli_hist(fq, slots, work)  # first run pays the page-in of a new binary
            runs = [run_product_cli_hist(fq, slots, work) for _ in range(3)]
            cli_wall, ours, ph = min(runs, key=lambda x: x[0])
            a_ = open(ours, "rb").read()
            same = a_ == ref_hist_bytes
            line["cli_file"] = {"wall_s": cli_wall, "reference_wall_s": r["wall"], "output_file_identical": bool(same),
                                "value": n_inst / cli_wall, "unit": "k-mers/s", "wall_s_all_runs": [round(x[0], 3) for x in runs],
                                "phases": ph, "reference_phases": r["phases"],
                                "what": f"kat_b200/bin/kat hist -m27 -H{slots} on the sample FASTQ ({os.path.getsize(fq)} bytes, in /dev/shm): "
                                        "process start, CUDA context, file mapped, FASTQ parsed on the GPU, count, bin, hist file written; best of 3 "
                                        "(katgpu_init = CUDA context creation of a fresh process, the floor of any CUDA program here)"}
            if not same:
                for la, lb in zip(a_.splitlines(), ref_hist_bytes.splitlines()):
                    if la != lb:
                        line["cli_file"]["first_difference"] = [la[:120].decode(errors="replace"), lb[:120].decode(errors="replace")]
                        break
        except Exception as ex:  # never lose the bench line over the extra measurement
            line["cli_file"] = {"error": str(ex)[:300]}
        # out-of-the-box behaviour: default -H 1e8, the hash doubles while counting
        try:
            r2 = run_reference_hist(fq, r["threads"], 0, work, n_inst, max_attempts=3)
            out["default_hash_size"] = {"value": n_inst / r2["wall"], "unit": "k-mers/s", "wall_s": r2["wall"], "phases": r2["phases"],
                                        "what": "same sample, reference run with its default -H 100000000 (table doubles during counting)",
                                        "attempts": r2["attempts"]}
        except Exception as ex:
            out["default_hash_size"] = {"error": str(ex)[:200]}
    except Exception as ex:
        out["error"] = (type(ex).__name__ + ": " + str(ex))[:300]
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return out


# ---------------------------------------------------------------------------------------------------
# The other two tools of the path (BASELINE.json configs[2] and, per GPU, configs[3]); the default run
# (the driver's) is `hist` above.  Same contract: one JSON line, value = k-mer instances / s resident,
# roofline for the WALK kernel of the tool (K3 gcp_kernel / K4 comp_kernel), the
# reference timed beside it on a bounded sample with its output compared bit for bit.
WORKLOADS = {
    # C4 per GPU: 50 Gbp / 8 = 6.25 Gbp of 150 bp reads from 1 Gbp / 8 of genome, k = 21
    "gcp": {"k": 21, "reads": 41_666_667, "genome": 125_000_000, "seeds": [(4, 5)], "slots_log2": 31,
            "what": "kat gcp -m21 (BASELINE configs[3] per GPU: 50 Gbp / 8)"},
    # C3: two read sets of 10 Gbp from the same 200 Mbp genome, k = 27
    "comp": {"k": 27, "reads": 66_666_667, "genome": 200_000_000, "seeds": [(1, 2), (1, 3)], "slots_log2": 31,
             "what": "kat comp -m27 reads-vs-reads (BASELINE configs[2])"},
}


def run_reference_tool(tool, k, fastqs, threads, slots, workdir, max_attempts=4):
    """`kat gcp|comp` with the unmodified reference.  Returns dict(wall, out prefix, attempts)."""
    import oracle as O
    prefix = os.path.join(workdir, "ref_" + tool)
    attempts, last = 0, ""
    t = max(1, threads)
    while True:
        for _ in range(max_attempts):
            attempts += 1
            if tool == "gcp":
                args = ["gcp", f"-m{k}", f"-t{t}", f"-H{slots}", "-o", prefix] + fastqs
            else:
                args = ["comp", f"-m{k}", f"-t{t}", f"-H{slots}", f"-I{slots}", "-o", prefix] + fastqs
            t0 = time.perf_counter()
            try:
                r = O.run_ref(args, cwd=workdir, timeout=3600)
            except subprocess.TimeoutExpired:
                last = "timeout"
                continue
            wall = time.perf_counter() - t0
            ok = r.returncode == 0 and os.path.exists(prefix + (".mx" if tool == "gcp" else "-main.mx"))
            if ok:
                return {"wall": wall, "prefix": prefix, "attempts": attempts, "threads": t}
            last = f"rc {r.returncode}"
        if t == 1:
            raise ReferenceFailed(f"reference kat {tool} failed {attempts} times ({last})")
        t = max(1, t // 2)


def spectra_workload(args, ctx, dist, rank, world, local_rank, sampler, emit):
    import numpy as np
    import torch
    from kat_b200 import binding as B
    wl = WORKLOADS[args.workload]
    k, L = wl["k"], READ_LEN
    n_reads = args.reads if args.reads != READS_PER_GPU else wl["reads"]
    slots_log2 = args.slots_log2 or wl["slots_log2"]
    n_bytes = n_reads * (L + 1)
    hbm_peak, peak_src = measured_peaks()
    genome_len = wl["genome"] * world
    d_bases = [ctx.device_alloc(n_bytes + 64) for _ in wl["seeds"]]
    for d, (sg, sr) in zip(d_bases, wl["seeds"]):
        ctx.synth_reads_device(d, genome_len, sg, sr, rank * n_reads, n_reads, L, SUB_RATE, N_RATE)

    def count(d):
        t = ctx.count_begin(k, True, 1 << slots_log2, True)
        t.feed_device(d, n_bytes).end()
        return t

    def job():
        tabs = [count(d) for d in d_bases]
        if args.workload == "gcp":
            res = tabs[0].gcp()
        else:
            res = B.comp(tabs[0], tabs[1])
        tot = sum(t.n_total for t in tabs)
        dis = [t.n_distinct for t in tabs]
        loc = [t.info()["n_distinct"] for t in tabs]
        for t in tabs:
            t.free()
        return res, tot, dis, loc

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        job()
    barrier()
    l0 = ctx.launch_count()
    ctx.profile_enable(True)
    sampler.mark_begin()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res, n_total, n_distinct, local_distinct = job()
    barrier()
    dt = time.perf_counter() - t0
    sampler.mark_end()
    fams = ("route", "bin1", "bin2", "tile", "insert", "memset", "gcp", "comp", "hist")
    prof = {f: ctx.profile_get(f) for f in fams}
    ctx.profile_enable(False)
    if dist is not None:
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    table_bytes = (1 << slots_log2) * 8.0
    walk = args.workload
    w_ms, w_n = prof[walk]
    per_job_ms = w_ms / max(args.steps, 1)
    if walk == "gcp":
        mx = res
        assert int(mx.sum()) <= n_distinct[0]
        alg_bytes = table_bytes  # 8 B per slot of THIS rank's table, read once (our slot is 8 B; SURVEY assumed 12)
        updates = local_distinct[0]
        kernel = "gcp_kernel: one pass over the table, GC count of the reconstructed k-mer, k x 1001 matrix in shared memory (warp-aggregated atomics), one flush per CTA"
    else:
        c = res["counters"]
        assert c["hash1_total"] + c["hash2_total"] == n_total and c["hash1_distinct"] == n_distinct[0] and c["hash2_distinct"] == n_distinct[1]
        assert int(res["main"].sum()) == c["hash1_distinct"] + c["hash2_only_distinct"]
        # per distinct k-mer of the walked table 12 B sequential + 12 B random probe, both passes
        alg_bytes = 24.0 * (local_distinct[0] + local_distinct[1])
        updates = local_distinct[0] + local_distinct[1]
        kernel = ("comp_kernel x2: walk hash1 probing hash2, walk hash2 probing hash1 (one 32-byte bucket per probe); matrix corner + spectra in "
                  "shared memory, the rest global atomics")
    achieved = alg_bytes / 1e9 / (per_job_ms / 1e3) if per_job_ms else None
    traffic = (sum(table_bytes for _ in d_bases) + (32.0 * updates if walk == "comp" else 0.0))
    line = {"metric": f"canonical k-mers/sec (k={k}) into {walk}", "value": n_total * args.steps / dt, "unit": "k-mers/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"{wl['what']}: {len(d_bases)} x {n_reads * L / 1e9:.2f} Gbp synthetic {L} bp reads per GPU ({n_reads} reads, genome {wl['genome'] / 1e6:.0f} Mbp/GPU, "
                                   f"sub {SUB_RATE}, N {N_RATE}, seeds {wl['seeds']})", "k": k, "canonical": True, "read_len": L, "reads_per_gpu": n_reads,
                       "table_slots_per_gpu": 1 << slots_log2, "parallelism": f"read-sharded x{world}, k-mers stored on their owner GPU over NVLink" if world > 1 else "1 GPU",
                       "l2": "tables (17 GB each) and base streams far larger than the 126 MB L2; no explicit flush"},
            "kmer_instances_per_step": n_total, "distinct_kmers": n_distinct, "gpu_launches": ctx.launch_count() - l0, "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak if achieved else None, "traffic": None,
                         "traffic_model_bytes": traffic, "peak_source": peak_src, "algorithmic_bytes": alg_bytes,
                         "kernel_ms_per_job": per_job_ms, "kernel_launches": w_n,
                         "matrix_updates_per_s": updates / (per_job_ms / 1e3) if per_job_ms else None,
                         "what_updates_are": "one matrix (+ spectra / counter) update per distinct k-mer walked: shared-memory atomics, warp-aggregated (match.any); "
                                             "cells outside the shared corner and the final flush are 64-bit global atomics",
                         "counting_kernels_ms_per_job": {f: prof[f][0] / max(args.steps, 1) for f in ("route", "bin1", "bin2", "tile", "insert", "memset") if prof[f][1]}}}
    # ---- the reference beside it: bounded sample, output compared bit for bit
    if not args.no_cpu_baseline and world == 1 and rank == 0:
        line["cpu_baseline"] = spectra_cpu_baseline(args, ctx, d_bases, wl, k, n_reads)
    for d in d_bases:
        ctx.device_free(d)
    if rank == 0:
        emit(line)


def spectra_cpu_baseline(args, ctx, d_bases, wl, k, n_reads):
    import numpy as np
    import oracle as O
    from kat_b200 import binding as B
    import golden_util as G
    threads = os.cpu_count() or 1
    out = {"value": None, "unit": "k-mers/s", "cores": threads, "kind": "reference"}
    if not O.have_ref():
        out["sample"] = "oracle/_ref/kat not present"
        return out
    sample = min(args.cpu_sample_reads // len(d_bases), n_reads)
    work = tempfile.mkdtemp(prefix="katref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        fqs, streams, n_inst = [], [], 0
        for i, d in enumerate(d_bases):
            s = np.empty(sample * (READ_LEN + 1), dtype=np.uint8)
            ctx.d2h(s, d)
            fq = os.path.join(work, f"sample{i + 1}.fastq")
            O.write_fastq(s, sample, READ_LEN, fq)
            fqs.append(fq)
            streams.append(s)
            n_inst += O.count_instances(s, k)
        slots = 1 << max(20, (sample * (READ_LEN - k + 1) - 1).bit_length())
        tabs = [ctx.count([s], k, True, slots) for s in streams]
        r = run_reference_tool(args.workload, k, fqs, threads, slots, work)
        out.update({"value": n_inst / r["wall"], "cores": r["threads"], "attempts": r["attempts"],
                    "sample": f"first {sample} reads of each set ({n_inst} k-mer instances) as FASTQ in /dev/shm; unmodified reference `kat {args.workload} -m{k} -t{r['threads']} -H{slots}` wall {r['wall']:.2f}s"})
        if args.workload == "gcp":
            _, ref_mx = G.parse_mx_text(open(r["prefix"] + ".mx", "rb").read())
            out["parity_on_sample"] = bool(np.array_equal(tabs[0].gcp(), ref_mx))
        else:
            g = B.comp(tabs[0], tabs[1])
            _, ref_mx = G.parse_mx_text(open(r["prefix"] + "-main.mx", "rb").read())
            want, _ = G.parse_stats_text(open(r["prefix"] + ".stats", "rb").read())
            out["parity_on_sample"] = bool(np.array_equal(g["main"], ref_mx) and g["counters"] == want)
        for t in tabs:
            t.free()
    except Exception as ex:
        out["error"] = (type(ex).__name__ + ": " + str(ex))[:300]
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return out


def main():
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner)
    # are sent to stderr, the result goes to the saved descriptor.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="katgpu", choices=["katgpu", "reference"])
    ap.add_argument("--workload", default="hist", choices=["hist", "gcp", "comp"],
                    help="hist = BASELINE configs[1] (the default, what the driver runs); gcp / comp = the other two tools of the path")
    ap.add_argument("--reads", type=int, default=READS_PER_GPU, help="reads per GPU (default: the 10 Gbp workload)")
    ap.add_argument("--slots-log2", type=int, default=0, help="table slots per GPU (log2)")
    ap.add_argument("--cpu-sample-reads", type=int, default=4_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.slots_log2 == 0 and args.workload == "hist":  # 1.2 G distinct 27-mers per GPU: load 0.56
        args.slots_log2 = 31
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_reads = args.reads
    n_kmers_nominal = n_reads * KMERS_PER_READ
    workload = (f"kat hist -m{K}: {n_reads * READ_LEN / 1e9:.2f} Gbp synthetic {READ_LEN} bp reads per GPU "
                f"({n_reads} reads, genome {GENOME_PER_GPU / 1e6:.0f} Mbp/GPU, sub {SUB_RATE}, N {N_RATE}, seeds {SEED_G}/{SEED_R})")
    config = {"workload": workload, "k": K, "canonical": True, "read_len": READ_LEN,
              "reads_per_gpu": n_reads, "table_slots_per_gpu": 1 << args.slots_log2,
              "parallelism": f"read-sharded x{world}, every k-mer stored on its owner GPU over NVLink (libkatgpu multi-GPU engine), one all-reduce of the bins" if world > 1 else "1 GPU",
              "l2": "inputs (10 GB base stream, >=17 GB table) far larger than the 126 MB L2; no explicit flush"}
    threads = os.cpu_count() or 1

    # ------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        import oracle as O
        if not O.have_ref():
            emit({"impl": "reference", "unavailable": "oracle/_ref/kat not built (run oracle/ref.mk where /root/reference exists)"})
            return 0
        sample = min(args.cpu_sample_reads, n_reads)
        work = tempfile.mkdtemp(prefix="katref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        base = {"impl": "reference", "metric": "canonical k-mers/sec (k=27) into hist", "unit": "k-mers/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic"}
        try:
            fq, n_inst, _ = reference_sample(sample, work)
            slots = sample_slots(sample)
            # one probing run sizes the sample so that warmup+steps runs end within a few minutes
            probe = run_reference_hist(fq, threads, slots, work, n_inst)
            budget_s = 200.0
            if probe["wall"] * (args.warmup + args.steps) > budget_s and sample > 500_000:
                sample = max(500_000, int(sample * budget_s / (probe["wall"] * (args.warmup + args.steps))) // 100_000 * 100_000)
                fq, n_inst, _ = reference_sample(sample, work)
                slots = sample_slots(sample)
            times, attempts, rejected, used_threads = [], 0, [], threads
            for i in range(args.warmup + args.steps):
                r = run_reference_hist(fq, used_threads, slots, work, n_inst)
                attempts += r["attempts"]
                rejected += r["rejected"]
                used_threads = r["threads"]
                if i >= args.warmup:
                    times.append((r["wall"], r["phases"]))
            total = sum(t[0] for t in times)
            val = n_inst * len(times) / total
            cfg = dict(config)
            cfg["workload"] = (f"kat hist -m{K} (unmodified reference, oracle/_ref/kat) on a bounded sample of the workload: the first {sample} reads "
                               f"({sample * READ_LEN / 1e9:.2f} Gbp, {n_inst} k-mer instances) of the {n_reads}-read / {n_reads * READ_LEN / 1e9:.2f} Gbp set, "
                               f"as 4-line FASTQ in /dev/shm; -t{used_threads} -H{slots}")
            cfg["sample_reads"] = sample
            cfg["table_slots_per_gpu"] = slots
            cfg["parallelism"] = f"{used_threads} host threads"
            base.update({"value": val, "ms_per_step": 1e3 * total / len(times), "config": cfg,
                         "cpu_baseline": {"value": val, "unit": "k-mers/s", "cores": used_threads, "kind": "reference",
                                          "sample": f"first {sample} reads ({n_inst} k-mer instances) of the workload as FASTQ; "
                                                    f"`kat hist -m27 -t{used_threads} -H{slots}` wall time per run incl. process start, parse, count, bin, write; "
                                                    f"last run phases {times[-1][1]}",
                                          "runs": args.warmup + args.steps, "attempts": attempts, "rejected_runs": rejected[:8],
                                          "host_cores": threads},
                         "e2e": {"value": val, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
            emit(base)
        except Exception as ex:  # the reference could not be made to produce a valid run: say so, exit 0
            base.update({"unavailable": ("reference run failed: " + str(ex))[:300]})
            emit(base)
        finally:
            shutil.rmtree(work, ignore_errors=True)
        return 0

    # ------------------------------------------------------------ CUDA arm
    import numpy as np
    import torch
    import kat_b200
    from kat_b200 import binding as B
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libkatgpu has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = kat_b200.Context(local_rank)
    sampler = ClockSampler(local_rank)
    sampler.start()
    hbm_peak, peak_src = measured_peaks()
    if world > 1:  # this context becomes one rank of the job; every table call below is then collective
        from kat_b200.multigpu import join_job
        join_job(ctx, dist, device=torch.device("cuda", local_rank))
    if args.workload != "hist":
        spectra_workload(args, ctx, dist, rank, world, local_rank, sampler, emit)
        sampler.stop()
        ctx.close()
        if dist is not None:
            dist.destroy_process_group()
        return 0

    n_bytes = n_reads * (READ_LEN + 1)
    genome_len = GENOME_PER_GPU * world
    first_read = rank * n_reads
    d_bases = ctx.device_alloc(n_bytes + 64)
    ctx.synth_reads_device(d_bases, genome_len, SEED_G, SEED_R, first_read, n_reads, READ_LEN, SUB_RATE, N_RATE)
    base, ceil_, nb = B.hist_geometry(1, 10000)


    call_ms = {}  # host wall time per C-ABI call of the job, summed over the timed steps (where a step's time goes)

    def timed_call(name, fn, *a):
        t0 = time.perf_counter()
        r = fn(*a)
        call_ms[name] = call_ms.get(name, 0.0) + 1e3 * (time.perf_counter() - t0)
        return r

    def job_resident():
        """One whole `kat hist` job, base stream resident in HBM.  Returns (bins, n_total, n_distinct[, local n_total])."""
        t = timed_call("count_begin", ctx.count_begin, K, True, 1 << args.slots_log2, True)
        timed_call("count_feed_device", t.feed_device, d_bases, n_bytes)
        timed_call("count_end", t.end)
        _, bins = timed_call("hist", t.hist, 1, 10000, 1)
        nt, nd = t.n_total, t.n_distinct  # on a job: the job's totals
        local = timed_call("table_get_info", t.info)["n_total"]  # k-mer instances inserted into THIS rank's table
        timed_call("table_free", t.free)
        return bins, nt, nd, local

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    step_ms = []

    def timed(fn, steps, warmup):
        out = None
        for _ in range(warmup):
            out = fn()
        barrier()
        l0 = ctx.launch_count()
        ctx.profile_enable(True)
        call_ms.clear()
        step_ms.clear()
        sampler.mark_begin()
        t0 = time.perf_counter()
        for _ in range(steps):
            t1 = time.perf_counter()
            out = fn()
            step_ms.append(round(1e3 * (time.perf_counter() - t1), 2))
        barrier()
        dt = time.perf_counter() - t0
        sampler.mark_end()
        clocks = sampler.summary()
        prof = {f: ctx.profile_get(f) for f in ("count", "partition", "sweep", "bin1", "bin2", "tile", "hist", "route", "insert", "memset", "rehash")}
        ctx.profile_enable(False)
        if dist is not None:
            tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        return dt, out, ctx.launch_count() - l0, clocks, prof

    dt, (bins, n_total_all, n_distinct, n_total), launches, clocks, prof = timed(job_resident, args.steps, args.warmup)
    resident_calls = {k_: round(v / args.steps, 2) for k_, v in call_ms.items()}
    resident_steps = list(step_ms)
    value = n_total_all * args.steps / dt
    # size-independent checks of the full-size result
    labels = np.arange(base, base + nb, dtype=np.uint64)
    # KATGPU_SWEEP_EXPERIMENT runs (measurement aids with deliberately wrong results) skip the checks
    nocheck = bool(os.environ.get("KATGPU_SWEEP_EXPERIMENT"))
    assert nocheck or int(bins.sum()) == n_distinct, "sum(bins) != distinct k-mers"
    if bins[-1] == 0:
        assert nocheck or int((bins * labels).sum()) == n_total_all, "sum(label*bins) != k-mer instances"

    # K1 (extract + canonicalise + insert) is either one kernel (extract_kernel<insert>,
    # family "count") or the pair partition_kernel + sweep_kernel; its algorithmic bytes are B_ALG per k-mer.
    if world > 1:
        fam = [f for f in ("route", "bin1", "bin2", "tile", "insert") if prof[f][1] > 0]
        kname = ("route_bases_kernel (extract + hash + peer stores of every key into its owner's HBM over NVLink) + bin_hashed_kernel + bin_keys_kernel + "
                 "tile_insert_kernel on the owner: K1 + K5 of this rank")
    elif prof["bin1"][1] > 0:
        fam = [f for f in ("bin1", "bin2", "tile", "insert") if prof[f][1] > 0]
        kname = "bin_bases_kernel + bin_keys_kernel + tile_insert_kernel (+ general insert of the side list): K1 as two binning passes and shared-memory table tiles"
    elif prof["partition"][1] > 0:
        fam, kname = ["partition", "sweep"], "partition_kernel + sweep_kernel (K1 split in two streaming phases)"
    else:
        fam, kname = ["count"], "extract_kernel<insert>"
    k_ms = sum(prof[f][0] for f in fam)
    # K1 runs once per step; on the pending-key paths one run is several launches (binning passes, groups of tiles)
    k_launches = sum(prof[f][1] for f in fam)
    k_n = args.steps if fam != ["count"] else max(prof[f][1] for f in fam)
    k_inst = n_total * args.steps  # instances inserted on this rank's table
    achieved = (k_inst * B_ALG / 1e9) / (k_ms / 1e3) if k_ms > 0 else None  # GB/s; per launch: bytes/launch over ms/launch
    roofline = {"bound": "hbm", "kernel": kname,
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak if achieved else None,
                "traffic": (TRAFFIC_B_PER_KMER_PARTITIONED * k_inst / args.steps) if fam == ["partition", "sweep"] else
                           (TRAFFIC_B_PER_KMER_TILED * k_inst / args.steps) if fam[:3] == ["bin1", "bin2", "tile"] and world == 1 else None,
                "traffic_source": "ncu dram__bytes_read.sum+dram__bytes_write.sum of the K1 kernels at 5 Gbp, scaled per k-mer to one step "
                                  "(profiles/r2_k1_walks_route_ncu.txt; r1_v2_partition_sweep_ncu.txt for the L2-sweep path)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_kmer": B_ALG, "kernel_ms_total": k_ms, "kernel_launches": k_launches, "k1_passes": k_n,
                "kernel_ms_by_family": {f: prof[f][0] for f in fam},
                "kmers_per_launch": k_inst / k_n if k_n else None,
                "input_only_GBps": (k_inst * (READ_LEN / KMERS_PER_READ) / 1e9) / (k_ms / 1e3) if k_ms > 0 else None,
                "other_kernels_ms": {f: prof[f][0] for f in prof if f not in fam}}

    # K2 (hist walk): 8 B per slot read once; one histogram update per distinct k-mer
    # (warp-aggregated shared-memory atomics, flushed with 64-bit global atomics)
    reduction = None
    try:
        h_ms, h_n = prof["hist"]
        if h_n:
            per = h_ms / h_n / 1e3
            table_gb = (1 << args.slots_log2) * 8 / 1e9
            reduction = {"kernel": "hist_kernel", "ms_per_launch": h_ms / h_n, "bound": "hbm",
                         "achieved": table_gb / per, "peak": hbm_peak, "unit": "GB/s", "frac": table_gb / per / hbm_peak,
                         "histogram_updates_per_s": n_distinct / per, "slots_per_s": (1 << args.slots_log2) / per}
    except Exception as ex:  # an extra, never worth the bench line
        reduction = {"error": str(ex)[:200]}
    if numa:
        config["host_numa"] = numa
    line = {"metric": "canonical k-mers/sec (k=27) into hist", "value": value, "unit": "k-mers/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
            "kmer_instances_per_step": n_total_all, "distinct_kmers": n_distinct, "kmer_instances_rank0_table": n_total,
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "reduction": reduction,
            "host_ms_per_step_by_call": resident_calls, "step_ms": resident_steps[:32]}

    # ------------------------------------------------------------ end to end
    if not args.no_e2e:
        hp = C_void_p_host(n_bytes)
        try:
            B._check(B.lib().katgpu_memcpy_d2h(ctx.h, hp, d_bases, n_bytes))

            def job_e2e():
                t = ctx.count_begin(K, True, 1 << args.slots_log2, True)
                t.feed_ptr(hp, n_bytes).end()
                _, b2 = t.hist(1, 10000, 1)
                nt = t.n_total
                t.free()
                return b2, nt, None, None
            dt2, (bins2, nt2, _, _), _, _, _ = timed(job_e2e, args.steps, min(args.warmup, 3))
            assert np.array_equal(bins2, bins) and nt2 == n_total_all
            chunks = (n_bytes + (128 << 20) - 1) // (128 << 20)
            line["e2e"] = {"value": nt2 * args.steps / dt2, "unit": "k-mers/s",
                           "h2d_bytes_per_step": (n_bytes + 32 * (chunks - 1)) * world, "d2h_bytes_per_step": nb * 8 * world,
                           "ms_per_step": 1e3 * dt2 / args.steps,
                           "path": "katgpu_count_begin -> katgpu_count_feed(pinned host base stream, one shard per rank) -> katgpu_count_end -> katgpu_hist"}
        finally:
            B.lib().katgpu_host_free(hp)

    # ------------------------------------------------------------ multi-GPU: exchange figures + parity against the oracle
    if world > 1:
        try:
            st = ctx.dist_stats()
            r_ms = prof["route"][0]
            runs = args.steps  # profiled jobs (the resident ones)
            recv_bytes = n_total * 8.0 * (world - 1) / world  # hashed keys that reached this rank from the others, per job
            line["nvlink"] = {"key_bytes_in_per_job_rank0": recv_bytes, "route_kernel_ms_per_job": r_ms / max(runs, 1),
                              "GBps_in_during_route": recv_bytes * runs / 1e9 / (r_ms / 1e3) if r_ms else None,
                              "peak_GBps_per_direction": 900.0, "rounds_total": st["rounds"],
                              "what": "8-byte hashed keys stored by the other ranks' route_bases_kernel straight into this rank's receive area (peer-mapped HBM); "
                                      "an equal amount leaves this rank in the same kernel time"}
        except Exception as ex:
            line["nvlink"] = {"error": str(ex)[:200]}
        line["parity_multi_gpu"] = multi_gpu_parity(ctx, dist, rank, world, local_rank)

    # ------------------------------------------------------------ CPU baseline (rank 0, N=1)
    if not args.no_cpu_baseline and world == 1 and rank == 0:
        line["cpu_baseline"] = cpu_baseline_leg(ctx, d_bases, n_reads, threads, args, line)

    ctx.device_free(d_bases)
    sampler.stop()
    if rank == 0:
        emit(line)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def C_void_p_host(nbytes):
    import ctypes as C
    from kat_b200 import binding as B
    p = C.c_void_p()
    B._check(B.lib().katgpu_host_alloc(nbytes, C.byref(p)))
    return p


if __name__ == "__main__":
    sys.exit(main())
