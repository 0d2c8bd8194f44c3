#!/usr/bin/env python
"""bench.py -- headline benchmark of the ovl hot path (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--genome BP --coverage X]

Workload (config.workload): BASELINE.json configs[1] -- "5 Mbp bacterial 50x HiFi-like reads
(ovlErrorRate 0.01)": uniform-random 5 Mbp genome, reads sampled from both strands with a log-normal
length distribution (mean ~11 kb, HiFi-like), 0.1 % per-read error (sub:ins:del 4:3:3), k=22,
--minlength 500, --maxerate 0.01, one hash block x one ref block (-h 1-N -r 1-N).  Synthetic, seeded.

One "step" = one pass of the hot path over the tile: k-mer index build over the hash reads, lookup +
seed-run emission + chaining for every ref read in both orientations, banded extension of every
candidate pair, overlap records packed on the device.

  value  read-pairs aligned per second (candidate oriented pairs entering Process_Matches =
         "Kmer hits with olaps + Kmer hits without olaps" of the -s file), device pipeline only,
         reads already resident in HBM (dp4-encoded) when the timed region starts.
  e2e    the same metric through the C ABI with HOST buffers: packed reads copied host->device for
         the hash and ref side, records copied device->host, inside the timed region (the ref batch's
         upload is issued before ovlb_build_index and runs on the library's copy stream beside it).
  roofline  the longest HBM-bound kernel launch of the step (by measured time per launch) against the measured
         HBM peak; `stage_rooflines` lists every HBM-bound stage the same way.  The extension kernel is integer-ALU
         bound (no GEMM shape, no tensor cores) and is reported under `extension` (HiFi tile of the step) and
         `extension_noisy` (a small C3-like tile: 3 % read error, --maxerate 0.06, where it is >99 % of the time).
  cpu_baseline  the UNMODIFIED reference overlapInCore (oracle/_ref, built by oracle/build_ref.sh) run
         with -t <all host cores> on a bounded sample of the same workload (smaller genome, same
         coverage / read model), rank 0 only.

--impl reference times that reference binary instead (all host threads), same metric and unit.
Under torchrun (N>1) every rank owns an independent tile of the same size (Canu's own job split:
tiles share nothing), no collective on the data path; scaling is "weak".
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 22
MINLEN = 500
ERATE = 0.01
READ_ERR = 0.001
REFBIN = os.path.join(ROOT, "oracle", "_ref", "bin")


def make_workload(genome_bp, coverage, seed):
    from canu_b200 import synth
    g = synth.make_genome(genome_bp, seed=seed)
    # HiFi-like: log-normal lengths, mean ~11 kb, clipped to [3 kb, 30 kb]
    reads = synth.simulate_reads(g, coverage, 3000, 30000, READ_ERR, seed=seed + 1, lognormal=(9.25, 0.3))
    return reads


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                               "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


def reference_run(reads, threads, workdir, tag, erate=None):
    """Run the unmodified reference overlapper on `reads`; returns (seconds, pairs, overlaps)."""
    erate = ERATE if erate is None else erate
    from canu_b200 import synth
    fa = os.path.join(workdir, tag + ".fasta")
    st = os.path.join(workdir, tag + ".seqStore")
    if not os.path.exists(st):
        synth.write_fasta(fa, reads)
        subprocess.check_call([os.path.join(REFBIN, "sqStoreCreate"), "-o", st, "-minlength", "1000",
                               "-pacbio-hifi", "lib", fa], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        os.remove(fa)
    n = len(reads)
    out = os.path.join(workdir, tag + ".ovb")
    stats = os.path.join(workdir, tag + ".stats")
    cmd = [os.path.join(REFBIN, "overlapInCore"), "-t", str(threads), "-k", str(K), "--hashbits", "23",
           "--hashload", "0.8", "--hashdatalen", str(10 ** 10), "--maxerate", str(erate), "--minlength", str(MINLEN),
           "-h", "1-%d" % n, "-r", "1-%d" % n, "-o", out, "-s", stats, st]
    t0 = time.perf_counter()
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    dt = time.perf_counter() - t0
    vals = {}
    for line in open(stats):
        k, v = line.split("=")
        vals[k.strip()] = int(v)
    pairs = vals["Kmer hits without olaps"] + vals["Kmer hits with olaps"]
    return dt, pairs, vals["Total overlaps produced"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome", type=int, default=5_000_000)
    ap.add_argument("--coverage", type=float, default=50.0)
    ap.add_argument("--sample-genome", type=int, default=1_200_000, help="genome size of the CPU-baseline sample")
    ap.add_argument("--noisy-genome", type=int, default=500_000, help="genome size of the C3-like extension measurement (0 = skip)")
    ap.add_argument("--noisy-sample-genome", type=int, default=60_000, help="genome size of its CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    workload = "C2: %.1f Mbp random genome, %gx HiFi-like reads (log-normal ~11 kb, %.1f%% read error), k=%d, --maxerate %g, --minlength %d, single hash x ref tile" % (
        args.genome / 1e6, args.coverage, READ_ERR * 100, K, ERATE, MINLEN)

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        if not os.path.exists(os.path.join(REFBIN, "overlapInCore")):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/bin/overlapInCore was not built (run oracle/build_ref.sh where /root/reference exists)"}))
            return 0
        wd = tempfile.mkdtemp(prefix="ovlbench_ref_")
        try:
            reads = make_workload(args.sample_genome, args.coverage, seed=1001)
            for _ in range(max(args.warmup, 0) and 1):          # one warm-up run is enough to warm the page cache
                reference_run(reads, cores, wd, "s")
            ts, pairs = [], 0
            for _ in range(args.steps):
                dt, pairs, _ = reference_run(reads, cores, wd, "s")
                ts.append(dt)
            t = float(np.mean(ts))
            v = pairs / t
            sample = "%.2f Mbp genome x %gx (%d reads, %d bases), whole tile per step, reference overlapInCore -t %d" % (
                args.sample_genome / 1e6, args.coverage, len(reads), sum(r.size for r in reads), cores)
            print(json.dumps({
                "impl": "reference", "metric": "ovl read-pairs aligned/sec", "value": v, "unit": "read-pairs/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": {"workload": workload, "sample": sample},
                "cpu_baseline": {"value": v, "unit": "read-pairs/s", "cores": cores, "kind": "reference", "sample": sample},
                "e2e": {"value": v, "unit": "read-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            }))
        finally:
            shutil.rmtree(wd, ignore_errors=True)
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import canu_b200
    from canu_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")

    reads = make_workload(args.genome, args.coverage, seed=2001 + 17 * rank)
    n_reads = len(reads)
    total_bases = int(sum(r.size for r in reads))
    prm = api.OverlapParams(kmer_len=K, max_erate=ERATE, min_olap_len=MINLEN, max_read_len=max(r.size for r in reads))
    ov = api.Overlapper(prm, device=local_rank)
    packed = api.PackedReads(reads, first_read_id=1, min_len=MINLEN)
    del reads

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps: index build + seeding + extension (reads already in HBM)
    ov.load_hash_reads(packed)
    ov.build_index()
    ov.stage_ref_batch(packed)

    def step():
        ov.build_index()
        return ov.run_staged()

    for _ in range(args.warmup):
        step()
    ov.reset_counters()
    launches0 = ov.kernel_launches()
    stage_ms = {}
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    # CUDA events on the library's own stream (the one every kernel is launched on): ovlb_timer_start/stop
    t0 = time.perf_counter()
    ov.timer_start()
    n_rec = 0
    for _ in range(args.steps):
        n_rec = step()
        t = ov.timings()
        for k2, v2 in t.items():
            stage_ms[k2] = stage_ms.get(k2, 0.0) + v2 / args.steps
    dev_ms = ov.timer_stop()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = ov.kernel_launches() - launches0
    ctr = ov.counters()
    wall_ms = wall * 1e3
    pairs_step = ctr["pairs"] / args.steps
    cells_step = ctr["dp_cells"] / args.steps

    # ---- end to end through the C ABI with HOST buffers: packed reads are copied host->device for the hash side and
    #      again for the ref side, overlap records device->host, every step; host buffers are page-locked
    packed.pin()
    rec_host = np.zeros(max(n_rec, 1) + 1024, dtype=api.RECORD_DTYPE)
    api._check(api.load_library().ovlb_host_register(rec_host.ctypes.data, rec_host.nbytes))

    def e2e_step():
        ov.load_hash_reads(packed)                 # H2D of the hash block + encode (compute stream)
        ov.stage_ref_batch(packed)                 # H2D of the ref batch + encode on the copy stream: overlaps the index build
        ov.build_index()
        ov.run_staged()                            # waits for the ref upload, then seeding + extension
        k = ov.fetch_records_into(rec_host)        # D2H of the records
        return rec_host[:k]

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    ov.timer_start()
    for _ in range(args.steps):
        recs = e2e_step()
    e2e_dev_ms = ov.timer_stop()
    barrier()
    e2e_wall = max(time.perf_counter() - t0, e2e_dev_ms * 1e-3)   # host packing/copies count too: take the larger
    h2d = 2 * (packed.packed_bytes + n_reads * (8 + 4 + 8 + 8))
    d2h = int(recs.nbytes)

    # ---- secondary: the extension kernel where it dominates (C3-like reads: 3 % error, --maxerate 0.06), rank 0, N=1 only
    noisy = None
    if rank == 0 and world == 1 and args.noisy_genome > 0:
        from canu_b200 import synth

        def noisy_tile(genome_bp, seed):
            g = synth.make_genome(genome_bp, seed=seed)
            rd = synth.simulate_reads(g, 40, 10000, 20000, 0.03, seed=seed + 1)
            prm_n = api.OverlapParams(kmer_len=K, max_erate=0.06, min_olap_len=MINLEN, max_read_len=max(r.size for r in rd))
            ovn = api.Overlapper(prm_n, device=local_rank)
            pk = api.PackedReads(rd, first_read_id=1, min_len=MINLEN)
            ovn.load_hash_reads(pk); ovn.build_index(); ovn.stage_ref_batch(pk)
            ovn.run_staged()                                   # warm-up
            ovn.reset_counters()
            torch.cuda.synchronize()
            ovn.timer_start()
            ovn.build_index(); ovn.run_staged()
            ms = ovn.timer_stop()
            t = ovn.timings(); cn = ovn.counters()
            ovn.close()
            return rd, ms, t, cn

        rd, ms, t, cn = noisy_tile(args.noisy_genome, 3001)
        noisy = {"workload": "C3-like tile: %.2f Mbp x 40x, 10-20 kb reads, 3%% read error, --maxerate 0.06" % (args.noisy_genome / 1e6),
                 "ms_per_step": ms, "read_pairs_per_s": cn["pairs"] / (ms * 1e-3), "extend_ms": t["extend_ms"],
                 "kernel_gcells_per_s": cn["dp_cells"] / 1e9 / (t["extend_ms"] * 1e-3), "cells_per_step": cn["dp_cells"],
                 "extend_calls": cn["extend_calls"], "pairs_per_step": cn["pairs"]}
        if not args.no_cpu_baseline and os.path.exists(os.path.join(REFBIN, "overlapInCore")) and args.noisy_sample_genome > 0:
            # same read model, smaller genome; the DP cell count of the sample comes from our counters (cell counts are
            # part of the parity tests), the time from the reference binary on all host cores
            srd, sms, st, scn = noisy_tile(args.noisy_sample_genome, 3101)
            wd = tempfile.mkdtemp(prefix="ovlbench_cpun_")
            try:
                dt, sp, _ = reference_run(srd, cores, wd, "n", erate=0.06)
                noisy["cpu_baseline"] = {"gcells_per_s": scn["dp_cells"] / 1e9 / dt, "read_pairs_per_s": sp / dt, "cores": cores,
                                         "kind": "reference", "seconds": dt,
                                         "sample": "%.3f Mbp x 40x (%d reads), whole tile, reference overlapInCore -t %d; same sample on the GPU: %.1f ms" % (
                                             args.noisy_sample_genome / 1e6, len(srd), cores, sms)}
                assert sp == scn["pairs"], ("candidate-pair count differs from the reference", sp, scn["pairs"])
            finally:
                shutil.rmtree(wd, ignore_errors=True)

    tt = torch.tensor([dev_ms / args.steps, e2e_wall * 1e3 / args.steps, pairs_step, cells_step, h2d, d2h], dtype=torch.float64, device="cuda")
    if dist is not None:
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_step, e2e_ms = tmax[0].item(), tmax[1].item()
        pairs_all, cells_all = tsum[2].item(), tsum[3].item()
        h2d, d2h = int(tsum[4].item()), int(tsum[5].item())        # whole job, like `value`
    else:
        ms_step, e2e_ms, pairs_all, cells_all = tt[0].item(), tt[1].item(), pairs_step, cells_step

    if rank == 0:
        peaks, which = measured_peaks()
        # Rooflines of the HBM-bound stages (DESIGN.md section 4 states the per-unit bytes).  Stage times are CUDA-event
        # brackets on the library's stream around each kernel (group) of the step; a stage of n identical launches
        # (the radix-sort passes) is divided by n: `achieved` is algorithmic bytes PER LAUNCH over time PER LAUNCH.
        hk, rk, sh, sr = (ctr[x] / args.steps for x in ("hash_kmers", "ref_kmers", "seed_hits", "seed_runs"))
        bbits = 8                                                   # bucket bits of the bucketed index build (ovl_build_index)
        while bbits < 2 * K and (int(hk) >> bbits) > 4096:
            bbits += 1
        sort_passes = (bbits + 7) // 8
        stage_def = {   # stage: (kernel, launches, algorithmic bytes per launch)
            "index_tuples_ms": ("k_hash_tuples_compact", 1, hk * (0.5 + 12)),             # dp4 base read + (key, position) tuple write
            "index_sort_ms": ("cub::DeviceRadixSortOnesweep (one 8-bit pass over the bucket bits)", sort_passes, hk * 12 * 2),   # every pass reads and writes every 12 B tuple
            "index_table_ms": ("k_bucket_group + path sort + k_path_slots", 1, hk * (12 + 4)),    # partitioned tuples read once, grouped positions written (+ 3 x 32 B per distinct k-mer, not counted)
            "probe_ms": ("k_ref_probe", 1, rk * (0.5 + 32)),                      # dp4 base + one 32 B slot per window
            "expand_ms": ("k_expand_small + k_expand_large", 1, sr * (4 + 16 + 16)),   # occurrence + bases compared + run record written
            "sort_ms": ("cub::DeviceRadixSortOnesweep (runs)", 8, sr * 16 * 2),
            "chain_ms": ("k_pair_scatter + k_chain_pairs", 1, sr * (16 + 12 + 12 + 16)),
        }
        stage_roof = {}
        for st_name, (kname, nl, ab) in stage_def.items():
            ms_l = stage_ms.get(st_name, 0.0) / nl
            if ms_l <= 0:
                continue
            ach = ab / (ms_l * 1e-3) / 1e9
            stage_roof[st_name.replace("_ms", "")] = {"kernel": kname, "launches": nl, "ms_per_launch": round(ms_l, 4),
                                                      "achieved": round(ach, 1), "frac": round(ach / peaks["hbm_gbs"], 4)}
        dom = max(stage_roof, key=lambda k2: stage_roof[k2]["ms_per_launch"])          # longest single HBM-bound launch
        dom_all = max(("index_tuples_ms", "index_sort_ms", "index_table_ms", "probe_ms", "expand_ms", "sort_ms", "chain_ms", "extend_ms"),
                      key=lambda k2: stage_ms.get(k2, 0.0))
        # DRAM traffic of that kernel per launch from the committed ncu --set full capture of this same workload
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and args.genome == 5_000_000 and args.coverage == 50.0:
            traffic = json.load(open(tpath)).get(dom)
        roofline = {"bound": "hbm", "kernel": stage_roof[dom]["kernel"], "stage": dom, "achieved": stage_roof[dom]["achieved"],
                    "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": stage_roof[dom]["frac"], "traffic": traffic, "peak_source": which,
                    "ms_per_launch": stage_roof[dom]["ms_per_launch"], "longest_stage_of_step": dom_all.replace("_ms", ""),
                    "note": "longest HBM-bound kernel launch of the step; the extension kernel (longest stage) is integer-ALU bound and is reported under 'extension' / 'extension_noisy'"}
        line = {
            "metric": "ovl read-pairs aligned/sec", "value": pairs_all / (ms_step * 1e-3), "unit": "read-pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload, "reads_per_gpu": n_reads, "bases_per_gpu": total_bases,
                       "l2": "inputs (%.0f MB dp4 + index) exceed the 126 MB L2" % (total_bases * 1.0 / 1e6)},
            "e2e": {"value": pairs_all / (e2e_ms * 1e-3), "unit": "read-pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "stage_rooflines": stage_roof,
            "extension": {"gcells_per_s": cells_all / 1e9 / (ms_step * 1e-3),
                          "kernel_gcells_per_s": cells_step / 1e9 / (stage_ms["extend_ms"] * 1e-3) if stage_ms.get("extend_ms") else None,
                          "cells_per_step": cells_all},
            "extension_noisy": noisy,
            "stages_ms": {k2: round(v2, 3) for k2, v2 in stage_ms.items()},
            "overlaps_per_step": int(n_rec), "pairs_per_step": pairs_all, "host_wall_ms_per_step": wall_ms / args.steps,
        }
        if not args.no_cpu_baseline and os.path.exists(os.path.join(REFBIN, "overlapInCore")):
            wd = tempfile.mkdtemp(prefix="ovlbench_cpu_")
            try:
                sreads = make_workload(args.sample_genome, args.coverage, seed=1001)
                reference_run(sreads, cores, wd, "s")
                dt, sp, _ = reference_run(sreads, cores, wd, "s")
                line["cpu_baseline"] = {"value": sp / dt, "unit": "read-pairs/s", "cores": cores, "kind": "reference",
                                        "sample": "%.2f Mbp genome x %gx (%d reads), whole tile, reference overlapInCore -t %d, %.1f s" % (
                                            args.sample_genome / 1e6, args.coverage, len(sreads), cores, dt)}
            finally:
                shutil.rmtree(wd, ignore_errors=True)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    ov.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
