#!/usr/bin/env python
"""bench.py -- headline benchmark of the ovl hot path (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--chroms C --chrom-bp BP]

Workload (config.workload): the read model of BASELINE.json configs[2] -- the one north_star quotes its target on
("C. elegans-size 40x trimmed CLR-like reads, ovlErrorRate 0.06") -- as ONE overlap job sized so that a step fits the
bench's time budget: a random genome of C chromosomes x BP bases (default 24 x 100 kbp = 2.4 Mbp; the full C3 is 100
Mbp, the same job 42x longer: cost is linear in genome size at fixed coverage), 40x coverage, reads 10-20 kb from both
strands, 3 % per-read error (sub:ins:del 4:3:3), read order shuffled; k=22, --minlength 500, --maxerate 0.06,
`-h 1-N -r 1-N`.  Synthetic, seeded: every rank (and the reference arm) regenerates the same reads.

One "step" = one pass of the hot path over the whole job: k-mer index build over the hash reads, lookup + seed-run
emission + chaining for every ref read in both orientations, banded extension of every candidate pair, overlap records
packed on the device, and the records of all ranks merged on rank 0.

--gpus N (torchrun, one rank per GPU): STRONG scaling of that one job, the way north_star describes it.  The job has
one hash block, so every GPU indexes it and the ref range is cut into N contiguous tiles of equal
estimated work (`ovlb_plan_balanced`: only refID < hashID pairs are computed, so equal-base tiles are unequal work),
assigned with `ovlb_assign_tiles`.  No collective on the data path; the only exchange is the gather of the 24-byte
records on rank 0 (the reference's "merged on the host"), inside the timed region.

  value  read-pairs aligned per second (candidate oriented pairs entering Process_Matches = "Kmer hits with olaps +
         Kmer hits without olaps" of the -s file) of the whole job, reads already resident in HBM (dp4-encoded) when
         the timed region starts; device time (CUDA events), max over ranks.
  e2e    the same through the C ABI with HOST buffers: every step copies the packed hash block and the rank's ref tile
         host->device (pinned), and the merged records device->host on rank 0, inside the timed region.
  parity rank 0 also runs the CPU-baseline sample (all reads of chromosome 0, a complete job of the same read model and
         a strict subset of the reads the GPU arm times) on the GPU and compares the records with the reference
         binary's: "identical" = same canonically sorted 24-byte records and same -s counters.
  records_sha256  hash of the canonically sorted records of the whole job: equal at every N.
  cpu_baseline / --impl reference: the UNMODIFIED reference overlapInCore (oracle/_ref, built by oracle/build_ref.sh),
         -t <all host cores>, on that same sample.
  hifi_tile (rank 0): the C2 tile of BASELINE.json configs[1] (5 Mbp x 50x HiFi-like, --maxerate 0.01), where the
         index build and the lookup are a large share of the step: per-stage HBM rooflines; `roofline` is its longest
         HBM-bound launch.  The job's own dominant kernel, k_extend_pairs, is integer-ALU / issue bound (no GEMM shape,
         no tensor cores): reported under `extension` as DP Gcell/s with the warp-busy fraction of the launch.
"""
import argparse
import ctypes as C
import hashlib
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
ERATE = 0.06            # the job: C3 read model
READ_ERR = 0.03
COVERAGE = 40.0
LEN_LO, LEN_HI = 10000, 20000
HIFI_ERATE, HIFI_READ_ERR = 0.01, 0.001   # the secondary C2 tile
REFBIN = os.path.join(ROOT, "oracle", "_ref", "bin")
OURBIN = os.path.join(ROOT, "canu_b200", "bin")


def make_job(n_chroms, chrom_bp, seed=7001):
    """(job reads in shuffled order, reads of chromosome 0 in generation order)."""
    from canu_b200 import synth
    per_chrom = []
    for c in range(n_chroms):
        g = synth.make_genome(chrom_bp, seed=seed + 10 * c)
        per_chrom.append(synth.simulate_reads(g, COVERAGE, LEN_LO, LEN_HI, READ_ERR, seed=seed + 10 * c + 1))
    allr = [r for rs in per_chrom for r in rs]
    perm = np.random.default_rng(seed + 5).permutation(len(allr))
    return [allr[i] for i in perm], per_chrom[0]


def make_hifi_tile(genome_bp, coverage, seed):
    from canu_b200 import synth
    g = synth.make_genome(genome_bp, seed=seed)
    return synth.simulate_reads(g, coverage, 3000, 30000, HIFI_READ_ERR, seed=seed + 1, lognormal=(9.25, 0.3))


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


def pick_threads(n_reads, cores):
    """A -t for which the reference does not drop the last ref read."""
    for t in range(cores, 0, -1):
        per = 1 + (n_reads - 1) // t // 8
        if (n_reads - 1) % per != 0:
            return t
    return 1


def reference_run(reads, cores, workdir, tag, erate, tech="-pacbio"):
    """Run the unmodified reference overlapper on `reads` as one job; returns (seconds, pairs, stats dict, ovb path)."""
    from canu_b200 import synth
    fa = os.path.join(workdir, tag + ".fasta")
    st = os.path.join(workdir, tag + ".seqStore")
    if not os.path.exists(st):
        synth.write_fasta(fa, reads)
        subprocess.check_call([os.path.join(REFBIN, "sqStoreCreate"), "-o", st, "-minlength", "1000",
                               tech, "lib", fa], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        os.remove(fa)
    n = len(reads)
    out = os.path.join(workdir, tag + ".ovb")
    stats = os.path.join(workdir, tag + ".stats")
    cmd = [os.path.join(REFBIN, "overlapInCore"), "-t", str(pick_threads(n, cores)), "-k", str(K), "--hashbits", "23",
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
    return dt, pairs, vals, out


def read_ovb(path):
    """Records of a reference-written .ovb as a structured array (through our ovltool's reader)."""
    from canu_b200 import api
    txt = subprocess.check_output([os.path.join(OURBIN, "ovltool"), "dump-ovb", path]).decode().split()
    a = np.zeros(len(txt) // 4, dtype=api.RECORD_DTYPE)
    a["a_iid"] = np.array(txt[0::4], dtype=np.uint64)
    a["b_iid"] = np.array(txt[1::4], dtype=np.uint64)
    a["w0"] = np.array([int(x, 16) for x in txt[2::4]], dtype=np.uint64)
    a["w1"] = np.array([int(x, 16) for x in txt[3::4]], dtype=np.uint64)
    return a


def canon(recs):
    return np.sort(recs, order=["a_iid", "b_iid", "w0", "w1"])


def sample_text(args, n, bases, cores):
    return ("all %d reads (%d bases) of chromosome 0 of the job's genome (%.0f kbp x %gx, same read model, a subset of "
            "the reads the GPU arm times), run as one complete job -h 1-%d -r 1-%d, reference overlapInCore -t %d" % (
                n, bases, args.chrom_bp / 1e3, COVERAGE, n, n, cores))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chroms", type=int, default=24, help="chromosomes of the job's genome")
    ap.add_argument("--chrom-bp", type=int, default=100_000, help="bases per chromosome (chromosome 0 is the CPU sample)")
    ap.add_argument("--ref-runs", type=int, default=2, help="cap on the timed reference runs (each is the same deterministic sample)")
    ap.add_argument("--hifi-genome", type=int, default=5_000_000, help="genome of the secondary C2 tile (0 = skip)")
    ap.add_argument("--hifi-coverage", type=float, default=50.0)
    ap.add_argument("--hifi-steps", type=int, default=5)
    ap.add_argument("--e2e-steps", type=int, default=5, help="cap on the steps of the end-to-end (host buffers) timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--emulate", default="", help="tuning only: 'W:r' runs rank r's tile of a W-way split on this one GPU (no NCCL); not a bench line")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    workload = ("C3 read model as one job: %d x %.0f kbp random genome (%.2f Mbp), %gx, %d-%d kb reads, %.0f%% read error, "
                "shuffled; k=%d, --maxerate %g, --minlength %d, -h 1-N -r 1-N; one hash block, ref range cut into n_gpus "
                "cost-balanced tiles" % (args.chroms, args.chrom_bp / 1e3, args.chroms * args.chrom_bp / 1e6, COVERAGE,
                                         LEN_LO // 1000, LEN_HI // 1000, READ_ERR * 100, K, ERATE, MINLEN))
    config = {"workload": workload, "l2": "inputs (dp4 reads + index) exceed the 126 MB L2"}
    have_ref = os.path.exists(os.path.join(REFBIN, "overlapInCore"))

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        if not have_ref:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/bin/overlapInCore was not built (run oracle/build_ref.sh where /root/reference exists)"}))
            return 0
        wd = tempfile.mkdtemp(prefix="ovlbench_ref_")
        try:
            _, sample = make_job(1, args.chrom_bp)
            if args.warmup > 0:
                reference_run(sample, cores, wd, "s", ERATE)            # one warm-up run: page cache, store built
            ts, pairs = [], 0
            runs = max(1, min(args.steps, args.ref_runs))
            for _ in range(runs):
                dt, pairs, _, _ = reference_run(sample, cores, wd, "s", ERATE)
                ts.append(dt)
            t = float(np.mean(ts))
            v = pairs / t
            stxt = sample_text(args, len(sample), int(sum(r.size for r in sample)), cores) + (
                "; %d timed runs of %.1f s (capped: the run is deterministic), %d pairs per run" % (runs, t, pairs))
            print(json.dumps({
                "impl": "reference", "metric": "ovl read-pairs aligned/sec", "value": v, "unit": "read-pairs/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": v, "unit": "read-pairs/s", "cores": cores, "kind": "reference", "sample": stxt},
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
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    reads, sample = make_job(args.chroms, args.chrom_bp)
    n_reads = len(reads)
    lens = [int(r.size) for r in reads]
    total_bases = int(sum(lens))
    max_len = max(lens)

    # the job's tile grid: one hash block (everything fits HBM), ref range cut into `world` cost-balanced tiles
    plan_world, plan_rank = world, rank
    if args.emulate:
        plan_world, plan_rank = (int(x) for x in args.emulate.split(":"))
        args.hifi_genome = 0; args.no_cpu_baseline = True
    tiles = api.plan_balanced(lens, MINLEN, plan_world)
    owner = api.assign_tiles(tiles, plan_world)
    mine = [t for t, o in zip(tiles, owner) if o == plan_rank]
    assert len(mine) <= 1, "one launch per GPU and hash block"
    prm = api.OverlapParams(kmer_len=K, max_erate=ERATE, min_olap_len=MINLEN, max_read_len=max_len)
    ov = api.Overlapper(prm, device=local_rank)
    hpacked = api.PackedReads(reads, first_read_id=1, min_len=MINLEN)
    if mine:
        rb, re_ = mine[0]["ref_bgn"], mine[0]["ref_end"]
    else:
        rb, re_ = 1, 0
    rpacked = api.PackedReads(reads[rb - 1:re_], first_read_id=rb, min_len=MINLEN)
    del reads

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    rec_cap = [1 << 16]
    rec_dev = [torch.empty(rec_cap[0] * 24, dtype=torch.uint8, device=dev)]
    merged = {}

    def fetch_to_device(n):
        if n > rec_cap[0]:
            rec_cap[0] = int(n * 1.25) + 1024
            rec_dev[0] = torch.empty(rec_cap[0] * 24, dtype=torch.uint8, device=dev)
        k = C.c_uint64()
        api._check(ov.L.ovlb_fetch_records(ov._h, rec_dev[0].data_ptr(), rec_cap[0], C.byref(k)))
        return k.value

    def merge_on_rank0(n):
        """Gather the ranks' records on rank 0 (device memory); returns (tensor of 24-byte rows, count) on rank 0."""
        if dist is None:
            return rec_dev[0], n
        cnt = torch.tensor([n], dtype=torch.int64, device=dev)
        cnts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(cnts, cnt)
        cl = [int(c.item()) for c in cnts]
        mx = max(max(cl), 1)
        if mx > rec_cap[0]:
            old = rec_dev[0]
            rec_cap[0] = int(mx * 1.25) + 1024
            rec_dev[0] = torch.empty(rec_cap[0] * 24, dtype=torch.uint8, device=dev)
            rec_dev[0][: n * 24] = old[: n * 24]
        send = rec_dev[0][: mx * 24]
        if rank == 0:
            key = ("g", mx)
            if key not in merged:
                merged.clear()
                merged[key] = [torch.empty(mx * 24, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.gather(send, merged[key], dst=0)
            out = torch.cat([merged[key][r][: cl[r] * 24] for r in range(world)])
            return out, sum(cl)
        dist.gather(send, None, dst=0)
        return None, 0

    # ---- device-resident steps: index build + seeding + extension + merge (reads already in HBM)
    ov.load_hash_reads(hpacked)
    ov.build_index()
    ov.stage_ref_batch(rpacked)

    def step():
        ov.build_index()
        n = ov.run_staged() if mine else 0
        n = fetch_to_device(n) if n else 0
        return merge_on_rank0(n)

    for _ in range(args.warmup):
        step()
    ov.reset_counters()
    launches0 = ov.kernel_launches()
    stage_ms = {}
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    lib_ms = 0.0
    for _ in range(args.steps):
        out_t, out_n = step()
        t = ov.timings()
        lib_ms += t["index_tuples_ms"] + t["index_sort_ms"] + t["index_table_ms"] + t["index_skip_ms"] + (t["total_ms"] if mine else 0.0)
        for k2, v2 in t.items():
            stage_ms[k2] = stage_ms.get(k2, 0.0) + v2 / args.steps
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)          # device clock around the K steps (library stream work is synchronous inside it)
    clocks = sampler.stop()
    launches = ov.kernel_launches() - launches0
    ctr = ov.counters()
    pairs_step = ctr["pairs"] / args.steps
    cells_step = ctr["dp_cells"] / args.steps

    job_hash, n_job_recs = None, 0
    if rank == 0:
        recs = np.frombuffer(out_t[: out_n * 24].cpu().numpy().tobytes(), dtype=api.RECORD_DTYPE)
        n_job_recs = int(recs.size)
        job_hash = hashlib.sha256(canon(recs).tobytes()).hexdigest()[:16]

    # ---- end to end through the C ABI with HOST buffers
    hpacked.pin()
    rpacked.pin()
    host_out = {"buf": None}

    def e2e_step():
        ov.load_hash_reads(hpacked)                # H2D of the hash block + encode
        if mine:
            ov.stage_ref_batch(rpacked)            # H2D of the ref tile on the copy stream: overlaps the index build
        ov.build_index()
        n = ov.run_staged() if mine else 0
        n = fetch_to_device(n) if n else 0
        t_, n_ = merge_on_rank0(n)
        if rank == 0:
            if host_out["buf"] is None or host_out["buf"].numel() < n_ * 24:
                host_out["buf"] = torch.empty(int(n_ * 24 * 1.25) + 1024, dtype=torch.uint8, pin_memory=True)
            host_out["buf"][: n_ * 24].copy_(t_[: n_ * 24], non_blocking=True)     # D2H of the merged records
            torch.cuda.synchronize()
        return n_

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    n_e2e = 0
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(e2e_steps):
        n_e2e = e2e_step()
    ev1.record()
    barrier()
    e2e_ms = max((time.perf_counter() - t0) * 1e3, ev0.elapsed_time(ev1)) * args.steps / e2e_steps     # host packing/copies count too: take the larger; scaled to K steps
    h2d = hpacked.packed_bytes + hpacked.n_reads * 28 + (rpacked.packed_bytes + rpacked.n_reads * 28 if mine else 0)
    d2h = n_e2e * 24 if rank == 0 else 0

    tt = torch.tensor([dev_ms / args.steps, e2e_ms / args.steps, pairs_step, cells_step, h2d, d2h, launches,
                       ctr["ext_busy_ns"], ctr["ext_capacity_ns"], stage_ms.get("extend_ms", 0.0), lib_ms / args.steps],
                      dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        per_rank = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(per_rank, tt)
    else:
        tmax, tsum, per_rank = tt, tt, [tt]
    ms_step, e2e_step_ms = tmax[0].item(), tmax[1].item()
    pairs_all, cells_all = tsum[2].item(), tsum[3].item()

    if rank != 0:
        ov.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peaks, which = measured_peaks()
    line = {
        "metric": "ovl read-pairs aligned/sec", "value": pairs_all / (ms_step * 1e-3), "unit": "read-pairs/s",
        "n_gpus": world, **({"emulated_rank_of_world": args.emulate} if args.emulate else {}), "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config,
        "e2e": {"value": pairs_all / (e2e_step_ms * 1e-3), "unit": "read-pairs/s",
                "h2d_bytes_per_step": int(tsum[4].item()), "d2h_bytes_per_step": int(tsum[5].item()),
                "timed_steps": max(1, min(args.steps, args.e2e_steps))},
        "gpu_launches": int(tsum[6].item()),
        "clocks": clocks,
        "job": {"reads": n_reads, "bases": total_bases, "pairs_per_step": pairs_all, "overlaps_per_step": n_job_recs,
                "dp_cells_per_step": cells_all, "records_sha256": job_hash,
                "tiles": [{k2: t[k2] for k2 in ("hash_bgn", "hash_end", "ref_bgn", "ref_end", "ref_bases")} | {"owner": o, "cost": round(t["cost"])}
                          for t, o in zip(tiles, owner)],
                "per_rank_ms": [round(p[0].item(), 2) for p in per_rank],
                "per_rank_extend_ms": [round(p[9].item(), 2) for p in per_rank],
                "per_rank_library_ms": [round(p[10].item(), 2) for p in per_rank],
                "host_wall_ms_per_step": wall * 1e3 / args.steps},
        "extension": {"kernel": "k_extend_pairs", "bound": "int32 ALU / issue (integer DP: no tensor cores, 0.25 B/cell to HBM)",
                      "gcells_per_s": cells_all / 1e9 / (ms_step * 1e-3),
                      "kernel_gcells_per_s_per_gpu": [round(p[3].item() / 1e9 / (p[9].item() * 1e-3), 2) if p[9].item() > 0 else None for p in per_rank],
                      "warp_busy_frac": round(tsum[7].item() / tsum[8].item(), 4) if tsum[8].item() > 0 else None,
                      "note": "warp_busy_frac = time the persistent warps spent between their first and last pair / (launched warps x launch duration): 1 - tail loss"},
        "stages_ms": {k2: round(v2, 3) for k2, v2 in stage_ms.items()},
    }
    ov.close()

    # ---- parity + CPU baseline: the sample (chromosome 0's reads) through the GPU path and through the reference binary
    wd = tempfile.mkdtemp(prefix="ovlbench_cpu_")
    try:
        if world > 1:
            raise StopIteration
        sm = max(r.size for r in sample)
        sprm = api.OverlapParams(kmer_len=K, max_erate=ERATE, min_olap_len=MINLEN, max_read_len=sm)
        srecs, sctr = api.overlap_in_core(sample, sprm, device=local_rank)
        stxt = sample_text(args, len(sample), int(sum(r.size for r in sample)), cores)
        if have_ref and not args.no_cpu_baseline:
            dt, sp, svals, ovb = reference_run(sample, cores, wd, "s", ERATE)
            want = canon(read_ovb(ovb))
            got = canon(srecs)
            same_recs = want.size == got.size and want.tobytes() == got.tobytes()
            same_stats = (svals["Kmer hits without olaps"] == sctr["kmer_hits_without_olap"] and
                          svals["Kmer hits with olaps"] == sctr["kmer_hits_with_olap"] and
                          svals["Multiple overlaps/pair"] == sctr["multi_overlap"] and
                          svals["Contained overlaps"] == sctr["contained"] and svals["Dovetail overlaps"] == sctr["dovetail"] and
                          svals["Total overlaps produced"] == sctr["total_overlaps"])
            line["parity"] = "identical" if (same_recs and same_stats) else "DIFFERENT"
            line["parity_detail"] = {"sample_records": int(got.size), "reference_records": int(want.size), "records_identical": bool(same_recs),
                                     "stats_identical": bool(same_stats), "sample_pairs": sp,
                                     "sha256_gpu": hashlib.sha256(got.tobytes()).hexdigest()[:16],
                                     "sha256_reference": hashlib.sha256(want.tobytes()).hexdigest()[:16]}
            line["cpu_baseline"] = {"value": sp / dt, "unit": "read-pairs/s", "cores": cores, "kind": "reference",
                                    "sample": stxt + "; %.1f s, %d pairs" % (dt, sp),
                                    "gcells_per_s": sctr["dp_cells"] / 1e9 / dt}
        else:
            line["parity"] = "unchecked (reference binary not built)" if not have_ref else "unchecked (--no-cpu-baseline)"
            line["cpu_baseline"] = None
    except StopIteration:
        line["parity"] = "records_sha256 equals the N=1 line's (the sample check against the reference binary runs at N=1)"
        line["cpu_baseline"] = None
    finally:
        shutil.rmtree(wd, ignore_errors=True)

    # ---- secondary: the C2 tile (HiFi-like reads), where index build and lookup matter: per-stage HBM rooflines
    roofline = None
    if args.hifi_genome > 0:                         # rank 0, every N: the other ranks wait at the final barrier
        hreads = make_hifi_tile(args.hifi_genome, args.hifi_coverage, seed=2001)
        hprm = api.OverlapParams(kmer_len=K, max_erate=HIFI_ERATE, min_olap_len=MINLEN, max_read_len=max(r.size for r in hreads))
        hov = api.Overlapper(hprm, device=local_rank)
        hp = api.PackedReads(hreads, first_read_id=1, min_len=MINLEN)
        hbases = int(sum(r.size for r in hreads))
        del hreads
        hov.load_hash_reads(hp); hov.build_index(); hov.stage_ref_batch(hp)
        for _ in range(3):
            hov.build_index(); hov.run_staged()
        hov.reset_counters()
        hst = {}
        torch.cuda.synchronize()
        hov.timer_start()
        for _ in range(args.hifi_steps):
            hov.build_index(); hn = hov.run_staged()
            t = hov.timings()
            for k2, v2 in t.items():
                hst[k2] = hst.get(k2, 0.0) + v2 / args.hifi_steps
        hms = hov.timer_stop() / args.hifi_steps
        hc = hov.counters()
        hov.close()
        hk, rk, sr = (hc[x] / args.hifi_steps for x in ("hash_kmers", "ref_kmers", "seed_runs"))
        hpairs = hc["pairs"] / args.hifi_steps
        # Rooflines of the HBM-bound stages (DESIGN.md section 4 states the per-unit bytes).  Stage times are CUDA-event
        # brackets on the library's stream around each kernel (group) of the step; a stage of n identical launches
        # is divided by n: `achieved` is algorithmic bytes PER LAUNCH over time PER LAUNCH.
        stage_def = {   # stage: (kernel, launches, algorithmic bytes per launch)
            "index_tuples_ms": ("k_part1 (k-mers -> tuples scattered into coarse partitions)", 1, hk * (0.5 + 16)),
            "index_sort_ms": ("k_part2 (coarse partition -> buckets, 8-byte tuples)", 1, hk * (16 + 8)),
            "index_table_ms": ("k_bucket_group2 (bulk-copy fed) + first-position bitmap + k_path_slots2", 1, hk * (8 + 4)),
            "probe_ms": ("k_ref_probe<0> + k_ref_probe<1> (forward windows, then reverse windows)", 2, rk / 2 * (0.5 + 32)),
            "expand_ms": ("k_expand_coop", 1, sr * (4 + 16 + 16)),
            "sort_ms": ("radix sort of the seed runs", 8, sr * 16 * 2),
            "chain_ms": ("k_pair_scatter + k_chain_pairs", 1, sr * (16 + 12 + 12 + 16)),
        }
        stage_roof = {}
        for st_name, (kname, nl, ab) in stage_def.items():
            ms_l = hst.get(st_name, 0.0) / nl
            if ms_l <= 0:
                continue
            ach = ab / (ms_l * 1e-3) / 1e9
            stage_roof[st_name.replace("_ms", "")] = {"kernel": kname, "launches": nl, "ms_per_launch": round(ms_l, 4),
                                                      "achieved": round(ach, 1), "frac": round(ach / peaks["hbm_gbs"], 4)}
        dom = max(stage_roof, key=lambda k2: stage_roof[k2]["ms_per_launch"] * stage_roof[k2]["launches"])   # longest stage
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and args.hifi_genome == 5_000_000 and args.hifi_coverage == 50.0:
            traffic = json.load(open(tpath)).get(dom)
        index_ms = hst.get("index_tuples_ms", 0) + hst.get("index_sort_ms", 0) + hst.get("index_table_ms", 0)
        roofline = {"bound": "hbm", "kernel": stage_roof[dom]["kernel"], "stage": dom, "achieved": stage_roof[dom]["achieved"],
                    "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": stage_roof[dom]["frac"], "traffic": traffic, "peak_source": which,
                    "ms_per_launch": stage_roof[dom]["ms_per_launch"], "measured_on": "hifi_tile",
                    "launches": stage_roof[dom]["launches"],
                    "note": "longest HBM-bound stage of the C2 tile (bytes and time per launch); the job's dominant kernel (k_extend_pairs) is integer-ALU bound, see 'extension'"}
        line["hifi_tile"] = {
            "workload": "C2: %.1f Mbp x %gx HiFi-like reads (log-normal ~11 kb, %.1f%% read error), --maxerate %g, single hash x ref tile, inputs resident" % (
                args.hifi_genome / 1e6, args.hifi_coverage, HIFI_READ_ERR * 100, HIFI_ERATE),
            "bases": hbases, "ms_per_step": hms, "read_pairs_per_s": hpairs / (hms * 1e-3), "pairs_per_step": hpairs, "overlaps_per_step": int(hn),
            "stages_ms": {k2: round(v2, 3) for k2, v2 in hst.items()},
            "stage_rooflines": stage_roof,
            "index_build": {"ms": round(index_ms, 3), "survey_8d_bytes_per_kmer": 17.25,
                            "achieved_gbs": round(hk * 17.25 / (index_ms * 1e-3) / 1e9, 1) if index_ms > 0 else None,
                            "frac": round(hk * 17.25 / (index_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4) if index_ms > 0 else None},
            "extension_kernel_gcells_per_s": hc["dp_cells"] / args.hifi_steps / 1e9 / (hst["extend_ms"] * 1e-3) if hst.get("extend_ms") else None,
        }
    line["roofline"] = roofline
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
