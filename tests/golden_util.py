"""Helpers shared by the parity tests: load golden cases, format records like
the reference's `overlapConvert -unaligned` (stores/ovOverlap.C:54-66)."""
import gzip
import json
import os
import struct

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_cases():
    with open(os.path.join(GOLDEN, "cases.json")) as f:
        return json.load(f)


def case_names():
    return [c["name"] for c in load_cases()["cases"]]


def get_case(name):
    for c in load_cases()["cases"]:
        if c["name"] == name:
            return c
    raise KeyError(name)


def load_dump_reads(store):
    """Reads as the reference overlapper sees them (sqStoreDumpFASTQ of the default version)."""
    reads, cur = [], None
    with gzip.open(os.path.join(GOLDEN, store + ".dump.fasta.gz"), "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if cur is not None:
                    reads.append(np.frombuffer(b"".join(cur), dtype=np.uint8).copy())
                cur = []
            else:
                cur.append(line.strip())
    if cur is not None:
        reads.append(np.frombuffer(b"".join(cur), dtype=np.uint8).copy())
    return reads


def load_golden_lines(case):
    with gzip.open(os.path.join(GOLDEN, case + ".ovl.txt.gz"), "rb") as f:
        return [ln for ln in f.read().decode().split("\n") if ln]


def load_golden_stats(case):
    out = {}
    with open(os.path.join(GOLDEN, case + ".stats")) as f:
        for line in f:
            k, v = line.split("=")
            out[k.strip()] = int(v)
    return out


def load_golden_oc(case):
    return read_oc(os.path.join(GOLDEN, case + ".oc"))


def read_oc(path):
    """.oc = uint64 nOlaps, uint32 oprMax, uint32 opr[oprMax] (stores/ovStoreFile.H:73-91)."""
    with open(path, "rb") as f:
        d = f.read()
    n_olaps, opr_max = struct.unpack_from("<QI", d, 0)
    opr = np.frombuffer(d, dtype="<u4", count=opr_max, offset=12).copy()
    return n_olaps, opr


def skip_kmers(name):
    ks = []
    with open(os.path.join(GOLDEN, name)) as f:
        for line in f:
            if line.startswith(">"):
                continue
            ks.append(line.split()[0])
    return ks


def flags_to_kwargs(flags):
    """Map reference command-line flags of a golden case to parameter keywords."""
    kw = dict(kmer_len=22, min_olap_len=500, hash_bits=22, hash_load=0.8)
    skip = None
    i = 0
    while i < len(flags):
        f = flags[i]
        if f == "--maxerate":
            kw["max_erate"] = float(flags[i + 1]); i += 1
        elif f == "-partial":
            kw["partial"] = True
        elif f == "-m":
            kw["unique"] = False
        elif f == "--minkmers":
            kw["min_kmers"] = True
        elif f == "-z":
            kw["no_hopeless"] = True
        elif f == "--hashdatalen":
            kw["hash_data_len"] = int(flags[i + 1]); i += 1
        elif f == "-k":
            skip = flags[i + 1]; i += 1
        else:
            raise ValueError(f)
        i += 1
    return kw, skip


def format_records(recs):
    """Text lines identical to `overlapConvert -unaligned` for a structured record array."""
    M = (1 << 21) - 1
    out = []
    for a, b, w0, w1 in zip(recs["a_iid"].tolist(), recs["b_iid"].tolist(), recs["w0"].tolist(), recs["w1"].tolist()):
        out.append("%10d %10d  %c  %6d  %6d %6d %6d %6d  %7.6f %s %s %s" % (
            a, b, "I" if (w0 >> 58) & 1 else "N", (w1 >> 42) & M,
            w0 & M, (w0 >> 21) & M, w1 & M, (w1 >> 21) & M, ((w0 >> 42) & 0xFFFF) / 100000.0,
            "OBT" if (w0 >> 59) & 1 else "   ", "DUP" if (w0 >> 60) & 1 else "   ", "UTG" if (w0 >> 61) & 1 else "   "))
    out.sort()
    return out


def oc_from_records(recs, n_reads):
    opr = np.zeros(n_reads + 1, dtype=np.uint32)
    np.add.at(opr, recs["a_iid"], 1)
    np.add.at(opr, recs["b_iid"], 1)
    return len(recs), opr


def stats_match(golden_stats, st, partial=False):
    """Compare the 6 live counters of the -s file (the two window counters are dead code)."""
    exp = {
        "Kmer hits without olaps": st["kmer_hits_without_olap"],
        "Kmer hits with olaps": st["kmer_hits_with_olap"],
        "Multiple overlaps/pair": st["multi_overlap"],
        "Total overlaps produced": st["total_overlaps"],
        "Contained overlaps": st["contained"],
        "Dovetail overlaps": st["dovetail"],
        "Rejected by short window": 0,
        "Rejected by long window": 0,
    }
    return exp == golden_stats, exp
