"""Worker of tests/test_tiles.py::test_tiles_over_two_gloo_ranks_reproduce_the_single_tile (run under torchrun,
backend gloo).  Same sharding as bench.py / the host driver: plan tiles, LPT-assign them to ranks, every rank
computes only its own tiles, results are gathered on rank 0 -- no collective on the data path itself."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import numpy as np
import torch.distributed as dist

import golden_util as gu
from canu_b200 import api
from oracle import oracle_py as op


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    case = gu.get_case("A_default")
    reads = gu.load_dump_reads("A")
    kw, _ = gu.flags_to_kwargs(case["flags"])
    tiles = api.plan_tiles([len(r) for r in reads], 500, 120000, 150000)
    owner = api.assign_tiles(tiles, world)
    mine = [t for t, o in zip(tiles, owner) if o == rank]
    recs, stats = [], {}
    for t in mine:
        o = op.Oracle(kmer_len=22, max_erate=kw["max_erate"], min_olap_len=500, hash_bits=20, hash_load=0.8)
        o.set_reads(reads)
        recs.append(o.run(hb=t["hash_bgn"], he=t["hash_end"], rb=t["ref_bgn"], re=t["ref_end"], threads=3))
        for k, v in o.stats().items():
            stats[k] = stats.get(k, 0) + v
        o.close()
    mine_recs = np.concatenate(recs) if recs else np.zeros(0, dtype=op.REC_DTYPE)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine_recs.tobytes(), stats, len(mine)))
    if rank == 0:
        allrec = np.concatenate([np.frombuffer(g[0], dtype=op.REC_DTYPE) for g in gathered])
        tot = {}
        for g in gathered:
            for k, v in g[1].items():
                tot[k] = tot.get(k, 0) + v
        got = gu.format_records(allrec)
        want = gu.load_golden_lines("A_default")
        ok_stats, _ = gu.stats_match(gu.load_golden_stats("A_default"), tot)
        json.dump({"world": world, "tiles": len(tiles), "tiles_per_rank": [g[2] for g in gathered],
                   "records_match_golden": got == want, "n_got": len(got), "n_want": len(want),
                   "stats_match_golden": bool(ok_stats)}, open(sys.argv[1], "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
