"""Tile planning: the planner behind the multi-GPU runs must cut the job grid exactly like
the reference's overlapInCorePartition, and the tiles, spread over ranks, must reproduce the single-tile result.
CPU only: the planner is host code of the C ABI; the per-tile overlaps here come from the ORACLE (the checker),
which lets the N>1 sharding logic be tested with world_size 2 on `gloo` without a GPU."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lens(store):
    return [len(r) for r in gu.load_dump_reads(store)]


def _opt_line(t):
    s = "-h %d-%d -r %d-%d" % (t["hash_bgn"], t["hash_end"], t["ref_bgn"], t["ref_end"])
    if t["has_hash_reads"]:
        s += " --hashdatalen %d" % t["hash_bases"]
    return s


@pytest.mark.parametrize("case", json.load(open(os.path.join(gu.GOLDEN, "partition.json"))),
                         ids=lambda c: "%s-hl%d-rl%d-ol%d" % (c["store"], c["hl"], c["rl"], c["ol"]))
def test_plan_matches_reference_partition(case):
    """Bit-exact against the .ovlopt lines the reference overlapInCorePartition wrote for the same store."""
    from canu_b200 import api
    tiles = api.plan_tiles(_lens(case["store"]), case["ol"], case["hl"], case["rl"], strict_reference=True)
    assert [_opt_line(t) for t in tiles] == case["ovlopt"]


@pytest.mark.parametrize("hl,rl", [(200000, 400000), (60000, 90000), (10 ** 9, 10 ** 9), (1, 1)])
def test_full_cover_plan_covers_every_pair_once(hl, rl):
    """strict_reference=False: every (ref < hash) read pair lies in exactly one tile."""
    from canu_b200 import api
    lens = _lens("A")
    n = len(lens)
    tiles = api.plan_tiles(lens, 500, hl, rl, strict_reference=False)
    cover = np.zeros((n + 1, n + 1), dtype=np.int32)
    for t in tiles:
        cover[t["ref_bgn"]:t["ref_end"] + 1, t["hash_bgn"]:t["hash_end"] + 1] += 1
    r, h = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    need = (r >= 1) & (h >= 1) & (r < h)
    assert (cover[need] == 1).all()


def test_assignment_is_balanced_and_deterministic():
    from canu_b200 import api
    tiles = api.plan_tiles(_lens("A"), 500, 40000, 60000)
    assert len(tiles) > 20
    for w in (1, 2, 4, 8):
        own = api.assign_tiles(tiles, w)
        assert own == api.assign_tiles(tiles, w)
        load = np.zeros(w)
        for t, o in zip(tiles, own):
            load[o] += t["cost"]
        assert set(own) == set(range(w))
        assert load.max() <= load.mean() * 1.25 + max(t["cost"] for t in tiles)


def test_tiles_over_two_gloo_ranks_reproduce_the_single_tile(tmp_path):
    """world_size 2 on gloo: each rank runs (with the oracle) only the tiles it owns; the gathered union must be
    the reference's golden output for the whole range, and the counters must add up."""
    script = os.path.join(ROOT, "tests", "gloo_tiles_worker.py")
    out = str(tmp_path / "result.json")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29571")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29571", script, out]
    r = subprocess.run(cmd, env=env, capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-3000:]
    res = json.load(open(out))
    assert res["world"] == 2 and res["tiles"] > 4
    assert min(res["tiles_per_rank"]) > 0
    assert res["records_match_golden"], res
    assert res["stats_match_golden"], res


@pytest.mark.parametrize("parts", [1, 2, 3, 4, 8])
def test_balanced_ref_cut_covers_once_and_is_balanced(parts):
    """ovlb_plan_balanced: the ref range of one hash block is cut into contiguous tiles that cover every (ref < hash)
    pair exactly once and whose TRUE triangular work (ref bases x hash bases behind the ref read) is even."""
    from canu_b200 import api
    lens = _lens("A")
    n = len(lens)
    tiles = api.plan_balanced(lens, 500, parts)
    assert len(tiles) == parts
    assert tiles[0]["ref_bgn"] == 1 and tiles[-1]["ref_end"] == n - 1
    for a, b in zip(tiles, tiles[1:]):
        assert b["ref_bgn"] == a["ref_end"] + 1
    L = np.array([0] + [x if x >= 500 else 0 for x in lens], dtype=np.float64)
    behind = np.concatenate([np.cumsum(L[::-1])[::-1][1:], [0.0]])      # hash bases with ID > r
    work = L * behind
    per = np.array([work[t["ref_bgn"]:t["ref_end"] + 1].sum() for t in tiles])
    assert per.max() <= per.mean() * 1.15 + work.max()
    own = api.assign_tiles(tiles, parts)
    assert sorted(own) == list(range(parts))                # one tile per worker


def test_balanced_ref_cut_subranges_and_degenerate():
    from canu_b200 import api
    lens = _lens("A")
    t = api.plan_balanced(lens, 500, 4, hash_range=(50, 120), ref_range=(10, 200))
    assert t[0]["ref_bgn"] == 10 and t[-1]["ref_end"] == 119 and all(x["hash_bgn"] == 50 and x["hash_end"] == 120 for x in t)
    assert api.plan_balanced(lens, 500, 4, hash_range=(5, 5), ref_range=(5, 9)) == []
    few = api.plan_balanced(lens, 500, 8, hash_range=(1, 4), ref_range=(1, 4))
    assert 1 <= len(few) <= 3 and few[-1]["ref_end"] == 3


def test_hash_block_model_follows_the_memory_budget():
    """ovlb_hash_block_bases: the re-blocking counterpart of Configure.pm's ovlHashBlockLength.  What the run side needs
    comes off the budget first, the rest is divided by 150 B per hash base: 734 Mbases for the 0.76 x 183 GB a context
    gets on a B200 (HiFi-like job), less when long noisy reads need a large extension scratch or the ref batches are
    larger, half of it (and a bit) for two contexts on one device; clamped to [1 Mbase, 1.5 Gbases]."""
    from canu_b200 import api
    b200 = int(183e9 * 0.95 * 0.8)
    hifi = api.hash_block_bases(b200, 30000, 0.01)
    assert 720e6 < hifi < 750e6
    noisy = api.hash_block_bases(b200, 60000, 0.12)
    assert noisy < api.hash_block_bases(b200, 20000, 0.06) < hifi
    assert api.hash_block_bases(b200, 30000, 0.01, ref_batch_bases=1_000_000_000) < hifi
    half = api.hash_block_bases(b200 // 2, 30000, 0.01)
    assert 0.4 * hifi < half < 0.5 * hifi
    assert api.hash_block_bases(10 ** 9, 30000, 0.01) == 1_000_000                # too little memory: the floor
    assert api.hash_block_bases(10 ** 13, 30000, 0.01) == 1_500_000_000           # the 2^32-position cap of one block
    grown = [api.hash_block_bases(g << 30, 30000, 0.01) for g in range(8, 200, 8)]
    assert grown == sorted(grown)

