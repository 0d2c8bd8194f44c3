"""The C restatement (oracle/) against every golden fixture minted from the
unmodified reference binary (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import oracle_py as op
import golden_util as gu


@pytest.mark.parametrize("case", gu.case_names())
def test_oracle_matches_reference(case):
    c = gu.get_case(case)
    reads = gu.load_dump_reads(c["store"])
    kw, skip = gu.flags_to_kwargs(c["flags"])
    o = op.Oracle(**kw)
    o.set_reads(reads)
    if skip:
        o.set_skip_kmers(gu.skip_kmers(skip))
    recs = o.run(hb=c["h"][0], he=c["h"][1], rb=c["r"][0], re=c["r"][1], threads=4)
    assert gu.format_records(recs) == gu.load_golden_lines(case)
    ok, exp = gu.stats_match(gu.load_golden_stats(case), o.stats())
    assert ok, (exp, gu.load_golden_stats(case))
    n, opr = gu.oc_from_records(recs, len(reads))
    gn, gopr = gu.load_golden_oc(case)
    assert n == gn and np.array_equal(opr, gopr)
    o.close()


def test_match_limit_tables_are_monotone_and_float_rounded():
    # --maxerate is parsed with strtof (overlapInCore.C:380): 0.06 -> 0.0599999986588954...
    o = op.Oracle(max_erate=0.06)
    assert o.p.max_erate == float(np.float32(0.06)) and o.p.max_erate < 0.06
    eml = o.edit_match_limit()
    assert eml[0] == 0 and eml[1] == 0 and np.all(np.diff(eml) >= 0)
    assert o.max_errors() == 1 + int(np.ceil(o.p.max_erate * ((1 << 21) - 1)))
    assert o.error_bound(1000) == int(np.ceil(1000 * o.p.max_erate))
    o.close()
