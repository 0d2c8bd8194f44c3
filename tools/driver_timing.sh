
cd $GRAFT_REPO_ROOT
python - <<'PY'
import sys, os, subprocess
sys.path.insert(0, '.')
from canu_b200 import synth
g = synth.make_genome(2_000_000, seed=11)
reads = synth.simulate_reads(g, 50, 3000, 30000, 0.001, seed=12, lognormal=(9.25, 0.3))
synth.write_fasta('/tmp/r.fasta', reads)
subprocess.check_call(['oracle/_ref/bin/sqStoreCreate', '-o', '/tmp/r.seqStore', '-minlength', '1000', '-pacbio-hifi', 'lib', '/tmp/r.fasta'], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
print(len(reads))
PY
for i in 1 2; do date +%s.%N; canu_b200/bin/overlapInCore -t 4 -k 22 --maxerate 0.01 --minlength 500 -o /tmp/o.ovb -s /tmp/o.stats /tmp/r.seqStore 2>&1; done
date +%s.%N
