mkdir -p gpurun_out/r2_e3 /tmp/vg; python - <<'P'
import sys; sys.path.insert(0,'.')
from workloads import vcfgen
print(vcfgen.generate('/tmp/vg/in', seed=2, contig_len=2_000_000, n_contigs=2))
P
mkdir -p /tmp/vg/out; cd /tmp/vg/out; VD_DROPIN_TIMES=1 $GRAFT_REPO_ROOT/oracle/_ref/vcfdist_b200cli /tmp/vg/in/query.vcf /tmp/vg/in/truth.vcf /tmp/vg/in/ref.fa -p /tmp/vg/out/ -v 1 -t 16 2>&1 | grep -E "GPU clustering|GPU prec|\[[0-9]\] " 
