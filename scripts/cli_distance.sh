# CLI with --distance: reference vs the three drop-ins on a synthetic pair (2 x $1 bp)
L=${1:-1000000}
mkdir -p /tmp/vd; python - <<P
import sys; sys.path.insert(0,'.')
from workloads import vcfgen
print(vcfgen.generate('/tmp/vd/in', seed=3, contig_len=$L, n_contigs=2))
P
for b in vcfdist_ref vcfdist_b200cli; do
  mkdir -p /tmp/vd/$b; cd /tmp/vd/$b
  echo "== $b"; VD_DROPIN_TIMES=1 $GRAFT_REPO_ROOT/oracle/_ref/$b /tmp/vd/in/query.vcf /tmp/vd/in/truth.vcf /tmp/vd/in/ref.fa -p /tmp/vd/$b/ -v 1 -t 16 --distance 2>&1 | grep -E "\[[0-9]\] |Total edit"
  cd $GRAFT_REPO_ROOT
done
for f in distance.tsv distance-summary.tsv edits.tsv superclusters.tsv precision-recall-summary.tsv; do cmp /tmp/vd/vcfdist_ref/$f /tmp/vd/vcfdist_b200cli/$f && echo "$f identical"; done
