# CLI with the default (biwfa) clustering on an SV-bearing synthetic pair: reference vs drop-ins.  usage: cli_sv_cluster.sh [contig_len] [sv_max] [extra CLI flags, e.g. --distance]
L=${1:-500000}; SV=${2:-1500}; shift; shift; EXTRA="$@"
mkdir -p /tmp/vs; python - <<P
import sys; sys.path.insert(0,'.')
from workloads import vcfgen
print(vcfgen.generate('/tmp/vs/in', seed=5, contig_len=$L, n_contigs=2, sv_rate=0.03, sv_max=$SV))
P
for b in vcfdist_b200cli ${SKIP_REF:+-} vcfdist_ref; do
  [ "$b" = "-" ] && break
  mkdir -p /tmp/vs/$b; cd /tmp/vs/$b
  echo "== $b"; SECONDS=0
  VD_DROPIN_TIMES=1 timeout 900 $GRAFT_REPO_ROOT/oracle/_ref/$b /tmp/vs/in/query.vcf /tmp/vs/in/truth.vcf /tmp/vs/in/ref.fa -p /tmp/vs/$b/ -v 1 -t 16 $EXTRA 2>&1 | grep -E "\[[0-9]\] |ERROR|GPU clustering|GPU precision|vd_wf_batch"
  echo "wall ${SECONDS} s"
  cd $GRAFT_REPO_ROOT
done
for f in superclusters.tsv precision-recall-summary.tsv query.tsv $( [ -f /tmp/vs/vcfdist_ref/distance.tsv ] && echo distance.tsv edits.tsv ); do cmp /tmp/vs/vcfdist_ref/$f /tmp/vs/vcfdist_b200cli/$f && echo "$f identical"; done
