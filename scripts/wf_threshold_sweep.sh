for cm in 3072 6144 12288; do
  for bm in 384; do
    echo "== cluster_min $cm block_min $bm"
    VD_WF_CLUSTER_MIN=$cm VD_WF_BLOCK_MIN=$bm SKIP_REF=1 VD_WF_STATS=1 bash scripts/cli_sv_cluster.sh 200000 4000 2>&1 | grep "vd_wf_batch mode\|\[3\]" | awk '/vd_wf_batch/ {ms=$(NF-1); tot[$3]+=ms} /\[3\]/ {print} END {for (k in tot) print "mode",k,tot[k],"ms"}'
  done
done
echo "== block_min 192"; VD_WF_BLOCK_MIN=192 SKIP_REF=1 VD_WF_STATS=1 bash scripts/cli_sv_cluster.sh 200000 4000 2>&1 | grep "vd_wf_batch mode\|\[3\]" | awk '/vd_wf_batch/ {ms=$(NF-1); tot[$3]+=ms} /\[3\]/ {print} END {for (k in tot) print "mode",k,tot[k],"ms"}'
echo "== block_min 768"; VD_WF_BLOCK_MIN=768 SKIP_REF=1 VD_WF_STATS=1 bash scripts/cli_sv_cluster.sh 200000 4000 2>&1 | grep "vd_wf_batch mode\|\[3\]" | awk '/vd_wf_batch/ {ms=$(NF-1); tot[$3]+=ms} /\[3\]/ {print} END {for (k in tot) print "mode",k,tot[k],"ms"}'
