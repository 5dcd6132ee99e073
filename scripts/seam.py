"""Timing at the reference's own seam (bench.run_seam) on a chosen workload size.  usage: seam.py [n_sc] [cli_contig_len]"""
import json, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_600_000
args = types.SimpleNamespace(seed=1, cli_contig_len=int(sys.argv[2]) if len(sys.argv) > 2 else 6_000_000,
                             cli_cluster_contig_len=int(sys.argv[3]) if len(sys.argv) > 3 else 2_000_000)
b, cells, total = bench.make_workload("wgs", n, 1, 0, 1, 10000)
print(json.dumps(bench.run_seam(args, b), indent=1))
