out=gpurun_out/$1; mkdir -p $out
VD_SERIAL=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_split.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-seam > $out/launches_split.log 2>&1
python scripts/launch_summary.py $out/launches_split.csv > $out/launch_summary_split.txt 2>&1; cat $out/launch_summary_split.txt
