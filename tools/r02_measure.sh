#!/bin/bash
# Round-2 evidence run (one gpurun call): GPU tests, smoke, bench lines of every config, ncu launch list, ncu --set full of K1/K2/K4.
# Usage on the box: bash tools/r02_measure.sh [tag]    -> everything lands in gpurun_out/<tag>_*
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $out/${tag}_pytest.log 2>&1
echo "pytest exit: $?" >> $out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1
timeout 600 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
timeout 300 python bench.py --config 1 > $out/${tag}_bench_c1.json 2> $out/${tag}_bench_c1.err
timeout 300 python bench.py --config 3 --skip-cpu > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
timeout 300 python bench.py --config 3 --hop --skip-cpu > $out/${tag}_bench_c3hop.json 2> $out/${tag}_bench_c3hop.err
timeout 300 python bench.py --config 4 --skip-cpu > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err
timeout 400 python bench.py --config 5 --sweep --skip-cpu > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err
# launch list of the default command (cold-cache, serialised times: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e > $out/${tag}_launches.log 2>&1
for k in k_spectrum3 k_detect k_track; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $out/${tag}_$k \
      python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e > $out/${tag}_ncu_$k.log 2>&1
done
ls -la $out
