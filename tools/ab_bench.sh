#!/bin/bash
# A/B of alternative builds of libb2s.so through bench.py (kernel times, CTA balance, value). Usage: bash tools/ab_bench.sh <tag> <config args> -- <lib>...
tag=$1; shift
args=()
while [ "$1" != "--" ]; do args+=("$1"); shift; done
shift
mkdir -p gpurun_out
out=gpurun_out/${tag}_abbench.txt
: > $out
for lib in "$@"; do
  if [ "$lib" = main ]; then path=rtl-sdr-scanner-cpp_b200/lib/libb2s.so; else path=$lib; fi
  for rep in 1 2; do
    echo "== $lib (run $rep)" >> $out
    B2S_LIB=$PWD/$path timeout 300 python bench.py --skip-cpu --skip-e2e "${args[@]}" 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception:
        print(line.rstrip()); continue
    r = d['roofline']
    print(json.dumps({'value': round(d['value']), 'ms_per_step': round(d['ms_per_step'], 4), 'k1_ms': round(r['kernel_ms'], 4), 'k2_ms': round(r['other_kernels_ms']['k_detect+list_ordering'], 4),
                      'k4_ms': round(r['other_kernels_ms']['k_track (beside the next step\'s K1)'], 4), 'cta_median': r['k_detect']['cta_median_ms'], 'cta_max': r['k_detect']['cta_max_ms'], 'tx': d['detections']['transmissions_after_last_step'], 'entries': d['detections']['detect_entries_last_step']}))
" >> $out
  done
done
cat $out
