#!/bin/bash
# The gpurun command behind profiles/kernel_constants.json (one GPU):
#   gpurun --timeout 1500 -- 'bash profiles/ncu_capture.sh r2'
# Each capture runs the same command plainly first (it must exit 0), then under
#   ncu --set full --metrics <pipe-cycle counters> --clock-control none --import-source on
# on the last launches of the driver script.  Reports land in gpurun_out/<tag>_<stem>.ncu-rep;
# `python profiles/make_kernel_constants.py --tag <tag>` (run in the authoring container) turns them into
# profiles/kernel_constants.json and profiles/summarise.py into the readable summaries.
set -u
TAG=${1:-r2}
KEEP=${2:-"gibbs_f32 predict_f32"}
ONLY=${3:-""}          # optional: capture only these stems
OUT=gpurun_out
mkdir -p $OUT
METRICS=$(python profiles/make_kernel_constants.py --print-metrics)
cap() {   # stem, kernel regex, launches to skip, launches to take, command...
    local stem=$1 kre=$2 skip=$3 take=$4; shift 4
    if [ -n "$ONLY" ] && ! echo " $ONLY " | grep -q " $stem "; then return 0; fi
    "$@" > $OUT/${TAG}_${stem}_plain.log 2>&1 || { echo "plain run of $stem failed"; tail -5 $OUT/${TAG}_${stem}_plain.log; return 1; }
    ncu --set full --metrics "$METRICS" --clock-control none --import-source on -k "regex:$kre" -s $skip -c $take \
        -f -o $OUT/${TAG}_${stem} "$@" > $OUT/${TAG}_${stem}_ncu.log 2>&1 || { echo "ncu of $stem failed"; tail -5 $OUT/${TAG}_${stem}_ncu.log; }
    tail -2 $OUT/${TAG}_${stem}_plain.log
    if [ -f $OUT/${TAG}_${stem}.ncu-rep ]; then
        ncu -i $OUT/${TAG}_${stem}.ncu-rep --page raw --csv > $OUT/${TAG}_${stem}_raw.csv 2>/dev/null
        ncu -i $OUT/${TAG}_${stem}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $OUT/${TAG}_${stem}_source.csv.gz
        if ! echo " $KEEP " | grep -q " $stem "; then rm -f $OUT/${TAG}_${stem}.ncu-rep; fi
    fi
}
cap gibbs_f32 'gibbs_conjugate_kernel' 2 1 python profiles/time_gibbs.py float32 10000 65536
cap gibbs_f64 'gibbs_conjugate_kernel' 2 1 python profiles/time_gibbs.py float64 10000 65536
cap gibbs_k64 'gibbs_conjugate_kernel' 2 1 python profiles/time_gibbs_k64.py float32 1000 16384
cap simplex_f32 'gibbs_simplex' 1 1 python profiles/prof_simplex.py
cap predict_f32 'predict_(pass_tc|select)_kernel' 8 8 python profiles/time_predict.py 100000 100000 16 3
ls -la $OUT/${TAG}_*
