#!/bin/bash
# usage (on the GPU box): tools/ncu_capture.sh <name> <kernel regex> <prof_op.py args...>
# one `ncu --set full` capture of the first matching launch after a warm-up launch, exported as
# gpurun_out/<name>_raw.csv (metrics), <name>_sass.csv (per-instruction) and <name>_src.csv (per source line)
NAME=$1; KRE=$2; shift 2
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k "regex:$KRE" -s 1 -c 1 -f -o gpurun_out/$NAME \
    python tools/prof_op.py "$@" > gpurun_out/${NAME}_ncu.log 2>&1
ncu -i gpurun_out/$NAME.ncu-rep --page raw --csv > gpurun_out/${NAME}_raw.csv 2>/dev/null
ncu -i gpurun_out/$NAME.ncu-rep --page source --csv --print-source sass > gpurun_out/${NAME}_sass.csv 2>/dev/null
ncu -i gpurun_out/$NAME.ncu-rep --page source --csv --print-source cuda > gpurun_out/${NAME}_src.csv 2>/dev/null
rm -f gpurun_out/$NAME.ncu-rep   # > 64 MiB of reports does not travel back; the CSV exports do
tail -2 gpurun_out/${NAME}_ncu.log
