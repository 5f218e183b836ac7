#!/bin/bash
# usage: tools/ncu_export.sh <report.ncu-rep>   (run on the GPU box, inside gpurun)
# Writes <report>.details.csv / .raw.csv / .source.csv (SASS view) / .cuda.csv (CUDA-C view, needs
# -lineinfo and --import-source on) next to it and drops the report when it is too large to travel
# back (gpurun_out/ is capped at 64 MiB).
rep="$1"; base="${rep%.ncu-rep}"
ncu -i "$rep" --page details --csv > "$base.details.csv" 2>/dev/null
ncu -i "$rep" --page raw --csv > "$base.raw.csv" 2>/dev/null
ncu -i "$rep" --page source --csv > "$base.source.csv" 2>/dev/null
ncu -i "$rep" --page source --print-source cuda --csv > "$base.cuda.csv" 2>/dev/null
sz=$(stat -c %s "$rep")
if [ "$sz" -gt 25000000 ]; then rm -f "$rep"; fi
