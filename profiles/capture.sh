#!/bin/bash
# One ncu --set full capture of one kernel launch, digested ON THE GPU BOX (a .ncu-rep is ~23 MB and gpurun
# brings back at most 64 MiB): key metrics + instruction mix, executed instructions per source line, stall
# samples per source line.  The report itself is deleted unless KEEP_REP=1.
# usage: profiles/capture.sh <out-prefix> <kernel-regex> <mangled-kernel-substring> <rays-per-launch> <command ...>
set -u
out=$1; kre=$2; ksub=$3; rays=$4; shift 4
ncu --set full --import-source on --clock-control none -k "regex:$kre" --launch-skip 1 -c 1 -f -o "$out" "$@" > "$out.ncu.log" 2>&1
rep="$out.ncu-rep"
[ -f "$rep" ] || { echo "no report for $out"; tail -5 "$out.ncu.log"; exit 1; }
python profiles/ncu_digest.py "$rep" "$rays" > "${out}_digest.csv" 2>&1
python profiles/ncu_hot.py "$rep" xicsrt_b200/libxrt.so "$ksub" "$rays" 80 > "${out}_hot_lines.csv" 2>&1
python profiles/ncu_stalls.py "$rep" xicsrt_b200/libxrt.so "$ksub" 50 > "${out}_stalls.csv" 2>&1
ncu -i "$rep" --page details --csv 2>/dev/null | grep -E "Stall|stall|Warp Cycles|No Eligible|Eligible|Issued Warp|Registers|Shared Memory|Achieved Occupancy|Theoretical Occupancy|L1/TEX Hit|L2 Hit|Branch" > "${out}_details.csv"
[ "${KEEP_REP:-0}" = 1 ] || rm -f "$rep"
