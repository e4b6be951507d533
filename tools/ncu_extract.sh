#!/bin/bash
# On the GPU box: turn an .ncu-rep into small text artefacts (raw metrics csv, gzipped source-page csv, opcode histogram,
# per-line stall attribution) and delete the report (gpurun merges back at most 64 MiB).
# usage: tools/ncu_extract.sh gpurun_out/prof_x.ncu-rep <object file> <kernel substring>
set -u
REP=$1; OBJ=$2; KERN=$3
B=${REP%.ncu-rep}
python tools/ncu_summary.py $REP ${B}.csv > ${B}_summary.txt 2>&1
python tools/ncu_opcodes.py $REP > ${B}_opcodes.txt 2>&1
python tools/ncu_by_line.py $REP $OBJ $KERN 60 > ${B}_by_line.txt 2>&1
ncu -i $REP --page raw --csv 2>/dev/null | gzip -9 > ${B}_raw.csv.gz
ncu -i $REP --page source --csv 2>/dev/null | gzip -9 > ${B}_source.csv.gz
rm -f $REP
