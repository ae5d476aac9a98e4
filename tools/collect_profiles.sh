#!/bin/bash
# Run ON the GPU box (gpurun): launch list + ncu --set full of the walk and the direct sum, round tag in $1.
# Every ncu run follows a plain run of the same command (B200_PROFILING.md).
tag=${1:-r02}
out=gpurun_out
python tools/build_prof.py > $out/${tag}_build_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_1m.csv python tools/build_prof.py > /dev/null 2>&1
python tools/walk_prof.py 2,1 1,0 > $out/${tag}_walk_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_walk -c 4 -o $out/${tag}_walk python tools/walk_prof.py 2,1 1,0 > $out/${tag}_walk_ncu.log 2>&1
python tools/direct_prof.py > $out/${tag}_direct_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_direct -c 2 -o $out/${tag}_direct python tools/direct_prof.py > $out/${tag}_direct_ncu.log 2>&1
cat $out/${tag}_build_plain.log $out/${tag}_walk_plain.log $out/${tag}_direct_plain.log
