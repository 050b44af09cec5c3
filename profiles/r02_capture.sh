#!/bin/bash
# Round-2 ncu captures (one B200).  Run as:  gpurun --timeout 1700 -- 'bash profiles/r02_capture.sh'
# Every profiled command first runs WITHOUT ncu and must exit 0 (B200_PROFILING.md); numbers printed under ncu are never bench values.
set -u
O=gpurun_out; mkdir -p $O
P8M="python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline"
ML8="python bench.py --workload p8m_ml --steps 1 --warmup 3 --no-cpu-baseline"
MLC="python bench.py --workload c2_ml --steps 1 --warmup 3 --no-cpu-baseline"
C4="python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline"
$P8M > $O/r02_plain_p8m.json 2> $O/r02_plain_p8m.err || { echo "plain p8m failed"; exit 1; }
$ML8 > $O/r02_plain_p8m_ml.json 2> $O/r02_plain_p8m_ml.err || { echo "plain p8m_ml failed"; exit 1; }
$C4 > $O/r02_plain_c4.json 2> $O/r02_plain_c4.err || { echo "plain c4 failed"; exit 1; }
# launch lists: one step's worth of launches after the warm-up steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 5950 -c 1980 --csv --log-file $O/r02_launches_p8m.csv $P8M > $O/r02_ncu_p8m.log 2>&1; echo "launch list p8m rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 4100 -c 1370 --csv --log-file $O/r02_launches_p8m_ml.csv $ML8 > $O/r02_ncu_p8m_ml.log 2>&1; echo "launch list p8m_ml rc=$?"
# full captures of the kernels VERDICT r1 names + the new ones
ncu --set full --clock-control none --import-source on -k regex:k_update_dot_tma -s 300 -c 3 -o $O/r02_prof_update_dot_tma $P8M > $O/r02_ncu_ud.log 2>&1; echo "update_dot_tma rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell -s 400 -c 2 -o $O/r02_prof_spmv_p8m $P8M > $O/r02_ncu_spmv.log 2>&1; echo "spmv rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_ilu_solve -s 20 -c 2 -o $O/r02_prof_ilu_solve $C4 > $O/r02_ncu_ilus.log 2>&1; echo "ilu solve rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_ilu_factor -s 1 -c 1 -o $O/r02_prof_ilu_factor $C4 > $O/r02_ncu_iluf.log 2>&1; echo "ilu factor rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:k_amg_(compress|merge|strength|mis_m1|restrict|csr)" -s 30 -c 12 -o $O/r02_prof_amg $ML8 > $O/r02_ncu_amg.log 2>&1; echo "amg rc=$?"
ls -la $O/*.ncu-rep
