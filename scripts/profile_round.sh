#!/bin/bash
# ncu evidence for profiles/: run on the GPU box via
#   gpurun --timeout 1500 -- 'bash scripts/profile_round.sh r2 [train|render|all]'
# Every ncu pass follows a plain run of the same command that exited 0.  Outputs land in gpurun_out/.
tag=${1:-r2}
what=${2:-all}
out=gpurun_out
mkdir -p $out
set -x
R="python bench.py --workload render --steps 3 --warmup 3 --no-cpu-baseline"
T="python bench.py --workload train --steps 3 --warmup 3"
C="python scripts/prof_composite.py"
if [ "$what" = "render" ] || [ "$what" = "all" ]; then
$R > $out/${tag}_plain_render.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_render.csv $R > $out/${tag}_ncu_render.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:FwdEpi -s 3 -c 1 -o $out/${tag}_prof_fwd $R > $out/${tag}_ncu_fwd.log 2>&1
fi
if [ "$what" = "train" ] || [ "$what" = "all" ]; then
$T > $out/${tag}_plain_train.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_train.csv $T > $out/${tag}_ncu_train.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_wgrad -s 3 -c 1 -o $out/${tag}_prof_wgrad $T > $out/${tag}_ncu_wgrad.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:DgradEpi -s 3 -c 1 -o $out/${tag}_prof_dgrad $T > $out/${tag}_ncu_dgrad.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:FwdEpi -s 3 -c 1 -o $out/${tag}_prof_fwdsave $T > $out/${tag}_ncu_fwdsave.log 2>&1
fi
if [ "$what" = "all" ]; then
$C > $out/${tag}_plain_composite.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:composite_ -s 4 -c 2 -o $out/${tag}_prof_composite $C > $out/${tag}_ncu_composite.log 2>&1
fi
ls -la $out | tail -20
