set -x
mkdir -p gpurun_out
for w in "ballbot 65536" "quadrotor_slq 32768" "manipulator 16384"; do set -- $w; python tools/prof_shape.py $1 $2 3 > gpurun_out/prof_$1.log 2>&1 || exit 1; done
for w in "ballbot 65536" "quadrotor_slq 32768" "manipulator 16384"; do set -- $w; ncu --set full --clock-control none --import-source on -k regex:rpl -s 2 -c 2 -f -o gpurun_out/rpl_$1 python tools/prof_shape.py $1 $2 3 > gpurun_out/ncu_$1.log 2>&1; done
ls -la gpurun_out/*.ncu-rep; cat gpurun_out/prof_*.log
