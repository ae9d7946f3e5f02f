set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_plain2.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:ilqr_wpp -s 3 -c 1 -f -o gpurun_out/wpp_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_wpp.log 2>&1
tail -3 gpurun_out/ncu_wpp.log; ls -la gpurun_out/*.ncu-rep
