set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
for w in ballbot quadrotor_slq manipulator; do python bench.py --workload $w --no-cpu-baseline 2>/dev/null | grep '^{' >> gpurun_out/bench_small.jsonl; done
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench.json; cut -c1-300 gpurun_out/bench_reference.json
