# One gpurun call that re-validates the tree the way the driver does at round end: GPU tests, smoke(), both bench arms.
set -x
python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu_final.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_final.log; tail -3 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final_reference.json
python bench.py > gpurun_out/bench_final_default.json 2>> gpurun_out/bench_final.err; tail -c 700 gpurun_out/bench_final_default.json
