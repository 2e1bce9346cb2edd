set -x
python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu_v11.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_v11.log; tail -3 gpurun_out/pytest_gpu_v11.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_v11.log 2>&1; tail -1 gpurun_out/smoke_v11.log
python bench.py > gpurun_out/bench_v11_default.json 2> gpurun_out/bench_v11.err; tail -c 600 gpurun_out/bench_v11_default.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_v11_reference.json 2>> gpurun_out/bench_v11.err; tail -c 300 gpurun_out/bench_v11_reference.json
python bench.py --workload degrade_fog_blur_unet --no-cpu-baseline > gpurun_out/bench_v11_cfg1.json 2>> gpurun_out/bench_v11.err
python bench.py --workload degrade_random_resunet --no-cpu-baseline > gpurun_out/bench_v11_cfg2.json 2>> gpurun_out/bench_v11.err
python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_v11.csv python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/ncu_bench_v11.log 2>&1
for hb in "224 256" "224 16384" "64 256" "64 16384"; do set -- $hb; python bench.py --hw $1 --batch $2 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/sweep_v11_extra.jsonl; done
wc -l gpurun_out/sweep_v11_extra.jsonl gpurun_out/launches_v11.csv
