#!/bin/bash
# one GPU session: tests, headline bench, the other BASELINE configs, launch list; outputs under gpurun_out/
tag=${1:-r2}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gpu_tests.log 2>&1; tail -3 gpurun_out/${tag}_gpu_tests.log
python bench.py --steps 8 --warmup 3 > gpurun_out/${tag}_bench_config3.json 2> gpurun_out/${tag}_bench_config3.err; echo "config3 rc=$?"
python bench.py --config 2 --steps 8 --warmup 3 --no-side > gpurun_out/${tag}_bench_config2.json 2> gpurun_out/${tag}_bench_config2.err; echo "config2 rc=$?"
python bench.py --config 5 --steps 4 --warmup 3 --no-side --cpu-sample 1024 > gpurun_out/${tag}_bench_config5.json 2> gpurun_out/${tag}_bench_config5.err; echo "config5 rc=$?"
python bench.py --config 1 --steps 20 --warmup 3 --no-side > gpurun_out/${tag}_bench_config1.json 2> gpurun_out/${tag}_bench_config1.err; echo "config1 rc=$?"
python bench.py --config 4 --cl-steps 500 --warmup 3 > gpurun_out/${tag}_bench_config4.json 2> gpurun_out/${tag}_bench_config4.err; echo "config4 rc=$?"
