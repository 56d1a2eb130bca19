cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 2> gpurun_out/r02_bench_n1.err | grep -a "^{" > gpurun_out/r02_bench_n1.json; python tools/show_bench.py gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/r02_bench_ref.err | grep -a "^{" > gpurun_out/r02_bench_reference_arm.json; head -c 600 gpurun_out/r02_bench_reference_arm.json; echo
python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-verify > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_c4.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-verify > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/r02_launches_bench_c4.csv
bash tools/_audit.sh 2>&1 | tail -25
