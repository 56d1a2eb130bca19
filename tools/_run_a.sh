cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 20 --warmup 3 2> gpurun_out/r02_bench_n1.err | grep -a "^{" > gpurun_out/r02_bench_n1.json; python tools/show_bench.py gpurun_out/r02_bench_n1.json | head -14
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/r02_bench_ref.err | grep -a "^{" > gpurun_out/r02_bench_reference_arm.json; head -c 300 gpurun_out/r02_bench_reference_arm.json; echo
