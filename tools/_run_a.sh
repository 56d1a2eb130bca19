cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 2> gpurun_out/r02_bench_n1.err | grep -a "^{" > gpurun_out/r02_bench_n1.json; python tools/show_bench.py gpurun_out/r02_bench_n1.json | head -13
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n1.json').read()); print(d.get('collapsed_opt_in'))"
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
