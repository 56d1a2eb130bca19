# usage: bash tools/_run_n.sh N   (inside gpurun --gpus N)
N=$1
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for w in c4 c3; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 --workload $w --no-configs --no-cpu-baseline 2> gpurun_out/r02_bench_${w}_n$N.err | grep -a "^{" > gpurun_out/r02_bench_${w}_n$N.json
  python tools/show_bench.py gpurun_out/r02_bench_${w}_n$N.json | head -14
done
