# multi-GPU bench job (developer helper): bash tests/gpu_job_multi_dev.sh <N>
N=$1
for w in g1_n21 g2_n18; do
  timeout -s KILL 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 --workload $w > gpurun_out/r1d_bench_${w}_N$N.json 2> gpurun_out/r1d_bench_${w}_N$N.err
  tail -1 gpurun_out/r1d_bench_${w}_N$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['e2e']['value'], d['config']['shard_config'], d['result_consistent'])"
done
