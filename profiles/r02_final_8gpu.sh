# Round-2 final 8-GPU run: sharded inference (config 5), fold-parallel sweep (config 4), data-parallel train step (config 3)
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 -m embrace_b200.infer --regions 50000000 > $O/r2w_infer50M_8.json 2> $O/r2w_infer50M_8.err
$TR --nproc-per-node 8 --master-port 29512 bench.py --workload infer --gpus 8 --cpu-baseline 0 > $O/r2w_bench_infer_8.json 2> $O/r2w_bench_infer_8.err
python -m embrace_b200.sweep --gpus 8 --rows 16384 --epochs 4 --out /tmp/sw8 > $O/r2w_sweep8.json 2> $O/r2w_sweep8.err
$TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --cpu-baseline 0 > $O/r2w_bench_dp8.json 2> $O/r2w_bench_dp8.err
$TR --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --cpu-baseline 0 > $O/r2w_bench_dp4.json 2> $O/r2w_bench_dp4.err
for f in infer50M_8 bench_infer_8 sweep8 bench_dp8 bench_dp4; do tail -n 1 $O/r2w_$f.json | cut -c1-420; done
