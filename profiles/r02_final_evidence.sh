# Round-2 final evidence run (one B200; SKIP_TESTS=1 skips the test suite): tests, smoke, the three bench workloads, the ncu launch list of one train step and the
# full-set captures the DESIGN.md numbers cite.  Everything lands in gpurun_out/ and is summarised into profiles/ afterwards.
set -x
O=gpurun_out
if [ -z "$SKIP_TESTS" ]; then (time python -m pytest tests -m gpu -q -x --durations=8) > $O/r2v_pytest.log 2>&1; tail -15 $O/r2v_pytest.log; fi
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2v_smoke.log 2>&1; tail -3 $O/r2v_smoke.log
python bench.py > $O/r2v_bench_L8192.json 2> $O/r2v_bench_L8192.err
python bench.py --workload S256 > $O/r2v_bench_S256.json 2> $O/r2v_bench_S256.err
python bench.py --workload infer > $O/r2v_bench_infer.json 2> $O/r2v_bench_infer.err
EMB_INFER_FUSE=0 python bench.py --workload infer --cpu-baseline 0 > $O/r2v_bench_infer_unfused.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2v_bench_reference.json 2> $O/r2v_bench_reference.err
for f in L8192 S256 infer infer_unfused reference; do python -c "
import json;d=json.loads(open('$O/r2v_bench_$f.json').read().strip().splitlines()[-1]);print('$f',d.get('value'),d.get('ms_per_step'),(d.get('e2e') or {}).get('value'),(d.get('roofline') or {}).get('frac'))"; done
(time python -m embrace_b200.sweep --gpus 1 --out /tmp/sw1) > $O/r2v_sweep1.log 2>&1; tail -4 $O/r2v_sweep1.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv --log-file $O/r2v_launches_L8192.csv python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --graph 0 > $O/r2v_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file $O/r2v_launches_infer.csv python bench.py --workload infer --steps 2 --warmup 3 --cpu-baseline 0 > $O/r2v_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"onehot_conv_pool_tc_kernel|tc_conv_pool_kernel" -c 3 -o $O/r02_infer_final python bench.py --workload infer --steps 2 --warmup 3 --cpu-baseline 0 > $O/r2v_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k2_fwd_packed_kernel|pool_bn_bwd_tma_kernel|bn_stats_v2_kernel|opt_step_kernel|onehot_conv" --launch-skip 40 -c 14 -o $O/r02_train_elem_final python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --graph 0 > $O/r2v_ncu4.log 2>&1
ncu --set full --clock-control none -k regex:"tc_gemm_kernel|tc_conv_reuse_kernel" --launch-skip 96 -c 32 -o $O/r02_train_gemm_final python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --graph 0 > $O/r2v_ncu5.log 2>&1
# gpurun brings back at most 64 MiB: export the tables here and drop the reports
for r in r02_infer_final r02_train_elem_final r02_train_gemm_final; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
done
ncu -i $O/r02_infer_final.ncu-rep --page source --csv > $O/r02_infer_final.source.csv 2>/dev/null
ls -la $O/*.ncu-rep; rm -f $O/*.ncu-rep; du -sh $O
