# bisection of the pooled inference kernels (EMB_CONV_DEBUG bits: 1 no pooling walk, 8 test_wait spin; bits 2 / 4 -- no MMAs / one MMA per
# block -- existed until the issue loops were specialised; results of the full bisection: profiles/r02_infer_bisect.txt)
for d in ${BISECT_MODES:-0 1 8}; do
echo "== EMB_CONV_DEBUG=$d"
EMB_CONV_DEBUG=$d timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pool" -c 3 python bench.py --workload infer --steps 1 --warmup 3 --cpu-baseline 0 2>&1 | grep -E "gpu__time" 
done
