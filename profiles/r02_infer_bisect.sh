# bisection of the pooled inference kernels (EMB_CONV_DEBUG bits: 1 no pooling walk, 2 no MMAs, 8 test_wait spin)
for d in ${BISECT_MODES:-0 8 3 11}; do
echo "== EMB_CONV_DEBUG=$d"
EMB_CONV_DEBUG=$d timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pool" -c 3 python bench.py --workload infer --steps 1 --warmup 3 --cpu-baseline 0 2>&1 | grep -E "gpu__time" 
done
