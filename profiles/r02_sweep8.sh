O=gpurun_out
python -m embrace_b200.sweep --gpus ${1:-8} --rows 16384 --epochs 4 --out /tmp/sw > $O/r2x_sweep_${1:-8}.json 2> $O/r2x_sweep_${1:-8}.err
tail -n 1 $O/r2x_sweep_${1:-8}.json | cut -c1-900
