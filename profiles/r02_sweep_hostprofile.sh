# where the host time of a sweep job goes (cProfile of the 1-GPU driver; 6 jobs)
python - <<'PY' > gpurun_out/r2y_sweep_hostprofile.txt 2>&1
import cProfile, pstats, io
from embrace_b200 import sweep
pr = cProfile.Profile()
pr.enable()
sweep.main(['--gpus', '1', '--datasets', '2', '--rows', '16384', '--epochs', '4', '--out', '/tmp/swp'])
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(45)
print(s.getvalue())
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumtime').print_stats(60)
print(s.getvalue())
PY
head -120 gpurun_out/r2y_sweep_hostprofile.txt
