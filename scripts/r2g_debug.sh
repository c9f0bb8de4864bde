mkdir -p gpurun_out
cat > /tmp/small.py <<'PY'
import sys, faulthandler; faulthandler.enable(); sys.path.insert(0, '.')
from adaptiveviscositysolver_b200 import Params, Solver, sphere_drop
sc = sphere_drop(32, 10)
print("create", flush=True); s = Solver(device=0)
out = [v.data.copy() for v in sc.vel]
print("solve", flush=True)
print(s.solve(sc, Params(octree_levels=4, tolerance=1e-6), out).iterations, flush=True)
PY
AVS_TRACE=1 python /tmp/small.py > gpurun_out/r2g_small.log 2>&1; tail -30 gpurun_out/r2g_small.log
if which gdb > /dev/null; then gdb -batch -ex run -ex bt --args python /tmp/small.py > gpurun_out/r2g_gdb.log 2>&1; tail -40 gpurun_out/r2g_gdb.log; fi
