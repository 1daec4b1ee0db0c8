set -x
O=gpurun_out/r2d
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
timeout 600 python tools/r2_tune_flow.py 20 6 128 0 256,1024,4096,65536 > $O/flow_cfg2.log 2>&1; cut -c1-400 $O/flow_cfg2.log
timeout 600 python tools/r2_tune_flow.py 20 5 128 1 256,4096 > $O/flow_rmat20_opt5bs1.log 2>&1; cut -c1-400 $O/flow_rmat20_opt5bs1.log
timeout 600 python tools/r2_tune_flow.py 22 7 64 0 256,4096 > $O/flow_cfg3.log 2>&1; cut -c1-400 $O/flow_cfg3.log
for v in -1 21 22; do VARIANT=$v timeout 600 python tools/r2_probe_cfg.py 24 5 128 1 65536 3 > $O/probe_cfg4_v$v.log 2>&1; tail -1 $O/probe_cfg4_v$v.log | cut -c1-700; done
python - <<'PY' > $O/cora_modes.log 2>&1
import sys, time; sys.path.insert(0, '.')
import numpy as np, force2vec_b200 as F
from force2vec_b200 import host
rp, ci = host.load_mtx('tests/golden/cora.mtx')
out = {}
for mode in (0, 2, 0, 2):
    alg = F.Algorithms(rp, ci, 'cora.mtx', '/tmp/', 128); alg.epoch_mode = mode
    alg.AlgoForce2VecNS(20, 0, 256, 5, 0.02, write=False)
    sec = alg.AlgoForce2VecNS(1200, 0, 256, 5, 0.02, write=False)[0]
    out.setdefault(mode, alg.nCoordinates.copy())
    print("cora option5 B256 it1200 mode", mode, "wall_s", sec, flush=True)
print("mode0 == mode2 (adaptive chunk):", np.array_equal(out[0], out[2]))
PY
cat $O/cora_modes.log
