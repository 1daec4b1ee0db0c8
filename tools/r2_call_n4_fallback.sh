set -x
O=gpurun_out/r2n4
mkdir -p $O
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $O/gpus.txt; free -g >> $O/gpus.txt; df -h /dev/shm >> $O/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
# cfg4 at N=8 (the headline workload of the scaling run)
timeout 600 $TR --master-port 29711 bench.py --gpus 4 --no-extra > $O/bench_cfg4_n4.json 2> $O/bench_cfg4_n4.err; tail -2 $O/bench_cfg4_n4.err; python -c "
import json;d=json.load(open('$O/bench_cfg4_n4.json'));print('cfg4 N8 ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'parity',d['parity']['bit_exact'])"
# cfg5: R-MAT-26, replicated tables (fits: 2 x 32 GiB + 9 GiB CSR per GPU), with the 1-GPU-vs-8-GPU check
timeout 1500 $TR --master-port 29712 bench.py --gpus 4 --workload cfg5 --steps 3 --warmup 2 --no-extra > $O/bench_cfg5_n4_replicated.json 2> $O/bench_cfg5_n4_replicated.err
rc=$?; tail -3 $O/bench_cfg5_n4_replicated.err; cut -c1-1500 $O/bench_cfg5_n4_replicated.json
if [ $rc -ne 0 ]; then
  timeout 1500 $TR --master-port 29714 bench.py --gpus 4 --workload cfg5 --steps 3 --warmup 2 --no-extra --multicast 0 > $O/bench_cfg5_n4_replicated_unicast.json 2> $O/bench_cfg5_n4_replicated_unicast.err; tail -3 $O/bench_cfg5_n4_replicated_unicast.err; cut -c1-1500 $O/bench_cfg5_n4_replicated_unicast.json
fi
# cfg5 row-sharded (what BASELINE names): each GPU stores 1/8 of both tables
timeout 1500 $TR --master-port 29713 bench.py --gpus 4 --workload cfg5 --sharded 1 --steps 2 --warmup 1 --no-extra > $O/bench_cfg5_n4_sharded.json 2> $O/bench_cfg5_n4_sharded.err; tail -3 $O/bench_cfg5_n4_sharded.err; cut -c1-1500 $O/bench_cfg5_n4_sharded.json
ls -la /dev/shm | head; du -sh $O
# the drop-in CLI on 1 and 8 GPUs of this node (one process, one host thread per GPU, NVLink multicast): whole-run
# wall as the reference reports it (Results.txt: init + epochs), R-MAT-20 from the binary CSR cache
python - <<'PY'
import sys; sys.path.insert(0, '.')
from force2vec_b200 import host
rp, ci = host.rmat_csr_cached(20, 16, 1)
import numpy as np
host.write_csr('/dev/shm/rmat20.f2vcsr', np.ascontiguousarray(rp), np.ascontiguousarray(ci))
PY
mkdir -p /tmp/cli1 /tmp/cli4
(cd /tmp/cli1 && $OLDPWD/bin/Force2Vec -input /dev/shm/rmat20.f2vcsr -output /tmp/cli1/ -iter 50 -batch 65536 -dim 128 -option 6 -chunk 64 > $OLDPWD/$O/cli_gpus1.log 2>&1; cat Results.txt >> $OLDPWD/$O/cli_results.txt)
(cd /tmp/cli4 && $OLDPWD/bin/Force2Vec -input /dev/shm/rmat20.f2vcsr -output /tmp/cli4/ -iter 50 -batch 65536 -dim 128 -option 6 -chunk 64 -gpus 4 > $OLDPWD/$O/cli_gpus4.log 2>&1; cat Results.txt >> $OLDPWD/$O/cli_results.txt)
cmp /tmp/cli1/*.embd /tmp/cli4/*.embd && echo "CLI -gpus 4 .embd == -gpus 1 .embd (byte-identical)" >> $O/cli_results.txt
cat $O/cli_results.txt; tail -3 $O/cli_gpus4.log
