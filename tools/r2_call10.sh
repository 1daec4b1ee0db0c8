set -x
O=gpurun_out/r2j
mkdir -p $O
F2V_BIG_WORLD=4 F2V_BIG_SCALE=20 timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -k "scale26" > $O/pytest_big_n4.log 2>&1; tail -3 $O/pytest_big_n4.log
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -k "equals_single_gpu and 4" > $O/pytest_mgpu_n4.log 2>&1; tail -3 $O/pytest_mgpu_n4.log; grep -i "differ\|max diff" $O/pytest_mgpu_n4.log | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 4 --no-extra > $O/bench_cfg4_n4.json 2> $O/bench_cfg4_n4.err; tail -2 $O/bench_cfg4_n4.err; python -c "
import json;d=json.load(open('$O/bench_cfg4_n4.json'));print('cfg4 N4 ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'parity',d['parity'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 4 --workload cfg2 --no-extra > $O/bench_cfg2_n4.json 2> $O/bench_cfg2_n4.err; python -c "
import json;d=json.load(open('$O/bench_cfg2_n4.json'));print('cfg2 N4 ms',d['ms_per_step'],'parity',d['parity']['bit_exact'])"
MC=1 SCALE=24 MODEL=5 DIM=128 BS=1 BATCHES=262144,1048576 CHUNKS=0 ORDERS=1 SIGS=2 FREE=0 TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29713 tools/mgpu_probe.py > $O/mgpu_cfg4_n4.log 2> $O/mgpu_cfg4_n4.err; tail -2 $O/mgpu_cfg4_n4.err; cut -c1-330 $O/mgpu_cfg4_n4.log
