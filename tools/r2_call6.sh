set -x
O=gpurun_out/r2f
mkdir -p $O
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -k "not scale26" > $O/pytest_mgpu.log 2>&1; tail -4 $O/pytest_mgpu.log; grep -i "differ\|MGPU\|max diff" $O/pytest_mgpu.log | head -20
for B in 65536 262144; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --batch $B --no-extra > $O/bench_cfg4_n2_B$B.json 2> $O/bench_cfg4_n2_B$B.err; tail -2 $O/bench_cfg4_n2_B$B.err; cut -c1-250 $O/bench_cfg4_n2_B$B.json; python -c "
import json;d=json.load(open('$O/bench_cfg4_n2_B$B.json'));print('N2 B$B ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'parity',d['parity']['bit_exact'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --workload cfg2 --no-extra > $O/bench_cfg2_n2.json 2> $O/bench_cfg2_n2.err; python -c "
import json;d=json.load(open('$O/bench_cfg2_n2.json'));print('cfg2 N2 ms',d['ms_per_step'],'parity',d['parity']['bit_exact'])"
# A/B: gathered rows through L1 (.ca) vs L2 only (.cg), one GPU each
(CUDA_VISIBLE_DEVICES=0 python tools/r2_tune_ring.py 24 5 128 1 65536 -1,-1 > $O/ab_cfg4_cg.log 2>&1; CUDA_VISIBLE_DEVICES=0 F2V_LIB=$PWD/force2vec_b200/lib/libf2v_ca.so python tools/r2_tune_ring.py 24 5 128 1 65536 -1,-1 > $O/ab_cfg4_ca.log 2>&1) &
(CUDA_VISIBLE_DEVICES=1 python tools/r2_tune_ring.py 20 6 128 0 65536 -1,-1 > $O/ab_cfg2_cg.log 2>&1; CUDA_VISIBLE_DEVICES=1 F2V_LIB=$PWD/force2vec_b200/lib/libf2v_ca.so python tools/r2_tune_ring.py 20 6 128 0 65536 -1,-1 > $O/ab_cfg2_ca.log 2>&1; CUDA_VISIBLE_DEVICES=1 python tools/r2_tune_ring.py 22 7 64 0 65536 -1 > $O/ab_cfg3_cg.log 2>&1; CUDA_VISIBLE_DEVICES=1 F2V_LIB=$PWD/force2vec_b200/lib/libf2v_ca.so python tools/r2_tune_ring.py 22 7 64 0 65536 -1 > $O/ab_cfg3_ca.log 2>&1) &
wait
for f in $O/ab_*.log; do echo $f; cut -c1-260 $f; done
