set -x
O=gpurun_out/r2c
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
./tools/probes/random_gather_probe 8 512 > $O/random_gather_512.log 2>&1; cat $O/random_gather_512.log
./tools/probes/random_gather_probe 8 256 > $O/random_gather_256.log 2>&1; cat $O/random_gather_256.log
python bench.py --no-extra --no-cpu-baseline --no-e2e > $O/bench_cfg4_n1_clampless.json 2> $O/bench_cfg4_n1.err; cut -c1-400 $O/bench_cfg4_n1_clampless.json
python -m pytest tests/test_multi_gpu.py -x -q -k "not scale26" > $O/pytest_mgpu.log 2>&1; tail -15 $O/pytest_mgpu.log
for B in 65536 262144; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --batch $B --no-extra > $O/bench_cfg4_n2_B$B.json 2> $O/bench_cfg4_n2_B$B.err; tail -3 $O/bench_cfg4_n2_B$B.err; cut -c1-3000 $O/bench_cfg4_n2_B$B.json
done
python bench.py --batch 262144 --no-extra --no-cpu-baseline --no-e2e > $O/bench_cfg4_n1_B262144.json 2>> $O/bench_cfg4_n1.err; cut -c1-400 $O/bench_cfg4_n1_B262144.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --workload cfg2 --no-extra > $O/bench_cfg2_n2.json 2> $O/bench_cfg2_n2.err; tail -3 $O/bench_cfg2_n2.err; cut -c1-1200 $O/bench_cfg2_n2.json
