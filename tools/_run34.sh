cd /root/repo
python -m pytest tests -m gpu -q > gpurun_out/r34_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r34_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"
R48_LIBRARY=tools/ab/libr48_fma.so bash tools/ncu_quick.sh fma; echo "ncu fma rc=$?"
