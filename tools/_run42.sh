cd /root/repo
bash tools/collect_profiles.sh > gpurun_out/collect.log 2>&1; echo "collect rc=$?"
R48_LIBRARY=tools/ab/libr48_fma.so bash tools/ncu_quick.sh fma; echo "ncu fma rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_ref_n1.json 2> gpurun_out/r2_ref_n1.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
rm -f gpurun_out/*.ncu-rep
