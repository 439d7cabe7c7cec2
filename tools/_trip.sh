mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/t46_bench_full.log 2>&1; echo "bench rc=$?" >> gpurun_out/t46_bench_full.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 3400 -c 1150 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/t46_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tma_kernel -s 70 -c 6 -f -o gpurun_out/r1b_tma python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/t46_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_tc_kernel -s 40 -c 5 -f -o gpurun_out/r1b_wgrad python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/t46_ncu3.log 2>&1; echo "ncu3 rc=$?"
tail -c 1500 gpurun_out/t46_bench_full.log; ls -la gpurun_out | tail -5; tail -2 gpurun_out/r1b_launches.csv | cut -c1-200
