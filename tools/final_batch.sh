cd /root/repo
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/t23.log 2>&1; echo pytest_rc=$?; tail -3 gpurun_out/t23.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke4.log 2>&1; echo smoke_rc=$?; tail -2 gpurun_out/smoke4.log
timeout 120 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r02d.csv python tools/one_step.py tf32x3 kitti_r18 > gpurun_out/ncu_list4.log 2>&1; echo ncu_rc=$?
python tools/ncu_summarise.py gpurun_out/launches_r02d.csv r02 "python tools/one_step.py tf32x3 kitti_r18 (one eager step between cudaProfilerStart/Stop; ncu --profile-from-start off; final build of the round)" > /dev/null 2>&1
timeout 200 python bench.py > gpurun_out/bench_final2_r18.log 2>&1; echo r18_rc=$?; tail -1 gpurun_out/bench_final2_r18.log | cut -c1-400
timeout 130 python bench.py --config kitti_r50 --no-cpu-baseline > gpurun_out/bench_final2_r50.log 2>&1; echo r50_rc=$?; tail -1 gpurun_out/bench_final2_r50.log | cut -c1-300
timeout 100 python bench.py --config nyu_r18 --no-cpu-baseline > gpurun_out/bench_final2_nyu.log 2>&1; echo nyu_rc=$?; tail -1 gpurun_out/bench_final2_nyu.log | cut -c1-300
timeout 60 python tools/conv_layers.py --mode tf32x3 > gpurun_out/layers_x3_final2.log 2>&1; tail -4 gpurun_out/layers_x3_final2.log
timeout 60 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_thin -c 1 -o gpurun_out/r02_wgrad_thin python tools/conv_layers.py --mode tf32x3 --only "dec 0_1" --pass wgrad --reps 1 > gpurun_out/ncu_full3.log 2>&1; echo ncufull_rc=$?
timeout 80 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_augment_gpu.py tests/test_nets_gpu.py -q -x -k "golden_vectors or (thin and case0) or (thin and case4)" > gpurun_out/sanitizer2.log 2>&1; echo san_rc=$?; tail -4 gpurun_out/sanitizer2.log
