mkdir -p gpurun_out
timeout 400 python tools/check_conv_tma.py --quick > gpurun_out/t45_tma_check.log 2>&1; rc=$?
echo "harness rc=$rc" >> gpurun_out/t45_tma_check.log
timeout 100 python tools/tma_profile.py > gpurun_out/t45_tma_roles.log 2>&1
if [ $rc -ne 0 ]; then export SCSFM_CONV_TMA=0; fi
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t45_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t45_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/t45_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/t45_bench.log
tail -3 gpurun_out/t45_tma_check.log; tail -3 gpurun_out/t45_pytest.log; tail -c 600 gpurun_out/t45_bench.log
