mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t48_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/t48_smoke.log
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/t48_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t48_pytest.log
timeout 120 python bench.py --no-cpu-baseline > gpurun_out/t48_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/t48_bench.log
tail -3 gpurun_out/t48_smoke.log; tail -4 gpurun_out/t48_pytest.log; tail -c 700 gpurun_out/t48_bench.log
