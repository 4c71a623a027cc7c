mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r2b_tests.log
( SWEEP_G=2e5 SWEEP_READS=4e6 timeout 300 python scripts/sweep_probe.py 2>&1 | tail -5 ) > gpurun_out/r2b_sweep_l2.log
cat gpurun_out/r2b_tests.log gpurun_out/r2b_sweep_l2.log
