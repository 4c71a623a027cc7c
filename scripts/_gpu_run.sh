mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r2i_tests.log
( SWEEP_G=2e5 SWEEP_READS=4e6 timeout 300 python scripts/sweep_probe.py 2>&1 | tail -5 ) > gpurun_out/r2i_sweep.log
( MLG_MZ_FBITS=34 SWEEP_G=2e5 SWEEP_READS=4e6 timeout 300 python scripts/sweep_probe.py 2>&1 | tail -5 ) >> gpurun_out/r2i_sweep.log
( MLG_MZ_FBITS=35 SWEEP_G=2e5 SWEEP_READS=4e6 timeout 300 python scripts/sweep_probe.py 2>&1 | tail -5 ) >> gpurun_out/r2i_sweep.log
( MLG_MZ_FBITS=32 SWEEP_G=2e5 SWEEP_READS=4e6 timeout 300 python scripts/sweep_probe.py 2>&1 | tail -5 ) >> gpurun_out/r2i_sweep.log
SWEEP_G=2e5 SWEEP_READS=4e6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_minimizer -s 2 -c 1 -f -o gpurun_out/r2i_k1mz python scripts/sweep_probe.py > gpurun_out/r2i_ncu.log 2>&1
cat gpurun_out/r2i_tests.log gpurun_out/r2i_sweep.log
