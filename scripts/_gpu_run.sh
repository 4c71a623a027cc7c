mkdir -p gpurun_out
( SWEEP_G=2e3,2e4,1e5,2e5,5e5 SWEEP_READS=1e7 timeout 900 python scripts/sweep_probe.py 2>&1 | tail -8 ) > gpurun_out/r2a_sweep_G.log
( SWEEP_G=1e6 SWEEP_READS=4e6 timeout 900 python scripts/sweep_probe.py 2>&1 | tail -3 ) >> gpurun_out/r2a_sweep_G.log
cat gpurun_out/r2a_sweep_G.log
