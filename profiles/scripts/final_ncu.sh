mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_b64_fused_final.csv python profiles/run_generate.py 64 14 1.0 nocodec > /dev/null 2>&1
ncu --kernel-name-base demangled -k regex:decode_step_fused --launch-skip 114 --launch-count 1 --set full --import-source on --clock-control none -f -o gpurun_out/fused_final python profiles/run_generate.py 64 120 1.0 nocodec > /dev/null 2>&1
ncu -i gpurun_out/fused_final.ncu-rep --page raw --csv > gpurun_out/r01_ncu_full_fused_b64_final.csv 2>/dev/null
ls -la gpurun_out | tail -5
