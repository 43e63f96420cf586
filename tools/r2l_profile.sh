set -x
timeout 200 ncu --set full --clock-control none --import-source on -k regex:sn_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r2l_sn_headline python bench.py --samples 200000 --steps 1 --warmup 3 --paths off --no-cpu-baseline > gpurun_out/r2l_ncu1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:.*sn_kernel<0, 0, 0, 0, 0, 1>.*' --launch-skip 2 -c 1 -f -o gpurun_out/r2l_sn_bolo python bench.py --samples 200000 --steps 1 --warmup 1 --paths on --only config1 --parity-samples 200 --no-cpu-baseline > gpurun_out/r2l_ncu2.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2l_ncu3.log 2>&1
ls -la gpurun_out/r2l_*; tail -2 gpurun_out/r2l_ncu1.log gpurun_out/r2l_ncu2.log | cut -c1-300; wc -l gpurun_out/r2l_launches.csv
