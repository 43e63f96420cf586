#!/bin/bash
# ncu --set full of the bolometric sn_kernel instantiation (BASELINE config 1: arnett_bolometric, E = 100)
timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:sn_kernel.*\(int\)1>' --launch-skip 2 -c 1 -f -o gpurun_out/r2m_sn_bolo \
  python bench.py --samples 200000 --steps 1 --warmup 1 --paths on --only config1 --parity-samples 200 --no-cpu-baseline \
  > gpurun_out/r2m_ncu_bolo.log 2>&1
ls -la gpurun_out/r2m_sn_bolo.ncu-rep; tail -n 3 gpurun_out/r2m_ncu_bolo.log | cut -c1-300
