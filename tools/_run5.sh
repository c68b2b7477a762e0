VPD_GRAPH=0 timeout 800 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:halo_kernel<\(int\)1, \(int\)2' -c 2 -o gpurun_out/r02y_dgrad1 python tools/prof_step.py 1 > gpurun_out/r02y_ncu.log 2>&1
tail -3 gpurun_out/r02y_ncu.log
ls -la gpurun_out/r02y_dgrad1.ncu-rep
