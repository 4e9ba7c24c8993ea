set -x
SM="python bench.py --steps 1 --warmup 3 --sims 24 --preroll 40 --no-cpu-baseline --no-vl --no-train"
timeout 300 $SM > gpurun_out/r3r_small.json 2> gpurun_out/r3r_small.err; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3r_launches.csv $SM > gpurun_out/r3r_ncu1.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv1_taps|conv_umma|fc_umma|advance" -s 200 -c 10 -o gpurun_out/r3r_full -f $SM > gpurun_out/r3r_ncu2.log 2>&1; echo rc=$?
ls -la gpurun_out/r3r_full.ncu-rep
