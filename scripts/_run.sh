python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:k_rad_chains -s 2 -c 1 -o gpurun_out/prof_chains -f python scripts/radial_time.py > gpurun_out/ncu_chains.log 2>&1
ncu -i gpurun_out/prof_chains.ncu-rep --page source --csv > gpurun_out/src_v13_chains.csv 2>/dev/null
ncu -i gpurun_out/prof_chains.ncu-rep --page raw --csv > gpurun_out/raw_v13_chains.csv 2>/dev/null
rm -f gpurun_out/prof_chains.ncu-rep
python scripts/frame_stage_time.py 50 2>&1 | tail -1
