set -x
python scripts/prof_run.py 2 2 > gpurun_out/prof_plain12.log 2>&1 || { tail -5 gpurun_out/prof_plain12.log; exit 1; }
LS3D_E2E_GRAPH=0 ncu --set full --clock-control none -k regex:"k_organized_count|k_map_cull_compact|k_copy_out|k_icp_match_packet|k_icp_stats|k_icp_sums|k_pack_ply_body|k_pack_transfer_body|k_triangles|k_chunk_scan" -c 24 -o gpurun_out/prof_r01_v12 -f python scripts/prof_run.py 2 2 > gpurun_out/ncu_v12.log 2>&1
tail -2 gpurun_out/ncu_v12.log
ncu -i gpurun_out/prof_r01_v12.ncu-rep --page raw --csv > gpurun_out/raw_v12.csv 2>/dev/null
ls -la gpurun_out/
LS3D_E2E_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"k_organized_count" -s 1 -c 1 -o gpurun_out/prof_r01_v12_org -f python scripts/prof_run.py 2 1 > gpurun_out/ncu_v12b.log 2>&1
ncu -i gpurun_out/prof_r01_v12_org.ncu-rep --page source --csv > gpurun_out/src_v12_org.csv 2>/dev/null
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain12.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 320 --csv --log-file gpurun_out/launches_r01_v12.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch12.log 2>&1
wc -l gpurun_out/launches_r01_v12.csv
rm -f gpurun_out/prof_r01_v12_org.ncu-rep
du -sh gpurun_out
