timeout 400 ncu --set full --clock-control none --import-source on -k regex:vx_mesh_emit -s 2 -c 1 -o gpurun_out/r02_emit python bench.py --region 256 --steps 1 --warmup 3 --device-only > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vx_mesh_count -s 2 -c 1 -o gpurun_out/r02_count python bench.py --region 256 --steps 1 --warmup 3 --device-only > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
