# one --set full capture of the count kernel and of the main pass of the tile kernel (second step of profile_step.py)
tag=$1; shift
ncu --set full --clock-control none --import-source on -k regex:count_kernel -s 1 -c 1 -f -o gpurun_out/${tag}_cnt python tools/profile_step.py --steps 2 "$@" > gpurun_out/${tag}_cnt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:msp_tile_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_tile python tools/profile_step.py --steps 2 "$@" > gpurun_out/${tag}_tile.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
