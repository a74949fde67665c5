# usage: bash tools/ncu_one.sh <tag> <kernel regex> <skip> [profile_step args]
tag=$1; re=$2; skip=$3; shift 3
ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c 1 -f -o gpurun_out/${tag} python tools/profile_step.py --steps 2 "$@" > gpurun_out/${tag}.log 2>&1
ls -la gpurun_out/${tag}.ncu-rep
