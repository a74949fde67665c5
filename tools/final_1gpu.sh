# round-end evidence on one GPU: tests, bench line, reference arm, ncu launch list, one --set full capture of the two dominant kernels
tag=${1:-r02_final}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -n 3 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python tools/show_bench.py gpurun_out/${tag}_bench.json
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; tail -c 600 gpurun_out/${tag}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c3 > gpurun_out/${tag}_ncu_bench.log 2>&1
bash tools/ncu_two.sh ${tag}
