# racecheck / memcheck of the small all-paths driver + the new tests
compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitize_small.py > gpurun_out/r02_race.log 2>&1; tail -n 4 gpurun_out/r02_race.log
compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_small.py > gpurun_out/r02_memcheck.log 2>&1; tail -n 3 gpurun_out/r02_memcheck.log
python -m pytest tests -m gpu -x -q -k "fastq or compress_graph or multi_rank" 2>&1 | tail -n 3
