# compute-sanitizer memcheck over the node-level graph operations and the multi-rank path (local transport, small inputs)
compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -x -q -k "compress_graph and not long_chains" > gpurun_out/r02_memcheck_graph.log 2>&1; tail -n 5 gpurun_out/r02_memcheck_graph.log
compute-sanitizer --tool memcheck --print-limit 10 python tools/multi_check.py --local 2 --reads 6000 > gpurun_out/r02_memcheck_multi.log 2>&1; tail -n 4 gpurun_out/r02_memcheck_multi.log
