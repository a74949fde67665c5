import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_debruijn_b200 as D
ctx = D.Context(0); L = ctx._L
R = 10_000_000
ss = D.SeqSet.synth(ctx, R, 1, 83886)
hw, hs, hl = ss.copy_out()
pw = torch.empty(len(hw), dtype=torch.int64, pin_memory=True); pw.numpy()[:] = hw.view(np.int64)
wp = pw.numpy().view(np.uint64)
import time
for i in range(4):
    gh = C.c_void_p()
    t0 = time.perf_counter()
    ctx.check(L.dbg_reads_to_graph_host_uniform(ctx._h, 31, C.c_void_p(wp.ctypes.data), len(wp), R, 150, None, 2, 0, 0, None, C.byref(gh)))
    t1 = time.perf_counter()
    s = ctx.stats()
    print(i, round((t1 - t0) * 1e3, 2), {k: round(v, 2) if isinstance(v, float) else v for k, v in s.items() if k in ("direct_partition", "ms_partition", "ms_count", "ms_k_partition", "ms_k_count", "ms_filter_total", "n_records")}, flush=True)
    L.dbg_graph_free(gh)
