"""Per-step wall times of the e2e path (pinned host reads in, BaseGraph arrays out) — debugging aid."""
import ctypes as C, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_debruijn_b200 as D
ctx = D.Context(0); L = ctx._L
R = 10_000_000
ss = D.SeqSet.synth(ctx, R, 1, 83886)
hw, hs, hl = ss.copy_out()
pw = torch.empty(len(hw), dtype=torch.int64, pin_memory=True); pw.numpy()[:] = hw.view(np.int64)
wp = pw.numpy().view(np.uint64)
if "--resident-first" in sys.argv:
    for i in range(4):
        g = D.reads_to_graph(ss, D.CountFilter(2), D.SimpleCompress(0), k=31); g.free()
out = dict(words=torch.empty(12_000_000, dtype=torch.int64, pin_memory=True), start=torch.empty(10_000_000, dtype=torch.int64, pin_memory=True),
           length=torch.empty(10_000_000, dtype=torch.int32, pin_memory=True), exts=torch.empty(10_000_000, dtype=torch.uint8, pin_memory=True),
           data=torch.empty(10_000_000, dtype=torch.int16, pin_memory=True))
for i in range(10):
    gh = C.c_void_p()
    t0 = time.perf_counter()
    ctx.check(L.dbg_reads_to_graph_host_uniform(ctx._h, 31, C.c_void_p(wp.ctypes.data), len(wp), R, 150, None, 2, 0, 0, None, C.byref(gh)))
    t1 = time.perf_counter()
    ctx.check(L.dbg_graph_copy_out(gh, *(C.c_void_p(out[k].data_ptr()) for k in ("words", "start", "length", "exts", "data"))))
    t2 = time.perf_counter()
    L.dbg_graph_free(gh)
    t3 = time.perf_counter()
    s = ctx.stats()
    print(i, "compute %.2f copy_out %.2f free %.2f | part %.2f count %.2f sort %.2f filt %.2f comp %.2f" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3,
          s["ms_partition"], s["ms_count"], s["ms_sort"], s["ms_filter_total"], s["ms_compress_total"]), flush=True)
