"""Multi-GPU form of the path through the C ABI (dbg_comm_* / dbg_multi_* in include/dbg_b200.h): the reference's sharded
flow (src/test.rs:418-470: msp_sequence -> per-shard filter_kmers -> compress -> combine + compress_graph) as ONE collective
call whose per-rank outputs, concatenated in rank order, are the single-GPU BaseGraph bit for bit.

All data-path communication (super-k-mer records stored straight into the owners' peer-mapped windows over NVLink; NCCL
all-to-alls of neighbour queries, walkers and finished nodes) happens inside libdbg_b200.so.  Python only hands every rank the same NCCL unique id:

    comm = Comm.from_torch(ctx)          # one process per GPU under torchrun: id broadcast through torch.distributed
    g = comm.reads_to_graph(seqs, CountFilter(2), SimpleCompress(SAT_ADD), k=31)

    mc = MultiContext([0, 1, 2, 3])      # one process, one host thread per GPU (communicators owned by the handle)
    graphs = mc.reads_to_graph([seqs_0, seqs_1, seqs_2, seqs_3], CountFilter(2), SimpleCompress(SAT_ADD), k=31)
"""
import ctypes as C

from . import _lib
from .api import BaseGraph, Context, CountFilter, ScmapCompress, SimpleCompress


class MultiGraph(BaseGraph):
    """This rank's run of nodes of the complete BaseGraph (or the complete graph when `replicated`): node0 / base0 place it,
    n_nodes_total / n_bases_total / n_valid_total describe the whole job, `info` is the dbg_multi_info of the call."""

    def __init__(self, ctx, handle, info):
        super().__init__(ctx, handle)
        self.info = info
        self.node0, self.base0 = info["node0"], info["base0"]
        self.n_nodes_total, self.n_bases_total, self.n_valid_total = info["n_nodes_total"], info["n_bases_total"], info["n_valid_total"]
        self.replicated = bool(info["replicated"])
        self.invariants = {"ok": bool(info["check_ok"]),
                           "what": "all-reduced inside the library every call: every valid k-mer covered by exactly one unitig, "
                                   "sum of node lengths = V + M(K-1)",
                           "n_valid_total": self.n_valid_total, "n_nodes_total": self.n_nodes_total,
                           "n_bases_total": self.n_bases_total}


def _check_args(summarizer, spec):
    if not isinstance(summarizer, CountFilter):
        raise TypeError("only CountFilter is on the accelerated path (SURVEY.md §8)")
    if not isinstance(spec, (SimpleCompress, ScmapCompress)):
        raise TypeError("only SimpleCompress / ScmapCompress are on the accelerated path (SURVEY.md §8)")


class Comm:
    """One rank of a multi-GPU job (dbg_comm): the ctx's device + an NCCL communicator owned by the library."""

    def __init__(self, ctx, handle):
        self.ctx, self._h = ctx, handle
        L = ctx._L
        self.rank, self.size = L.dbg_comm_rank(handle), L.dbg_comm_size(handle)
        self.transport = L.dbg_comm_transport(handle).decode()

    @staticmethod
    def unique_id(ctx):
        buf = C.create_string_buffer(128)
        st = ctx._L.dbg_comm_unique_id(buf)
        if st != 0:
            raise _lib.DbgError(st, "dbg_comm_unique_id failed: libnccl.so.2 not loadable (the multi-GPU path needs NCCL)")
        return buf.raw

    @staticmethod
    def create(ctx, n_ranks, rank, unique_id):
        h = C.c_void_p()
        ctx.check(ctx._L.dbg_comm_create(ctx._h, n_ranks, rank, C.c_char_p(unique_id), C.byref(h)))
        return Comm(ctx, h)

    @staticmethod
    def from_torch(ctx, group=None):
        """Rendezvous through an initialised torch.distributed process group: rank 0 draws the id, everybody gets it."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device("cuda", ctx.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t = torch.frombuffer(bytearray(Comm.unique_id(ctx)), dtype=torch.uint8).to(dev)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return Comm.create(ctx, world, rank, bytes(t.cpu().numpy().tobytes()))

    def reads_to_graph(self, seqs, summarizer, spec, stranded=False, k=31):
        """filter_kmers + compress_kmers_with_hash over ALL ranks' sequences (collective)."""
        _check_args(summarizer, spec)
        info, gh = _lib.MultiInfo(), C.c_void_p()
        self.ctx.check(self.ctx._L.dbg_reads_to_graph_multi(self._h, k, seqs._h, summarizer.min_kmer_obs, int(bool(stranded)),
                                                            spec.func, C.byref(info), C.byref(gh)))
        return MultiGraph(self.ctx, gh, info.as_dict())

    def close(self):
        if self._h:
            self.ctx._L.dbg_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _BorrowedContext(Context):
    """A ctx owned by a dbg_multi handle (never destroyed from here)."""

    def __init__(self, L, handle, device):
        self._L, self._h, self.device = L, C.c_void_p(handle), device

    def close(self):
        self._h = None


class MultiContext:
    """One process driving several ranks, one host thread per rank inside the library (dbg_multi).  devices[i] = CUDA device of
    rank i; a repeated device selects the host-staged "local" transport (several ranks on one GPU: test configuration)."""

    def __init__(self, devices):
        self._L = _lib.lib()
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        st = self._L.dbg_multi_create(arr, len(devices), C.byref(h))
        if st != 0:
            raise _lib.DbgError(st, f"dbg_multi_create({list(devices)}) failed")
        self._h = h
        self.size = len(devices)
        self.transport = self._L.dbg_multi_transport(h).decode()
        self.ctxs = [_BorrowedContext(self._L, self._L.dbg_multi_ctx(h, r), devices[r]) for r in range(self.size)]

    def reads_to_graph(self, seqs_per_rank, summarizer, spec, stranded=False, k=31):
        _check_args(summarizer, spec)
        n = self.size
        if len(seqs_per_rank) != n:
            raise ValueError("one sequence set per rank")
        sp = (C.c_void_p * n)(*[s._h for s in seqs_per_rank])
        infos = (_lib.MultiInfo * n)()
        gs = (C.c_void_p * n)()
        st = self._L.dbg_multi_reads_to_graph(self._h, k, sp, summarizer.min_kmer_obs, int(bool(stranded)), spec.func, infos, gs)
        if st != 0:
            msgs = [self._L.dbg_last_error(c._h).decode() for c in self.ctxs]
            raise _lib.DbgError(st, " | ".join(m for m in msgs if m))
        return [MultiGraph(self.ctxs[r], C.c_void_p(gs[r]), infos[r].as_dict()) for r in range(n)]

    def close(self):
        if self._h:
            for c in self.ctxs:
                c.close()
            self._L.dbg_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
