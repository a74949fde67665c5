"""Host-side mirror of the reference interface for the read -> unitig path, over the C ABI.

Names, argument meaning and error behaviour follow the crate:
    filter::filter_kmers            src/filter.rs:139-231      -> filter_kmers()
    filter::CountFilter             src/filter.rs:40-63        -> CountFilter
    compression::SimpleCompress     src/compression.rs:40-65   -> SimpleCompress (fixed menu of reduce closures)
    compression::compress_kmers_with_hash  :588-594            -> compress_kmers_with_hash()
    compression::compress_kmers     :598-615                   -> compress_kmers()
    graph::BaseGraph                src/graph.rs:44-114        -> BaseGraph (len / sequences / exts / data / stranded)
    dna_string::PackedDnaStringSet  src/dna_string.rs:763-822  -> PackedDnaStringSet
The reference panics on misuse; here the same conditions raise DbgError with the ABI status.
All compute happens in libdbg_b200.so on the GPU; numpy is only used to hold host copies."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import DbgError

SAT_ADD, WRAP_ADD, ADD_MOD_65535, MAX = 0, 1, 2, 3
_SCMAP = 4


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context:
    """One CUDA device + stream + memory pool (dbg_ctx)."""

    def __init__(self, device=0):
        self._L = _lib.lib()
        h = C.c_void_p()
        st = self._L.dbg_ctx_create(device, C.byref(h))
        if st != 0:
            raise DbgError(st, f"dbg_ctx_create(device={device}) failed: no usable CUDA device (there is no CPU fallback)")
        self._h = h
        self.device = device

    def check(self, st):
        if st != 0:
            raise DbgError(st, self._L.dbg_last_error(self._h).decode())

    def set_param(self, name, value):
        self.check(self._L.dbg_ctx_set_param(self._h, name.encode(), int(value)))

    def stats(self):
        s = _lib.Stats()
        self.check(self._L.dbg_stats_get(self._h, C.byref(s)))
        return s.as_dict()

    def stream_ptr(self):
        return self._L.dbg_ctx_stream(self._h)

    def synchronize(self):
        self.check(self._L.dbg_ctx_synchronize(self._h))

    def close(self):
        if self._h:
            self._L.dbg_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class Exts:
    """src/lib.rs:577-749 (value type; only what the path's callers touch)."""

    def __init__(self, val=0):
        self.val = int(val) & 0xff

    @staticmethod
    def empty():
        return Exts(0)

    def __eq__(self, o):
        return isinstance(o, Exts) and o.val == self.val

    def __repr__(self):
        return "".join("ACGT"[i] for i in range(4) if self.val & (1 << i)) + "|" + \
               "".join("ACGT"[i] for i in range(4) if self.val & (16 << i))


class SeqSet:
    """Device-resident `&[(V, Exts, D1)]` (dbg_seqset)."""

    def __init__(self, ctx, handle):
        self.ctx, self._h = ctx, handle

    @staticmethod
    def upload(ctx, words, start, length, seq_exts=None):
        words = np.ascontiguousarray(words, np.uint64)
        start = np.ascontiguousarray(start, np.uint64)
        length = np.ascontiguousarray(length, np.uint32)
        if seq_exts is not None:
            seq_exts = np.ascontiguousarray(seq_exts, np.uint8)
        h = C.c_void_p()
        ctx.check(ctx._L.dbg_seqset_upload(ctx._h, _ptr(words), len(words), _ptr(start), _ptr(length), _ptr(seq_exts),
                                           len(start), C.byref(h)))
        return SeqSet(ctx, h)

    @staticmethod
    def from_ascii(ctx, ascii_seqs, seq_exts=None):
        """DnaString::from_acgt_bytes (src/dna_string.rs:224-250) for a list of ASCII sequences (bytes), packed on the
        device like PackedDnaStringSet::add.  Non-ACGT characters become A; their number is left in .n_invalid."""
        buf = b"".join(bytes(x) for x in ascii_seqs)
        length = np.array([len(x) for x in ascii_seqs], np.uint32)
        start = np.zeros(len(ascii_seqs), np.uint64)
        if len(ascii_seqs) > 1:
            start[1:] = np.cumsum(length[:-1], dtype=np.uint64)
        arr = np.frombuffer(buf, np.uint8) if buf else np.zeros(0, np.uint8)
        if seq_exts is not None:
            seq_exts = np.ascontiguousarray(seq_exts, np.uint8)
        h, bad = C.c_void_p(), C.c_uint64()
        ctx.check(ctx._L.dbg_seqset_from_ascii(ctx._h, _ptr(arr) if len(arr) else None, len(arr), _ptr(start), _ptr(length),
                                               _ptr(seq_exts), len(ascii_seqs), C.byref(bad), C.byref(h)))
        ss = SeqSet(ctx, h)
        ss.n_invalid = bad.value
        return ss

    @staticmethod
    def from_ascii_hashn(ctx, ascii_seqs, read_names, seq_exts=None):
        """DnaString::from_acgt_bytes_hashn (src/dna_string.rs:254-278) for a list of ASCII sequences with their read names:
        a non-ACGT character becomes the repeatable base DefaultHasher(name, position) % 4.  .n_invalid counts them."""
        if len(read_names) != len(ascii_seqs):
            raise ValueError("one read name per sequence")
        buf = b"".join(bytes(x) for x in ascii_seqs)
        nbuf = b"".join(bytes(x) for x in read_names)
        length = np.array([len(x) for x in ascii_seqs], np.uint32)
        nlen = np.array([len(x) for x in read_names], np.uint32)
        start = np.zeros(len(ascii_seqs), np.uint64)
        nstart = np.zeros(len(ascii_seqs), np.uint64)
        if len(ascii_seqs) > 1:
            start[1:] = np.cumsum(length[:-1], dtype=np.uint64)
            nstart[1:] = np.cumsum(nlen[:-1], dtype=np.uint64)
        arr = np.frombuffer(buf, np.uint8) if buf else np.zeros(1, np.uint8)
        narr = np.frombuffer(nbuf, np.uint8) if nbuf else np.zeros(1, np.uint8)
        if seq_exts is not None:
            seq_exts = np.ascontiguousarray(seq_exts, np.uint8)
        h, bad = C.c_void_p(), C.c_uint64()
        ctx.check(ctx._L.dbg_seqset_from_ascii_hashn(ctx._h, _ptr(arr), len(buf), _ptr(start), _ptr(length), _ptr(narr), len(nbuf),
                                                     _ptr(nstart), _ptr(nlen), _ptr(seq_exts), len(ascii_seqs), C.byref(bad), C.byref(h)))
        ss = SeqSet(ctx, h)
        ss.n_invalid = bad.value
        return ss

    @staticmethod
    def parse_fastq(data):
        """4-line FASTQ records (bytes or a path) -> (read names without '@' up to the first blank, sequence lines).  The crate has no
        FASTQ reader of its own: this is the caller-side loop that feeds DnaString::from_acgt_bytes[_hashn]."""
        if not isinstance(data, (bytes, bytearray)):
            with open(data, "rb") as f:
                data = f.read()
        lines = bytes(data).split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()
        if len(lines) % 4:
            raise ValueError("FASTQ: number of lines is not a multiple of 4")
        names, seqs = [], []
        for i in range(0, len(lines), 4):
            h, sq, plus = lines[i].rstrip(b"\r"), lines[i + 1].rstrip(b"\r"), lines[i + 2]
            if not h.startswith(b"@") or not plus.startswith(b"+"):
                raise ValueError("FASTQ: malformed record at line %d" % (i + 1))
            names.append(h[1:].split(None, 1)[0] if len(h) > 1 else b"")
            seqs.append(sq)
        return names, seqs

    @staticmethod
    def from_fastq(ctx, data, hashn=True):
        """A FASTQ file (path or bytes) as a device sequence set: sequences packed on the device, non-ACGT characters replaced as in
        DnaString::from_acgt_bytes_hashn (hash of read name + position; hashn=False: from_acgt_bytes, 'A')."""
        names, seqs = SeqSet.parse_fastq(data)
        return SeqSet.from_ascii_hashn(ctx, seqs, names) if hashn else SeqSet.from_ascii(ctx, seqs)

    @staticmethod
    def upload_uniform(ctx, words, n_seqs, read_len, seq_exts=None, pipelined=False):
        """pipelined=True: asynchronous chunked upload overlapping the partition stage of the next call; `words` (pinned
        for real overlap) must stay untouched until that call has returned (the SeqSet keeps a reference)."""
        words = np.ascontiguousarray(words, np.uint64)
        if seq_exts is not None:
            seq_exts = np.ascontiguousarray(seq_exts, np.uint8)
        h = C.c_void_p()
        fn = ctx._L.dbg_seqset_upload_uniform_async if pipelined else ctx._L.dbg_seqset_upload_uniform
        ctx.check(fn(ctx._h, _ptr(words), len(words), n_seqs, read_len, _ptr(seq_exts), C.byref(h)))
        ss = SeqSet(ctx, h)
        ss._keep = (words, seq_exts) if pipelined else None   # the asynchronous copies read these buffers
        return ss

    @staticmethod
    def synth(ctx, n_reads, seed=1, err_thr=0):
        h = C.c_void_p()
        ctx.check(ctx._L.dbg_seqset_synth(ctx._h, n_reads, seed, err_thr, C.byref(h)))
        return SeqSet(ctx, h)

    def __len__(self):
        return self.ctx._L.dbg_seqset_len(self._h)

    def copy_out(self):
        n, nw = len(self), self.ctx._L.dbg_seqset_n_words(self._h)
        words, start, length = np.zeros(nw, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint32)
        self.ctx.check(self.ctx._L.dbg_seqset_copy_out(self._h, _ptr(words), _ptr(start), _ptr(length)))
        return words, start, length

    def free(self):
        if self._h:
            if self.ctx._h:
                self.ctx._L.dbg_seqset_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CountFilter:
    """src/filter.rs:40-63: keep k-mers observed at least `min_kmer_obs` times; data = count capped at 65535."""

    def __init__(self, min_kmer_obs):
        self.min_kmer_obs = int(min_kmer_obs)
        if self.min_kmer_obs < 0:
            raise ValueError("CountFilter: min_kmer_obs must be >= 0")
        # counts saturate at 65535 (filter.rs:57): any larger threshold censors everything, and the ABI takes a u32
        self.min_kmer_obs = min(self.min_kmer_obs, 65536)


class CountFilterSet:
    """src/filter.rs:68-101 with D = u8: keep k-mers observed at least `min_kmer_obs` times; the summary of a k-mer is the sorted,
    deduplicated set of the labels (one per input sequence, < 64) it was observed with."""

    def __init__(self, min_kmer_obs):
        self.min_kmer_obs = int(min_kmer_obs)
        if not 0 <= self.min_kmer_obs <= 65535:
            raise ValueError("CountFilterSet: min_kmer_obs must be in [0, 65535]")


class SimpleCompress:
    """src/compression.rs:40-65.  The closure cannot cross to the GPU: `func` is one of the four
    reductions the reference's tests use (SAT_ADD, WRAP_ADD, ADD_MOD_65535, MAX); join_test is always true."""

    def __init__(self, func=SAT_ADD):
        if func not in (SAT_ADD, WRAP_ADD, ADD_MOD_65535, MAX):
            raise ValueError("SimpleCompress: func must be one of SAT_ADD, WRAP_ADD, ADD_MOD_65535, MAX")
        self.func = func


class ScmapCompress:
    """src/compression.rs:66-98: join_test = equality of the two k-mers' data, reduce keeps the (common) value."""
    func = _SCMAP


class KmerTable:
    """Device-resident stand-in for BoomHashMap2<K, Exts, u16>: ascending k-mers with exts and counts."""

    def __init__(self, ctx, handle):
        self.ctx, self._h = ctx, handle

    def __len__(self):  # BoomHashMap2::len
        return self.ctx._L.dbg_table_len(self._h)

    @property
    def k(self):
        return self.ctx._L.dbg_table_k(self._h)

    @property
    def n_input(self):
        return self.ctx._L.dbg_table_n_input(self._h)

    def to_host(self):
        L = self.ctx._L
        n, na, k = len(self), L.dbg_table_all_len(self._h), self.k
        two = k > 32
        out = dict(k=k, lo=np.zeros(n, np.uint64), hi=np.zeros(n if two else 0, np.uint64), exts=np.zeros(n, np.uint8),
                   counts=np.zeros(n, np.uint16), all_lo=np.zeros(na, np.uint64),
                   all_hi=np.zeros(na if two else 0, np.uint64), n_input=self.n_input)
        self.ctx.check(L.dbg_table_copy_out(self._h, _ptr(out["lo"]), _ptr(out["hi"]) if two else None,
                                            _ptr(out["exts"]), _ptr(out["counts"]), _ptr(out["all_lo"]),
                                            _ptr(out["all_hi"]) if two else None))
        return out

    def colorsets(self):
        """CountFilterSet tables: per k-mer the sorted list of labels (the reference's Vec<u8> summary), from the 64-bit masks."""
        n = len(self)
        masks = np.zeros(n, np.uint64)
        self.ctx.check(self.ctx._L.dbg_table_colorsets(self._h, _ptr(masks)))
        return [[c for c in range(64) if (int(m) >> c) & 1] for m in masks]

    def iter(self):
        """(kmer, exts, count) in table order, like BoomHashMap2::iter (order here: ascending k-mer)."""
        t = self.to_host()
        for i in range(len(t["lo"])):
            km = int(t["lo"][i]) | ((int(t["hi"][i]) << 64) if t["k"] > 32 else 0)
            yield km, Exts(t["exts"][i]), int(t["counts"][i])

    def free(self):
        if self._h:
            if self.ctx._h:   # a destroyed context already released the device (its pool went with it)
                self.ctx._L.dbg_table_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PackedDnaStringSet:
    """src/dna_string.rs:763-822: `sequence` words + `start` + `length`."""

    def __init__(self, words, start, length, n_bases):
        self.sequence, self.start, self.length, self.n_bases = words, start, length, n_bases

    def __len__(self):
        return len(self.start)

    def get(self, i):
        """Bases (0..3, uint8) of sequence i."""
        idx = np.arange(int(self.start[i]), int(self.start[i]) + int(self.length[i]), dtype=np.uint64)
        w = self.sequence[(idx >> np.uint64(5)).astype(np.int64)]
        return ((w >> (np.uint64(62) - np.uint64(2) * (idx & np.uint64(31)))) & np.uint64(3)).astype(np.uint8)


def format_gfa(k, g, target, flags):
    """Text of DebruijnGraph::write_gfa (src/graph.rs:538-614) from host node arrays and find_edges results."""
    words, start, length = g["words"], g["start"], g["length"]
    lines = ["H\tVN:Z:debruijn-rs"]
    for n in range(int(g["n_nodes"])):
        idx = np.arange(int(start[n]), int(start[n]) + int(length[n]), dtype=np.uint64)
        b = (words[(idx >> np.uint64(5)).astype(np.int64)] >> (np.uint64(62) - np.uint64(2) * (idx & np.uint64(31)))) & np.uint64(3)
        lines.append(f"S\t{n}\t" + "".join("ACGT"[int(x)] for x in b))
        for d, sign, keep in ((0, "-", lambda t: t >= n), (1, "+", lambda t: t > n)):   # l_edges then r_edges
            for i in range(4):
                t = int(target[n, d, i])
                if t != 0xffffffff and keep(t):
                    lines.append(f"L\t{n}\t{sign}\t{t}\t{'-' if flags[n, d, i] & 1 else '+'}\t{k - 1}M")
    return "\n".join(lines) + "\n"


class BaseGraph:
    """src/graph.rs:44-114.  Device-resident (dbg_graph) until to_host() is called."""

    def __init__(self, ctx, handle):
        self.ctx, self._h = ctx, handle
        self._host = None

    def __len__(self):  # BaseGraph::len, graph.rs:63-65
        return self.ctx._L.dbg_graph_len(self._h)

    def is_empty(self):
        return len(self) == 0

    @property
    def stranded(self):
        return bool(self.ctx._L.dbg_graph_stranded(self._h))

    def to_host(self):
        if self._host is None:
            L = self.ctx._L
            m, nb, nw = len(self), L.dbg_graph_n_bases(self._h), L.dbg_graph_n_words(self._h)
            g = dict(n_nodes=m, n_bases=nb, words=np.zeros(nw, np.uint64), start=np.zeros(m, np.uint64),
                     length=np.zeros(m, np.uint32), exts=np.zeros(m, np.uint8), data=np.zeros(m, np.uint16),
                     stranded=self.stranded)
            self.ctx.check(L.dbg_graph_copy_out(self._h, _ptr(g["words"]), _ptr(g["start"]), _ptr(g["length"]),
                                                _ptr(g["exts"]), _ptr(g["data"])))
            self._host = g
        return self._host

    @staticmethod
    def combine(graphs):
        """BaseGraph::combine (src/graph.rs:71-100): one graph holding the nodes of `graphs` in order."""
        graphs = list(graphs)
        if not graphs:
            raise ValueError("combine needs at least one graph")
        ctx = graphs[0].ctx
        arr = (C.c_void_p * len(graphs))(*[g._h for g in graphs])
        h = C.c_void_p()
        ctx.check(ctx._L.dbg_graph_combine(ctx._h, arr, len(graphs), C.byref(h)))
        return BaseGraph(ctx, h)

    def finish(self):
        """BaseGraph::finish (src/graph.rs:116-142).  The first/last-k-mer maps are rebuilt on the device by the calls that need them
        (edges, fix_exts, is_compressed, compress_graph), so the DebruijnGraph is this object."""
        return self

    def edges(self):
        """DebruijnGraph::find_edges for every (node, side) after BaseGraph::finish (src/graph.rs:116-142, 223-291):
        (target[M, 2, 4] uint32, 0xffffffff = none; flags[M, 2, 4] uint8: bit 0 incoming side, bit 1 rc)."""
        m = len(self)
        target, flags = np.zeros(8 * m, np.uint32), np.zeros(8 * m, np.uint8)
        self.ctx.check(self.ctx._L.dbg_graph_edges(self.ctx._h, self._h, _ptr(target), _ptr(flags)))
        return target.reshape(m, 2, 4), flags.reshape(m, 2, 4)

    def fix_exts(self, valid_nodes=None):
        """DebruijnGraph::fix_exts (src/graph.rs:337-343), in place on the device; valid_nodes: optional bool/uint8 array."""
        if valid_nodes is not None:
            valid_nodes = np.ascontiguousarray(np.asarray(valid_nodes).astype(np.uint8))
            if len(valid_nodes) != len(self):
                raise ValueError("valid_nodes needs one entry per node")
        self.ctx.check(self.ctx._L.dbg_graph_fix_exts(self.ctx._h, self._h, _ptr(valid_nodes)))
        self._host = None

    def is_compressed(self, spec=None):
        """DebruijnGraph::is_compressed (src/graph.rs:296-334) on the device: None if no two nodes could be collapsed, else the
        first (node, next_node) pair in the reference's iteration order.  spec: SimpleCompress (join_test always true, the
        default) or ScmapCompress (join_test = data equality)."""
        pr = C.c_int64(-1)
        self.ctx.check(self.ctx._L.dbg_graph_is_compressed(self.ctx._h, self._h, int(isinstance(spec, ScmapCompress)), C.byref(pr)))
        return None if pr.value < 0 else (int(pr.value >> 32), int(pr.value & 0xffffffff))

    def to_bincode(self):
        """bincode 1.x image of the crate's serde-derived BaseGraph<K, u16> (field order src/graph.rs:43-50): bytes."""
        L = self.ctx._L
        n = C.c_uint64()
        self.ctx.check(L.dbg_graph_serialize(self._h, None, 0, C.byref(n)))
        buf = np.zeros(n.value, np.uint8)
        self.ctx.check(L.dbg_graph_serialize(self._h, _ptr(buf), n.value, C.byref(n)))
        return buf.tobytes()

    @staticmethod
    def from_bincode(data, k, ctx=None):
        """Inverse of to_bincode (K is a type parameter in the crate, so it is an argument here)."""
        ctx = ctx or default_context()
        buf = np.frombuffer(bytes(data), np.uint8)
        h = C.c_void_p()
        ctx.check(ctx._L.dbg_graph_deserialize(ctx._h, k, _ptr(buf), len(buf), C.byref(h)))
        return BaseGraph(ctx, h)

    @staticmethod
    def host_image(g):
        """bincode 1.x image (serde field order of the crate's BaseGraph<K, u16>, src/graph.rs:43-50) of host arrays: dict with words /
        start / length / exts / data / n_bases / stranded.  Pure host code."""
        def vec(x, dt):
            x = np.ascontiguousarray(x, dt)
            return np.uint64(len(x)).tobytes() + x.tobytes()
        nw = (int(g["n_bases"]) + 31) // 32
        return (vec(np.asarray(g["words"])[:nw], "<u8") + np.uint64(int(g["n_bases"])).tobytes() + vec(g["start"], "<u8") +
                vec(g["length"], "<u4") + vec(g["exts"], "u1") + vec(g["data"], "<u2") + bytes([1 if g.get("stranded") else 0]))

    @staticmethod
    def from_host(g, k, ctx=None):
        """A device BaseGraph from host arrays, through the bincode image."""
        return BaseGraph.from_bincode(BaseGraph.host_image(g), k, ctx=ctx)

    def write_gfa(self, out):
        """DebruijnGraph::write_gfa (src/graph.rs:538-614): header, one S line per node, L lines for the left edges with
        target >= node and the right edges with target > node (edge direction '+' = enters the target through its left
        side), overlap (K-1)M.  `out`: path or text file object.  Edges come from dbg_graph_edges on the device."""
        g = self.to_host()
        target, flags = self.edges()
        text = format_gfa(self.ctx._L.dbg_graph_k(self._h), g, target, flags)
        if hasattr(out, "write"):
            out.write(text)
        else:
            with open(out, "w") as f:
                f.write(text)

    @property
    def sequences(self):
        g = self.to_host()
        return PackedDnaStringSet(g["words"], g["start"], g["length"], g["n_bases"])

    @property
    def exts(self):
        return [Exts(v) for v in self.to_host()["exts"]]

    @property
    def data(self):
        return self.to_host()["data"]

    def free(self):
        if self._h:
            if self.ctx._h:
                self.ctx._L.dbg_graph_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def remove_censored_exts(stranded, table):
    """filter::remove_censored_exts (src/filter.rs:280-306): in place on the device table."""
    table.ctx.check(table.ctx._L.dbg_remove_censored_exts(table.ctx._h, table._h, int(bool(stranded)), 0))


def remove_censored_exts_sharded(stranded, table):
    """filter::remove_censored_exts_sharded (src/filter.rs:238-276); all_kmers = the table's own (report_all_kmers)."""
    table.ctx.check(table.ctx._L.dbg_remove_censored_exts(table.ctx._h, table._h, int(bool(stranded)), 1))


def filter_kmers(seqs, summarizer, stranded, report_all_kmers, memory_size, k=31, ctx=None, labels=None):
    """filter::filter_kmers (src/filter.rs:139-148).

    seqs: a SeqSet, or (words, start, length[, seq_exts]) host arrays in PackedDnaStringSet layout.
    summarizer: CountFilter, or CountFilterSet with `labels` (one u8 < 64 per sequence: the D1 of the reference's tuples).
    Returns (KmerTable, all_kmers) — all_kmers is empty unless report_all_kmers."""
    if isinstance(summarizer, CountFilterSet):
        if labels is None or report_all_kmers:
            raise ValueError("CountFilterSet needs labels (and does not report all_kmers)")
        own = None
        if isinstance(seqs, SeqSet):
            ss, ctx = seqs, seqs.ctx
        else:
            ctx = ctx or default_context()
            ss = own = SeqSet.upload(ctx, *seqs)
        lab = np.ascontiguousarray(labels, np.uint8)
        if len(lab) != len(ss):
            raise ValueError("one label per sequence")
        try:
            h = C.c_void_p()
            ctx.check(ctx._L.dbg_filter_kmers_colorset(ctx._h, k, ss._h, _ptr(lab), summarizer.min_kmer_obs, int(bool(stranded)),
                                                       int(memory_size), C.byref(h)))
        finally:
            if own is not None:
                own.free()
        return KmerTable(ctx, h), np.zeros(0, np.uint64)
    if not isinstance(summarizer, CountFilter):
        raise TypeError("only CountFilter / CountFilterSet are on the accelerated path (SURVEY.md §8)")
    own = None
    if isinstance(seqs, SeqSet):
        ss, ctx = seqs, seqs.ctx
    else:
        ctx = ctx or default_context()
        ss = own = SeqSet.upload(ctx, *seqs)
    try:
        h = C.c_void_p()
        ctx.check(ctx._L.dbg_filter_kmers(ctx._h, k, ss._h, summarizer.min_kmer_obs, int(bool(stranded)),
                                          int(bool(report_all_kmers)), int(memory_size), C.byref(h)))
    finally:
        if own is not None:
            own.free()
    table = KmerTable(ctx, h)
    all_kmers = np.zeros(0, np.uint64)
    if report_all_kmers:
        t = table.to_host()
        all_kmers = t["all_lo"] if k <= 32 else np.stack([t["all_lo"], t["all_hi"]], axis=1)
    return table, all_kmers


def compress_kmers_with_hash(stranded, spec, index):
    """compression::compress_kmers_with_hash (src/compression.rs:588-594)."""
    if not isinstance(spec, (SimpleCompress, ScmapCompress)):
        raise TypeError("only SimpleCompress / ScmapCompress are on the accelerated path (SURVEY.md §8)")
    ctx = index.ctx
    h = C.c_void_p()
    ctx.check(ctx._L.dbg_compress_kmers_with_hash(ctx._h, index._h, int(bool(stranded)), spec.func, C.byref(h)))
    return BaseGraph(ctx, h)


def compress_graph(stranded, spec, old_graph, censor_nodes=None):
    """compression::compress_graph (src/compression.rs:338-349): merge the unbranched runs of a (partially compressed) graph,
    optionally leaving out `censor_nodes` (node ids).  Returns the new graph; old_graph is not modified."""
    if not isinstance(spec, (SimpleCompress, ScmapCompress)):
        raise TypeError("only SimpleCompress / ScmapCompress are on the accelerated path (SURVEY.md §8)")
    ctx = old_graph.ctx
    cn = None if censor_nodes is None else np.ascontiguousarray(np.asarray(list(censor_nodes), dtype=np.uint64))
    h = C.c_void_p()
    ctx.check(ctx._L.dbg_compress_graph(ctx._h, old_graph._h, int(bool(stranded)), spec.func, _ptr(cn) if cn is not None and len(cn) else None,
                                        0 if cn is None else len(cn), C.byref(h)))
    return BaseGraph(ctx, h)


def table_from_host(k, lo, hi, exts, counts, ctx=None):
    ctx = ctx or default_context()
    lo = np.ascontiguousarray(lo, np.uint64)
    hi = np.ascontiguousarray(hi, np.uint64) if k > 32 else None
    exts = np.ascontiguousarray(exts, np.uint8)
    counts = np.ascontiguousarray(counts, np.uint16)
    h = C.c_void_p()
    ctx.check(ctx._L.dbg_table_from_host(ctx._h, k, len(lo), _ptr(lo), _ptr(hi), _ptr(exts), _ptr(counts), C.byref(h)))
    return KmerTable(ctx, h)


def compress_kmers(stranded, spec, kmer_exts, k=31, ctx=None):
    """compression::compress_kmers (src/compression.rs:598-615): kmer_exts = (lo, hi, exts, data) arrays."""
    lo, hi, exts, data = kmer_exts
    t = table_from_host(k, lo, hi, exts, data, ctx)
    try:
        return compress_kmers_with_hash(stranded, spec, t)
    finally:
        t.free()


def msp_kmer_buckets(seqs, k, p, stranded=False):
    """msp::msp_sequence (src/msp.rs:279-324, identity permutation, rc = !stranded): MspIntervalP::bucket() of the
    interval every k-mer falls in, sequence-major, one u32 per k-mer start."""
    ctx = seqs.ctx
    _, _, length = seqs.copy_out()
    n = int(np.maximum(length.astype(np.int64) - k + 1, 0).sum())
    out = np.zeros(n, np.uint32)
    ctx.check(ctx._L.dbg_msp_kmer_buckets(ctx._h, k, p, seqs._h, int(bool(stranded)), _ptr(out), n))
    return out


def msp_sequence(seqs, k, p, permutation=None, rc=True):
    """msp::msp_sequence (src/msp.rs:279-324) for every sequence of a SeqSet: dict of arrays (seq, start, len, bucket, exts), one
    entry per MSP interval in scan order; the reference's Vmer of an interval is bases start .. start+len of sequence seq."""
    ctx = seqs.ctx
    perm = None if permutation is None else np.ascontiguousarray(permutation, np.uint32)
    if perm is not None and len(perm) != 4 ** p:
        raise ValueError("permutation needs 4^p entries")
    n = C.c_uint64()
    ctx.check(ctx._L.dbg_msp_sequence(ctx._h, k, p, seqs._h, int(bool(rc)), _ptr(perm), 0, C.byref(n), None, None, None, None, None))
    m = n.value
    out = dict(seq=np.zeros(m, np.uint32), start=np.zeros(m, np.uint32), len=np.zeros(m, np.uint32), bucket=np.zeros(m, np.uint32),
               exts=np.zeros(m, np.uint8))
    if m:
        ctx.check(ctx._L.dbg_msp_sequence(ctx._h, k, p, seqs._h, int(bool(rc)), _ptr(perm), m, C.byref(n), _ptr(out["seq"]), _ptr(out["start"]),
                                          _ptr(out["len"]), _ptr(out["bucket"]), _ptr(out["exts"])))
    return out


def reads_to_graph(seqs, summarizer, spec, stranded=False, k=31, keep_table=False):
    """Fused filter_kmers -> compress_kmers_with_hash with the table kept on the device."""
    ctx = seqs.ctx
    th, gh = C.c_void_p(), C.c_void_p()
    ctx.check(ctx._L.dbg_reads_to_graph(ctx._h, k, seqs._h, summarizer.min_kmer_obs, int(bool(stranded)), spec.func,
                                        C.byref(th) if keep_table else None, C.byref(gh)))
    g = BaseGraph(ctx, gh)
    return (KmerTable(ctx, th), g) if keep_table else g
