"""ctypes binding of the CPU oracle (oracle/oracle.cpp) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The shipped path (rust_debruijn_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def _cpu_id():
    """Identity of the host CPU (model + ISA flags): the library is built -march=native, so a copy built on another
    machine (this repo travels to the GPU box as a snapshot) must be rebuilt before it is loaded."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            lines = [ln for ln in f if ln.startswith(("model name", "flags"))][:2]
        return hashlib.sha1("".join(lines).encode()).hexdigest()
    except OSError:
        return "unknown"


def build(force=False):
    src = os.path.join(_HERE, "oracle.cpp")
    stamp = os.path.join(_HERE, ".build_cpu")
    cpu = _cpu_id()
    try:
        with open(stamp) as f:
            built_for = f.read().strip()
    except OSError:
        built_for = None
    stale = not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src) or built_for != cpu
    if force or stale:
        try:
            subprocess.check_call(["make", "-C", _HERE, "-B", "-s"], stdout=subprocess.DEVNULL)
        except (subprocess.CalledProcessError, OSError):
            if os.path.exists(_SO) and built_for is None:
                return _SO   # no compiler here: keep the shipped library (built x86-64-v2 compatible by the fallback below)
            subprocess.check_call(["make", "-C", _HERE, "-B", "-s", "ARCH=-march=x86-64-v2 -mtune=generic"],
                                  stdout=subprocess.DEVNULL)
        with open(stamp, "w") as f:
            f.write(cpu)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u64p, u32p, u16p, u8p = (C.POINTER(t) for t in (C.c_uint64, C.c_uint32, C.c_uint16, C.c_uint8))
        L.orc_kmer_rc.restype = C.c_uint64
        L.orc_kmer_rc.argtypes = [C.c_int, C.c_uint64]
        L.orc_kmer_rc128.argtypes = [C.c_int, u64p, u64p]
        L.orc_kmer_extend_left.restype = C.c_uint64
        L.orc_kmer_extend_left.argtypes = [C.c_int, C.c_uint64, C.c_int]
        L.orc_kmer_extend_right.restype = C.c_uint64
        L.orc_kmer_extend_right.argtypes = [C.c_int, C.c_uint64, C.c_int]
        L.orc_pack_bases.argtypes = [u8p, C.c_uint64, u64p]
        L.orc_filter_kmers.restype = C.c_void_p
        L.orc_filter_kmers.argtypes = [C.c_int, u64p, u64p, u32p, u8p, C.c_uint64, C.c_uint32, C.c_int, C.c_int,
                                       C.c_uint64, C.c_int]
        for f in ("orc_table_len", "orc_table_all_len", "orc_table_n_input"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_table_passes.argtypes = [C.c_void_p]
        L.orc_table_copy.argtypes = [C.c_void_p, u64p, u64p, u8p, u16p, u64p, u64p]
        L.orc_table_free.argtypes = [C.c_void_p]
        L.orc_compress_kmers.restype = C.c_void_p
        L.orc_compress_kmers.argtypes = [C.c_int, C.c_uint64, u64p, u64p, u8p, u16p, C.c_int, C.c_int, u32p]
        L.orc_graph_error.argtypes = [C.c_void_p]
        L.orc_compress_graph.restype = C.c_void_p
        L.orc_compress_graph.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, u64p, u64p, u32p, u8p, u16p, u8p]
        for f in ("orc_graph_n_nodes", "orc_graph_n_bases"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_graph_copy.argtypes = [C.c_void_p, u64p, u64p, u32p, u8p, u16p]
        L.orc_graph_free.argtypes = [C.c_void_p]
        L.orc_msp_scan.restype = C.c_int64
        L.orc_msp_scan.argtypes = [C.c_int, C.c_int, u8p, C.c_uint32, u64p, C.c_int, u32p, u32p, u32p, u64p, u64p, u8p]
        L.orc_synth_genome_bases.restype = C.c_uint64
        L.orc_synth_genome_bases.argtypes = [C.c_uint64]
        L.orc_synth_reads.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, u64p]
        L.orc_graph_edges.restype = C.c_int64
        L.orc_graph_edges.argtypes = [C.c_int, C.c_int, C.c_uint64, u64p, u64p, u32p, u8p, u32p, u8p]
        L.orc_remove_censored_exts.argtypes = [C.c_int, C.c_uint64, u64p, u64p, u8p, C.c_int, C.c_uint64, u64p, u64p, C.c_int]
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


SAT_ADD, WRAP_ADD, ADD_MOD_65535, MAX, SCMAP = 0, 1, 2, 3, 4   # SCMAP: ScmapCompress (compression.rs:66-98)
ERR_THR_NOISY = 83886


def pack_bases(bases):
    """0..3 bases (uint8 array) -> DnaString words (src/dna_string.rs:383-399)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    words = np.zeros((len(bases) + 31) // 32, dtype=np.uint64)
    if len(bases):
        lib().orc_pack_bases(_p(bases, C.c_uint8), len(bases), _p(words, C.c_uint64))
    return words


def acgt_to_bits(ascii_bytes):
    """lib.rs:65-73 base_to_bits over a byte array: A/a=0 C/c=1 G/g=2 T/t=3, anything else 0.  Returns (bits, valid)."""
    lut = np.zeros(256, np.uint8)
    ok = np.zeros(256, bool)
    for ch, v in ((b"A", 0), (b"C", 1), (b"G", 2), (b"T", 3)):
        for c in (ch[0], ch.lower()[0]):
            lut[c], ok[c] = v, True
    a = np.frombuffer(bytes(ascii_bytes), np.uint8) if not isinstance(ascii_bytes, np.ndarray) else ascii_bytes
    return lut[a], ok[a]


def from_acgt_bytes(ascii_seqs):
    """DnaString::from_acgt_bytes (dna_string.rs:224-250; AVX2 twin bitops_avx2.rs:48-132: same mapping, invalid -> A)
    for every sequence, appended with PackedDnaStringSet::add (dna_string.rs:811-821).
    Returns (words, start, length, n_invalid)."""
    bits, n_bad = [], 0
    for sq in ascii_seqs:
        b, ok = acgt_to_bits(sq)
        bits.append(b)
        n_bad += int((~ok).sum())
    w, st, ln = seqset_from_lists(bits)
    return w, st, ln, n_bad


def siphash13(data, k0=0, k1=0):
    """SipHash-1-3 (the function behind Rust's DefaultHasher, zero key) in plain Python integers — test infrastructure."""
    M = (1 << 64) - 1
    v0, v1, v2, v3 = k0 ^ 0x736f6d6570736575, k1 ^ 0x646f72616e646f6d, k0 ^ 0x6c7967656e657261, k1 ^ 0x7465646279746573

    def rotl(x, b):
        return ((x << b) | (x >> (64 - b))) & M

    def rnd(v0, v1, v2, v3):
        v0 = (v0 + v1) & M; v1 = rotl(v1, 13); v1 ^= v0; v0 = rotl(v0, 32)
        v2 = (v2 + v3) & M; v3 = rotl(v3, 16); v3 ^= v2
        v0 = (v0 + v3) & M; v3 = rotl(v3, 21); v3 ^= v0
        v2 = (v2 + v1) & M; v1 = rotl(v1, 17); v1 ^= v2; v2 = rotl(v2, 32)
        return v0, v1, v2, v3

    data = bytes(data)
    n = len(data)
    for i in range(0, n - n % 8, 8):
        m = int.from_bytes(data[i:i + 8], "little")
        v3 ^= m
        v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
        v0 ^= m
    b = int.from_bytes(data[n - n % 8:], "little") | ((n & 0xff) << 56)
    v3 ^= b
    v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
    v0 ^= b
    v2 ^= 0xff
    for _ in range(3):
        v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
    return v0 ^ v1 ^ v2 ^ v3


def from_acgt_bytes_hashn(ascii_seqs, read_names):
    """DnaString::from_acgt_bytes_hashn (dna_string.rs:254-278) for every (sequence, read name): a non-ACGT position becomes
    DefaultHasher{read_name.hash(); pos.hash()}.finish() % 4 — SipHash-1-3, zero key, over [len(name) u64 LE][name][pos u64 LE]
    (Hash for [u8] writes the length prefix and the bytes, usize its 8 native-endian bytes).  Returns (words, start, length, n_invalid)."""
    bits, n_bad = [], 0
    for sq, name in zip(ascii_seqs, read_names):
        b, ok = acgt_to_bits(sq)
        b = b.copy()
        pre = len(name).to_bytes(8, "little") + bytes(name)
        for pos in np.nonzero(~ok)[0]:
            b[pos] = siphash13(pre + int(pos).to_bytes(8, "little")) % 4
        bits.append(b)
        n_bad += int((~ok).sum())
    w, st, ln = seqset_from_lists(bits)
    return w, st, ln, n_bad


def graph_to_bincode(g):
    """bincode 1.x image of BaseGraph<K, u16> (serde field order: graph.rs:43-50, dna_string.rs:72-76, 762-767): storage Vec<u64>,
    len usize, start Vec<usize>, length Vec<u32>, exts Vec<u8>, data Vec<u16>, stranded bool; u64 little-endian lengths."""
    def vec(a, dt):
        a = np.ascontiguousarray(a, dt)
        return int(len(a)).to_bytes(8, "little") + a.tobytes()
    return (vec(g["words"], "<u8") + int(g["n_bases"]).to_bytes(8, "little") + vec(g["start"], "<u8") + vec(g["length"], "<u4") +
            vec(g["exts"], "u1") + vec(g["data"], "<u2") + (b"\x01" if g.get("stranded") else b"\x00"))


def seqset_from_lists(seqs):
    """list of uint8 base arrays -> (words, start, length) in PackedDnaStringSet layout."""
    length = np.array([len(s) for s in seqs], dtype=np.uint32)
    start = np.zeros(len(seqs), dtype=np.uint64)
    if len(seqs):
        start[1:] = np.cumsum(length.astype(np.uint64))[:-1]
    cat = np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs]) if len(seqs) else np.zeros(0, np.uint8)
    return pack_bases(cat), start, length


def synth_reads(R, seed=1, err_thr=0):
    """synth-v1: R reads x 150 bases, packed contiguously."""
    words = np.zeros((150 * R + 31) // 32, dtype=np.uint64)
    lib().orc_synth_reads(R, seed, err_thr, _p(words, C.c_uint64))
    start = np.arange(R, dtype=np.uint64) * np.uint64(150)
    length = np.full(R, 150, dtype=np.uint32)
    return words, start, length


def filter_kmers(k, words, start, length, seq_exts=None, min_obs=1, stranded=False, report_all=False, memory_gb=4,
                 threads=1):
    """src/filter.rs:139-231 with CountFilter.  Returns dict(kmers_lo, kmers_hi, exts, counts, all_lo, all_hi, ...)."""
    L = lib()
    words = np.ascontiguousarray(words, np.uint64)
    start = np.ascontiguousarray(start, np.uint64)
    length = np.ascontiguousarray(length, np.uint32)
    if seq_exts is not None:
        seq_exts = np.ascontiguousarray(seq_exts, np.uint8)
    h = L.orc_filter_kmers(k, _p(words, C.c_uint64), _p(start, C.c_uint64), _p(length, C.c_uint32),
                           _p(seq_exts, C.c_uint8), len(start), min_obs, int(stranded), int(report_all), memory_gb,
                           threads)
    n, na = L.orc_table_len(h), L.orc_table_all_len(h)
    two = k > 32
    out = dict(k=k, lo=np.zeros(n, np.uint64), hi=np.zeros(n if two else 0, np.uint64), exts=np.zeros(n, np.uint8),
               counts=np.zeros(n, np.uint16), all_lo=np.zeros(na, np.uint64),
               all_hi=np.zeros(na if two else 0, np.uint64), n_input=L.orc_table_n_input(h),
               passes=L.orc_table_passes(h))
    L.orc_table_copy(h, _p(out["lo"], C.c_uint64), _p(out["hi"], C.c_uint64), _p(out["exts"], C.c_uint8),
                     _p(out["counts"], C.c_uint16), _p(out["all_lo"], C.c_uint64), _p(out["all_hi"], C.c_uint64))
    L.orc_table_free(h)
    return out


def graph_edges(k, g, stranded=False):
    """BaseGraph::finish + DebruijnGraph::find_edges for every (node, side) (src/graph.rs:116-142, 223-291) and
    is_compressed (:296-334, join_test = true).  Returns (target[M,2,4] uint32 with 0xffffffff = none,
    flags[M,2,4] uint8: bit0 incoming side, bit1 rc, collapsible pair or None)."""
    m = int(g["n_nodes"])
    words = np.ascontiguousarray(np.concatenate([g["words"], np.zeros(1, np.uint64)]), np.uint64)
    start = np.ascontiguousarray(g["start"], np.uint64)
    length = np.ascontiguousarray(g["length"], np.uint32)
    exts = np.ascontiguousarray(g["exts"], np.uint8)
    target = np.zeros(m * 8, np.uint32)
    flags = np.zeros(m * 8, np.uint8)
    r = lib().orc_graph_edges(k, int(stranded), m, _p(words, C.c_uint64), _p(start, C.c_uint64), _p(length, C.c_uint32),
                              _p(exts, C.c_uint8), _p(target, C.c_uint32), _p(flags, C.c_uint8))
    pair = None if r < 0 else (int(r >> 32), int(r & 0xffffffff))
    return target.reshape(m, 2, 4), flags.reshape(m, 2, 4), pair


def write_gfa(k, g, stranded=False):
    """DebruijnGraph::write_gfa / node_to_gfa (src/graph.rs:538-614) as text: "H\\tVN:Z:debruijn-rs", then per node its S line,
    the L lines of l_edges with target >= node_id ("-" out of the node), the L lines of r_edges with target > node_id ("+");
    the target's orientation is "+" when it is entered through its Left side (Dir::Left => "+"), overlap (K-1)M."""
    target, flags, _ = graph_edges(k, g, stranded=stranded)
    out = ["H\tVN:Z:debruijn-rs\n"]
    for n in range(int(g["n_nodes"])):
        seq = "".join("ACGT"[int(b)] for b in unpack_bases(g["words"], int(g["start"][n]), int(g["length"][n])))
        out.append("S\t%d\t%s\n" % (n, seq))
        for i in range(4):
            t = int(target[n, 0, i])
            if t != 0xffffffff and t >= n:
                out.append("L\t%d\t-\t%d\t%s\t%dM\n" % (n, t, "-" if flags[n, 0, i] & 1 else "+", k - 1))
        for i in range(4):
            t = int(target[n, 1, i])
            if t != 0xffffffff and t > n:
                out.append("L\t%d\t+\t%d\t%s\t%dM\n" % (n, t, "-" if flags[n, 1, i] & 1 else "+", k - 1))
    return "".join(out)


def graph_fix_exts(k, g, stranded=False, valid_nodes=None):
    """DebruijnGraph::fix_exts / get_valid_exts (src/graph.rs:337-377): an extension survives iff find_link resolves it
    (and the target is in valid_nodes when given).  Returns the new node Exts array."""
    target, _, _ = graph_edges(k, g, stranded=stranded)
    ok = target != 0xffffffff
    if valid_nodes is not None:
        vn = np.asarray(valid_nodes).astype(bool)
        ok &= vn[np.where(ok, target, 0)]
    bits = (1 << np.arange(8, dtype=np.uint32)).reshape(2, 4)      # Exts bit 4 * dir + base (lib.rs:609-618)
    return (ok * bits).sum(axis=(1, 2)).astype(np.uint8)


def remove_censored_exts(k, t, stranded=False, sharded=False):
    """filter::remove_censored_exts (src/filter.rs:280-306) or, with sharded=True, remove_censored_exts_sharded
    (:238-276, needs the table's all_kmers).  Returns the new exts array (the table dict is not modified)."""
    lo = np.ascontiguousarray(t["lo"], np.uint64)
    hi = np.ascontiguousarray(t["hi"], np.uint64)
    exts = np.array(t["exts"], np.uint8, copy=True)
    alo = np.ascontiguousarray(t["all_lo"], np.uint64) if sharded else np.zeros(0, np.uint64)
    ahi = np.ascontiguousarray(t["all_hi"], np.uint64) if sharded else np.zeros(0, np.uint64)
    lib().orc_remove_censored_exts(k, len(lo), _p(lo, C.c_uint64), _p(hi, C.c_uint64), _p(exts, C.c_uint8), int(stranded),
                                   len(alo), _p(alo, C.c_uint64), _p(ahi, C.c_uint64), int(sharded))
    return exts


def compress_kmers(k, lo, hi, exts, counts, stranded=False, reduce_op=SAT_ADD, seed_order=None):
    """src/compression.rs:545-583 (CompressFromHash + SimpleCompress).  Returns BaseGraph arrays."""
    L = lib()
    lo = np.ascontiguousarray(lo, np.uint64)
    hi = np.ascontiguousarray(hi, np.uint64) if k > 32 else None
    exts = np.ascontiguousarray(exts, np.uint8)
    counts = np.ascontiguousarray(counts, np.uint16)
    if seed_order is not None:
        seed_order = np.ascontiguousarray(seed_order, np.uint32)
    h = L.orc_compress_kmers(k, len(lo), _p(lo, C.c_uint64), _p(hi, C.c_uint64), _p(exts, C.c_uint8),
                             _p(counts, C.c_uint16), int(stranded), reduce_op, _p(seed_order, C.c_uint32))
    err = L.orc_graph_error(h)
    m, nb = L.orc_graph_n_nodes(h), L.orc_graph_n_bases(h)
    g = dict(error=err, n_nodes=m, n_bases=nb, words=np.zeros((nb + 31) // 32, np.uint64), start=np.zeros(m, np.uint64),
             length=np.zeros(m, np.uint32), exts=np.zeros(m, np.uint8), data=np.zeros(m, np.uint16), stranded=stranded)
    L.orc_graph_copy(h, _p(g["words"], C.c_uint64), _p(g["start"], C.c_uint64), _p(g["length"], C.c_uint32),
                     _p(g["exts"], C.c_uint8), _p(g["data"], C.c_uint16))
    L.orc_graph_free(h)
    return g


def _graph_out(h, stranded):
    L = lib()
    err = L.orc_graph_error(h)
    m, nb = L.orc_graph_n_nodes(h), L.orc_graph_n_bases(h)
    g = dict(error=err, n_nodes=m, n_bases=nb, words=np.zeros((nb + 31) // 32, np.uint64), start=np.zeros(m, np.uint64),
             length=np.zeros(m, np.uint32), exts=np.zeros(m, np.uint8), data=np.zeros(m, np.uint16), stranded=stranded)
    L.orc_graph_copy(h, _p(g["words"], C.c_uint64), _p(g["start"], C.c_uint64), _p(g["length"], C.c_uint32),
                     _p(g["exts"], C.c_uint8), _p(g["data"], C.c_uint16))
    L.orc_graph_free(h)
    return g


def combine_graphs(graphs):
    """BaseGraph::combine (src/graph.rs:71-100): the nodes of every graph appended in order (PackedDnaStringSet::add re-packs the
    bases contiguously); mixing stranded and unstranded graphs panics there -> ValueError here."""
    st = [bool(g["stranded"]) for g in graphs]
    if any(st) and not all(st):
        raise ValueError("attempted to combine stranded and unstranded graphs")
    bases, start, length, pos = [], [], [], 0
    for g in graphs:
        for i in range(int(g["n_nodes"])):
            n = int(g["length"][i])
            bases.append(unpack_bases(g["words"], int(g["start"][i]), n))
            start.append(pos)
            length.append(n)
            pos += n
    allb = np.concatenate(bases) if bases else np.zeros(0, np.uint8)
    return dict(error=0, n_nodes=len(start), n_bases=pos, words=pack_bases(allb), start=np.array(start, np.uint64),
                length=np.array(length, np.uint32), exts=np.concatenate([np.asarray(g["exts"], np.uint8) for g in graphs]) if graphs else np.zeros(0, np.uint8),
                data=np.concatenate([np.asarray(g["data"], np.uint16) for g in graphs]) if graphs else np.zeros(0, np.uint16),
                stranded=all(st) if st else False)


def compress_graph(k, g, stranded=False, reduce_op=SAT_ADD, censor_nodes=None):
    """compression::compress_graph (src/compression.rs:291-349): fix_exts(Some(available)) -> greedy node walk in node order ->
    finish + fix_exts(None).  censor_nodes: iterable of node ids.  Returns BaseGraph arrays (error: 1 "unreachable", 3 "No kmer")."""
    L = lib()
    m = int(g["n_nodes"])
    words = np.ascontiguousarray(np.concatenate([g["words"], np.zeros(1, np.uint64)]), np.uint64)
    start = np.ascontiguousarray(g["start"], np.uint64)
    length = np.ascontiguousarray(g["length"], np.uint32)
    exts = np.ascontiguousarray(g["exts"], np.uint8)
    data = np.ascontiguousarray(g["data"], np.uint16)
    censor = None
    if censor_nodes is not None:
        censor = np.zeros(m, np.uint8)
        censor[np.asarray(list(censor_nodes), np.int64)] = 1
    h = L.orc_compress_graph(k, int(stranded), int(bool(g.get("stranded", stranded))), reduce_op, m, _p(words, C.c_uint64), _p(start, C.c_uint64), _p(length, C.c_uint32),
                             _p(exts, C.c_uint8), _p(data, C.c_uint16), _p(censor, C.c_uint8))
    return _graph_out(h, stranded)


def msp_scan(k, p, seq, perm=None, rc=True):
    """src/msp.rs:207-324.  seq: uint8 0..3.  Returns dict of interval arrays."""
    L = lib()
    seq = np.ascontiguousarray(seq, np.uint8)
    m = len(seq)
    cap = max(m - k + 1, 1)
    o = dict(start=np.zeros(cap, np.uint32), len=np.zeros(cap, np.uint32), minpos=np.zeros(cap, np.uint32),
             minimizer=np.zeros(cap, np.uint64), bucket=np.zeros(cap, np.uint64), exts=np.zeros(cap, np.uint8))
    if perm is not None:
        perm = np.ascontiguousarray(perm, np.uint64)
    n = L.orc_msp_scan(k, p, _p(seq, C.c_uint8), m, _p(perm, C.c_uint64), int(rc), _p(o["start"], C.c_uint32),
                       _p(o["len"], C.c_uint32), _p(o["minpos"], C.c_uint32), _p(o["minimizer"], C.c_uint64),
                       _p(o["bucket"], C.c_uint64), _p(o["exts"], C.c_uint8))
    return {kk: v[:n] for kk, v in o.items()}


# ---- checksums used by the golden anchors ---------------------------------
def xor_valid(t):
    lo = int(np.bitwise_xor.reduce(t["lo"])) if len(t["lo"]) else 0
    if t["k"] > 32:
        hi = int(np.bitwise_xor.reduce(t["hi"])) if len(t["hi"]) else 0
        return (hi << 64) | lo
    return lo


def mix_valid(t):
    e = t["exts"].astype(np.uint64)
    s = (t["lo"] * (e + np.uint64(1)) + t["counts"].astype(np.uint64))
    if t["k"] > 32:
        s = s + t["hi"] * (e + np.uint64(3))
    return int(np.add.reduce(s, dtype=np.uint64)) if len(s) else 0


def unpack_bases(words, start, n):
    idx = np.arange(start, start + n, dtype=np.uint64)
    return ((words[(idx >> np.uint64(5)).astype(np.int64)] >> (np.uint64(62) - np.uint64(2) * (idx & np.uint64(31)))) & np.uint64(3)).astype(np.uint8)


def fnv_nodes(g):
    """FNV-1a-64 over, per node in order: one byte per base, then exts, data lo, data hi."""
    parts = []
    for i in range(g["n_nodes"]):
        parts.append(unpack_bases(g["words"], int(g["start"][i]), int(g["length"][i])))
        d = int(g["data"][i])
        parts.append(np.array([g["exts"][i], d & 0xff, d >> 8], dtype=np.uint8))
    buf = np.concatenate(parts).tobytes() if parts else b""
    h = 0xcbf29ce484222325
    for b in buf:
        h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h
