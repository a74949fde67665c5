"""CPU tests of the host-side mirror (rust_debruijn_b200/api.py) that need no device: FASTQ record splitting, the bincode image
builder, argument checks of the reference-facing classes."""
import numpy as np
import pytest

import rust_debruijn_b200 as D


def test_parse_fastq_records():
    fq = b"@r1 some description\nACGTN\n+\nIIIII\n@r2\nGG\r\n+r2\nII\n@\n\n+\n\n"
    names, seqs = D.SeqSet.parse_fastq(fq)
    assert names == [b"r1", b"r2", b""] and seqs == [b"ACGTN", b"GG", b""]
    assert D.SeqSet.parse_fastq(b"") == ([], [])
    for bad in (b"@r\nACGT\n+\n", b"r\nACGT\n+\nIIII\n", b"@r\nACGT\n-\nIIII\n"):
        with pytest.raises(ValueError):
            D.SeqSet.parse_fastq(bad)


def test_parse_fastq_from_path(tmp_path):
    p = tmp_path / "x.fastq"
    p.write_bytes(b"@a\nAC\n+\nII\n")
    assert D.SeqSet.parse_fastq(str(p)) == ([b"a"], [b"AC"])


def test_host_image_equals_the_oracle_bincode_layout(orc):
    """BaseGraph.host_image (what BaseGraph.from_host feeds dbg_graph_deserialize) is byte for byte the oracle's bincode image
    (tests/test_oracle.py::test_bincode_image_layout pins that one to the serde field order by hand)."""
    w, s, l = orc.synth_reads(300, 1, orc.ERR_THR_NOISY)
    for k, stranded in ((31, False), (63, True)):
        t = orc.filter_kmers(k, w, s, l, min_obs=2, stranded=stranded)
        g = orc.compress_kmers(k, t["lo"], t["hi"], t["exts"], t["counts"], stranded=stranded)
        assert D.BaseGraph.host_image(g) == orc.graph_to_bincode(g)
    empty = dict(n_nodes=0, n_bases=0, words=np.zeros(0, np.uint64), start=np.zeros(0, np.uint64), length=np.zeros(0, np.uint32),
                 exts=np.zeros(0, np.uint8), data=np.zeros(0, np.uint16), stranded=False)
    assert D.BaseGraph.host_image(empty) == orc.graph_to_bincode(empty)


def test_summarizer_and_spec_argument_checks():
    assert D.CountFilter(2).min_kmer_obs == 2
    assert D.CountFilter(1 << 40).min_kmer_obs == 65536       # counts saturate at 65535: censors everything, never wraps a u32
    with pytest.raises(ValueError):
        D.CountFilter(-1)
    with pytest.raises(ValueError):
        D.CountFilterSet(70000)
    assert D.SimpleCompress(D.MAX).func == D.MAX and D.ScmapCompress().func == 4
    with pytest.raises(TypeError):
        D.compress_graph(False, object(), None)
