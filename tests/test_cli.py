"""The CLI `host/kmercamel` (C++ host over the C ABI) against the reference CLI's observable behaviour: header lines
(src/parser.h:167-179, src/masks.h:27-37), the two-line .msfa output, flag validation (src/main.cpp:277-308) and the
outputs recorded from the unmodified reference (tests/golden/*.json).  The text conversions and the flag errors need no
GPU; everything else is marked gpu."""
import gzip
import json
import os
import subprocess

import pytest

from conftest import GOLDEN_DIR, ROOT, md5

EXE = os.path.join(ROOT, "host", "kmercamel")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", ROOT, "host/kmercamel"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.fixture(scope="module")
def gm():
    return json.load(open(os.path.join(GOLDEN_DIR, "golden_masks.json")))


def run(args, stdin=None):
    return subprocess.run([EXE, *args], capture_output=True, input=stdin)


def unzip(tmp_path, name):
    path = tmp_path / name.replace(".gz", "")
    path.write_bytes(gzip.open(os.path.join(GOLDEN_DIR, name)).read())
    return str(path)


# ---- no GPU needed ---------------------------------------------------------------------------------------------------
def test_cli_conversions_vs_reference(gm, tmp_path):
    for i, case in enumerate(gm["conversions"][:12]):
        ms_path = tmp_path / "ms.fa"
        ms_path.write_text(">ms\n%s\n" % case["ms"])
        m, s = str(tmp_path / "m.txt"), str(tmp_path / "s.txt")
        assert run(["ms2mssep", "-m", m, "-s", s, str(ms_path)]).returncode == 0
        assert open(m).read() == case["mask"] and open(s).read() == case["superstring"]
        assert run(["mssep2ms", "-m", m, "-s", s]).stdout.decode() == case["joined"]
        assert run(["ms2spss", "-k", str(case["k"]), str(ms_path)]).stdout.decode() == case["spss"]
        rec_path = tmp_path / "rec.fa"
        rec_path.write_text("".join(">x%d\n%s\n" % (j, r) for j, r in enumerate(case["records"])))
        p = run(["spss2ms", "-k", str(case["k"]), str(rec_path)])
        assert p.stdout.decode() == ">superstring %s\n%s\n" % (rec_path, case["spss2ms"])


def test_cli_flag_errors(tmp_path):
    fa = tmp_path / "x.fa"
    fa.write_text(">r\nACGTACGT\n")
    for args, msg in [(["compute", str(fa)], "Required parameter k not set."),
                      (["compute", "-k", "-3", str(fa)], "k must be positive."),
                      (["compute", "-k", "128", str(fa)], "k > 127 not supported"),
                      (["compute", "-k", "5", "-z", "300", str(fa)], "Minimum frequency '-z' must be between 1 and 255."),
                      (["compute", "-k", "5", "-z", "2", "-S", str(fa)], "not compatible with frequency filterring"),
                      (["compute", "-k", "5", "-d", "2", str(fa)], "Unsupported argument d"),
                      (["compute", "-k", "5", "-a", "streaming", "-S", str(fa)], "Assuming simplitigs is only supported"),
                      (["compute", "-k", "5", "-a", "streaming", "-M", "m.fa", str(fa)], "maximum number of ones is only supported"),
                      (["compute", "-k", "5", "-a", "local-greedy", str(fa)], "not part of the GPU compute path"),
                      (["maskopt", str(fa)], "Required parameter k not set."),
                      (["maskopt", "-k", "3", "-t", "min-run", str(fa)], "not recognized"),
                      (["ms2spss", str(fa)], "Required parameter k not set."),
                      (["ms2mssep", str(fa)], "Cannot have both superstring and mask redirected to stdout."),
                      (["frobnicate"], "Unknown sub-command")]:
        p = run(args)
        assert p.returncode == 1 and msg in p.stderr.decode(), (args, p.stderr.decode()[-300:])
    assert run(["-h"]).returncode == 0


def test_cli_fails_loudly_without_a_gpu(tmp_path):
    """No CPU fallback: on a machine without a CUDA device `compute` must refuse to run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    fa = tmp_path / "x.fa"
    fa.write_text(">r\nACGTACGT\n")
    p = run(["compute", "-k", "5", str(fa)])
    assert p.returncode == 1 and "cannot initialise CUDA device" in p.stderr.decode()


# ---- GPU -------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cli_compute_spneumoniae(golden, tmp_path):
    sp = unzip(tmp_path, "spneumoniae.fa.gz")
    want = golden["spneumoniae_compute"]["k31"]
    mo = str(tmp_path / "mo.fa")
    p = run(["compute", "-k", "31", "-M", mo, sp])
    assert p.returncode == 0, p.stderr.decode()
    head, line, rest = p.stdout.split(b"\n", 2)
    assert head.decode() == ">maskedsuperstring dataset='%s' k=31 alg=greedy mask=min-one mode=bidirectional" % sp and rest == b""
    assert sum(1 for c in line if c <= 90) == want["ones"] and abs(len(line) - want["length"]) <= want["length"] // 1000
    mh, ml, mr = open(mo, "rb").read().split(b"\n", 2)
    assert mh.decode() == ">maskedsuperstring dataset='%s' k=31 alg=greedy mask=max-one mode=bidirectional" % sp and mr == b""
    assert ml.upper() == line.upper() and abs(sum(1 for c in ml if c <= 90) - want["maxone_ones"]) <= want["maxone_ones"] // 1000
    assert "Finished collecting k-mers: %d 31-mers." % want["n_kmers"] in p.stderr.decode()


@pytest.mark.gpu
def test_cli_compute_S_and_lowerbound(golden, tmp_path):
    sim = unzip(tmp_path, "simplitigs-k31.fa.gz")
    want = golden["simplitigs_S"]["k31u"]
    out = str(tmp_path / "o.fa")
    p = run(["compute", "-k", "31", "-S", "-u", "-o", out, sim])
    assert p.returncode == 0 and p.stdout == b""
    head, line, rest = open(out, "rb").read().split(b"\n", 2)
    assert head.decode().endswith("k=31 alg=greedy mask=min-one mode=unidirectional") and rest == b""
    assert md5(line + b"\n") == want["md5"]
    p = run(["lowerbound", "-k", "31", "-S", sim])
    assert p.returncode == 0 and int(p.stdout.split()[0]) == golden["lowerbound"]["simplitigs_S"]["k31"]
    sp = unzip(tmp_path, "spneumoniae.fa.gz")
    p = run(["lowerbound", "-k", "31", "-z", "2", sp])
    assert p.returncode == 0 and int(p.stdout.split()[0]) == golden["lowerbound"]["spneumoniae"]["k31z2"]


@pytest.mark.gpu
def test_cli_streaming_and_maskopt(gm, tmp_path):
    sp = unzip(tmp_path, "spneumoniae.fa.gz")
    for flags, name in [([], "sp_k31"), (["-z", "2"], "sp_k31z2")]:
        p = run(["compute", "-a", "streaming", "-k", "31", *flags, sp])
        assert p.returncode == 0, p.stderr.decode()
        head, line, rest = p.stdout.split(b"\n", 2)
        row = gm["streaming"][name]
        assert head.decode() == ">maskedsuperstring dataset='%s' %s" % (sp, row["header_tail"]) and rest == b""
        assert (len(line), md5(line)) == (row["length"], row["md5"])
    # maskopt on the streaming superstring written above (k = 31, z = 1): header reprinted as src/masks.h:27-37 does
    stream = tmp_path / "stream.fa"
    stream.write_bytes(run(["compute", "-a", "streaming", "-k", "31", sp]).stdout)
    for t in ("max-one", "min-one"):
        p = run(["maskopt", "-t", t, "-k", "31", str(stream)])
        assert p.returncode == 0, p.stderr.decode()
        head, line, rest = p.stdout.split(b"\n", 2)
        assert head.decode() == ">maskedsuperstring reoptimized=%s dataset='%s' k=31 alg=streaming mask=min-one mode=bidirectional" % (t, sp)
        want = gm["maskopt"]["sp_stream_k31_" + t]
        assert rest == b"" and (sum(1 for c in line if c <= 90), md5(line)) == (want["ones"], want["md5"])
    empty = tmp_path / "empty.fa"
    empty.write_text(">r\nACG\n")
    p = run(["compute", "-k", "5", str(empty)])
    assert p.returncode == 1 and "contains no k-mers" in p.stderr.decode()  # src/main.cpp:155-158


@pytest.mark.gpu
def test_cli_several_gpus_and_verify(tmp_path):
    """`compute -g a,b,...` (one process, one host thread per rank; ranks may share a GPU): byte-identical to `-g 0`, the
    reference's log lines, and -V (the verify.py check as a device-side digest) passes."""
    import torch
    from kmercamel_b200 import synth
    n = torch.cuda.device_count()
    recs = synth.human_like_genome(3_000_000, 77)
    fa = tmp_path / "g.fa"
    fa.write_bytes(synth.fasta_bytes(recs))
    one = run(["compute", "-k", "31", "-V", str(fa)])
    assert one.returncode == 0, one.stderr.decode()
    assert "Verification passed" in one.stderr.decode() and "simplitigs)." in one.stderr.decode()
    for ranks in (2, 4):
        devs = ",".join(str(r % n) for r in range(ranks))
        p = run(["compute", "-k", "31", "-V", "-g", devs, str(fa)])
        assert p.returncode == 0, p.stderr.decode()
        assert p.stdout == one.stdout
        err = p.stderr.decode()
        assert "Verification passed" in err and "sharded by hash range over %d GPUs" % ranks in err
    p = run(["compute", "-k", "31", "-g", "0-1x", str(fa)])
    assert p.returncode == 1 and "-g takes CUDA device ordinals" in p.stderr.decode()
    # -S on several devices runs on the first one (with a note) and gives the single-GPU answer
    sim = tmp_path / "s.fa"
    sim.write_bytes(synth.fasta_bytes([r[:5000] for r in recs[:6]]))
    a = run(["compute", "-k", "31", "-S", str(sim)])
    b = run(["compute", "-k", "31", "-S", "-g", "0,0", str(sim)])
    assert a.returncode == 0 and b.returncode == 0 and a.stdout == b.stdout and "run on one GPU" in b.stderr.decode()
