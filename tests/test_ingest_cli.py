"""The input stage (csrc/ingest.cpp: line reader + FASTQ/FASTA framing, reference src/pipeline/sketch.go:40-161)
and the `hulk` front end (csrc/cli/hulk_main.cpp: flags and log lines of cmd/root.go, cmd/sketch.go).

CPU tests cover the reader against the Python restatement of the reference's framing (hulk_b200.seqio) and the
front end's argument handling; the GPU tests run `hulk sketch` end to end on the reference's own fixture."""
import gzip
import json
import os
import re
import subprocess

import numpy as np
import pytest

import hulk_b200
from conftest import GOLDEN, has_gpu, random_reads

FIXTURE = os.path.join(GOLDEN, "c1_reads.fq.gz")
TS = r"^\d{4}/\d{2}/\d{2} \d{2}:\d{2}:\d{2} "


def _native(paths, fasta=False, batch_bytes=0):
    with hulk_b200.NativeReader([str(p) for p in paths], fasta=fasta, batch_bytes=batch_bytes) as rd:
        return rd.reads()


def test_native_reader_on_reference_fixture(fixture_reads):
    assert _native([FIXTURE]) == fixture_reads
    # tiny batches: every batch boundary falls between two records
    assert _native([FIXTURE], batch_bytes=4096) == fixture_reads


def test_native_reader_framing_quirks(tmp_path):
    # empty lines are nil in the reference and re-fill the same slot; CRLF; last line without newline
    p = tmp_path / "q.fq"
    p.write_bytes(b"@r1\n\nACGT\n+\nIIII\n@r2\r\nGGCC\r\n+\r\nIIII")
    assert _native([p]) == hulk_b200.read_fastq(str(p)) == [b"ACGT", b"GGCC"]
    # an unfinished record at the end is never emitted; '@' is the only check and only on line 1
    q = tmp_path / "t.fq"
    q.write_bytes(b"@r1\nAC\nwhatever\n!!\n@r2\nGG\n+\n")
    assert _native([q]) == [b"AC"]
    bad = tmp_path / "bad.fq"
    bad.write_bytes(b"@ok\nAAAA\n+\nIIII\nr1\nACGT\n+\nIIII\n")
    with pytest.raises(ValueError, match="read ID in fastq file does not begin with @: r1"):
        _native([bad])
    # the line stream runs across files: a record may straddle two files, and a final line
    # without newline does not join the next file's first line
    a, b = tmp_path / "a.fq", tmp_path / "b.fq"
    a.write_bytes(b"@r1\nACGT\n+\nIIII\n@r2\nTTTT")
    b.write_bytes(b"+\nIIII\n@r3\nCC\n+\nII\n")
    assert _native([a, b]) == [b"ACGT", b"TTTT", b"CC"]


def test_native_reader_gzip_multimember_and_long_lines(tmp_path):
    reads = random_reads(5000, 150, seed=3, ragged=100)
    rec = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)) for i, r in enumerate(reads))
    half = len(rec) // 2
    cut = rec.index(b"\n@", half) + 1
    gz = tmp_path / "m.fastq.gz"
    gz.write_bytes(gzip.compress(rec[:cut]) + gzip.compress(rec[cut:]))      # two gzip members (Go: multistream)
    assert _native([gz], batch_bytes=1 << 16) == reads
    notgz = tmp_path / "n.fq.gz"
    notgz.write_bytes(rec[:1000])
    with pytest.raises(ValueError, match="gzip: invalid header"):
        _native([notgz])
    # bufio.Scanner: a line of 64 KiB or more is fatal, 64 KiB - 1 is fine
    ok = tmp_path / "ok.fq"
    ok.write_bytes(b"@r\n" + b"A" * 65535 + b"\n+\n" + b"I" * 65535 + b"\n")
    assert _native([ok], batch_bytes=4096) == [b"A" * 65535]                # also: one read larger than a batch
    long = tmp_path / "long.fq"
    long.write_bytes(b"@r\n" + b"A" * 65536 + b"\n+\n" + b"I" * 10 + b"\n")
    with pytest.raises(ValueError, match="token too long"):
        _native([long])
    with pytest.raises(ValueError, match="token too long"):
        hulk_b200.read_fastq(str(long))


def test_native_reader_gzip_errors_follow_go(tmp_path):
    """compress/gzip.Reader (multistream): garbage behind a member is gzip.ErrHeader -- zlib's gzread would stop
    silently --, a short header or a member cut short is io.ErrUnexpectedEOF, an empty file is io.EOF from
    gzip.NewReader, a CRC / length mismatch is gzip.ErrChecksum; log.Fatal(err) prints exactly these."""
    reads = random_reads(300, 120, seed=9)
    rec = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)) for i, r in enumerate(reads))
    good = gzip.compress(rec, 1)

    def run(name, data, **env):
        f = tmp_path / name
        f.write_bytes(data)
        return _native_env([f], env)

    assert run("ok.fq.gz", good).startswith("OK 300 ")
    assert run("two.fq.gz", good + gzip.compress(b"", 1)).startswith("OK 300 ")          # an empty member is fine
    for name, data, msg in (("garbage.fq.gz", good + b"trailing garbage, more than ten bytes", "gzip: invalid header"),
                            ("zeros.fq.gz", good + b"\0" * 512, "gzip: invalid header"),     # Go does not skip padding
                            ("short.fq.gz", good + b"\x1f\x8b\x08", "unexpected EOF"),
                            ("cut.fq.gz", good[:len(good) // 2], "unexpected EOF"),
                            ("cut8.fq.gz", good[:-3], "unexpected EOF"),
                            ("empty.fq.gz", b"", "EOF"),
                            ("one.fq.gz", b"\x1f", "unexpected EOF"),
                            ("crc.fq.gz", good[:-8] + bytes([good[-8] ^ 1]) + good[-7:], "gzip: invalid checksum"),
                            ("len.fq.gz", good[:-4] + bytes([good[-4] ^ 1]) + good[-3:], "gzip: invalid checksum"),
                            ("method.fq.gz", good[:2] + b"\x07" + good[3:], "gzip: invalid header")):
        got = run(name, data)
        assert got.startswith("ERR") and got.endswith(" " + msg), (name, got)
    # the reads in front of the bad spot were already handed on, like through Go's streaming reader
    assert run("garbage2.fq.gz", good + b"x" * 64).split()[1] == "300"
    # BGZF: a file that is only the EOF marker block holds no reads and is not an error
    assert run("eof.fq.gz", _bgzf(b"")) == run("eof2.fq.gz", gzip.compress(b"", 1))
    assert run("eof.fq.gz", _bgzf(b"")).startswith("OK 0 ")


def test_native_reader_fasta(tmp_path):
    fa = tmp_path / "x.fa"
    fa.write_bytes(b"stray\n>c1 desc\nACGT\nTTAA\n>c2\n>c3\nGG\n\n>c4\nAAAA\n")
    want = hulk_b200.read_fasta(str(fa))
    assert want == [b"ACGTTTAA", b"", b"GG"]                                # stops at the empty line
    assert _native([fa], fasta=True) == want
    big = tmp_path / "g.fna"
    seq = random_reads(1, 300000, seed=9)[0]
    big.write_bytes(b">chr\n" + b"\n".join(seq[i:i + 70] for i in range(0, len(seq), 70)) + b"\n")
    assert _native([big], fasta=True, batch_bytes=8192) == [seq]
    # chromosome-sized records (what the sliced scan of k1_long.cuh is for) arrive as one sequence each, plain and gzipped
    rng = np.random.default_rng(10)
    chrom = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, 6_000_000)].tobytes()
    body = b">chr1 six megabases\n" + b"\n".join(chrom[i:i + 60] for i in range(0, len(chrom), 60)) + b"\n>plasmid\nACGTACGT\nAC\n"
    for name, data in (("chrom.fa", body), ("chrom.fa.gz", gzip.compress(body, 1))):
        path = tmp_path / name
        path.write_bytes(data)
        got = _native([path], fasta=True)
        assert len(got) == 2 and got[0] == chrom and got[1] == b"ACGTACGTAC"
    none = tmp_path / "none.fa"
    none.write_bytes(b"ACGT\n")
    with pytest.raises(ValueError):
        _native([none], fasta=True)


def _hulk(*args, stdin=None, env=None):
    assert os.path.exists(hulk_b200.CLI_PATH), "build the front end first (python -m hulk_b200.build)"
    return subprocess.run([hulk_b200.CLI_PATH, *args], capture_output=True, text=True, timeout=600, stdin=stdin,
                          env=(dict(os.environ, **env) if env else None))


def test_cli_version_and_flag_errors(tmp_path):
    r = _hulk("version")
    assert r.returncode == 0 and r.stdout == "1.0.0\n"
    r = _hulk("sketch", "--nope")
    assert r.returncode == 1 and "unknown flag: --nope" in r.stdout
    r = _hulk("sketch", "-k", "x")
    assert r.returncode == 1 and 'invalid argument "x" for "-k, --kmerSize" flag' in r.stdout
    # helpers.CheckFile / CheckExt messages through helpers.ErrorCheck
    r = _hulk("sketch", "-f", str(tmp_path / "missing.fq"), "-o", str(tmp_path / "o"))
    assert r.returncode == 1 and re.search(TS + r"ERROR---> file does not exist: .*missing.fq$", r.stdout.strip().split("\n")[-1])
    odd = tmp_path / "reads.txt"
    odd.write_text("@r\nACGT\n+\nIIII\n")
    r = _hulk("sketch", "-f", str(odd), "-o", str(tmp_path / "o"))
    assert r.returncode == 1 and "ERROR---> file does not have recognised extension" in r.stdout
    r = _hulk("sketch", "-o", str(tmp_path / "o"), stdin=subprocess.DEVNULL)
    assert r.returncode == 1 and "ERROR---> no STDIN found" in r.stdout
    # the banner lines of cmd/sketch.go:90-122 come before any GPU work
    lines = _hulk("sketch", "-f", str(odd), "-o", str(tmp_path / "o")).stdout.split("\n")
    assert re.match(TS + r"this is hulk \(version 1\.0\.0\)$", lines[0])
    assert lines[2].endswith("starting the sketch subcommand") and lines[3].endswith("checking parameters...")


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_cli_fails_loudly_without_gpu(tmp_path):
    r = _hulk("sketch", "-f", FIXTURE, "-o", str(tmp_path / "o"))
    assert r.returncode == 1 and "ERROR---> CUDA error: no CUDA device (no CPU fallback exists)" in r.stdout
    assert not os.path.exists(str(tmp_path / "o.json"))


# ---- GPU: the front end end to end ------------------------------------------------------------------
EXPECTED_LOG = [
    r"this is hulk \(version 1\.0\.0\)", r"please cite Rowe et al\. 2019, doi: https://doi\.org/10\.1186/s40168-019-0653-2",
    r"starting the sketch subcommand", r"checking parameters\.\.\.", r"\tmode: FASTQ", r"\tno\. processors: 1",
    r"\tminimizer k-mer size: 21", r"\tminimizer window size: 9", r"\tsketch size: 50", r"\tstreaming: disabled",
    r"\tconcept drift: disabled", r"\tnumber of bins in k-mer spectrum: 194481", r"\tadding KHF sketch: false",
    r"\tadding KMV sketch: false", r"initialising sketching pipeline\.\.\.", r"\tinitialising the processes",
    r"\tconnecting data streams", r"\tnumber of processes added to the sketching pipeline: 4",
    r"\tnumber of minions in the sketching pool: 1", r"finding minimizers\.\.\.",
    r"generating final histosketch of k-mer spectra\.\.\.", r"\tprocessed 1000 sequences in total",
    r"\tmean sequence length: 100", r"\tfound 17040 minimizers", r"\thistosketching across 194481 bins",
    r"cleaning up\.\.\.", r"\twritten sketch to disk: .*\.json", r"finished in [0-9.]+(µs|ms|s)",
]


def _same_sketch(doc, want, filename):
    gj, wj = json.loads(doc), json.loads(want)
    assert gj["filename"] == filename
    gs, ws = gj["signatures"][0]["Sketch"], wj["signatures"][0]["Sketch"]
    assert gs["mins"] == ws["mins"] and gs["md5sum"] == ws["md5sum"]
    np.testing.assert_allclose(gs["weights"], ws["weights"], rtol=1e-12)
    strip = lambda d: [ln for ln in d.split("\n") if "e" not in ln and "." not in ln and "filename" not in ln]
    assert strip(doc) == strip(want)                         # byte-identical layout, key order, integers


@pytest.mark.gpu
def test_cli_sketch_c1_matches_golden(tmp_path):
    # BASELINE config C1 (.travis.yml:22 of the reference): hulk sketch -f testing/test-reads-small.fq.gz -k 21 -s 50
    out = str(tmp_path / "run" / "c1")                       # the output directory is created (cmd/sketch.go:188-195)
    r = _hulk("sketch", "-f", FIXTURE, "-k", "21", "-s", "50", "-o", out)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.rstrip("\n").split("\n")
    assert len(lines) == len(EXPECTED_LOG), r.stdout
    for ln, pat in zip(lines, EXPECTED_LOG):
        assert re.match(TS + pat + "$", ln), (ln, pat)
    _same_sketch(open(out + ".json").read(), open(os.path.join(GOLDEN, "c1_k21_s50.json")).read(), FIXTURE + ",")


@pytest.mark.gpu
def test_cli_intervals_drift_stream_and_stdin(tmp_path):
    out = str(tmp_path / "c1x")
    r = _hulk("sketch", "--fastq=" + FIXTURE, "-x", "0.2", "-i", "250", "-s50", "--stream", "-b", "lbl", "-p", "2", "-o", out)
    assert r.returncode == 0 and r.stdout == ""             # --stream moves the log to <outFile>.log
    log = open(out + ".log").read()
    assert "\tconcept drift: enabled\n" in log and "\tdecay ratio: 0.20\n" in log and "\tstreaming: enabled\n" in log
    assert [m for m in re.findall(r"\treached interval (\d+) -> histosketching", log)] == ["1", "2", "3", "4"]
    assert "merging sketches and cleaning up..." in log
    doc = open(out + ".json").read()
    _same_sketch(doc.replace('"banner_label": "lbl"', '"banner_label": "blank"'),
                 open(os.path.join(GOLDEN, "c1_k21_s50_x02_i250.json")).read(), FIXTURE + ",")
    assert '"banner_label": "lbl"' in doc and '"concept_drift": true' in doc
    # STDIN (plain FASTQ through a pipe) gives the same sketch, filename "STDIN"
    raw = gzip.open(FIXTURE, "rb").read()
    p = subprocess.run([hulk_b200.CLI_PATH, "sketch", "-k", "21", "-s", "50", "-o", str(tmp_path / "pipe")],
                       input=raw, capture_output=True, timeout=600)
    assert p.returncode == 0, p.stdout
    assert b"\tinput file: using STDIN" in p.stdout
    _same_sketch(open(str(tmp_path / "pipe.json")).read(), open(os.path.join(GOLDEN, "c1_k21_s50.json")).read(), "STDIN")


@pytest.mark.gpu
def test_cli_gpus_flag_gives_the_same_json(tmp_path):
    """`hulk sketch --gpus N` (hulk_b200_group_*: reads split per interval, spectra summed over NVLink inside the flush,
    slots sharded) writes the golden sketches.  With the members on one device the mechanism runs on a single-GPU box;
    with two or more GPUs present the real thing runs as well."""
    import torch
    runs = [("3", {"HULK_B200_GPUS_ON_ONE_DEVICE": "1", "CUDA_DEVICE_MAX_CONNECTIONS": "32"})]
    if torch.cuda.device_count() >= 2:
        runs.append((str(min(torch.cuda.device_count(), 4)), {}))
    for gpus, env in runs:
        for name, extra in (("c1_k21_s50.json", []), ("c1_k21_s50_x02_i250.json", ["-x", "0.2", "-i", "250"])):
            out = str(tmp_path / ("g" + gpus + name))
            r = _hulk("sketch", "-f", FIXTURE, "-k", "21", "-s", "50", "--gpus", gpus, "-o", out, *extra, env=env)
            assert r.returncode == 0, r.stdout + r.stderr
            assert "\tfound 17040 minimizers" in r.stdout and "\tprocessed 1000 sequences in total" in r.stdout
            _same_sketch(open(out + ".json").read(), open(os.path.join(GOLDEN, name)).read(), FIXTURE + ",")
    r = _hulk("sketch", "-f", FIXTURE, "--gpus", "0", "-o", str(tmp_path / "bad"))
    assert r.returncode == 1 and "--gpus must be between 1 and 16" in r.stdout + r.stderr


@pytest.mark.gpu
def test_cli_with_tables_drawn_on_the_device(tmp_path):
    """HULK_B200_CWS_DEVICE=1: `hulk sketch` draws the CWS tables on the GPU (also with --gpus: every member its own rows).
    Same mins and md5 as the golden sketch; the weights agree to 1e-12 (tables equal to the last bit or two)."""
    gold = json.load(open(os.path.join(GOLDEN, "c1_k21_s50.json")))
    for gpus, env in (("1", {}), ("2", {"HULK_B200_GPUS_ON_ONE_DEVICE": "1", "CUDA_DEVICE_MAX_CONNECTIONS": "32"})):
        out = str(tmp_path / ("dev" + gpus))
        r = _hulk("sketch", "-f", FIXTURE, "-k", "21", "-s", "50", "--gpus", gpus, "-o", out,
                  env=dict(env, HULK_B200_CWS_DEVICE="1"))
        assert r.returncode == 0, r.stdout + r.stderr
        got = json.load(open(out + ".json"))
        g, w = got["signatures"][0]["Sketch"], gold["signatures"][0]["Sketch"]
        assert g["mins"] == w["mins"] and g["md5sum"] == w["md5sum"]
        np.testing.assert_allclose(g["weights"], w["weights"], rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_cli_reference_fatals(tmp_path):
    short = tmp_path / "short.fq"
    short.write_bytes(b"@r1\n" + b"ACGT" * 30 + b"\n+\n" + b"I" * 120 + b"\n@r2\nACGTACGT\n+\nIIIIIIII\n")
    r = _hulk("sketch", "-f", str(short), "-o", str(tmp_path / "s"))
    assert r.returncode == 1 and "ERROR---> sequence length must be >= w + k - 1" in r.stdout
    r = _hulk("sketch", "-f", FIXTURE, "-k", "32", "-o", str(tmp_path / "s"))
    assert r.returncode == 1 and "ERROR---> k size must be: 0 < k < 32" in r.stdout
    # fewer than 1 % of the bins used at the flush: "not used yet" (src/kmerspectrum/kmerspectrum.go:94-96)
    r = _hulk("sketch", "-f", FIXTURE, "-i", "10", "-o", str(tmp_path / "s"))
    assert r.returncode == 1 and "ERROR---> not used yet" in r.stdout
    empty = tmp_path / "empty.fq"
    empty.write_bytes(b"")
    r = _hulk("sketch", "-f", str(empty), "-o", str(tmp_path / "s"))
    assert r.returncode == 1 and "ERROR---> no sequences received" in r.stdout
    assert not os.path.exists(str(tmp_path / "s.json"))


@pytest.mark.gpu
def test_cli_khf_and_kmv_flags_behave_like_the_reference(tmp_path):
    """The reference builds both MinHash sketches but never feeds them (src/pipeline/boss.go:18-19,70-71):
    --khf writes a signature of s MaxUint64 values behind the histosketch, --kmv dies in sketchio.Add."""
    out = str(tmp_path / "khf")
    r = _hulk("sketch", "-f", FIXTURE, "-k", "21", "-s", "50", "--khf", "-o", out)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "\tadding KHF sketch: true\n" in r.stdout and "\tadding KMV sketch: false\n" in r.stdout
    doc = json.loads(open(out + ".json").read())
    assert [g["Algorithm"] for g in doc["signatures"]] == ["histosketch", "khf"]
    gs = doc["signatures"][0]["Sketch"]
    ws = json.loads(open(os.path.join(GOLDEN, "c1_k21_s50.json")).read())["signatures"][0]["Sketch"]
    assert gs["mins"] == ws["mins"] and gs["md5sum"] == ws["md5sum"] and list(gs) == list(ws)
    np.testing.assert_allclose(gs["weights"], ws["weights"], rtol=1e-12)
    khf = doc["signatures"][1]["Sketch"]
    assert list(khf) == ["ksize", "md5sum", "mins", "num"] and khf["ksize"] == 21 and khf["num"] == 50
    assert khf["mins"] == [2 ** 64 - 1] * 50 and khf["md5sum"] == hulk_b200.md5_mins(np.full(50, 2 ** 64 - 1, dtype=np.uint64))
    m, w, _ = hulk_b200.load_sketch(out + ".json", 21, "khf")
    assert m.tolist() == khf["mins"] and w.size == 0
    for flags in (("--kmv",), ("--kmv", "--khf")):
        out = str(tmp_path / "kmv")
        r = _hulk("sketch", "-f", FIXTURE, "-k", "21", "-s", "50", *flags, "-o", out)
        assert r.returncode == 1 and "ERROR---> no sketch was generated by the kmv algorithm" in r.stdout
        assert "cleaning up..." in r.stdout and not os.path.exists(out + ".json")


@pytest.mark.gpu
def test_cli_feeds_khf_and_kmv_when_asked(tmp_path, oracle, fixture_reads):
    """HULK_B200_FEED_MINHASH=1 (an extension, not a reference flag): every minimizer the collector receives also goes to
    KMVsketch.AddHash / KHFsketch.AddHash (src/minhash/kmv.go:40-71, khf.go:35-45); the histosketch signature is
    untouched and the two MinHash signatures follow in the reference's order (src/pipeline/sketch.go:227-234)."""
    out = str(tmp_path / "fed")
    r = _hulk("sketch", "-f", FIXTURE, "-k", "21", "-s", "50", "--kmv", "--khf", "-o", out,
              env={"HULK_B200_FEED_MINHASH": "1"})
    assert r.returncode == 0, r.stdout + r.stderr
    doc = json.loads(open(out + ".json").read())
    assert [g["Algorithm"] for g in doc["signatures"]] == ["histosketch", "kmv", "khf"]
    ws = json.loads(open(os.path.join(GOLDEN, "c1_k21_s50.json")).read())["signatures"][0]["Sketch"]
    assert doc["signatures"][0]["Sketch"]["mins"] == ws["mins"]
    keys = np.concatenate([oracle.minimizers(21, 9, rd) for rd in fixture_reads]).astype(np.uint64)
    kmv, khf = doc["signatures"][1]["Sketch"], doc["signatures"][2]["Sketch"]
    assert kmv["mins"] == np.sort(keys)[:50].tolist() and kmv["num"] == 50 and kmv["ksize"] == 21
    assert khf["mins"] == [int((keys * np.uint64(i + 1)).min()) for i in range(50)] and khf["num"] == 50
    assert kmv["md5sum"] == hulk_b200.md5_mins(np.array(kmv["mins"], dtype=np.uint64))
    gold = json.loads(open(os.path.join(GOLDEN, "c1_k21_s50_minhash.json")).read())
    assert kmv["mins"] == gold["kmv"] and khf["mins"] == gold["khf"]
    assert kmv["md5sum"] == gold["kmv_md5"] and khf["md5sum"] == gold["khf_md5"]


# ---- the parallel parse of plain FASTQ files (csrc/ingest.cpp produce_parallel) -----------------------
def _native_env(paths, env):
    """Read through the native reader in a subprocess (the mode switches are read from the environment at open)."""
    code = ("import sys, hashlib; sys.path.insert(0, %r); import hulk_b200\n"
            "try:\n"
            "    with hulk_b200.NativeReader(sys.argv[1:]) as rd:\n"
            "        n = 0; h = hashlib.md5(); hl = hashlib.md5()      # independent of how reads are batched\n"
            "        for b, offs in rd:\n"
            "            n += len(offs) - 1; h.update(b.tobytes()); hl.update((offs[1:] - offs[:-1]).tobytes())\n"
            "    print('OK', n, h.hexdigest(), hl.hexdigest())\n"
            "except ValueError as e:\n"
            "    print('ERR', n, h.hexdigest(), hl.hexdigest(), str(e))\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([os.sys.executable, "-c", code, *[str(p) for p in paths]], capture_output=True, text=True, env=e, timeout=300)
    assert r.returncode == 0, r.stderr
    return r.stdout.strip()


@pytest.mark.parametrize("chunk", ["16", "61", "257", "4096", "100000"])
def test_parallel_reader_equals_sequential(tmp_path, chunk):
    # every chunk size puts the task boundaries somewhere else: inside headers, sequences, quality lines,
    # runs of empty lines, CRLF pairs; records straddle tasks and files
    rng = np.random.default_rng(int(chunk))
    reads = random_reads(3000, 40, seed=5, ragged=120)
    parts = []
    for i, r in enumerate(reads):
        nl = b"\r\n" if rng.random() < 0.1 else b"\n"
        rec = b"@r%d" % i + nl + r + nl + b"+" + nl + bytes(rng.integers(33, 74, len(r)).astype(np.uint8)) + nl
        if rng.random() < 0.05:
            rec = rec.replace(nl, nl + b"\n" * int(rng.integers(1, 4)), 1)       # empty lines inside a record
        parts.append(rec)
    blob = b"".join(parts)
    cut1 = blob.index(b"\n", len(blob) // 3) + 1                 # file 1 ends inside a record (after some line)
    cut2 = 2 * len(blob) // 3                                     # file 2 ends in the middle of a line: no final newline
    cut2 = blob.index(b"\n", cut2)                                # ... exactly in front of a newline, which opens file 3
    files = [tmp_path / "a.fq", tmp_path / "b.fastq", tmp_path / "c.fq"]
    files[0].write_bytes(blob[:cut1])
    files[1].write_bytes(blob[cut1:cut2])
    files[2].write_bytes(blob[cut2 + 1:])
    seq = _native_env(files, {"HULK_B200_PARALLEL_READER": "0"})
    par = _native_env(files, {"HULK_B200_PARALLEL_READER": "1", "HULK_B200_PARALLEL_CHUNK": chunk})
    assert seq.startswith("OK 3000 ") and par == seq


def test_parallel_reader_errors_equal_sequential(tmp_path):
    reads = random_reads(400, 60, seed=8)
    recs = [b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)) for i, r in enumerate(reads)]
    bad = list(recs)
    bad[250] = bad[250].replace(b"@r250", b"r250")                # read ID without '@' in record 250
    f1 = tmp_path / "bad.fq"
    f1.write_bytes(b"".join(bad))
    long_ = list(recs)
    long_[300] = b"@r300\n" + b"A" * 70000 + b"\n+\n" + b"I" * 10 + b"\n"        # bufio.Scanner: token too long
    f2 = tmp_path / "long.fq"
    f2.write_bytes(b"".join(long_))
    for f, msg in ((f1, "read ID in fastq file does not begin with @: r250"), (f2, "token too long")):
        seq = _native_env([f], {"HULK_B200_PARALLEL_READER": "0"})
        for chunk in ("64", "1000", "50000"):
            par = _native_env([f], {"HULK_B200_PARALLEL_READER": "1", "HULK_B200_PARALLEL_CHUNK": chunk})
            assert par == seq and msg in par and par.startswith("ERR")


def test_parallel_reader_long_line_behind_a_chunk_boundary(tmp_path):
    """A sequence line far longer than a token (and than a worker's batch) followed by SHORT '+' and quality lines: the job
    that starts behind it looks back for the open record's lines and must refuse the long one instead of copying it
    (the copy used to run past the batch: 700 KB into a 384 KB buffer with the default 8 MB chunks scaled down here)."""
    reads = random_reads(60, 50, seed=9)
    recs = [b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)) for i, r in enumerate(reads)]
    for big, qual in ((70000, b"I"), (716800, b"II"), (65536, b"I" * 40)):
        body = list(recs)
        body[30] = b"@r30\n" + b"A" * big + b"\n+\n" + qual + b"\n"
        f = tmp_path / ("long_%d.fq" % big)
        f.write_bytes(b"".join(body))
        seq = _native_env([f], {"HULK_B200_PARALLEL_READER": "0"})
        assert seq.startswith("ERR") and "token too long" in seq
        for chunk in ("100", "777", "4096", "70000", "300000"):
            par = _native_env([f], {"HULK_B200_PARALLEL_READER": "1", "HULK_B200_PARALLEL_CHUNK": chunk})
            assert par == seq, (big, chunk)


# ---- BGZF (bgzip) input: members found from their headers, inflated on several threads ----------------
def _bgzf(data, block=0xff00, level=6):
    """htslib's blocked gzip: one member per <= 64 KiB of input, size in the 'BC' extra subfield, EOF marker."""
    import struct
    import zlib
    out = bytearray()
    for a in list(range(0, len(data), block)) + [None]:
        chunk = b"" if a is None else data[a:a + block]
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(comp) + 8
        out += struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 66, 67, 2, bsize - 1)
        out += comp + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk))
    return bytes(out)


def test_bgzf_bad_member_delivers_what_came_before(tmp_path):
    """A damaged member in the middle of a BGZF file: the parallel path hands on the reads of the members in front of it
    and then fails with the text the one-thread path (compress/gzip's rules) gives -- not the whole window dropped, not
    every failure called a checksum error."""
    import struct
    reads = random_reads(6000, 150, seed=33)
    rec = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)) for i, r in enumerate(reads))
    good = bytearray(_bgzf(rec))
    # member boundaries from the BSIZE fields
    offs, o = [], 0
    while o < len(good):
        offs.append(o)
        o += struct.unpack_from("<H", good, o + 16)[0] + 1
    assert len(offs) > 12
    m = offs[7]
    size = offs[8] - offs[7]
    cases = {}
    crc = bytearray(good); crc[m + size - 8] ^= 1                                  # CRC-32 of member 7
    cases["crc"] = (bytes(crc), "gzip: invalid checksum")
    isz = bytearray(good); isz[m + size - 4] ^= 1                                  # ISIZE of member 7
    cases["isize"] = (bytes(isz), "gzip: invalid checksum")
    body = bytearray(good); body[m + 18 + 40] ^= 0xFF; body[m + 18 + 41] ^= 0xFF   # deflate data of member 7
    cases["body"] = (bytes(body), None)
    for name, (data, want_msg) in cases.items():
        f = tmp_path / (name + ".fastq.gz")
        f.write_bytes(data)
        one = _native_env([f], {"HULK_B200_PARALLEL_READER": "0"}).split(" ", 4)
        par = _native_env([f], {"HULK_B200_PARALLEL_READER": "1", "HULK_B200_BGZF_WINDOW": "300000"}).split(" ", 4)
        assert one[0] == par[0] == "ERR", (name, one, par)
        n_one, n_par = int(one[1]), int(par[1])
        assert n_par > 0 and abs(n_one - n_par) <= 450, (name, n_one, n_par)      # at most one member's worth of reads apart
        if want_msg:
            assert one[4] == par[4] == want_msg, (name, one[4], par[4])
        else:
            assert par[4] in ("gzip: invalid checksum", "unexpected EOF") or par[4].startswith("flate: corrupt input"), par[4]


def test_bgzf_reader_equals_plain_gzip(tmp_path):
    reads = random_reads(20000, 150, seed=21, ragged=50)
    rec = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)) for i, r in enumerate(reads))
    assert gzip.decompress(_bgzf(rec)) == rec                       # the writer above makes valid multi-member gzip
    bg = tmp_path / "b.fastq.gz"
    bg.write_bytes(_bgzf(rec))
    gz = tmp_path / "g.fastq.gz"
    gz.write_bytes(gzip.compress(rec, 1))
    want = _native_env([gz], {})
    assert want.startswith("OK 20000 ")
    assert _native_env([bg], {}) == want                            # parallel BGZF path
    assert _native_env([bg], {"HULK_B200_PARALLEL_READER": "0"}) == want      # zlib path on the same file
    # several windows in flight (the next one is inflated while the current one is framed): window ends fall inside
    # lines and records, the last window can be the empty EOF marker alone
    for win in ("1", "65280", "200000", "1000000"):
        assert _native_env([bg], {"HULK_B200_BGZF_WINDOW": win}) == want
    # a BGZF file with an ordinary gzip member appended (cat a.gz b.gz): zlib takes over where BGZF ends,
    # in the middle of a record and of a line
    cut = rec.index(b"\n", len(rec) // 2) - 7
    mixed = tmp_path / "m.fq.gz"
    mixed.write_bytes(_bgzf(rec[:cut])[:-28] + gzip.compress(rec[cut:], 1))     # (-28: drop the EOF marker block)
    assert _native_env([mixed], {}) == want
    # a corrupted block is an error, not silence
    raw = bytearray(_bgzf(rec))
    raw[len(raw) // 2] ^= 0x55
    bad = tmp_path / "bad.fq.gz"
    bad.write_bytes(bytes(raw))
    assert _native_env([bad], {}).startswith("ERR")
    assert _native_env([bad], {"HULK_B200_BGZF_WINDOW": "300000"}).startswith("ERR")
    # the mixed file again with small windows, and a member with a broken header after good blocks
    assert _native_env([mixed], {"HULK_B200_BGZF_WINDOW": "100000"}) == want
    good = _bgzf(rec)
    third = good.index(b"\x1f\x8b\x08\x04", len(good) // 3)
    hdr = tmp_path / "hdr.fq.gz"
    hdr.write_bytes(good[:third] + b"XX" + good[third + 2:])
    for env in ({}, {"HULK_B200_BGZF_WINDOW": "100000"}):
        got = _native_env([hdr], env)
        assert got.startswith("ERR") and "gzip: invalid header" in got


# ---- randomised line soups: every reader path against the pure-Python restatement of the Go framing -----
def _soup(rnd, fasta):
    """A file of record-like runs with stray empty lines, CRLF, a missing final newline and -- in some files --
    lines dropped or junk inserted, which shifts the never-resynchronising FASTQ framing."""
    lines = []
    damage = rnd.choice((0.0, 0.0, 0.03, 0.15))
    for i in range(rnd.randint(0, 40)):
        seq = [bytes(rnd.choice(b"ACGTNacgt") for _ in range(rnd.randint(1, 90))) for _ in range(rnd.randint(1, 3) if fasta else 1)]
        rec = [b">s%d x" % i] + seq if fasta else [b"@r%d" % i, seq[0], b"+", b"I" * len(seq[0])]
        for ln in rec:
            if rnd.random() < 0.1 and not fasta:
                lines.append(b"")                      # nil line: skipped by the FASTQ filler (it ends a FASTA stream)
            t = rnd.random()
            if t < damage / 2:
                continue                               # line lost
            if t < damage:
                lines.append(bytes(rnd.choice(b"@>+!I# \tACGT") for _ in range(rnd.randint(1, 12))))
            lines.append(ln)
    if fasta and lines and rnd.random() < 0.3:
        lines.insert(rnd.randrange(len(lines)), b"")
    nl = b"\r\n" if rnd.random() < 0.2 else b"\n"
    data = nl.join(lines)
    if lines and rnd.random() < 0.7:
        data += nl
    return data


def _outcome(fn):
    try:
        return ("OK", fn())
    except ValueError as e:
        return ("ERR", str(e))


@pytest.mark.parametrize("fasta", [False, True])
def test_native_reader_random_soups_equal_python_restatement(tmp_path, monkeypatch, fasta):
    import random
    rnd = random.Random(11 + fasta)
    ref = hulk_b200.read_fasta if fasta else hulk_b200.read_fastq
    n_err = n_reads = 0
    for case in range(120):
        data = _soup(rnd, fasta)
        if fasta and not data.lstrip(b"\r\n").startswith(b">"):
            data = b">first\n" + data                  # (a FASTA stream that starts without a header panics in the reference)
        enc = rnd.choice(("plain", "gz", "bgzf"))
        path = tmp_path / ("s%03d.%s" % (case, "fa" if fasta else "fq") + ("" if enc == "plain" else ".gz"))
        path.write_bytes(data if enc == "plain" else gzip.compress(data, 1) if enc == "gz" else _bgzf(data, rnd.choice((64, 1000))))
        want = _outcome(lambda: ref(str(path)))
        modes = [{"HULK_B200_PARALLEL_READER": "0"},
                 {"HULK_B200_PARALLEL_READER": "1", "HULK_B200_PARALLEL_CHUNK": str(rnd.choice((1, 7, 50, 300))),
                  "HULK_B200_BGZF_WINDOW": str(rnd.choice((1, 100, 5000)))}]
        for env in modes:
            for key, val in env.items():
                monkeypatch.setenv(key, val)
            got = _outcome(lambda: _native([path], fasta=fasta, batch_bytes=rnd.choice((0, 256))))
            assert got == want, (case, enc, env, data)
        n_err += want[0] == "ERR"
        n_reads += len(want[1]) if want[0] == "OK" else 0
    assert n_reads > 500 and (fasta or n_err > 5)       # the soups exercise both outcomes


# ---- ordinary gzip on several threads (csrc/pgzip.h): guessed block starts, marker decoding, stitching ---------
def test_parallel_gzip_equals_zlib_path(tmp_path):
    """One deflate stream inflated on several threads must hand on exactly the bytes of the one-thread path: compression
    levels 1/6/9, several members (one empty), sync/full flush points, fixed-Huffman and stored blocks, a stream that is
    almost all back-references; tiny chunks put many stitch points, member ends and false block guesses into the run."""
    import zlib
    reads = random_reads(30000, 150, seed=31, ragged=60)
    rec = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, r, bytes(33 + (7 * j + i) % 40 for j in range(len(r)))) for i, r in enumerate(reads))
    files = {"l1": gzip.compress(rec, 1), "l6": gzip.compress(rec, 6), "l9": gzip.compress(rec[:2_000_000], 9)}
    third = rec.index(b"\n@", len(rec) // 3) + 1
    files["members"] = gzip.compress(rec[:third], 6) + gzip.compress(b"", 6) + gzip.compress(rec[third:], 2)
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    parts = []
    for i, a in enumerate(range(0, len(rec), 70_000)):
        parts += [co.compress(rec[a:a + 70_000]), co.flush(zlib.Z_FULL_FLUSH if i % 3 == 0 else zlib.Z_SYNC_FLUSH)]
    files["flush"] = b"".join(parts) + co.flush()
    co = zlib.compressobj(6, zlib.DEFLATED, 31, 9, zlib.Z_FIXED)
    files["fixed"] = co.compress(rec[:1_500_000]) + co.flush()
    co = zlib.compressobj(0, zlib.DEFLATED, 31)
    files["stored"] = co.compress(rec[:1_500_000]) + co.flush()
    rep = b"@r\n" + b"ACGT" * 40 + b"\n+\n" + b"I" * 160 + b"\n"
    files["repeats"] = gzip.compress(rep * 40000, 6)
    for name, blob in files.items():
        f = tmp_path / (name + ".fq.gz")
        f.write_bytes(blob)
        want = _native_env([f], {"HULK_B200_PARALLEL_READER": "0"})
        assert want.startswith("OK "), (name, want)
        for chunk, threads in (("100000", "4"), ("9000", "7"), ("1500", "3")):
            got = _native_env([f], {"HULK_B200_PGZ_MIN": "0", "HULK_B200_PGZ_CHUNK": chunk, "HULK_B200_PGZ_THREADS": threads})
            assert got == want, (name, chunk, threads)
    # the same errors in the same words as the one-thread path (Go's)
    good = files["l6"]
    for name, data, msg in (("garbage", good + b"trailing garbage, more than ten bytes", "gzip: invalid header"),
                            ("cut", good[:len(good) // 2], "unexpected EOF"),
                            ("crc", good[:-8] + bytes([good[-8] ^ 1]) + good[-7:], "gzip: invalid checksum"),
                            ("len", good[:-4] + bytes([good[-4] ^ 1]) + good[-3:], "gzip: invalid checksum"),
                            ("short", good + b"\x1f\x8b", "unexpected EOF")):
        f = tmp_path / (name + ".bad.fq.gz")
        f.write_bytes(data)
        for env in ({"HULK_B200_PARALLEL_READER": "0"}, {"HULK_B200_PGZ_MIN": "0", "HULK_B200_PGZ_CHUNK": "50000"}):
            got = _native_env([f], env)
            assert got.startswith("ERR") and got.endswith(" " + msg), (name, env, got)
    # damage in the middle of the deflate data: both paths refuse the file
    raw = bytearray(good)
    raw[len(raw) // 2] ^= 0x10
    f = tmp_path / "flip.fq.gz"
    f.write_bytes(bytes(raw))
    for env in ({"HULK_B200_PARALLEL_READER": "0"}, {"HULK_B200_PGZ_MIN": "0", "HULK_B200_PGZ_CHUNK": "50000"}):
        assert _native_env([f], env).startswith("ERR")
