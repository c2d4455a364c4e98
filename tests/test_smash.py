"""`hulk smash`: the JSON read side (sketchio.LoadHULKdata / FindSketch) and the
all-pairs similarity matrix (HULKdata.GetDistance, cmd/smash.go makeMatrix) against the Python restatement in
oracle/pyref.py.  CPU tests cover loading and the front end's checks; the matrix itself runs on the GPU."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

import hulk_b200
from conftest import GOLDEN

GOLD = os.path.join(GOLDEN, "c1_k21_s50.json")


def _write_sketch(path, mins, weights, k=21, banner="blank", D=194481):
    doc = hulk_b200.sketch_json(str(path) + ",", k, np.asarray(mins, dtype=np.uint64), np.asarray(weights, dtype=np.float64),
                                D, False, banner_label=banner)
    open(path, "w").write(doc)


def _pile(tmp_path, n=5, s=50, seed=3):
    """n related sketches: a base sketch with a growing number of slots replaced (some weights negative)."""
    rng = np.random.default_rng(seed)
    base_m = rng.integers(0, 194481, s).astype(np.uint64)
    base_w = rng.normal(0, 3, s)
    mins, weights = [], []
    for i in range(n):
        m, w = base_m.copy(), base_w + rng.normal(0, 0.1, s)
        idx = rng.choice(s, size=i * 7 % s, replace=False)
        m[idx] = rng.integers(0, 194481, idx.size).astype(np.uint64)
        mins.append(m)
        weights.append(w)
        _write_sketch(tmp_path / ("s%02d.json" % i), m, w, banner="b%d" % i)
    return np.array(mins), np.array(weights)


def test_load_sketch_reads_what_the_writer_wrote():
    m, w, banner = hulk_b200.load_sketch(GOLD)
    doc = json.load(open(GOLD))["signatures"][0]["Sketch"]
    assert m.tolist() == doc["mins"] and banner == "blank"
    np.testing.assert_array_equal(w, np.array(doc["weights"]))


def test_load_sketch_checks_mirror_reference(tmp_path):
    good = json.load(open(GOLD))

    def variant(name, edit):
        d = json.loads(json.dumps(good))
        edit(d)
        p = tmp_path / name
        p.write_text(json.dumps(d))
        return str(p)

    cases = [
        (variant("v.json", lambda d: d.update(version="0.9.9")), "different version of HULK: 0.9.9"),
        (variant("c.json", lambda d: d.update({"class": "other"})), "JSON not created by HULK"),
        (variant("m.json", lambda d: d["signatures"][0]["Sketch"]["mins"].__setitem__(0, 7)), "md5sum mismatch"),
        (variant("n.json", lambda d: d.update(signatures=[])), "no signatures found"),
        (variant("a.json", lambda d: d["signatures"][0].update(Algorithm="bloom")), "unknown sketching algorithm: bloom"),
        (str(tmp_path / "missing.json"), "file does not exist"),
    ]
    for path, msg in cases:
        with pytest.raises(hulk_b200.HulkError, match=msg):
            hulk_b200.load_sketch(path)
    # hostile documents are refused, not crashed on (tools/asan/run.sh runs the same parser under ASan/UBSan)
    for name, data in (("deep.json", b"[" * 300000), ("deepobj.json", b'{"a":' * 200000), ("cut.json", b'{"class": "hulk_sk'),
                       ("esc.json", b'"\\u12'), ("types.json", b'{"class": "hulk_sketch", "version": "1.0.0", "signatures": '
                                                  b'[{"Algorithm": "histosketch", "Sketch": {"mins": 5, "weights": "x", "ksize": []}}]}')):
        p = tmp_path / name
        p.write_bytes(data)
        with pytest.raises(hulk_b200.HulkError):
            hulk_b200.load_sketch(str(p))
    with pytest.raises(hulk_b200.HulkError, match=r"specified k-mer size \(31\) not found"):
        hulk_b200.load_sketch(GOLD, k=31)
    with pytest.raises(hulk_b200.HulkError, match="no sketches were produced using the kmv algorithm"):
        hulk_b200.load_sketch(GOLD, algo="kmv")


def _hulk(*args):
    return subprocess.run([hulk_b200.CLI_PATH, *args], capture_output=True, text=True, timeout=300)


def test_cli_smash_checks(tmp_path):
    r = _hulk("smash", "-d", str(tmp_path / "nowhere"), "-o", str(tmp_path / "o"))
    assert r.returncode == 1 and "ERROR---> directory does not exist" in r.stdout
    r = _hulk("smash", "-d", str(tmp_path), "-o", str(tmp_path / "o"))
    assert r.returncode == 1 and "ERROR---> no JSON files found in supplied directory" in r.stdout
    _write_sketch(tmp_path / "one.json", [1, 2, 3], [0.5, -1.0, 2.0])
    r = _hulk("smash", "-d", str(tmp_path), "-o", str(tmp_path / "o"))
    assert r.returncode == 1 and "1 sketches found in the supplied directory, HULK needs at least 2 to smash!" in r.stdout
    r = _hulk("smash", "-d", str(tmp_path), "-m", "cosine", "-o", str(tmp_path / "o"))
    assert r.returncode == 1 and "supplied distance metric is not available: cosine" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("metric", ["jaccard", "weightedjaccard"])
def test_smash_matrix_bit_exact(tmp_path, metric):
    from oracle import pyref as P
    mins, weights = _pile(tmp_path, n=6, s=64)
    mins[2, 5] = mins[3, 5] = np.uint64(2 ** 63 + 1)      # compared as float64, like the reference
    mins[3, 5] += np.uint64(1)
    sim = hulk_b200.smash(mins, weights, metric)
    want, _ = P.smash_matrix(mins, weights, metric)
    np.testing.assert_array_equal(sim, np.array(want))


@pytest.mark.gpu
def test_cli_smash_end_to_end(tmp_path):
    from oracle import pyref as P
    d = tmp_path / "pile"
    (d / "sub").mkdir(parents=True)
    mins, weights = _pile(d, n=4, s=50)
    _write_sketch(d / "sub" / "deep.json", mins[0], weights[0])
    for metric in ("jaccard", "weightedjaccard"):
        out = str(tmp_path / ("m_" + metric))
        r = _hulk("smash", "-d", str(d), "-m", metric, "-o", out, "--bannerMatrix")
        assert r.returncode == 0, r.stdout + r.stderr
        assert "\tnumber of sketch objects: 4" in r.stdout and "HULK SMASH!" in r.stdout
        rows = [ln.split(",") for ln in open(out + ".hulk-matrix.csv").read().strip().split("\n")]
        names = sorted(str(d / ("s%02d.json" % i)) for i in range(4))
        assert rows[0] == names
        _, want = P.smash_matrix(mins, weights, metric)
        assert rows[1:] == want
        banner = [ln.split(",") for ln in open(out + ".banner-matrix.csv").read().strip().split("\n")]
        assert [b[-1] for b in banner] == ["b0", "b1", "b2", "b3"] and banner[1][:-1] == [str(int(x)) for x in mins[1]]
    r = _hulk("smash", "-d", str(d), "--recursive", "-o", str(tmp_path / "rec"))
    assert r.returncode == 0 and "\tnumber of sketch objects: 5" in r.stdout
