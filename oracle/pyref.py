"""Second, independent restatement of the `hulk sketch` hot path in plain Python.

TEST INFRASTRUCTURE ONLY (same rule as oracle/hulk_oracle.c).  Pure-Python loops: use for
small cases.  Deliberately structured differently from the C oracle so the two can check each
other: the sliding-window minimum is a brute-force min over the window (no deque), the
per-read set is a Python set, the count-min sketch is a dict of counters.

Also holds the oracle side of the output stage: Go-compatible JSON (encoding/json
MarshalIndent, src/sketchio/sketchio.go:78-97) and the md5 of the mins
(src/helpers/helpers.go:156-166).
"""
from __future__ import annotations

import hashlib
import math
import struct

M64 = (1 << 64) - 1


def nt4(b: int) -> int:
    """src/minimizer/minimizer.go:13-30"""
    if b < 4:
        return b
    ch = chr(b).upper()
    return {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}.get(ch, 4)


def hash64(key: int, mask: int) -> int:
    """src/minimizer/minimizer.go:33-42"""
    key = ((~key & M64) + ((key << 21) & M64)) & M64 & mask
    key ^= key >> 24
    key = (key + ((key << 3) & M64) + ((key << 8) & M64)) & M64 & mask
    key ^= key >> 14
    key = (key + ((key << 2) & M64) + ((key << 4) & M64)) & M64 & mask
    key ^= key >> 28
    key = (key + ((key << 31) & M64)) & M64 & mask
    return key


def jump(key: int, n: int) -> int:
    """dgryski/go-jump Hash (Lamping & Veach); call sites kmerspectrum.go:70, countmin.go:125"""
    b, j = -1, 0
    while j < n:
        b = j
        key = (key * 2862933555777941757 + 1) & M64
        j = int(float(b + 1) * (float(1 << 31) / float((key >> 33) + 1)))
    return b


def position_values(k: int, w: int, seq: bytes):
    """X_i for every position that reaches the queue (None where skipped); minimizer.go:109-159"""
    mask = (1 << (2 * k)) - 1
    shift = 2 * (k - 1)
    fwd = rev = 0
    out = []
    for i, ch in enumerate(seq):
        c = nt4(ch)
        span = i - w + 2 if (i - w + 2) < k else k
        fwd = ((fwd << 2) | c) & mask
        rev = ((rev >> 2) | ((3 ^ c) << shift)) & M64
        if i < k - 1 or fwd == rev:
            out.append(None)
            continue
        canon = rev if fwd > rev else fwd
        x = ((hash64(canon, mask) << 8) & M64) | (span & M64)      # uint64(int32 span) sign-extends
        out.append(x)
    return out


def minimizers(k: int, w: int, seq: bytes) -> set:
    """The per-read set: { min X over the last w positions : every emitting position }"""
    if w > 256:
        raise ValueError("w must be: 0 < w < 257")
    if k > 31:
        raise ValueError("k size must be: 0 < k < 32")
    if len(seq) < 1:
        raise ValueError("sequence length must be > 0")
    if len(seq) < w + k - 1:
        raise ValueError("sequence length must be >= w + k - 1")
    xs = position_values(k, w, seq)
    res = set()
    for i, x in enumerate(xs):
        if x is None or i - w + 1 < 0:
            continue
        window = [v for v in xs[max(0, i - w + 1): i + 1] if v is not None]
        res.add(min(window))
    return res


def histogram(k: int, w: int, D: int, reads) -> list:
    hist = [0] * D
    for r in reads:
        for m in minimizers(k, w, r):
            hist[jump(m, D)] += 1
    return hist


class CountMin:
    """src/countmin/countmin.go:28-57,103-147 with lazily materialised counters"""

    def __init__(self, decay: float):
        self.width = math.ceil(2 / 0.001)
        self.depth = math.ceil(math.log(1 - 0.99) / math.log(0.5))
        self.q = {}
        self.scaling = 0.0 < decay < 1.0
        self.weight = math.exp(-decay) if self.scaling else 0.0

    def add(self, element: int, inc: float) -> float:
        if self.scaling:
            for key in self.q:
                self.q[key] = self.q[key] * self.weight
        cur = 1.7976931348623157e308
        for d in range(self.depth):
            g = jump((element + d * element) & M64, self.width)
            v = self.q.get((d, g), 0.0)
            if inc != 0.0:
                v += inc
                self.q[(d, g)] = v
            cur = min(cur, v)
        return cur


class HistoSketch:
    """src/histosketch/histosketch.go:50-92,129-155; r, c, b indexable as [slot][bin]"""

    def __init__(self, k, s, D, decay, r, c, b):
        if k > 31:
            raise ValueError("histosketching only supports k <= 31")
        if decay < 0.0 or decay > 1.0:
            raise ValueError("decay ratio must be between 0.0 and 1.0")
        if D < 2:
            raise ValueError("histogram must have at least 2 bins")
        self.s, self.D = s, D
        self.r, self.c, self.b = r, c, b
        self.sketch = [0] * s
        self.weights = [1.7976931348623157e308] * s
        self.cms = CountMin(decay)
        self.drift = decay != 1.0

    def add_element(self, bin_id: int, value: float) -> float:
        f = self.cms.add(bin_id, value)
        for j in range(self.s):
            yka = math.exp(math.log(f) - self.b[j][bin_id])
            a = self.c[j][bin_id] / (yka * math.exp(self.r[j][bin_id]))
            if self.drift:
                wgt = self.weights[j]
                dw = self.cms.weight
                if dw == 0.0:       # Go float division by zero: +-Inf / NaN, no exception
                    cur = math.nan if wgt == 0.0 else math.copysign(math.inf, wgt)
                else:
                    cur = wgt / dw
            else:
                cur = self.weights[j]
            if a < cur:
                self.sketch[j] = bin_id
                self.weights[j] = a
        return f

    def flush(self, hist) -> None:
        used = sum(1 for v in hist if v != 0)
        if used == 0:
            return
        if used / self.D < 0.01:
            raise ValueError("not used yet")
        for i, v in enumerate(hist):
            if v != 0:
                self.add_element(i, float(v))


# ------------------------------------------------------------------------------------------
# output stage
# ------------------------------------------------------------------------------------------
def md5_of_mins(mins) -> str:
    """src/helpers/helpers.go:156-166 + histosketch.go:167-170 ("%x" of the 16 bytes)"""
    return hashlib.md5(b"".join(struct.pack("<Q", int(m)) for m in mins)).hexdigest()


def go_float(f: float) -> str:
    """encoding/json floatEncoder: shortest round-trip digits, 'e' form iff |f| < 1e-6 or >= 1e21,
    and "e-0X" rewritten to "e-X"."""
    if f != f or f in (math.inf, -math.inf):
        raise ValueError("json: unsupported value")
    if f == 0:
        return "-0" if math.copysign(1.0, f) < 0 else "0"
    a = abs(f)
    r = repr(a)
    # shortest digits + decimal exponent from Python's repr (same shortest-round-trip rule)
    if "e" in r:
        mant, exp = r.split("e")
        exp = int(exp)
    else:
        mant, exp = r, 0
    if "." in mant:
        ip, fp = mant.split(".")
    else:
        ip, fp = mant, ""
    digits = (ip + fp).lstrip("0")
    # decimal point position relative to the start of `ip+fp`
    point = len(ip) + exp - (len(ip + fp) - len((ip + fp).lstrip("0")))
    digits = digits.rstrip("0") or "0"
    # value = 0.d1d2... * 10^point
    if a < 1e-6 or a >= 1e21:
        e = point - 1
        s = digits[0] + ("." + digits[1:] if len(digits) > 1 else "") + "e" + ("-" if e < 0 else "+")
        ae = abs(e)
        s += "%02d" % ae                                   # strconv: at least two exponent digits
        # encoding/json: clean up e-09 to e-9
        if len(s) >= 4 and s[-4] == "e" and s[-3] == "-" and s[-2] == "0":
            s = s[:-2] + s[-1]
    else:
        if point <= 0:
            s = "0." + "0" * (-point) + digits
        elif point >= len(digits):
            s = digits + "0" * (point - len(digits))
        else:
            s = digits[:point] + "." + digits[point:]
    return ("-" if f < 0 else "") + s


def go_json_string(s: str) -> str:
    """encoding/json string escaping (HTML-safe: <, >, & as \\u00XX; U+2028/2029 escaped)"""
    out = ['"']
    for ch in s:
        o = ord(ch)
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\t":
            out.append("\\t")
        elif o < 0x20 or ch in "<>&":
            out.append("\\u%04x" % o)
        elif o in (0x2028, 0x2029):
            out.append("\\u%04x" % o)
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def sketch_json(filename, k, mins, weights, D, drift, banner="blank", kmv=None, khf=None) -> str:
    """json.MarshalIndent(HULKdata, "", "    ") -- sketchio.go:20-34,86; histosketch.go:36-47;
    optional minhash signatures (kmv.go:12-21, khf.go:11-17) in the order pipeline/sketch.go:227-234 appends them"""
    I = "    "
    L = []
    L.append("{")
    L.append(f'{I}"class": "hulk_sketch",')
    L.append(f'{I}"filename": {go_json_string(filename)},')
    L.append(f'{I}"hash_function": "ntHash",')
    L.append(f'{I}"license": "CC0",')
    L.append(f'{I}"signatures": [')
    L.append(f"{I*2}{{")
    L.append(f'{I*3}"Algorithm": "histosketch",')
    L.append(f'{I*3}"Sketch": {{')
    L.append(f'{I*4}"ksize": {k},')
    L.append(f'{I*4}"md5sum": "{md5_of_mins(mins)}",')
    if len(mins):
        L.append(f'{I*4}"mins": [')
        L.extend(f"{I*5}{int(m)}" + ("," if i + 1 < len(mins) else "") for i, m in enumerate(mins))
        L.append(f"{I*4}],")
        L.append(f'{I*4}"weights": [')
        L.extend(f"{I*5}{go_float(float(x))}" + ("," if i + 1 < len(weights) else "")
                 for i, x in enumerate(weights))
        L.append(f"{I*4}],")
    else:
        L.append(f'{I*4}"mins": [],')
        L.append(f'{I*4}"weights": [],')
    L.append(f'{I*4}"num": {len(mins)},')
    L.append(f'{I*4}"num_histogram_bins": {D},')
    L.append(f'{I*4}"concept_drift": {"true" if drift else "false"}')
    L.append(f"{I*3}}}")
    for algo, mh in (("kmv", kmv), ("khf", khf)):
        if mh is None:
            continue
        if len(mh) == 0:
            raise ValueError(f"no sketch was generated by the {algo} algorithm")     # sketchio.go:59-61
        L.append(f"{I*2}}},")
        L.append(f"{I*2}{{")
        L.append(f'{I*3}"Algorithm": "{algo}",')
        L.append(f'{I*3}"Sketch": {{')
        L.append(f'{I*4}"ksize": {k},')
        L.append(f'{I*4}"md5sum": "{md5_of_mins(mh)}",')
        L.append(f'{I*4}"mins": [')
        L.extend(f"{I*5}{int(m)}" + ("," if i + 1 < len(mh) else "") for i, m in enumerate(mh))
        L.append(f"{I*4}],")
        L.append(f'{I*4}"num": {len(mh)}')
        L.append(f"{I*3}}}")
    L.append(f"{I*2}}}")
    L.append(f"{I}],")
    L.append(f'{I}"version": "1.0.0",')
    L.append(f'{I}"banner_label": {go_json_string(banner)}')
    L.append("}")
    return "\n".join(L)


# ---- `hulk smash` (test infrastructure, like everything in oracle/) ----------------------------------
def get_distance(mins_a, mins_b, weights_a, metric: str) -> float:
    """HULKdata.GetDistance (src/sketchio/sketchio.go:262-306) + distances.GetDistance / GetWJD
    (src/distances/distances.go:12-72).  For "weightedjaccard" the reference takes BOTH weight vectors from the
    subject (sketchio.go:293-299 asserts hsB from subjectSketchObj), so only weights_a is used."""
    set_a = [float(int(x)) for x in mins_a]
    set_b = [float(int(x)) for x in mins_b]
    if len(set_a) != len(set_b):
        raise ValueError("sketch length mismatch: %d vs %d" % (len(set_a), len(set_b)))
    if metric == "jaccard":
        intersect = 0.0
        for a, b in zip(set_a, set_b):
            if a == b:
                intersect += 1
        return 1.0 - (intersect / float(len(set_a)))
    if metric == "weightedjaccard":
        intersect, union = 0.0, 0.0
        for i in range(len(set_a)):
            w = float(weights_a[i])
            wa = max(max(w, 0.0), max(-w, 0.0))
            wb = wa
            if set_a[i] == set_b[i]:
                if wa < wb:
                    intersect += wa
                    union += wb
                else:
                    intersect += wb
                    union += wa
            else:
                union += wa if wa > wb else wb
        return 1 - (intersect / union)
    raise ValueError("unknown distance metric: %s" % metric)


def smash_matrix(mins, weights, metric: str):
    """makeMatrix (cmd/smash.go:183-226): similarity = 100 - distance * 100, formatted 'f', 2."""
    n = len(mins)
    sim = [[100 - (get_distance(mins[i], mins[j], weights[i], metric) * 100) for j in range(n)] for i in range(n)]
    return sim, [["%.2f" % v for v in row] for row in sim]


# ---- src/minhash: the two MinHash side sketches ------------------------------------------------------
# The reference constructs both next to the k-mer spectrum (src/pipeline/boss.go:70-71) and never feeds them
# (boss.go:18-19, "not used yet").  These are the types as written, for the feed the library can switch on
# (hulk_b200_minhash_enable): every minimizer the collector receives goes to AddHash.
class KHFsketch:
    """src/minhash/khf.go:11-60"""

    def __init__(self, k: int, s: int):
        self.k, self.s = k, s
        self.sketch = [M64] * s                                   # khf.go:20-32

    def add_hash(self, hv: int):                                  # khf.go:35-45
        for i in range(self.s):
            val = (hv + ((i * hv) & M64)) & M64
            if val < self.sketch[i]:
                self.sketch[i] = val

    def merge(self, other: "KHFsketch"):                          # khf.go:47-55
        for i, m in enumerate(other.sketch):
            if m < self.sketch[i]:
                self.sketch[i] = m

    def get_sketch(self):
        return self.sketch

    def similarity(self, other) -> float:                         # khf.go:82-100
        a, b = self.get_sketch(), other.get_sketch()
        n = min(len(a), len(b))
        return sum(1.0 for i in range(n) if a[i] == b[i]) / float(n)


class KMVsketch:
    """src/minhash/kmv.go:11-176 with container/heap over IntHeap (src/minhash/heap.go: Less is '>', so the LARGEST
    value sits at index 0)."""

    def __init__(self, k: int, s: int):
        self.k, self.s = k, s
        self.heap = []
        self.multiplicity_sum = 0

    # container/heap's up/down for a heap whose Less(i, j) is h[i] > h[j]
    def _up(self, j: int):
        h = self.heap
        while True:
            i = (j - 1) // 2
            if i == j or j <= 0 or not (h[j] > h[i]):
                break
            h[i], h[j] = h[j], h[i]
            j = i

    def _down(self, i0: int, n: int) -> bool:
        h = self.heap
        i = i0
        while True:
            j1 = 2 * i + 1
            if j1 >= n or j1 < 0:
                break
            j = j1
            j2 = j1 + 1
            if j2 < n and h[j2] > h[j1]:
                j = j2
            if not (h[j] > h[i]):
                break
            h[i], h[j] = h[j], h[i]
            i = j
        return i > i0

    def add_hash(self, hv: int):                                  # kmv.go:40-71
        self.multiplicity_sum += 1
        if len(self.heap) < self.s:
            self.heap.append(hv)                                  # heap.Push
            self._up(len(self.heap) - 1)
        elif hv < self.heap[0]:
            self.heap[0] = hv
            if not self._down(0, len(self.heap)):                 # heap.Fix(h, 0)
                self._up(0)

    def get_sketch(self):                                         # kmv.go:160-176
        return sorted(self.heap)

    def similarity(self, other: "KMVsketch") -> float:            # kmv.go:118-157
        longer, shorter = (self.heap, other.heap) if len(self.heap) > len(other.heap) else (other.heap, self.heap)
        counts = {}
        for v in longer:
            counts[v] = counts.get(v, 0) + 1
        intersect = 0
        for m in shorter:
            if counts.get(m, 0) > 0:
                counts[m] -= 1
                intersect += 1
        return float(intersect) / float(len(longer))
