#!/usr/bin/env python3
"""Corpus for tools/asan/host_fuzz: valid JSON sketches and FASTQ/FASTA/gzip/BGZF files plus seeded mutations of
them (byte flips, truncations, duplicated and deleted spans, hostile hand-written cases).  usage: make_corpus.py DIR [N]"""
import gzip, json, os, random, struct, sys, zlib

out = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
rnd = random.Random(7)
for sub in ("json", "fastq", "fasta"):
    os.makedirs(os.path.join(out, sub), exist_ok=True)


def mutate(b: bytes) -> bytes:
    b = bytearray(b)
    for _ in range(rnd.randint(1, 6)):
        if not b:
            break
        op = rnd.randrange(6)
        i = rnd.randrange(len(b))
        if op == 0:
            b[i] ^= 1 << rnd.randrange(8)
        elif op == 1:
            b[i] = rnd.choice(b'{}[]",:\\\n\r@>+-e.0 \x00\xff')
        elif op == 2:
            del b[i:i + rnd.randint(1, 40)]
        elif op == 3:
            j = rnd.randrange(len(b))
            b[i:i] = b[j:j + rnd.randint(1, 60)]
        elif op == 4:
            del b[i:]
        else:
            b[i:i] = bytes(rnd.choice(b'[{"\\\n') for _ in range(rnd.randint(1, 30)))
    return bytes(b)


def bgzf(data, block=0xff00):
    o = bytearray()
    for a in list(range(0, len(data), block)) + [None]:
        chunk = b"" if a is None else data[a:a + block]
        co = zlib.compressobj(1, zlib.DEFLATED, -15)
        comp = co.compress(chunk) + co.flush()
        o += struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 66, 67, 2, 12 + 6 + len(comp) + 8 - 1)
        o += comp + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk))
    return bytes(o)


# ---- JSON sketches ----
here = os.path.dirname(os.path.abspath(__file__))
gold = os.path.join(here, "..", "..", "tests", "golden")
seeds = [open(os.path.join(gold, f), "rb").read() for f in sorted(os.listdir(gold)) if f.startswith("c1_k21") and f.endswith(".json")]
doc = json.loads(seeds[0])
doc["signatures"].append({"Algorithm": "khf", "Sketch": {"ksize": 21, "md5sum": "x", "mins": [2 ** 64 - 1] * 4, "num": 4}})
seeds.append(json.dumps(doc, indent=4).encode())
hostile = [
    b"", b"{", b"[" * 200000, b"{\"a\":" * 100000, b'"\\u', b'"\\ud800\\u', b'{"signatures": [{"Algorithm": "histosketch", "Sketch": 3}]}',
    b'{"class": "hulk_sketch", "version": "1.0.0", "signatures": [{"Algorithm": "histosketch", "Sketch": {"mins": 5, "weights": "x", "ksize": []}}]}',
    b'{"class": "hulk_sketch", "version": "1.0.0", "signatures": [{"Algorithm": "histosketch", "Sketch": {"mins": [[1]], "md5sum": 7, "ksize": {"a": 1}}}]}',
    b'{"signatures": {"Algorithm": 1}}', b"nul", b"-", b"1e", b'{"a" 1}', b'{"a": 1,}', b"[1 2]", b'"' + b"\\" * 99999,
    b'{"signatures": [null, 3, "x", []]}', b'\xef\xbb\xbf{}', b'{"class": "hulk_sketch", "signatures": [{"Algorithm": null}]}',
]
n = 0
for s in seeds + hostile:
    open(os.path.join(out, "json", "c%04d.json" % n), "wb").write(s)
    n += 1
for i in range(N):
    open(os.path.join(out, "json", "c%04d.json" % n), "wb").write(mutate(rnd.choice(seeds)))
    n += 1

# ---- FASTQ ----
def fastq(nrec, crlf=False, blank=0.0):
    nl = b"\r\n" if crlf else b"\n"
    o = bytearray()
    for i in range(nrec):
        L = rnd.randint(20, 300)
        seq = bytes(rnd.choice(b"ACGTN") for _ in range(L))
        o += b"@r%d" % i + nl + seq + nl + b"+" + nl + b"I" * L + nl
        if rnd.random() < blank:
            o += nl
    return bytes(o)


fq = [fastq(400), fastq(50, crlf=True), fastq(100, blank=0.3), fastq(3)[:-1], b"", b"\n\n\n", b"@only\n", b"A" * 70000 + b"\n",
      b"@r\n" + b"A" * 65535 + b"\n+\n" + b"I" * 65535 + b"\n", b"@r\n" + b"A" * 65536 + b"\n+\nI\n", b"\r\n\r\r\n", b"x\ny\nz\nw\n" * 10]
# lines of a token's length or more in each of the four record slots, next to the chunk boundaries of the parallel
# parser (HULK_B200_PARALLEL_CHUNK=777 in run.sh): short neighbours so that the record straddles a boundary
def long_line_cases():
    for slot in range(4):
        for big in (65535, 65536, 70000, 200000):
            for lead in (0, 3, 97):
                rec = [b"@r", b"ACGT" * 8, b"+", b"I" * 32]
                rec[slot] = (b"@" if slot == 0 else b"") + (b"A" if slot != 3 else b"I") * big
                yield fastq(lead) + b"\n".join(rec) + b"\n" + fastq(4)
    yield b"@r\n" + b"A" * 70000 + b"\n+\nII\n@s\nACGT\n+\nIIII\n"          # long sequence, short quality line
    yield fastq(7) + b"@r\n" + b"A" * 716800 + b"\n+\nI\n" + fastq(9)


fq += list(long_line_cases())
n = 0
for s in fq:
    for enc in ("plain", "gz", "bgzf"):
        name = os.path.join(out, "fastq", "c%04d.fq" % n) + ("" if enc == "plain" else ".gz")
        open(name, "wb").write(s if enc == "plain" else gzip.compress(s, 1) if enc == "gz" else bgzf(s, 4096))
        n += 1
base = fastq(300)
for i in range(N):
    enc = rnd.choice(("plain", "gz", "bgzf", "gz2"))
    if enc == "plain":
        data, ext = mutate(base), ""
    elif enc == "gz":
        data, ext = mutate(gzip.compress(base, 1)), ".gz"
    elif enc == "gz2":
        data, ext = gzip.compress(mutate(base)[:5000], 1) + mutate(gzip.compress(base[:3000], 1)), ".gz"
    else:
        data, ext = mutate(bgzf(base, rnd.choice((512, 4096, 0xff00)))), ".gz"
    open(os.path.join(out, "fastq", "c%04d.fq%s" % (n, ext)), "wb").write(data)
    n += 1

# ---- FASTA ----
def fasta(nrec):
    o = bytearray()
    for i in range(nrec):
        o += b">s%d desc\n" % i
        for _ in range(rnd.randint(1, 6)):
            o += bytes(rnd.choice(b"ACGTN") for _ in range(rnd.randint(1, 80))) + b"\n"
    return bytes(o)


fa = [fasta(50), fasta(5) + b"\n" + fasta(3), b">a\n", b">\n>\n>\n", b"ACGT\n>x\nAC\n", b"", b">x\n" + b"A" * 65536 + b"\n"]
n = 0
for s in fa:
    open(os.path.join(out, "fasta", "c%04d.fa" % n), "wb").write(s)
    n += 1
    open(os.path.join(out, "fasta", "c%04d.fa.gz" % n), "wb").write(gzip.compress(s, 1))
    n += 1
base = fasta(80)
for i in range(N // 2):
    open(os.path.join(out, "fasta", "c%04d.fa" % n), "wb").write(mutate(base))
    n += 1
print("corpus written to", out)
