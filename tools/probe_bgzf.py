#!/usr/bin/env python3
"""Throughput of the native reader on a BGZF FASTQ file (written here with zlib, htslib layout) and on the same data as
one ordinary gzip stream, each against the one-thread zlib path on the same file.  Usage: probe_bgzf.py [n_reads] [dir]"""
import os, struct, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hulk_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d = sys.argv[2] if len(sys.argv) > 2 else "/tmp"
path = os.path.join(d, "probe_bgzf.fq.gz")
reads = hulk_b200.synthetic_reads(n, 150, seed=5)
lut = np.frombuffer(b"ACGT", dtype=np.uint8)
rec = np.empty((n, 4 + 151 + 2 + 151), dtype=np.uint8)
rec[:, 0:4] = np.frombuffer(b"@r1\n", dtype=np.uint8)
rec[:, 4:154] = reads if reads.max() > 3 else lut[reads]
rec[:, 154] = 10
rec[:, 155:157] = np.frombuffer(b"+\n", dtype=np.uint8)
rec[:, 157:307] = rng = np.random.default_rng(1).integers(35, 75, (n, 150), dtype=np.uint8)   # quality lines that do not compress to nothing
rec[:, 307] = 10
data = rec.tobytes()
t0 = time.time()
with open(path, "wb") as fh:
    for a in list(range(0, len(data), 0xff00)) + [None]:
        chunk = b"" if a is None else data[a:a + 0xff00]
        co = zlib.compressobj(1, zlib.DEFLATED, -15)
        comp = co.compress(chunk) + co.flush()
        fh.write(struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 66, 67, 2, 12 + 6 + len(comp) + 8 - 1))
        fh.write(comp + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
print("wrote %.1f MB of FASTQ as %.1f MB BGZF in %.1f s" % (len(data) / 1e6, os.path.getsize(path) / 1e6, time.time() - t0), flush=True)
gz = os.path.join(d, "probe_plain.fq.gz")
t0 = time.time()
co = zlib.compressobj(4, zlib.DEFLATED, 31)
with open(gz, "wb") as fh:
    for a in range(0, len(data), 1 << 24):
        fh.write(co.compress(data[a:a + (1 << 24)]))
    fh.write(co.flush())
print("wrote the same as %.1f MB ordinary gzip (one stream, level 4) in %.1f s" % (os.path.getsize(gz) / 1e6, time.time() - t0), flush=True)


def run(label, file, env):
    os.environ["HULK_B200_PARALLEL_READER"] = env
    t0 = time.time()
    got = 0
    with hulk_b200.NativeReader([file]) as rd:
        for b, offs in rd:
            got += len(offs) - 1
    dt = time.time() - t0
    print("%-28s %d reads in %.2f s: %.2f GB/s of FASTQ, %.1f M reads/s" % (label, got, dt, len(data) / dt / 1e9, got / dt / 1e6), flush=True)


for rep in range(2):
    run("bgzf, parallel inflate", path, "1")
    run("bgzf, zlib one thread", path, "0")
    run("gzip, parallel (pgzip.h)", gz, "1")
    run("gzip, zlib one thread", gz, "0")
os.remove(path)
os.remove(gz)
