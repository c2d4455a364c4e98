#!/bin/bash
# AddressSanitizer + UBSan pass over the host-only parsers (JSON sketches, FASTQ/FASTA/gzip/BGZF reader).
# Runs on a CPU-only box.  usage: tools/asan/run.sh [corpus size]   -> profiles/<tag>_asan_host.txt is the caller's job
set -u
here="$(cd "$(dirname "$0")" && pwd)"
root="$here/../.."
work="${TMPDIR:-/tmp}/hulk_b200_asan"
rm -rf "$work" && mkdir -p "$work"
g++ -std=c++17 -O1 -g -fno-omit-frame-pointer -fsanitize=address,undefined -fno-sanitize-recover=undefined \
    -I "$root/include" "$here/host_fuzz.cpp" "$root/hulk_b200/csrc/sketch_json.cpp" "$root/hulk_b200/csrc/ingest.cpp" \
    "$root/hulk_b200/csrc/host_io.cpp" -o "$work/host_fuzz" -lz -lpthread || exit 1
python "$here/make_corpus.py" "$work/corpus" "${1:-300}" || exit 1
export ASAN_OPTIONS=detect_leaks=1:abort_on_error=0
rc=0
"$work/host_fuzz" json "$work"/corpus/json/* || rc=1
for par in 0 1; do
    # par=1 must reach produce_parallel: that needs the default batch size (HOST_FUZZ_BATCH=0), see host_fuzz.cpp
    HOST_FUZZ_BATCH=$((par ? 0 : 65536)) \
    HULK_B200_PARALLEL_READER=$par HULK_B200_PARALLEL_CHUNK=777 HULK_B200_BGZF_WINDOW=5000 HOST_FUZZ_VERBOSE=1 \
        "$work/host_fuzz" fastq "$work"/corpus/fastq/* > "$work/verdict$par.txt" || rc=1
    tail -1 "$work/verdict$par.txt"
done
# ordinary gzip through the multi-threaded single-stream inflater (pgzip.h), forced on with tiny chunks
HULK_B200_PARALLEL_READER=1 HULK_B200_PGZ_MIN=0 HULK_B200_PGZ_CHUNK=700 HULK_B200_PGZ_THREADS=5 HULK_B200_BGZF_WINDOW=5000 \
    HOST_FUZZ_VERBOSE=1 "$work/host_fuzz" fastq "$work"/corpus/fastq/* > "$work/verdict2.txt" || rc=1
tail -1 "$work/verdict2.txt"
# every path must reach the same verdict (accepted / rejected) on every file
for v in 1 2; do
    if ! diff <(awk '{print $1, $3}' "$work/verdict0.txt") <(awk '{print $1, $3}' "$work/verdict$v.txt") > "$work/diff$v.txt"; then
        echo "verdicts differ between the one-thread reader and path $v:"; head -20 "$work/diff$v.txt"; rc=1
    fi
done
"$work/host_fuzz" fasta "$work"/corpus/fasta/* || rc=1
echo "asan/ubsan exit status: $rc"
exit $rc
