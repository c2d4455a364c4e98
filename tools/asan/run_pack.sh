#!/bin/bash
# Sanitizer passes over the host packer and its worker pool (csrc/pack.cpp): ASan + UBSan, then TSan.  CPU-only box.
set -u
here="$(cd "$(dirname "$0")" && pwd)"
root="$here/../.."
work="${TMPDIR:-/tmp}/hulk_b200_pack_san"
rm -rf "$work" && mkdir -p "$work"
rc=0
for san in "address,undefined" "thread"; do
    g++ -std=c++17 -O1 -g -fno-omit-frame-pointer -fsanitize=$san -I "$root/include" "$here/pack_stress.cpp" \
        "$root/hulk_b200/csrc/pack.cpp" -o "$work/pack_stress_${san%%,*}" -lpthread || exit 1
    for isa in scalar avx2 avx512; do
        echo "== -fsanitize=$san HULK_B200_PACK_ISA=$isa"
        HULK_B200_PACK_ISA=$isa "$work/pack_stress_${san%%,*}" || rc=1
    done
done
echo "pack sanitizer exit status: $rc"
exit $rc
