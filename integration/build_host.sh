#!/bin/bash
# integration/build_host.sh -- builds the reference host twice from a scratch copy of /root/reference (nothing is copied into the repo):
#   integration/_build/mmseqs_avx2   the unmodified reference, AVX2, CPU only        (baseline arm of the wall-clock comparison)
#   integration/_build/mmseqs_b200   the same sources + integration/mmseqs_b200.patch, -DENABLE_B200=1, linked to libb200align.so
# Both get integration/no_rust.patch (this image has no cargo: the Rust block-aligner is replaced by integration/block_aligner_stub.c,
# SURVEY.md 8c / T7).  integration/_build/ is git-ignored but travels to the GPU box with gpurun.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF:-/root/reference}"
SCRATCH="${SCRATCH:-/tmp/mmseqs_host_build}"
OUT="$HERE/_build"
WHAT="${1:-both}"
export CC=/usr/bin/gcc CXX=/usr/bin/g++       # /opt/gcc/bin has no libgomp.spec
mkdir -p "$OUT" "$SCRATCH"

prepare() {   # $1 = tree name, $2 = apply the drop-in patch (0/1)
    local tree="$SCRATCH/$1"
    if [ ! -f "$tree/.prepared" ]; then
        rm -rf "$tree"; mkdir -p "$tree"
        (cd "$REF" && tar cf - --exclude=.git .) | (cd "$tree" && tar xf -)
        chmod -R u+w "$tree"
        (cd "$tree" && patch -s -p1 < "$HERE/no_rust.patch")
        if [ "$2" = 1 ]; then (cd "$tree" && patch -s -p1 < "$HERE/mmseqs_b200.patch"); fi
        touch "$tree/.prepared"
    fi
}

build() {     # $1 = tree name, $2 = output binary name, rest = extra cmake args
    local tree="$SCRATCH/$1" bdir="$SCRATCH/$1-build" name="$2"; shift 2
    cmake -S "$tree" -B "$bdir" -G Ninja -DCMAKE_BUILD_TYPE=Release -DHAVE_AVX2=1 -DHAVE_TESTS=0 -DHAVE_SHELLCHECK=0 \
          -DBLOCK_ALIGNER_STUB="$HERE/block_aligner_stub.c" "$@" > "$bdir.cmake.log" 2>&1 || { tail -30 "$bdir.cmake.log"; exit 1; }
    ninja -C "$bdir" mmseqs > "$bdir.ninja.log" 2>&1 || { grep -B2 -A12 "error" "$bdir.ninja.log" | head -80; exit 1; }
    cp "$bdir/src/mmseqs" "$OUT/$name"
    echo "built $OUT/$name"
}

# BASELINE config[0] input (data, not source): travels to the GPU box inside the git-ignored _build/
mkdir -p "$OUT/examples" && cp "$REF/examples/QUERY.fasta" "$REF/examples/DB.fasta" "$OUT/examples/"

if [ "$WHAT" = avx2 ] || [ "$WHAT" = both ]; then
    prepare ref 0
    build ref mmseqs_avx2
fi
if [ "$WHAT" = b200 ] || [ "$WHAT" = both ]; then
    python3 -c "import sys; sys.path.insert(0, '$ROOT'); from mmseqs2_b200 import build; build.build()"
    rm -f "$SCRATCH/b200/.prepared"     # the patch is the thing under development: always re-apply
    prepare b200 1
    build b200 mmseqs_b200 -DENABLE_B200=1 -DB200_ROOT="$ROOT" -DCMAKE_BUILD_RPATH='$ORIGIN/../../mmseqs2_b200' -DCMAKE_SKIP_BUILD_RPATH=OFF
fi
