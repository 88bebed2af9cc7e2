#!/bin/bash
# build_variant.sh NAME "EXTRA_NVFLAGS" — an A/B build of csrc/ with other -D flags into variants/NAME/libgsrast.so
# (variants/ is git-ignored but travels to the GPU box); select it with GSRAST_LIB=variants/NAME/libgsrast.so
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
D="$ROOT/variants/$1"
rm -rf "$D"; mkdir -p "$D/csrc"
cp "$ROOT"/gaussiansplatting.jl_b200/csrc/*.cu "$ROOT"/gaussiansplatting.jl_b200/csrc/*.cuh "$ROOT"/gaussiansplatting.jl_b200/csrc/*.cpp "$ROOT"/gaussiansplatting.jl_b200/csrc/Makefile "$D/csrc/"
# the sources include ../../include/gsrast.h relative to csrc/: point the copy at the real header
sed -i "s#\.\./\.\./include/gsrast.h#$ROOT/include/gsrast.h#g" "$D/csrc/Makefile"
sed -i "s#\"../../include/gsrast.h\"#\"$ROOT/include/gsrast.h\"#g" "$D"/csrc/*.cu "$D"/csrc/*.cuh "$D"/csrc/*.cpp 2>/dev/null || true
make -C "$D/csrc" -j8 EXTRA_NVFLAGS="$2" > "$D/build.log" 2>&1 || { tail -20 "$D/build.log"; exit 1; }
mv "$D/csrc/libgsrast.so" "$D/libgsrast.so"
cp "$D"/csrc/render.o.ptxas.log "$D/" 2>/dev/null || true
rm -rf "$D/csrc"
echo "built $D/libgsrast.so"
