#!/usr/bin/env bash
# Builds the checker libraries under oracle/:
#   oracle/_build/libuvrt_oracle.so  -- the plain-C port (always)
#   oracle/_ref/libuvrt_ref.so       -- the reference's OWN sources (bvh.cpp + cl/*.cl),
#                                       only when /root/reference (or $UVRT_REFERENCE) is present.
# Reference text is never copied into the repo: the transformed kernel sources live in
# oracle/_ref/gen/ (git-ignored) and differ from the originals only by two vector-literal
# rewrites and the dropped OpenCL pragma.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${UVRT_REFERENCE:-/root/reference}"
CFLAGS="-O2 -fPIC -fopenmp -ffp-contract=off -fno-fast-math -msse4.1"

mkdir -p "$here/_build"
gcc -std=c11 $CFLAGS -shared -o "$here/_build/libuvrt_oracle.so" "$here/uvrt_oracle.c" -lm

if [ -d "$ref/cl" ] && [ -f "$ref/bvh.cpp" ]; then
    mkdir -p "$here/_ref/gen"
    for k in generate extend accumulate shade reset; do
        sed -e 's/(float3)(/make_f3(/g' -e 's/(double2)(/make_d2(/g' -e '/#pragma OPENCL/d' \
            "$ref/cl/$k.cl" > "$here/_ref/gen/$k.cl.inc"
    done
    g++ -std=c++17 $CFLAGS -fpermissive -w -shared -Wl,-Bsymbolic -o "$here/_ref/libuvrt_ref.so" \
        -I"$here/ref_shim" -I"$here/_ref" -I"$ref" \
        "$here/ref_shim/ref_kernels.cpp" "$here/ref_shim/ref_bvh.cpp" "$ref/bvh.cpp"
    echo "built oracle/_ref/libuvrt_ref.so from $ref"
    # The reference's own loaders: mesh.cpp + bvh.cpp unmodified, tinygltf / tinyxml2 from its lib/,
    # LoadRoute / SaveRoute / UpdatePhotonsPerLight cut out of raytracer.cpp (the rest of that file is OpenCL plumbing).
    # Slow to compile (json.hpp): rebuilt only when missing or older than its sources.
    lib="$here/_ref/libuvrt_ref_loader.so"
    if [ -d "$ref/lib/tinygltf-master" ] && { [ ! -f "$lib" ] || [ "$here/ref_shim/loader/ref_loader.cpp" -nt "$lib" ] || [ "$here/ref_shim/loader/precomp.h" -nt "$lib" ]; }; then
        sed -n -e '61,64p' -e '228,300p' "$ref/raytracer.cpp" > "$here/_ref/gen/raytracer_route.inc"
        g++ -std=c++17 -O1 -fPIC -fopenmp -ffp-contract=off -msse4.1 -fpermissive -w -shared -Wl,-Bsymbolic -DGLM_ENABLE_EXPERIMENTAL \
            -o "$lib" -I"$here/ref_shim/loader" -I"$here/_ref" -I"$ref" -I"$ref/lib/tinygltf-master" -I"$ref/lib/tinyxml2" -I"$ref/lib/glm" \
            "$here/ref_shim/loader/ref_loader.cpp" "$here/ref_shim/loader/tinygltf_impl.cpp" "$ref/mesh.cpp" "$ref/bvh.cpp" \
            "$ref/lib/tinyxml2/tinyxml2.cpp" && echo "built oracle/_ref/libuvrt_ref_loader.so (the reference's GLB / route loaders)" \
            || echo "WARNING: the reference's loaders did not compile here; tests/test_host.py skips the T0 cross-check"
    fi
else
    echo "reference sources not found at $ref: keeping any prebuilt oracle/_ref/libuvrt_ref.so"
fi
