#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE (see oracle/__init__.py).
# Compiles the reference's own retinaface/RetinaFace.cpp, UNMODIFIED and from where it lies
# under /root/reference, against the compile-only stub headers in oracle/shim plus the fake
# engine in oracle/ref_driver.cpp.  Output: oracle/_ref/libref_postproc.so (git-ignored; it
# travels to the GPU box with the snapshot).  No reference source is copied anywhere.
# The reference's CMake build is NOT used (needs OpenCV/Caffe/TensorRT, all absent).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${RF_REFERENCE_ROOT:-/root/reference}/retinaface"
out="$here/_ref"
if [ ! -f "$ref/RetinaFace.cpp" ]; then
  echo "build_ref: $ref/RetinaFace.cpp not found (no reference tree on this machine) -- keeping prebuilt $out" >&2
  exit 0
fi
mkdir -p "$out"
# -DUSE_TENSORRT selects the TensorRT branch of the reference (the only one that compiles as
# shipped); -I shim comes first so <cuda_runtime_api.h>, "NvInfer.h",
# <opencv2/opencv.hpp>, <caffe/caffe.hpp> resolve to the stubs.
g++ -std=c++14 -O2 -fPIC -shared -w -DUSE_TENSORRT \
    -I"$here/shim" -I"$ref" \
    "$ref/RetinaFace.cpp" "$here/ref_driver.cpp" \
    -o "$out/libref_postproc.so"
echo "build_ref: built $out/libref_postproc.so"
