#!/bin/bash
# compute-sanitizer evidence (run under gpurun): memcheck and racecheck over the smoke pass (stem, tcgen05 convs incl. the
# fused tails / decode epilogues, SPPF pool, NMS, packing, CUDA-graph path) and over conv / attention / NMS test subsets.
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name, tool, command...
  local name=$1 tool=$2; shift 2
  timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 86 "$@" > gpurun_out/sanitizer_${tool}_$name.log 2>&1
  echo "$tool $name rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_$name.log | tail -1)"
}
run smoke memcheck python __graft_entry__.py smoke
run conv memcheck python -m pytest tests/test_gpu_conv.py -q -x -k "cta_pair and (3x3_c128_40x40 or 3x3s2_c64_n128 or 1x1_upadd or 3x3_views_res or 3x3_flat_20x20_n512 or 1x1_k768) or tail1x1"
run pool_nms memcheck python -m pytest tests/test_gpu_path.py -q -x -k "sppf or nms or decode"
run attention memcheck python -m pytest tests/test_gpu_attention.py -q -x -k "fused or cbam"
run smoke racecheck python __graft_entry__.py smoke
run conv racecheck python -m pytest tests/test_gpu_conv.py -q -x -k "cta_pair and (3x3_c128_40x40-fwd or 3x3_views_res-fwd or 1x1_upadd-fwd)"
run pool_nms racecheck python -m pytest tests/test_gpu_path.py -q -x -k "sppf or nms_"
