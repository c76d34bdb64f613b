#!/bin/bash
# phase timers of the forward engine: a CSB_PROF build of the library in a scratch copy
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
cp cswin-simam-unet_b200/libcsb200.so /tmp/libcsb200.keep
make -C cswin-simam-unet_b200/csrc -j16 EXTRA=-DCSB_PROF -B > gpurun_out/r2_prof_build.log 2>&1
echo "build rc=$?"
python benchmarks/debug/prof_bwd.py > gpurun_out/r2_prof_bwd.txt 2>&1
cat gpurun_out/r2_prof_bwd.txt
cp /tmp/libcsb200.keep cswin-simam-unet_b200/libcsb200.so
