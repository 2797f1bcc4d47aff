#!/bin/bash
# A/B timing of FFT kernel variants (tools/time_fft.py prints the per-kernel table): environment switches of the default library
# (NAME=VALUE arguments, one run each) and libraries built by tools/build_variant.sh (other arguments)
out=gpurun_out/fft_variants.log
: > $out
run() { echo "=== $1" >> $out; shift; env "$@" timeout 120 python tools/time_fft.py --shape 840 1440 --tiles 8 --freqs 2 --scales 16 >> $out 2>&1; }
run "default" X=1
for v in "$@"; do
  if [[ $v == *=* ]]; then run "$v" $v; else run "variant $v" NEMO_B200_LIB=build_exp/lib_$v.so; fi
done
cat $out
