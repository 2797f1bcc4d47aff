#!/bin/bash
# A/B timing of FFT kernel variants (tools/time_fft.py prints the per-kernel table): the default library, the TMA-fed column
# kernels switched on (all / plain / apply), everything register-staged, and libraries built by tools/build_variant.sh
out=gpurun_out/fft_variants.log
: > $out
run() { echo "=== $1" >> $out; shift; env "$@" timeout 120 python tools/time_fft.py --shape 840 1440 --tiles 8 --freqs 2 --scales 16 >> $out 2>&1; }
run "default (TMA inverse rows, register-staged forward rows and columns)" X=1
run "+ TMA columns: filter application only" NEMO_B200_SFFT_COLS_TMA=apply
run "+ TMA columns: plain / window only" NEMO_B200_SFFT_COLS_TMA=plain
run "all register-staged" NEMO_B200_SFFT_NO_TMA=1
for v in "$@"; do run "variant $v" NEMO_B200_LIB=build_exp/lib_$v.so; done
cat $out
