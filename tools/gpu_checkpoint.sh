#!/bin/bash
# One GPU-box visit: GPU test suite, default bench with the per-kernel table, the CPU arm, the ncu launch list of the same bench
# command, and a --set full capture of the DR5 step's kernels (exported to csv on the box; the report itself only comes back
# when it is small -- gpurun drops the whole directory above 64 MiB).  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash tools/gpu_checkpoint.sh r2f'
tag=${1:-ck}
what=${2:-tests,bench,ref,launches,full}
out=gpurun_out
mkdir -p $out
if [[ $what == *tests* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_gputests.log 2>&1; echo "tests rc $?"
  tail -3 $out/${tag}_gputests.log
fi
if [[ $what == *bench* ]]; then
  timeout 600 python bench.py --steps 10 --warmup 3 --kernel-table $out/${tag}_kernels_dr5.json > $out/${tag}_bench.log 2> $out/${tag}_bench.err; echo "bench rc $?"
fi
if [[ $what == *ref* ]]; then
  timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $out/${tag}_bench_reference.log 2>> $out/${tag}_bench.err
fi
if [[ $what == *launches* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 800 --csv --log-file $out/${tag}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-e2e > $out/${tag}_ncu_launches.log 2>&1
fi
if [[ $what == *full* ]]; then
  rep=/tmp/${tag}_dr5_full
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-sfft_|build_filter|cell_clip|plane_values|edge_cols}" -s ${NCU_SKIP:-60} -c ${NCU_COUNT:-16} -o $rep -f \
      python tools/time_config.py --shape 840 1440 --tiles 8 --freqs 2 --scales 16 --steps 1 > $out/${tag}_ncu_full.log 2>&1
  ncu -i $rep.ncu-rep --page raw --csv > $out/${tag}_dr5_full_raw.csv 2>/dev/null
  ncu -i $rep.ncu-rep --page source --csv > $out/${tag}_dr5_full_source.csv 2>/dev/null
  sz=$(stat -c %s $rep.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 0 ] && [ "$sz" -lt 30000000 ]; then cp $rep.ncu-rep $out/; fi
fi
if [[ $what == *rsprof* ]]; then
  rep=/tmp/${tag}_realspace_full
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv2d|gauss_" -s 4 -c 6 -o $rep -f \
      python tools/time_realspace_batched.py --tiles 16 --steps 1 > $out/${tag}_ncu_realspace.log 2>&1
  ncu -i $rep.ncu-rep --page raw --csv > $out/${tag}_realspace_full_raw.csv 2>/dev/null
  ncu -i $rep.ncu-rep --page source --csv > $out/${tag}_realspace_full_source.csv 2>/dev/null
fi
du -sh $out; ls -la $out | grep ${tag}_
