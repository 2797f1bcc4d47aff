python tools/time_k6.py 2>&1 | tail -3
for t in "$@"; do echo "== $t"; NEMO_B200_LIB=build_exp/lib_$t.so python tools/time_k6.py 2>&1 | tail -4; done
