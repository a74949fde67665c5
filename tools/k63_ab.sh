export DBG_B200_LIB=$PWD/rust_debruijn_b200/variants/libdbg_w2half.so
echo "== K=63 w2half occ 3300"; python tools/profile_step.py --steps 3 --k 63 --bucket-occ 3300 2>&1 | tail -n 1 | cut -c1-900
unset DBG_B200_LIB
echo "== K=63 base dedup 2"; python tools/profile_step.py --steps 3 --k 63 --dedup 2 2>&1 | tail -n 1 | cut -c1-900
