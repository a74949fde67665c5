# usage: bash tools/ab.sh [variant ...]   ("base" = the in-tree library); prints the stage times of the last 2 of 4 steps
for v in "$@"; do
  if [ "$v" != "base" ]; then export DBG_B200_LIB=$PWD/rust_debruijn_b200/variants/libdbg_$v.so; else unset DBG_B200_LIB; fi
  echo "== variant [$v]"; python tools/profile_step.py --steps 4 $AB_ARGS 2>&1 | tail -n 2 | python -c "
import sys,ast
for l in sys.stdin:
    d=ast.literal_eval(l.strip()); print({k:d[k] for k in ('ms_k_partition','ms_k_count','ms_count','ms_sort','ms_filter_total','ms_links','ms_rank','ms_emit','ms_compress_total','n_bucket_splits','bucket_bits','msp_p','n_records','n_distinct','n_valid')})"
done
