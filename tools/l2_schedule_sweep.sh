for l in 16 18; do
  for cfg in "8 4" "12 4" "16 4" "16 3" "24 3" "32 2" "48 2"; do
    set -- $cfg
    CFFT_B200_L2_CHUNK_MB=$1 CFFT_B200_L2_STREAMS=$2 timeout 100 python tools/cmp_variants.py $l 9 | sed "s/^/chunk=$1MB streams=$2 /"
  done
done
