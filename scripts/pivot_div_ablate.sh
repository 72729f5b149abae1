for cfg in "64 12,6,3" "128 8,4,2" "128 10,5,3" "96 10,5,3"; do set -- $cfg; echo "div=$1 ladder=$2"; MMSIM_PIVOT_DIV=$1 python scripts/ladder_ablate.py 128 $2 2>&1 | grep random; done
