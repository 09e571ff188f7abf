for cfg in "x x" "2 x" "x k" "2 k"; do set -- $cfg; FAST_B200_RHS=$1 FAST_B200_SENSOR=$2 timeout 200 python tools/dbg_tgv.py 48 2 2>&1 | tail -7; done
