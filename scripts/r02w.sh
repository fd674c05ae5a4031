set -x
python scripts/dump_e2e.py new
B200_TC4_RES_SMEM=0 python scripts/dump_e2e.py noressmem
