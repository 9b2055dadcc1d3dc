set -x
bash tools/ncu_refresh.sh 2>&1 | tail -n 15
bash tools/sanitize.sh 2>&1 | tail -n 30
