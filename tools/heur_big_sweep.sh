# K3 headline shape (148 fields of 4096^2, one CTA per field): cost-window width sweep
run() { echo "$1: $(env $1 python tools/heur_tune.py 4096 148 2 2>&1 | tail -1 | grep -o 'ms/batch=[0-9.]*')"; }
run "X=0"; for d in 128 160 192 256 384 768; do run "HPP_HEUR_DELTA=$d"; done; run "HPP_HEUR_DELTA=192 HPP_HEUR_DELTA_IN=48"; run "X=1"
