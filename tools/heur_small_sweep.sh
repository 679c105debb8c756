# K3 on small maps (the cfg5 field batches): default window against others.  PROBE_MAP=parking selects the cfg5 map class.
run() { echo "$1 side=$2: $(env $1 python tools/heur_tune.py $2 2048 3 2>&1 | tail -1 | grep -o 'ms/batch=[0-9.]*')"; }
for mp in parking cfg3; do export PROBE_MAP=$mp; echo "== $mp"
  for side in 256 512; do run "X=0" $side; run "HPP_HEUR_DELTA=64" $side; run "HPP_HEUR_DELTA=128" $side; run "HPP_HEUR_DELTA=512" $side; done
done
