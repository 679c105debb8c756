#!/usr/bin/env bash
# Round-2 ncu captures (run on the GPU box through gpurun; reports land in gpurun_out/).
#   launch list of a short bench run, then one `--set full` capture of each hot kernel: K3 (32 fields at 4096^2), K4 over
#   the compact layer, the K1 kernels, the planner kernel (one cfg5 map).
set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-extras --quick > gpurun_out/r2_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --quick > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:heur_warp_kernel -c 1 -o gpurun_out/r2_k3_heur \
    python tools/heur_tune.py 4096 32 > gpurun_out/r2_ncu_k3.log 2>&1
echo "k3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:collide_circles_u8_kernel -s 2 -c 1 -o gpurun_out/r2_k4_collide \
    python tools/collide_probe.py 1 > gpurun_out/r2_ncu_k4.log 2>&1
echo "k4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"edt_|build_cls" -s 8 -c 5 -o gpurun_out/r2_k1_layers \
    python tools/layers_probe.py 4096 1 > gpurun_out/r2_ncu_k1.log 2>&1
echo "k1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 1 -c 1 -o gpurun_out/r2_plan_kernel \
    python tools/cfg5_probe.py 1 > gpurun_out/r2_ncu_plan.log 2>&1
echo "plan rc=$?"
ls -la gpurun_out/*.ncu-rep
