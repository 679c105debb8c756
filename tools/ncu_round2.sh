#!/usr/bin/env bash
# Round-2 ncu captures (run on the GPU box through gpurun; reports land in gpurun_out/).
#   launch list of a short bench run, then one `--set full` capture of each hot kernel: K3 (32 fields at 4096^2), K4 over
#   the compact layer (float32 filter + exact list kernel), the K1 kernels, the pipelined planner's two passes on cfg5.
#   SKIP=k3,k1 skips captures (comma list of k3 k4 k1 plan list).
set -u
skip() { case ",${SKIP:-}," in *",$1,"*) return 0;; esac; return 1; }
mkdir -p gpurun_out
skip list || { python bench.py --steps 2 --warmup 3 --no-extras --quick > gpurun_out/r2_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --quick > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list rc=$?"; }
skip k3 || ncu --set full --clock-control none --import-source on -k regex:heur_warp_kernel -c 1 -o gpurun_out/r2_k3_heur \
    python tools/heur_tune.py 4096 32 > gpurun_out/r2_ncu_k3.log 2>&1
echo "k3 rc=$?"
skip k4 || ncu --set full --clock-control none --import-source on -k regex:"collide_filter_kernel|collide_u8_list_kernel" -s 4 -c 2 -o gpurun_out/r2_k4_filter \
    python tools/collide_probe.py 1 > gpurun_out/r2_ncu_k4.log 2>&1
echo "k4 rc=$?"
skip k1 || ncu --set full --clock-control none --import-source on -k regex:"edt_|build_cls" -s 8 -c 5 -o gpurun_out/r2_k1_layers \
    python tools/layers_probe.py 4096 1 > gpurun_out/r2_ncu_k1.log 2>&1
echo "k1 rc=$?"
skip plan || ncu --set full --clock-control none --import-source on -k regex:plan_pipe_kernel -c 2 -o gpurun_out/r2_plan_pipe_cfg5 \
    python tools/cfg5_probe.py > gpurun_out/r2_ncu_plan.log 2>&1
echo "plan rc=$?"
ls -la gpurun_out/*.ncu-rep
