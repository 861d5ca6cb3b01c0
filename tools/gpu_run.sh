#!/bin/bash
# usage: tools/gpu_run.sh <tag> [what...]   (runs on the GPU box through gpurun; outputs under gpurun_out/)
tag=$1; shift
what="$*"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi.txt 2>&1
for w in $what; do
  case $w in
    pytest) timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" ;;
    pytestall) timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" ;;
    fused) timeout 900 python -m pytest tests/test_fused_blocks_gpu.py -q --timeout 600 > gpurun_out/${tag}_fused.log 2>&1; echo "fused rc=$?" ;;
    benchunfused) MS_FUSED_BLOCKS=0 timeout 600 python bench.py --workload train > gpurun_out/${tag}_bench_train_unfused.json 2> gpurun_out/${tag}_bench_train_unfused.err; echo "benchunfused rc=$?" ;;
    bench3) timeout 900 python bench.py --workload config3 --steps 10 > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; echo "bench3 rc=$?" ;;
    phases) MS_PHASE_TS=1 timeout 300 python tools/chain_phases.py > gpurun_out/${tag}_chain_phases.txt 2>&1; MS_PHASE_TS=1 timeout 300 python tools/chain_phases.py --eval > gpurun_out/${tag}_chain_phases_eval.txt 2>&1; echo "phases rc=$?" ;;
    breakdownside) for n in 8 16 24; do echo "MS_SIDE_SMS=$n"; MS_SIDE_SMS=$n timeout 600 python tools/step_breakdown.py 2>&1 | grep "graph replay"; done > gpurun_out/${tag}_breakdown_side.txt 2>&1; echo "breakdownside rc=$?" ;;
    breakdown) timeout 600 python tools/step_breakdown.py --order > gpurun_out/${tag}_breakdown.txt 2>&1; echo "breakdown rc=$?" ;;
    breakdown128) timeout 600 python tools/step_breakdown.py --batch 128 --speakers 8 > gpurun_out/${tag}_breakdown128.txt 2>&1; echo "breakdown128 rc=$?" ;;
    smoke) timeout 600 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" ;;
    bench) timeout 1200 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" ;;
    benchtrain) timeout 600 python bench.py --workload train > gpurun_out/${tag}_bench_train.json 2> gpurun_out/${tag}_bench_train.err; echo "benchtrain rc=$?" ;;
    benchref) timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "benchref rc=$?" ;;
    ncufull) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_chain|wgrad_acc_multi" --launch-skip 60 -c 22 -f -o gpurun_out/${tag}_chain_full python bench.py --workload train --no-graphs --steps 2 --warmup 4 > gpurun_out/${tag}_ncufull.log 2>&1; echo "ncufull rc=$?" ;;
    launchesg) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --graph-profiling node -c 6000 --csv --log-file gpurun_out/${tag}_launches_graph.csv python bench.py --workload train --steps 2 --warmup 4 > gpurun_out/${tag}_launches_graph.log 2>&1; echo "launchesg rc=$?" ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --workload train --no-graphs --steps 2 --warmup 4 > gpurun_out/${tag}_launches.log 2>&1; echo "launches rc=$?" ;;
    *) echo "unknown $w" ;;
  esac
done
ls -la gpurun_out | tail -20
