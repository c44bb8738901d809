#!/bin/bash
# One gpurun call = one invocation of this script with a list of steps, e.g.
#   gpurun --timeout 1500 -- 'bash tools/gpu_run.sh tests bench snake flow5 refcuda'
# Every step runs under its own `timeout`, logs into gpurun_out/ and never aborts the following steps.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for step in "$@"; do
  case "$step" in
    build)   python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 ;;
    smoke)   timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" ;;
    snaketest) timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "snake" --timeout 300 > gpurun_out/test_snake.log 2>&1; echo "snaketest rc=$?"; tail -5 gpurun_out/test_snake.log ;;
    tests)   timeout 1500 python -m pytest tests -q -m gpu -s --timeout 600 > gpurun_out/test_gpu.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|error" gpurun_out/test_gpu.log | tail -3 ;;
    testsx)  timeout 1500 python -m pytest tests -q -m gpu -x -s --timeout 600 -k "${SVB_K:-}" > gpurun_out/test_gpu_k.log 2>&1; echo "testsx rc=$?"; tail -15 gpurun_out/test_gpu_k.log ;;
    bench)   timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cat gpurun_out/bench_default.json ;;
    benchq)  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "benchq rc=$?"; python tools/bench_brief.py gpurun_out/bench_quick.json ;;
    benchq0) SVB_RB_PERSIST=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick0.json 2> gpurun_out/bench_quick0.err; echo "benchq0 rc=$?"; python tools/bench_brief.py gpurun_out/bench_quick0.json ;;
    ab)      # A/B of environment switches: SVB_AB="VAR=1;VAR2=0 VAR3=5" -> one quick bench per ';'-separated setting
             IFS=';' read -ra SETS <<< "${SVB_AB:-}"
             for e in "${SETS[@]}"; do
               tag=$(echo "$e" | tr -c 'A-Za-z0-9=_\n' '_')
               env $e timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > "gpurun_out/bench_ab_$tag.json" 2> "gpurun_out/bench_ab_$tag.err"
               echo "ab [$e] rc=$?"; python tools/bench_brief.py "gpurun_out/bench_ab_$tag.json" | head -4
             done ;;
    e2etrace) timeout 300 python tools/e2e_trace.py > gpurun_out/e2e_trace.log 2>&1; echo "e2etrace rc=$?"; cat gpurun_out/e2e_trace.log ;;
    rbskew)  for cfg in "16 3" "16 7" "16 11" "32 3" "32 7" "64 7" "64 3"; do timeout 120 tools/bench_rbskew $cfg; done > gpurun_out/rbskew.log 2>&1; echo "rbskew rc=$?"; grep -E "per launch|period|conv" gpurun_out/rbskew.log | head -60 ;;
    snake)   timeout 600 python bench.py --vocoder nsf-snake-hifigan --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_snake.json 2> gpurun_out/bench_snake.err; echo "snake rc=$?"; cat gpurun_out/bench_snake.json ;;
    flow5)   timeout 300 python bench.py --workload flow5 --steps 20 --warmup 3 > gpurun_out/bench_flow5.json 2> gpurun_out/bench_flow5.err; echo "flow5 rc=$?"; cat gpurun_out/bench_flow5.json ;;
    refcuda) timeout 600 python bench.py --impl reference-cuda --steps 3 --warmup 2 > gpurun_out/bench_refcuda.json 2> gpurun_out/bench_refcuda.err; echo "refcuda rc=$?"; cat gpurun_out/bench_refcuda.json ;;
    refcpu)  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_refcpu.json 2> gpurun_out/bench_refcpu.err; echo "refcpu rc=$?"; cat gpurun_out/bench_refcpu.json ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; echo "launches rc=$?" ;;
    ncu_snake) timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:convn_tc_kernel<\\(int\\)128, \\(int\\)1, \\(int\\)2, \\(bool\\)1>" -s 6 -c 3 -f -o gpurun_out/prof_snake128 python bench.py --vocoder nsf-snake-hifigan --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_snake.log 2>&1; echo "ncu_snake rc=$?" ;;
    ncu_kern) timeout 900 ncu --set full --clock-control none --import-source on -k "regex:resblock_skew_kernel|pair_tc_kernel" -s 7 -c 7 -f -o gpurun_out/prof_kern python tools/prof_kernels.py > gpurun_out/ncu_kern.log 2>&1; echo "ncu_kern rc=$?" ;;
    ncu_ups) for spec in "ups32:\(int\)32, \(int\)2, \(int\)2, \(bool\)0, \(int\)1:0" "ffn2:\(int\)384, \(int\)1, \(int\)1, \(bool\)0, \(int\)0:3" "ups128:\(int\)128, \(int\)2, \(int\)2, \(bool\)0, \(int\)1:0"; do
               tag=${spec%%:*}; rest=${spec#*:}; pat=${rest%:*}; skip=${rest##*:}
               timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:convn_tc_kernel<$pat" -s $skip -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1; echo "ncu_ups $tag rc=$?"
             done ;;
    ncu_main) timeout 900 ncu --set full --clock-control none --import-source on -k "regex:flow_layer_kernel|resblock_skew_kernel" -c 13 -f -o gpurun_out/prof_main python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_main.log 2>&1; echo "ncu_main rc=$?" ;;
    ncu_prefix) timeout 900 ncu --set full --clock-control none --import-source on -k "regex:convn_tc_kernel|attn_rel_kernel" -s 60 -c 9 -f -o gpurun_out/prof_prefix python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_prefix.log 2>&1; echo "ncu_prefix rc=$?" ;;
    *) echo "unknown step $step" ;;
  esac
done
