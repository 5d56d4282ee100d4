#!/bin/bash
# One GPU lease, several independent stages, each bounded by `timeout` so that a hanging kernel costs one stage and
# not the call.  Everything lands in gpurun_out/.
# usage: scripts/gpu_session.sh [stage ...]   (default: all)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STAGES=${@:-"ubench light tests bench ncu sanitizer"}
OUT=gpurun_out
for st in $STAGES; do
  echo "=== stage $st $(date +%T) ==="
  case $st in
    ubench)
      # build here: (cd scripts/ubench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../baybe_b200/csrc -I../../include mma_rate.cu -o mma_rate)
      timeout 100 scripts/ubench/mma_rate > $OUT/mma_rate_cg1.txt 2>&1; echo "rc=$?" >> $OUT/mma_rate_cg1.txt
      grep -E "elect=1|rc=" $OUT/mma_rate_cg1.txt | grep -E "bg=0|rc=" ;;
    light)
      timeout 200 python scripts/ts_first_light.py > $OUT/ts_first_light.txt 2>&1; echo "light rc=$?"
      tail -45 $OUT/ts_first_light.txt ;;
    tests)
      FK=""  # only when the first-light stage ran IN THIS SESSION and failed, fall back to the fused_tc kernels
      if [ -f $OUT/ts_first_light.txt ] && ! grep -q FIRST_LIGHT_DONE $OUT/ts_first_light.txt; then FK="tc"; echo "TS kernel not healthy: tests run with BB_FORCE_KERNEL=tc"; fi
      [ -n "$BB_FORCE_KERNEL" ] && FK="$BB_FORCE_KERNEL"
      BB_FORCE_KERNEL=$FK timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread -p no:cacheprovider > $OUT/pytest_gpu.txt 2>&1; echo "tests rc=$?"
      tail -60 $OUT/pytest_gpu.txt ;;
    bench)
      timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?"
      cat $OUT/bench_n1.json | cut -c1-3000; tail -5 $OUT/bench_n1.err ;;
    benchtc)
      BB_FORCE_KERNEL=tc timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_n1_tc.json 2> $OUT/bench_n1_tc.err; echo "bench rc=$?"
      cat $OUT/bench_n1_tc.json | cut -c1-1500 ;;
    ncu)
      BB_OVERLAPPED_HOST_PASS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1; echo "ncu list rc=$?"
      timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused_ts -s 3 -c 1 -f -o $OUT/prof_fused_ts_r02 python scripts/profile_target.py fused > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
      tail -3 $OUT/ncu_full.log ;;
    sanitizer)
      timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/ts_first_light.py n64_d4 cfg1 task4 > $OUT/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"
      tail -8 $OUT/r02_sanitizer_memcheck.txt
      timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/ts_first_light.py n64_d4 cfg1 > $OUT/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"
      tail -8 $OUT/r02_sanitizer_racecheck.txt ;;
    ab)   # same-box A/B of library builds (scripts/build_variant.sh), two rounds to see drift
      for r in 1 2; do for v in $(ls baybe_b200/_C/variants/ | sed 's/.so$//') ""; do BB_LIB_VARIANT=$v timeout 120 python scripts/ab_kernel.py 2>&1 | grep -E "variant|Error" | tee -a $OUT/ab_kernel.txt; done; done ;;
    e2e)
      for m in 1 0; do BB_GATE_PUBLISH=$m timeout 200 python scripts/time_e2e.py 1 > $OUT/time_e2e_m$m.txt 2>&1; echo "e2e mode $m rc=$?"; grep -v Warn $OUT/time_e2e_m$m.txt | tail -7; done
      timeout 200 python scripts/time_e2e.py 0 > $OUT/time_e2e_blocks.txt 2>&1; echo "e2e blocks rc=$?"; tail -7 $OUT/time_e2e_blocks.txt ;;
    hybrid)
      timeout 600 python -m pytest tests/test_gpu_zz_hybrid.py -m gpu -q --timeout 300 -p no:cacheprovider > $OUT/pytest_gpu_hybrid.txt 2>&1; echo "hybrid rc=$?"
      tail -30 $OUT/pytest_gpu_hybrid.txt ;;
    campaign)
      timeout 600 python -m pytest tests/test_gpu_campaign.py -m gpu -q --timeout 300 -p no:cacheprovider > $OUT/pytest_gpu_campaign.txt 2>&1; echo "campaign rc=$?"
      tail -15 $OUT/pytest_gpu_campaign.txt ;;
    trace)
      timeout 200 python scripts/trace_timeline.py score ts > $OUT/r02_pipeline_trace_fused_ts.txt 2>&1; echo "trace rc=$?"
      head -130 $OUT/r02_pipeline_trace_fused_ts.txt ;;
    kmat)
      timeout 300 python scripts/time_kmat.py > $OUT/time_kmat.txt 2>&1; echo "kmat rc=$?"; cat $OUT/time_kmat.txt | tail -8
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_kmat_ts -s 1 -c 1 -f -o $OUT/prof_kmat_ts_r02 python scripts/profile_target.py kmat > $OUT/ncu_kmat.log 2>&1; echo "ncu kmat rc=$?"
      tail -2 $OUT/ncu_kmat.log ;;
    bench45)
      timeout 600 python bench.py --config 5 --steps 10 --warmup 3 > $OUT/bench_cfg5_n1.json 2> $OUT/bench_cfg5_n1.err; echo "cfg5 rc=$?"; cat $OUT/bench_cfg5_n1.json | cut -c1-1800; tail -3 $OUT/bench_cfg5_n1.err
      timeout 900 python bench.py --config 4 --steps 5 --warmup 3 > $OUT/bench_cfg4_n1.json 2> $OUT/bench_cfg4_n1.err; echo "cfg4 rc=$?"; cat $OUT/bench_cfg4_n1.json | cut -c1-1800; tail -3 $OUT/bench_cfg4_n1.err ;;
    bench2)
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench2 rc=$?"
      cat $OUT/bench_n2.json | cut -c1-3000; tail -15 $OUT/bench_n2.err ;;
    benchN)   # NG GPUs (gpurun --gpus NG): headline weak + strong scaling line
      timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus ${NG:-2} --steps 20 --warmup 3 > $OUT/bench_n${NG:-2}.json 2> $OUT/bench_n${NG:-2}.err; echo "benchN rc=$?"
      cat $OUT/bench_n${NG:-2}.json | cut -c1-2500; tail -8 $OUT/bench_n${NG:-2}.err ;;
    cfg45N)   # BASELINE configs 4 (10M fingerprints sharded) and 5 (4 x 250k, top-k merge) on NG GPUs
      timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29523 bench.py --config 5 --gpus ${NG:-2} --steps 10 --warmup 3 > $OUT/bench_cfg5_n${NG:-2}.json 2> $OUT/bench_cfg5_n${NG:-2}.err; echo "cfg5 rc=$?"; cat $OUT/bench_cfg5_n${NG:-2}.json | cut -c1-1500; tail -3 $OUT/bench_cfg5_n${NG:-2}.err
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29525 bench.py --config 4 --gpus ${NG:-2} --steps 5 --warmup 3 > $OUT/bench_cfg4_n${NG:-2}.json 2> $OUT/bench_cfg4_n${NG:-2}.err; echo "cfg4 rc=$?"; cat $OUT/bench_cfg4_n${NG:-2}.json | cut -c1-1500; tail -3 $OUT/bench_cfg4_n${NG:-2}.err ;;
    testsN)
      timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29527 scripts/multi_gpu_check.py > $OUT/multi_gpu_check_n${NG:-2}.txt 2>&1; echo "testsN rc=$?"
      tail -12 $OUT/multi_gpu_check_n${NG:-2}.txt ;;
    tests2)
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 scripts/multi_gpu_check.py > $OUT/multi_gpu_check.txt 2>&1; echo "tests2 rc=$?"
      tail -30 $OUT/multi_gpu_check.txt ;;
  esac
done
echo "=== done $(date +%T) ==="
