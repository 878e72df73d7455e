#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, ncu --set full of the top kernels.
# Usage (under gpurun): [BATCH=8] bash tools/gpu_round.sh <tag> [tests|smoke|bench|benchref|exp|in2|shapes|launches|full ...]
tag=${1:-r01}; shift
what=${*:-tests bench launches full}
B=${BATCH:-1}
out=gpurun_out/$tag
mkdir -p $out
# parity subset used to validate an opt-in variant (whole iterations incl. CUDA-graph replay, every operator case, the
# pair kernel); FULLSUITE=1 runs the whole -m gpu suite under each variant instead
if [ "${FULLSUITE:-0}" = "1" ]; then CORE="tests"; else CORE="tests/test_ops_gpu.py tests/test_cyclegan_gpu.py tests/test_pair_gpu.py tests/test_pix2pix_gpu.py"; fi
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/gpu.txt 2>&1
python -c "import os;print('cpus',os.cpu_count())" >> $out/gpu.txt
for w in $what; do
case $w in
tests)
  timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
  tail -5 $out/pytest_gpu.log ;;
smoke)
  timeout 600 python __graft_entry__.py smoke > $out/smoke.log 2>&1; tail -3 $out/smoke.log ;;
bench)
  timeout 900 python bench.py --batch $B > $out/bench_b$B.json 2> $out/bench_b$B.err; tail -c 3500 $out/bench_b$B.json
  tail -3 $out/bench_b$B.err ;;
exp)
  # opt-in kernel paths that have not run on a B200 yet (cta_group::2 convolution, gradient-side pixel windows):
  # every case in its own process under a timeout (tests/test_cg2_gpu.py), then the A/B microbench and a bench line
  GB_EXPERIMENTAL=1 timeout 1500 python -m pytest tests/test_cg2_gpu.py tests/test_strided_window_gpu.py tests/test_random_conv_gpu.py -m gpu -q \
     > $out/pytest_exp.log 2>&1; echo "pytest exit $?" >> $out/pytest_exp.log; tail -15 $out/pytest_exp.log
  timeout 600 python tools/conv_microbench.py 8 --layers 0,1,2,3,4,5 --variants 2,7,8,9,10 --what fwd,dgrad > $out/microbench_cg2.txt 2>&1
  tail -40 $out/microbench_cg2.txt
  GB_KNOBS=16=1 timeout 900 python bench.py --batch $B --no-cpu-baseline > $out/bench_cg2_b$B.json 2> $out/bench_cg2_b$B.err
  tail -c 1500 $out/bench_cg2_b$B.json; tail -3 $out/bench_cg2_b$B.err
  GB_KNOBS=16=2 timeout 900 python bench.py --batch $B --no-cpu-baseline > $out/bench_persist_b$B.json 2> $out/bench_persist_b$B.err
  tail -c 1500 $out/bench_persist_b$B.json; tail -3 $out/bench_persist_b$B.err
  # programmatic dependent launch (knob 20): the whole parity suite under it, then a bench line
  GB_KNOBS=20=1 timeout 1500 python -m pytest $CORE -m gpu -q -x > $out/pytest_pdl.log 2>&1; echo "pytest exit $?" >> $out/pytest_pdl.log
  tail -4 $out/pytest_pdl.log
  GB_KNOBS=20=1 timeout 900 python bench.py --batch $B --no-cpu-baseline --no-roofline > $out/bench_pdl_b$B.json 2> $out/bench_pdl_b$B.err
  tail -c 600 $out/bench_pdl_b$B.json; tail -3 $out/bench_pdl_b$B.err
  GB_BWD_WINDOW=1 timeout 900 python bench.py --batch $B --no-cpu-baseline --no-roofline > $out/bench_bwdwin_b$B.json 2> $out/bench_bwdwin_b$B.err
  tail -c 600 $out/bench_bwdwin_b$B.json; tail -3 $out/bench_bwdwin_b$B.err
  # parameter gradients accumulated by the kernels into param.grad (no AccumulateGrad add launches)
  GB_DIRECT_PARAM_GRAD=1 timeout 1500 python -m pytest $CORE -m gpu -q -x > $out/pytest_direct.log 2>&1; echo "pytest exit $?" >> $out/pytest_direct.log
  tail -4 $out/pytest_direct.log
  GB_DIRECT_PARAM_GRAD=1 timeout 900 python bench.py --batch $B --no-cpu-baseline --no-roofline > $out/bench_direct_b$B.json 2> $out/bench_direct_b$B.err
  tail -c 600 $out/bench_direct_b$B.json; tail -3 $out/bench_direct_b$B.err
  # second-generation weight pack / weight-gradient unpack (32-bit index math, 16-byte stores)
  GB_KNOBS=28=1 timeout 1500 python -m pytest $CORE -m gpu -q -x > $out/pytest_pack2.log 2>&1; echo "pytest exit $?" >> $out/pytest_pack2.log
  tail -4 $out/pytest_pack2.log
  GB_KNOBS=28=1 timeout 900 python bench.py --batch $B --no-cpu-baseline > $out/bench_pack2_b$B.json 2> $out/bench_pack2_b$B.err
  python - "$out/bench_pack2_b$B.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("pack2:", round(d["value"], 1), d["unit"], {k: (v["avg_us"], v["launches_per_step"]) for k, v in d.get("roofline_detail", {}).items() if k in ("pack", "unpack")})
except Exception as e:
    print("no bench line:", e)
PY
  # e2e leg with input prefetch on a copy stream and lagged loss read-back (train.input_prefetch)
  timeout 900 python bench.py --batch $B --e2e-pipeline --no-cpu-baseline --no-roofline > $out/bench_e2epipe_b$B.json 2> $out/bench_e2epipe_b$B.err
  tail -c 700 $out/bench_e2epipe_b$B.json; tail -3 $out/bench_e2epipe_b$B.err ;;
in2)
  # second-generation InstanceNorm backward (knob 22; tests/test_in_bwd_v2_emul.py already runs its thread body on the CPU):
  # parity per variant in its own process, A/B table, then the whole suite and a bench line with it as the default
  GB_EXPERIMENTAL=1 timeout 1500 python -m pytest tests/test_in_bwd_v2_gpu.py -m gpu -q -s > $out/pytest_in2.log 2>&1; echo "pytest exit $?" >> $out/pytest_in2.log
  grep -E "OK  |FAIL|passed|failed|exit" $out/pytest_in2.log | tail -60
  timeout 900 python tools/in_microbench.py 8 > $out/in_microbench_b8.txt 2>&1; tail -45 $out/in_microbench_b8.txt
  GB_KNOBS=22=1 timeout 1500 python -m pytest $CORE -m gpu -q -x > $out/pytest_in2_suite.log 2>&1; echo "pytest exit $?" >> $out/pytest_in2_suite.log
  tail -4 $out/pytest_in2_suite.log
  GB_KNOBS=22=1 timeout 900 python bench.py --batch $B --no-cpu-baseline --no-roofline > $out/bench_in2_b$B.json 2> $out/bench_in2_b$B.err
  tail -c 600 $out/bench_in2_b$B.json; tail -3 $out/bench_in2_b$B.err
  # on-chip cluster kernel (knob 24) for maps of <= 8192 pixels, second generation backward (22) and forward (26) for the rest
  GB_KNOBS=24=1,22=1,26=1 timeout 1500 python -m pytest $CORE -m gpu -q -x > $out/pytest_in3_suite.log 2>&1; echo "pytest exit $?" >> $out/pytest_in3_suite.log
  tail -4 $out/pytest_in3_suite.log
  GB_KNOBS=24=1,22=1,26=1 timeout 900 python bench.py --batch $B --no-cpu-baseline > $out/bench_in3_b$B.json 2> $out/bench_in3_b$B.err
  tail -c 2500 $out/bench_in3_b$B.json; tail -3 $out/bench_in3_b$B.err ;;
shapes)
  # throughput of every BASELINE.json configuration (one short bench line each; not the headline number)
  for wl in pix2pix_resnet pix2pix_unet cut cyclegan3d revgan3d revgan_piresnet3d; do
    # (with the per-kernel-family breakdown: roofline_detail says where each workload's time goes)
    timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 > $out/bench_$wl.json 2> $out/bench_$wl.err
    python - "$out/bench_$wl.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["metric"], round(d["value"], 2), d["unit"], "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2))
    for k, v in sorted(d.get("roofline_detail", {}).items(), key=lambda kv: -kv[1]["share_of_kernel_time"]):
        print("   %-12s share %.3f  %8.1f %s (%.2f of peak)  %d launches/step avg %.1f us" % (k, v["share_of_kernel_time"], v["achieved"], v["unit"], v["frac"], v["launches_per_step"], v["avg_us"]))
except Exception as e:
    print("no bench line:", e)
PY
    tail -2 $out/bench_$wl.err
  done
  # CUT as replayed CUDA-graph segments (capture path written without a GPU: first run)
  timeout 900 python bench.py --workload cut --graph --steps 5 --warmup 3 --no-roofline > $out/bench_cut_graph.json 2> $out/bench_cut_graph.err
  tail -c 700 $out/bench_cut_graph.json; tail -3 $out/bench_cut_graph.err ;;
benchref)
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 --batch $B > $out/bench_ref_b$B.json 2>> $out/bench_b$B.err; cat $out/bench_ref_b$B.json ;;
launches)
  # (WORKLOAD=cut ... lists another configuration's step; BATCH then is that workload's batch)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $out/launches_b$B.csv \
     python bench.py --workload ${WORKLOAD:-cyclegan2d} --batch $B --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-roofline --no-batch1 --single-stream > $out/launches_run.log 2>&1
  python tools/launch_summary.py $out/launches_b$B.csv 5 --md > $out/launch_summary_b$B.md 2>&1; head -42 $out/launch_summary_b$B.md ;;
full)
  # FULLSPECS="regex:skip:count ..." -- one ncu --set full capture per spec (kept small: reports come home)
  i=0
  for spec in ${FULLSPECS:-"igemm|in_fwd|in_bwd:300:12"}; do
    i=$((i+1))
    rx=${spec%%:*}; rest=${spec#*:}; sk=${rest%%:*}; ct=${rest#*:}
    rep=$out/full${i}_b$B
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip $sk -c $ct -o $rep -f \
       python bench.py --batch $B --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-roofline --no-batch1 --single-stream > $out/full_run.log 2>&1
    python tools/ncu_summary.py $rep.ncu-rep > $rep.md 2>&1
    sz=$(stat -c %s $rep.ncu-rep 2>/dev/null || echo 0)
    if [ "$sz" -gt 30000000 ]; then rm -f $rep.ncu-rep; echo "ncu-rep too large ($sz), removed" ; fi
    grep -E "^###|duration|DRAM read|DRAM write|DRAM throughput|tensor pipe|occupancy|stall" $rep.md | head -60
  done ;;
esac
done
python tools/summarize_round.py $out > $out/SUMMARY.txt 2>&1; cat $out/SUMMARY.txt
du -sh gpurun_out
