#!/bin/bash
# One gpurun session: probe, GPU tests, smoke, bench, error table, sanitizer.  Every step has its own timeout and log
# under gpurun_out/ so that one failure does not lose the others.  Usage: tools/gpu_session.sh [steps...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STEPS="${@:-probe tests smoke bench errtab}"
for s in $STEPS; do
  echo "=== $s $(date +%T)"
  case $s in
    probe)
      { nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv; nproc;
        python - <<'PY'
import importlib
for m in ("must3r", "croco", "dust3r", "curope", "xformers", "asmk"):
    try:
        importlib.import_module(m); print(m, "IMPORTABLE")
    except Exception as e:
        print(m, "absent:", type(e).__name__, e)
PY
        pip download must3r --no-deps -d /tmp/x 2>&1 | tail -1; find / -iname '*must3r*' -not -path '*/proc/*' 2>/dev/null | grep -v "$PWD" | head; } > gpurun_out/probe.log 2>&1 ;;
    tests)   timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" ;;
    testsall) timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" ;;
    precise) timeout 900 python -m pytest tests/test_gpu_precise.py -m gpu -q -s --timeout 600 > gpurun_out/pytest_precise.log 2>&1; echo "rc=$?" ;;
    smoke)   timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" ;;
    bench)   timeout 900 python bench.py > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err; echo "bench rc=$?" ;;
    bench16) timeout 900 python bench.py --head-precision bf16 --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_v1_bf16head.json 2> gpurun_out/bench_v1_bf16head.err ;;
    benchv2) timeout 900 python bench.py --variant v2 --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err ;;
    benchc3) timeout 900 python bench.py --views 64 --keyframes 8 --no-gpu-reference --no-cpu-baseline --steps 5 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err ;;
    benchref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ;;
    errtab)  timeout 1200 python tools/error_table.py > gpurun_out/errtab.log 2>&1; echo "errtab rc=$?" ;;
    memcheck) timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_precise.py tests/test_gpu_kernels.py -m gpu -q -x -k "not full_size and not 12288 and not 4864" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" ;;
    racecheck) timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "layernorm or mask_helpers or patchify or rope" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" ;;
    launches) timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv python tools/profile_step.py step > gpurun_out/launches_r02.log 2>&1 ;;
    ab)  # A/B on the same box, back to back: this tree (all-bf16 head; optional paths toggled) vs the round-1 tree (_r1/)
      one() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'])"; }
      for cfg in "" ""; do
        echo "--- this tree, bf16 head [$cfg]" >> gpurun_out/ab.log
        env $cfg timeout 300 python bench.py --head-precision bf16 --no-gpu-reference --no-cpu-baseline --no-bf16-head --steps 20 2>/dev/null | one >> gpurun_out/ab.log 2>&1
      done
      if [ -d _r1 ]; then
        for i in 1 2; do
          echo "--- round-1 tree" >> gpurun_out/ab.log
          (cd _r1 && timeout 300 python bench.py --no-cpu-baseline --steps 20 2>/dev/null | one) >> gpurun_out/ab.log 2>&1
        done
      fi
      echo "--- this tree, fp32-grade head" >> gpurun_out/ab.log
      timeout 300 python bench.py --no-gpu-reference --no-cpu-baseline --no-bf16-head --steps 20 2>/dev/null | one >> gpurun_out/ab.log 2>&1 ;;
    ksums) { python tools/kernel_sums.py . bf16; [ -d _r1 ] && python tools/kernel_sums.py _r1; python tools/kernel_sums.py . fp32; } > gpurun_out/kernel_sums.log 2>&1 ;;
    dist2) timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x --timeout 800 > gpurun_out/pytest_dist.log 2>&1; echo "dist rc=$?" ;;
    bench2|bench4|bench8) n=${s#bench}; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err; echo "bench$n rc=$?" ;;
    ncufull) timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r02_key_kernels python tools/profile_step.py kernels > gpurun_out/ncufull.log 2>&1
             ncu -i gpurun_out/r02_key_kernels.ncu-rep --page raw --csv > gpurun_out/r02_key_kernels_raw.csv 2>/dev/null; echo "ncufull rc=$?" ;;
    attnb) timeout 300 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1; [ -d _r1 ] && (cd _r1 && timeout 300 python ../tools/attn_bench.py) > gpurun_out/attn_bench_r1.log 2>&1 ;;
    attntr) timeout 300 python tools/attn_trace.py gpurun_out/attn_trace.md > gpurun_out/attn_trace.log 2>&1; echo "attntr rc=$?" ;;
    attnt) timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" --timeout 300 > gpurun_out/pytest_attn.log 2>&1; echo "attn tests rc=$?"; tail -5 gpurun_out/pytest_attn.log ;;
    memattn) timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention and not 12288" > gpurun_out/sanitizer_memcheck_attention.log 2>&1; echo "memcheck attention rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_attention.log ;;
    gemmt) timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm" --timeout 300 > gpurun_out/pytest_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -4 gpurun_out/pytest_gemm.log ;;
    sk) one() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'])"; }
        for v in 1 0 1 0; do echo "--- PST3R_SPLITK=$v" >> gpurun_out/sk.log
          PST3R_SPLITK=$v timeout 300 python bench.py --head-precision bf16 --no-gpu-reference --no-cpu-baseline --no-bf16-head --steps 20 2>/dev/null | one >> gpurun_out/sk.log 2>&1; done ;;
    stages) timeout 600 python tools/stage_times.py 16 v1 0 > gpurun_out/stages_r02_v1.json 2> gpurun_out/stages_r02.err; echo "stages rc=$?" ;;
    mmarate) timeout 120 tools/_bin/mma_rate > gpurun_out/mma_rate.md 2>&1; echo "mmarate rc=$?" ;;
    lazy) timeout 240 python -m pytest tests/test_lazy_masks.py -m gpu -q -s --timeout 200 > gpurun_out/pytest_lazy.log 2>&1; echo "lazy tests rc=$?"; tail -25 gpurun_out/pytest_lazy.log
          timeout 150 python tools/lazy_masks_bench.py time > gpurun_out/lazy_time.json 2> gpurun_out/lazy_time.err; echo "lazy time rc=$?"; cat gpurun_out/lazy_time.json
          timeout 240 ncu --profile-from-start off --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/lazy_ncu.csv python tools/lazy_masks_bench.py ncu > gpurun_out/lazy_ncu.log 2>&1; echo "lazy ncu rc=$?"
          python tools/lazy_masks_bench.py parse gpurun_out/lazy_ncu.csv gpurun_out/lazy_ncu_segments.json > gpurun_out/lazy_ncu.md 2>&1; cat gpurun_out/lazy_ncu.md ;;
    benchq) timeout 600 python bench.py --no-gpu-reference --no-cpu-baseline --steps 10 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "benchq rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "bf16_head", "ids_only", "clocks")})
PY
      ;;
    lazyfull) timeout 300 ncu --profile-from-start off --set full --cache-control none --clock-control none --import-source on -c 4 -f -o gpurun_out/r02_lazy_kernels python tools/lazy_masks_bench.py ncufull > gpurun_out/lazyfull.log 2>&1; echo "lazyfull rc=$?"
              ncu -i gpurun_out/r02_lazy_kernels.ncu-rep --page raw --csv > gpurun_out/r02_lazy_kernels_raw.csv 2>/dev/null; ls -la gpurun_out/r02_lazy_kernels* ;;
    headab) timeout 300 python tools/stage_times.py 16 v1 head > gpurun_out/head_ab.json 2> gpurun_out/head_ab.err; echo "headab rc=$?"; cat gpurun_out/head_ab.json; tail -3 gpurun_out/head_ab.err ;;
    *) echo "unknown step $s" ;;
  esac
done
echo "=== done $(date +%T)"
tail -5 gpurun_out/pytest.log 2>/dev/null
