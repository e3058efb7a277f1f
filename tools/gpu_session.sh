#!/bin/bash
# One gpurun session: GPU test tier, smoke, the headline bench and the three ViT-Lens recipe benches.  Outputs -> gpurun_out/.
# usage: tools/gpu_session.sh <tag> [steps...]   (steps: test smoke bench cfg2 cfg3 cfg4; default all)
tag=${1:-s}; shift
steps=${@:-test smoke bench cfg2 cfg3 cfg4}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$tag.txt 2>&1
for s in $steps; do
  case $s in
    test)  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" ;;
    testx) timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest -x rc=$?" ;;
    smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?" ;;
    bench) timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?" ;;
    cfg2|cfg3|cfg4) c=${s#cfg}
           timeout 600 python bench.py --config $c --steps 4 --no-cpu-baseline > gpurun_out/bench_${tag}_cfg$c.json 2> gpurun_out/bench_${tag}_cfg$c.err
           rc=$?; echo "cfg$c rc=$rc"
           if [ $rc -ne 0 ]; then
             timeout 600 python bench.py --config $c --steps 4 --batch 256 --no-cpu-baseline > gpurun_out/bench_${tag}_cfg${c}_b256.json 2> gpurun_out/bench_${tag}_cfg${c}_b256.err
             echo "cfg$c batch 256 rc=$?"
           fi ;;
  esac
done
tail -5 gpurun_out/pytest_$tag.log 2>/dev/null
mv gpurun_out/parity_report.jsonl gpurun_out/parity_$tag.jsonl 2>/dev/null
true
