#!/bin/bash
# Round-2 evidence session: ViT-Lens recipe benches (configs 2-4), ncu --set full of the new attention backward, launch list of one
# headline step.  Outputs -> gpurun_out/.
tag=${1:-q}
mkdir -p gpurun_out
for c in 2 3 4; do
  timeout 600 python bench.py --config $c --steps 4 --no-cpu-baseline > gpurun_out/bench_${tag}_cfg$c.json 2> gpurun_out/bench_${tag}_cfg$c.err; echo "cfg$c rc=$?"
done
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:attn_bwd3 -s 2 -c 1 -f -o gpurun_out/${tag}_attn_bwd3 python tools/ncu_one.py attn_bwd > gpurun_out/${tag}_attn_bwd3.log 2>&1; echo "ncu attn_bwd3 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 900 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/${tag}_* gpurun_out/bench_${tag}_* | head -20
true
