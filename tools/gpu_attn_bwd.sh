#!/bin/bash
# attention-backward bring-up session: parity probe (80 checks), timing, clock64 timeline
tag=${1:-a}
mkdir -p gpurun_out
timeout 60 python tools/probe_attn.py bwd > gpurun_out/probe_bwd_$tag.log 2>&1; rc=$?; echo "probe rc=$rc"
if [ $rc -ne 0 ]; then tail -3 gpurun_out/probe_bwd_$tag.log; exit 0; fi
echo "OK: $(grep -c 'OK ' gpurun_out/probe_bwd_$tag.log)  BAD: $(grep -c BAD gpurun_out/probe_bwd_$tag.log)"
grep -B4 BAD gpurun_out/probe_bwd_$tag.log | head -40
timeout 120 python tools/probe_attn.py time > gpurun_out/probe_time_$tag.log 2>&1; echo "time rc=$?"
grep bwd gpurun_out/probe_time_$tag.log
timeout 90 python tools/probe_attn_bwd3_timeline.py > gpurun_out/bwd3_timeline_$tag.log 2>&1; echo "timeline rc=$?"
head -46 gpurun_out/bwd3_timeline_$tag.log; grep -A24 issuer gpurun_out/bwd3_timeline_$tag.log; grep prologue: gpurun_out/bwd3_timeline_$tag.log
