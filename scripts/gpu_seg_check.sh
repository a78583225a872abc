#!/bin/bash
# One bounded GPU session for the streamed-table probe (gx_k_runjoin_seg): parity, A/B timing, then — only if the
# variant is both correct and faster — the bench line, the whole GPU suite under the switch and one ncu capture.
# Every step has its own timeout and writes under gpurun_out/, most important first, so a session that is cut short
# still leaves the earlier results.
#   gpurun --timeout 540 -- 'bash scripts/gpu_seg_check.sh'
set +e
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/seg_steps.log; }

stamp "1 parity of gx_k_runjoin_seg vs the oracle"
GX_RUNJOIN_SEG=1 timeout 170 python -m pytest tests/test_gpu_runjoin_seg.py -x -q > $O/seg_parity.log 2>&1
PARITY=$?
stamp "  rc=$PARITY $(tail -1 $O/seg_parity.log)"

stamp "2 A/B at SF100: gathering kernel vs ring depths 3 (default) / 2 / 4"
timeout 150 python scripts/profile_shapes.py --sf 100 --iters 10 --shapes config3 \
    --envs "GX_RUNJOIN_SEG=0;GX_RUNJOIN_SEG=1;GX_RUNJOIN_SEG=1,GX_RUNJOIN_SEG_BUFS=2;GX_RUNJOIN_SEG=1,GX_RUNJOIN_SEG_BUFS=4;GX_RUNJOIN_SEG=0;GX_RUNJOIN_SEG=1" \
    > $O/seg_ab.log 2>&1
stamp "  rc=$?"; tail -8 $O/seg_ab.log | tee -a $O/seg_steps.log
# faster = the best probe_agg time under the switch beats the best time without it by 2 %
WIN=$(python - <<'E'
import re
best = {}
for line in open("gpurun_out/seg_ab.log"):
    m = re.match(r"config3 \[(.*?)\]: .*'probe_agg': ([0-9.]+)", line)
    if m:
        on = "GX_RUNJOIN_SEG=1" in m.group(1)
        best[on] = min(best.get(on, 1e9), float(m.group(2)))
print(1 if (True in best and False in best and best[True] < 0.98 * best[False]) else 0)
E
)
stamp "  parity rc $PARITY, faster: $WIN"
if [ "$PARITY" != "0" ] || [ "$WIN" != "1" ]; then
    stamp "the variant stays off: one short ncu capture for the record, then stop"
    GX_RUNJOIN_SEG=1 timeout 150 ncu --set full --clock-control none -k regex:gx_k_runjoin_seg -c 1 -f \
        -o $O/r02_seg_ncu python scripts/ncu_probe.py 100 1 > $O/seg_ncu.log 2>&1
    timeout 60 ncu -i $O/r02_seg_ncu.ncu-rep --page raw --csv > $O/r02_seg_ncu_raw.csv 2>> $O/seg_ncu.log
    stamp "done (off)"
    exit 0
fi

stamp "3 bench line with the switch on"
GX_RUNJOIN_SEG=1 timeout 240 python bench.py --steps 20 --warmup 5 > $O/seg_bench.json 2> $O/seg_bench.err
stamp "  rc=$? $(head -c 200 $O/seg_bench.json)"

stamp "4 the whole GPU suite with the switch on"
GX_RUNJOIN_SEG=1 timeout 240 python -m pytest tests -x -q -m gpu > $O/seg_suite.log 2>&1
stamp "  rc=$? $(tail -1 $O/seg_suite.log)"

stamp "5 ncu --set full of the new kernel"
GX_RUNJOIN_SEG=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:gx_k_runjoin_seg -c 1 -f \
    -o $O/r02_seg_ncu python scripts/ncu_probe.py 100 1 > $O/seg_ncu.log 2>&1
stamp "  rc=$?"
timeout 60 ncu -i $O/r02_seg_ncu.ncu-rep --page raw --csv > $O/r02_seg_ncu_raw.csv 2>> $O/seg_ncu.log
stamp "done (on)"
exit 0
