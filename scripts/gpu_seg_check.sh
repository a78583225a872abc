#!/bin/bash
# One bounded GPU session for the cp.async.bulk/mbarrier variants of the config-3 probe (gx_k_runjoin_seg: join table
# streamed through a ring; gx_k_runjoin_tma: outer rows delivered by the copy engine): parity, A/B timing, then — only
# for a variant that is both correct and faster — the bench line, the whole GPU suite under its switch and one ncu
# capture.  Every step has its own timeout and writes under gpurun_out/, most important first, so a session that is
# cut short still leaves the earlier results.
#   gpurun --timeout 540 -- 'bash scripts/gpu_seg_check.sh'
set +e
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/seg_steps.log; }

stamp "1 parity of both variants vs the oracle"
timeout 200 python -m pytest tests/test_gpu_runjoin_seg.py -q --maxfail 6 > $O/seg_parity.log 2>&1
stamp "  rc=$? $(tail -1 $O/seg_parity.log)"

stamp "2 A/B at SF100"
timeout 150 python scripts/profile_shapes.py --sf 100 --iters 10 --shapes config3 \
    --envs "GX_RUNJOIN_SEG=0;GX_RUNJOIN_TMA=1;GX_RUNJOIN_TMA=2;GX_RUNJOIN_SEG=1;GX_RUNJOIN_SEG=0;GX_RUNJOIN_TMA=1;GX_RUNJOIN_TMA=2" \
    > $O/seg_ab.log 2>&1
stamp "  rc=$?"; tail -8 $O/seg_ab.log | tee -a $O/seg_steps.log
# a variant wins when its tests passed and its best probe_agg time beats the best time of the stock kernel by 2 %
WIN=$(python - <<'E'
import re
best = {}
for line in open("gpurun_out/seg_ab.log"):
    m = re.match(r"config3 \[(.*?)\]: .*'probe_agg': ([0-9.]+)", line)
    if m:
        v = "tma" if "GX_RUNJOIN_TMA=1" in m.group(1) else "tma2" if "GX_RUNJOIN_TMA=2" in m.group(1) else ("seg" if "GX_RUNJOIN_SEG=1" in m.group(1) else "base")
        best[v] = min(best.get(v, 1e9), float(m.group(2)))
log = open("gpurun_out/seg_parity.log").read()
ok = {v: (("[%s-" % v) not in "".join(l for l in log.splitlines(True) if l.startswith("FAILED") or l.startswith("ERROR"))) and " passed" in log
      for v in ("seg", "tma", "tma2")}
cands = [(best[v], v) for v in ("seg", "tma", "tma2") if v in best and "base" in best and ok[v] and best[v] < 0.98 * best["base"]]
print(min(cands)[1] if cands else "none")
E
)
stamp "  winner: $WIN"
if [ "$WIN" = "none" ]; then
    stamp "no variant is both correct and faster: one ncu capture of gx_k_runjoin_tma for the record, then stop"
    GX_RUNJOIN_TMA=2 timeout 150 ncu --set full --clock-control none -k regex:gx_k_runjoin_tma -c 1 -f \
        -o $O/r02_tma_ncu python scripts/ncu_probe.py 100 1 > $O/tma_ncu.log 2>&1
    timeout 60 ncu -i $O/r02_tma_ncu.ncu-rep --page raw --csv > $O/r02_tma_ncu_raw.csv 2>> $O/tma_ncu.log
    stamp "done (off)"
    exit 0
fi
if [ "$WIN" = "tma" ]; then export GX_RUNJOIN_TMA=1; KRE=gx_k_runjoin_tma; elif [ "$WIN" = "tma2" ]; then export GX_RUNJOIN_TMA=2; KRE=gx_k_runjoin_tma; else export GX_RUNJOIN_SEG=1; KRE=gx_k_runjoin_seg; fi

stamp "3 bench line with $WIN on"
timeout 240 python bench.py --steps 20 --warmup 5 > $O/${WIN}_bench.json 2> $O/${WIN}_bench.err
stamp "  rc=$? $(head -c 200 $O/${WIN}_bench.json)"

stamp "4 the whole GPU suite with $WIN on"
timeout 240 python -m pytest tests -x -q -m gpu > $O/${WIN}_suite.log 2>&1
stamp "  rc=$? $(tail -1 $O/${WIN}_suite.log)"

stamp "5 ncu --set full of the winner"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:$KRE -c 1 -f \
    -o $O/r02_${WIN}_ncu python scripts/ncu_probe.py 100 1 > $O/${WIN}_ncu.log 2>&1
stamp "  rc=$?"
timeout 60 ncu -i $O/r02_${WIN}_ncu.ncu-rep --page raw --csv > $O/r02_${WIN}_ncu_raw.csv 2>> $O/${WIN}_ncu.log
stamp "done ($WIN on)"
exit 0
