#!/bin/bash
# One bounded GPU session for the cp.async.bulk/mbarrier variants of the config-3 probe: parity of every variant,
# A/B timing at SF100, then the bench line, the whole GPU suite, smoke and the launch list with the variant that won
# (the built-in default unless GX_RUNJOIN_TMA=3 beats it by 1 %).  Every step has its own timeout and writes under
# gpurun_out/, most important first, so a session that is cut short still leaves the earlier results.
#   gpurun --timeout 420 -- 'bash scripts/gpu_seg_check.sh'
set +e
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/seg_steps.log; }

stamp "1 parity of the variants vs the oracle"
timeout 200 python -m pytest tests/test_gpu_runjoin_seg.py -q --maxfail 6 > $O/seg_parity.log 2>&1
stamp "  rc=$? $(tail -1 $O/seg_parity.log)"

stamp "2 A/B at SF100"
timeout 150 python scripts/profile_shapes.py --sf 100 --iters 10 --shapes config3 \
    --envs "GX_RUNJOIN_TMA=0;GX_RUNJOIN_TMA=2;GX_RUNJOIN_TMA=3;GX_RUNJOIN_TMA=0;GX_RUNJOIN_TMA=2;GX_RUNJOIN_TMA=3" \
    > $O/seg_ab.log 2>&1
stamp "  rc=$?"; tail -8 $O/seg_ab.log | tee -a $O/seg_steps.log
WIN=$(python - <<'PY'
import re
best = {}
for line in open("gpurun_out/seg_ab.log"):
    m = re.match(r"config3 \[GX_RUNJOIN_TMA=(\d)\]: .*'probe_agg': ([0-9.]+)", line)
    if m:
        best[m.group(1)] = min(best.get(m.group(1), 1e9), float(m.group(2)))
log = open("gpurun_out/seg_parity.log").read()
bad = "".join(l for l in log.splitlines(True) if l.startswith("FAILED") or l.startswith("ERROR"))
ok3 = "[tma3-" not in bad and " passed" in log
print("3" if ("3" in best and "2" in best and ok3 and best["3"] < 0.99 * best["2"]) else "default")
PY
)
stamp "  variant for the rest of the session: $WIN"
if [ "$WIN" = "3" ]; then export GX_RUNJOIN_TMA=3; fi

stamp "3 bench line"
timeout 240 python bench.py --steps 20 --warmup 5 > $O/final_bench.json 2> $O/final_bench.err
stamp "  rc=$? $(head -c 200 $O/final_bench.json)"

stamp "4 the whole GPU suite + smoke"
timeout 240 python -m pytest tests -x -q -m gpu > $O/final_suite.log 2>&1
stamp "  rc=$? $(tail -1 $O/final_suite.log)"
timeout 60 python -c "import __graft_entry__ as e; e.smoke()" > $O/final_smoke.log 2>&1
stamp "  smoke rc=$? $(tail -1 $O/final_smoke.log)"

stamp "5 launch list of one bench command"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/final_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > $O/final_launches_bench.log 2>&1
stamp "  rc=$?"
if [ "$WIN" = "3" ]; then
    stamp "6 ncu --set full of the winner"
    timeout 150 ncu --set full --clock-control none -k regex:gx_k_runjoin_tma -c 1 -f \
        -o $O/r02_tma3_ncu python scripts/ncu_probe.py 100 1 > $O/tma3_ncu.log 2>&1
    timeout 60 ncu -i $O/r02_tma3_ncu.ncu-rep --page raw --csv > $O/r02_tma3_ncu_raw.csv 2>> $O/tma3_ncu.log
fi
stamp "done ($WIN)"
exit 0
