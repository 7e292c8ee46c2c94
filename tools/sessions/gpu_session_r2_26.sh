set -x
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 100 > /dev/null &
SMI=$!
STEPS=40 timeout 600 python tools/e2e_steps.py 5 2>&1 | grep "per-step" | cut -c1-400
kill $SMI
