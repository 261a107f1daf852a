python -m pytest tests/test_gpu_parity.py -x -q -k "bitwise or edge_cases or mid_scene or known_answer" 2>&1 | tail -2
for opt in "" "6=1"; do
LIDAR_RT_B200_OPTIONS="$opt" python bench.py --no-cpu-baseline --steps 30 > gpurun_out/ab.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/ab.log") if x.startswith("{")][-1]
j=json.loads(l)
print("opt=[$opt]", round(j["value"],2), round(j["ms_per_step"],3), round(j["e2e"]["value"],2), {k:round(v["ms_per_step"],3) for k,v in j["roofline"]["kernels"].items() if v["ms_per_step"]>0.1})
PY
done
