run() {
env $1 timeout 300 python bench.py --steps 20 --warmup 5 --no-training --no-cpu-baseline --no-extras > gpurun_out/exp.json 2> gpurun_out/exp.err || tail -3 gpurun_out/exp.err
python - <<PY
import json
d=json.load(open("gpurun_out/exp.json"))
print("$1", "fps", round(d["value"]), "norm ms", round(d["roofline_hbm_passes"]["cin_apply_bf16"]["ms_per_step"],4), "trunk us", round(d["roofline"]["avg_launch_ms"]*1e3,2), "checksum", d["checksum"])
PY
}
run "RST_L2_HINTS=13"
run "RST_L2_HINTS=15"
run "RST_L2_HINTS=13"
run "RST_L2_HINTS=15"
run "RST_L2_HINTS=13"
run "RST_L2_HINTS=15"
