run() {
env $1 timeout 300 python bench.py --steps 20 --warmup 5 --no-training --no-cpu-baseline --no-extras > gpurun_out/exp.json 2> gpurun_out/exp.err || tail -3 gpurun_out/exp.err
python - <<PY
import json
d=json.load(open("gpurun_out/exp.json"))
print("$1", "fps", round(d["value"]), "checksum", d["checksum"])
PY
}
run "RST_PDL=1"
run "RST_PDL=1"
run "RST_PDL=1"
