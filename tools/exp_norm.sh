run() {
env $1 timeout 300 python bench.py --steps 20 --warmup 5 --no-training --no-cpu-baseline --no-extras > gpurun_out/exp.json 2> gpurun_out/exp.err || tail -3 gpurun_out/exp.err
python - <<PY
import json
d=json.load(open("gpurun_out/exp.json"))
print("$1", "fps", round(d["value"]), "checksum", d["checksum"])
PY
}
for lib in pass conv; do
cp gpurun_out_lib_$lib.so realtime_style_transfer_b200/csrc/librst_sm100.so
echo "== trigger in $lib kernels"
run "RST_PDL=0"
run "RST_PDL=1"
run "RST_PDL=0"
run "RST_PDL=1"
done
