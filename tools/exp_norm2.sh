echo "== register 65536"; tests/cuda/norm_pass_bench
echo "== register 102400"; RST_NORM_PPB=102400 tests/cuda/norm_pass_bench
echo "== register 32768"; RST_NORM_PPB=32768 tests/cuda/norm_pass_bench
echo "== bulk 102400"; RST_NORM_BULK=1 RST_NORM_PPB=102400 tests/cuda/norm_pass_bench
echo "== bulk 51200"; RST_NORM_BULK=1 RST_NORM_PPB=51200 tests/cuda/norm_pass_bench
echo "== bulk 51200 2x2048"; RST_NORM_BULK=1 RST_NORM_PPB=51200 RST_NORM_STAGES=2 RST_NORM_CHUNK=2048 RST_NORM_STAGES_RES=1 tests/cuda/norm_pass_bench
RST_NORM_BULK=1 RST_NORM_PPB=102400 ncu --cache-control none --clock-control none --set full -k regex:cin_apply -s 60 -c 2 -o gpurun_out/norm_bulk_warm -f tests/cuda/norm_pass_bench 128 28800 8 > gpurun_out/ncu_norm.log 2>&1
RST_NORM_PPB=65536 ncu --cache-control none --clock-control none --set full -k regex:cin_apply -s 60 -c 2 -o gpurun_out/norm_reg_warm -f tests/cuda/norm_pass_bench 128 28800 8 >> gpurun_out/ncu_norm.log 2>&1
tail -5 gpurun_out/ncu_norm.log
