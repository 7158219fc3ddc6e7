#!/bin/bash
# Round evidence: full GPU test suite, bench lines (both arms), ncu launch list, DRAM bytes, --set full of the top kernels.
# usage (under gpurun): bash tools/evidence.sh r02x [full|nofull]
T=${1:-r02x}; O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest_gpu.log
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref.json 2>> $O/${T}_bench.err
python bench.py --ncu-mode > $O/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_ncu_launches_bf16_1024rays.csv python bench.py --ncu-mode > $O/${T}_ncu1.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_ncu_dram_bytes.csv python bench.py --ncu-mode > $O/${T}_ncu2.log 2>&1
if [ "${2:-full}" = full ]; then
ncu --set full --import-source on --clock-control none -k regex:'mlp_bf16_fwd|mlp_bf16_bwd|chain2_kernel|encode_bwd|wgrad|shader_mid' -s 40 -c 24 -o $O/${T}_top python bench.py --ncu-mode > $O/${T}_ncu3.log 2>&1
python tools/ncu_summary.py $O/${T}_top.ncu-rep > $O/${T}_ncu_full_top_kernels_summary.txt 2>&1
fi
python -c "
import sys; sys.path.insert(0,'tools')
import launch_summary as L
L.main('$O/${T}_ncu_launches_bf16_1024rays.csv', full=True)" > $O/${T}_launch_shares_bf16_1024rays.txt 2>&1
python tools/traffic_from_csv.py $O/${T}_ncu_dram_bytes.csv $T > $O/${T}_traffic.log 2>&1
head -c 400 $O/${T}_bench.json; echo; head -30 $O/${T}_launch_shares_bf16_1024rays.txt
python tools/timeline.py > $O/${T}_timeline.txt 2>/dev/null; head -1 $O/${T}_timeline.txt
