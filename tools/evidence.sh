#!/bin/bash
# Round evidence: bench lines (both arms, config 3), ncu launch list, DRAM bytes, --set full of the top kernels.
# usage (under gpurun): bash tools/evidence.sh r01j [full|nofull]
T=${1:-r01x}; O=gpurun_out
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
python bench.py --impl reference --steps 3 --warmup 3 > $O/${T}_bench_ref.json 2>> $O/${T}_bench.err
python bench.py --workload config3 > $O/${T}_bench_c3.json 2>> $O/${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_ncu_launches_bf16_1024rays.csv python bench.py --ncu-mode > $O/${T}_ncu1.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_ncu_dram_bytes.csv python bench.py --ncu-mode > $O/${T}_ncu2.log 2>&1
if [ "${2:-full}" = full ]; then
ncu --set full --import-source on --clock-control none -k regex:'mlp_bf16_fwd|mlp_bf16_bwd|chain_kernel|density_normals_bwd|encode_bwd|wgrad' -s 48 -c 24 -o $O/${T}_top python bench.py --ncu-mode > $O/${T}_ncu3.log 2>&1
python tools/ncu_summary.py $O/${T}_top.ncu-rep > $O/${T}_ncu_full_all.txt 2>&1
fi
python tools/launch_summary.py $O/${T}_ncu_launches_bf16_1024rays.csv > $O/${T}_launch_shares.txt 2>&1
head -c 600 $O/${T}_bench.json; echo; tail -n 30 $O/${T}_launch_shares.txt
