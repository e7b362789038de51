#!/bin/bash
# Full GPU evidence pass (run on the B200 box through gpurun):  scripts/gpu_evidence.sh <tag>
# Writes under gpurun_out/<tag>_*: pytest log, smoke log, one bench line per workload / decoder (B200 arm, plus the
# reference arm of the headline workload), the ncu launch list of the headline bench command and one
# `ncu --set full` capture per kernel family.  Numbers printed by runs under ncu are never bench values.
set -u
TAG=${1:-ev}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
b() { name=$1; shift; timeout 300 python bench.py --cpu-seconds ${CPU_S:-4} "$@" > $OUT/${TAG}_bench_${name}.json 2>> $OUT/${TAG}_bench.err; echo "bench $name rc=$?"; }
CPU_S=12 b hgp_34_n1600
b hgp_34_n1600_reference --impl reference --steps 2 --warmup 1
b hgp_34_n1600_f32 --precision f32
b hgp_34_n225 --workload hgp_34_n225
b hgp_34_n6400 --workload hgp_34_n6400 --shots 262144
b lp_882 --workload lp_882
b qclp_1054 --workload qclp_1054
b a5lp_2700 --workload a5lp_2700 --shots 524288
b st_hgp225_T3 --workload st_hgp225_T3
b st_hgp1600_T3 --workload st_hgp1600_T3 --shots 262144
b st_hgp1600_T3_bpc --workload st_hgp1600_T3 --shots 131072 --kernel cluster
b st_hgp1600_T1 --workload st_hgp1600_T1 --shots 524288
b hgp_34_n1600_p0.01 --p 0.01
b hgp_34_n1600_p0.02 --p 0.02
b hgp_34_n1600_ps --bp-method ps
b ssf_hgp_34_n1600 --decoder ssf --shots 524288
b ssf_hgp_34_n225 --decoder ssf --workload hgp_34_n225
b ssf_a5lp_2700 --decoder ssf --workload a5lp_2700 --shots 262144
b bposd_n1600_p0.05 --decoder bposd --shots 262144
b bposd_n1600_p0.02 --decoder bposd --p 0.02 --shots 524288
b bposd_cs7_n1600_p0.05 --decoder bposd --osd-method osd_cs --osd-order 7 --shots 262144
b bposd_n225_p0.05 --decoder bposd --workload hgp_34_n225
SHORT="--steps 2 --warmup 3 --shots 131072 --no-cpu-baseline"
python bench.py $SHORT > $OUT/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_n1600.csv \
    python bench.py $SHORT > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
cap() { name=$1; k=$2; shift 2; ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $OUT/${TAG}_${name} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/${TAG}_ncu_${name}.log 2>&1; echo "ncu $name rc=$?"; }
cap bp7_n1600 bp7 --shots 131072
cap osd_n1600 osd --decoder bposd --shots 16384
cap ssf_n1600 ssf --decoder ssf --shots 131072
cap bp7c_st1600 bp7 --workload st_hgp1600_T3 --shots 16384
cap bpc_st1600 bpc --workload st_hgp1600_T3 --shots 16384 --kernel cluster
cap syn128_n1600 syndrome128 --shots 1048576
python scripts/microbench/syndrome_bw.py > $OUT/${TAG}_syndrome_microbench.txt 2>&1
