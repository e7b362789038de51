#!/bin/bash
# Short GPU evidence pass (about 12 minutes of box time):  scripts/gpu_evidence_c.sh <tag>
# pytest -m gpu, smoke, the bench lines of the BASELINE workloads and the round's changed kernels, the ncu launch list of
# the headline command and three `ncu --set full` captures.  Every step runs under its own timeout.
set -u
TAG=${1:-ev}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/${TAG}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
b() { name=$1; shift; timeout 240 python bench.py --cpu-seconds ${CPU_S:-2} "$@" > $OUT/${TAG}_bench_${name}.json 2>> $OUT/${TAG}_bench.err; echo "bench $name rc=$?"; }
CPU_S=10 b hgp_34_n1600
b hgp_34_n1600_reference --impl reference --steps 2 --warmup 1
b hgp_34_n1600_p0.01 --p 0.01
b hgp_34_n1600_p0.02 --p 0.02
b hgp_34_n1600_ps --bp-method ps
b hgp_34_n1600_f32 --precision f32
b hgp_34_n225 --workload hgp_34_n225
b hgp_34_n6400 --workload hgp_34_n6400 --shots 262144
b lp_882 --workload lp_882
b qclp_1054 --workload qclp_1054
b a5lp_2700 --workload a5lp_2700 --shots 524288
b st_hgp225_T3 --workload st_hgp225_T3
b st_hgp1600_T3 --workload st_hgp1600_T3 --shots 262144
b st_hgp1600_T3_bpc --workload st_hgp1600_T3 --shots 131072 --kernel cluster --steps 3
b st_hgp1600_T1 --workload st_hgp1600_T1 --shots 524288
b ssf_hgp_34_n1600 --decoder ssf --shots 524288
b bposd_n1600_p0.05 --decoder bposd --shots 262144
b bposd_n1600_p0.02 --decoder bposd --p 0.02 --shots 524288
b bposd_cs7_n1600_p0.05 --decoder bposd --osd-method osd_cs --osd-order 7 --shots 262144
SHORT="--steps 2 --warmup 3 --shots 131072 --no-cpu-baseline"
timeout 200 python bench.py $SHORT > $OUT/${TAG}_plain.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_n1600.csv \
    python bench.py $SHORT > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
cap() { name=$1; k=$2; shift 2; timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $OUT/${TAG}_${name} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/${TAG}_ncu_${name}.log 2>&1; echo "ncu $name rc=$?"
    # the reports are large (gpurun brings back at most 64 MiB): summarise here, keep the text only
    { python scripts/ncu_summary.py $OUT/${TAG}_${name}.ncu-rep; echo "--- hottest source lines (samples / instructions / shared wavefronts)";
      python scripts/ncu_lines.py $OUT/${TAG}_${name}.ncu-rep 24; } > $OUT/${TAG}_${name}_ncu_full_summary.txt 2>&1
    rm -f $OUT/${TAG}_${name}.ncu-rep; }
cap bp7_n1600 bp7 --shots 131072
cap bp7c_st1600 bp7 --workload st_hgp1600_T3 --shots 16384
cap osd_n1600 osd --decoder bposd --shots 16384
du -sh $OUT
