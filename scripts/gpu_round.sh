#!/bin/bash
# One GPU evidence pass (run on the B200 box through gpurun):
#   scripts/gpu_round.sh <tag> [workload] [extra bench flags...]
# writes under gpurun_out/: pytest log, bench lines (B200 arm + reference arm), the ncu launch list
# of the bench command and one `ncu --set full` capture of the dominant kernel.
# Numbers printed by the runs under ncu are never bench values.
set -u
TAG=${1:-dev}
WL=${2:-hgp_34_n1600}
shift 2 2>/dev/null || true
EXTRA="$*"
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
python bench.py --workload $WL $EXTRA > $OUT/${TAG}_bench_${WL}.json 2> $OUT/${TAG}_bench_${WL}.err
echo "bench rc=$?"
python bench.py --impl reference --workload $WL --steps 2 --warmup 1 > $OUT/${TAG}_bench_${WL}_reference.json 2>> $OUT/${TAG}_bench_${WL}.err
echo "reference rc=$?"
# launch list of a short bench (same command, fewer steps) — only after the plain run exited 0
python bench.py --workload $WL $EXTRA --steps 2 --warmup 3 --shots 131072 --no-cpu-baseline > $OUT/${TAG}_plain_${WL}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_${WL}.csv \
    python bench.py --workload $WL $EXTRA --steps 2 --warmup 3 --shots 131072 --no-cpu-baseline > $OUT/${TAG}_ncu_launches_${WL}.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:bp -s 3 -c 1 -f -o $OUT/${TAG}_bp_${WL} \
    python bench.py --workload $WL $EXTRA --steps 2 --warmup 3 --shots 131072 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_full_${WL}.log 2>&1
echo "ncu full rc=$?"
