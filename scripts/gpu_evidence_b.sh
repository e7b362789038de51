set -u
TAG=r1s4b
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
b() { name=$1; shift; python bench.py "$@" > $OUT/${TAG}_bench_${name}.json 2>> $OUT/${TAG}_bench.err; echo "bench $name rc=$?"; }
b hgp_34_n1600
b st_hgp1600_T3 --workload st_hgp1600_T3 --shots 262144
b bposd_n1600_p0.05 --decoder bposd --shots 262144
b bposd_n1600_p0.02 --decoder bposd --p 0.02 --shots 524288
b bposd_cs7_n1600_p0.05 --decoder bposd --osd-method osd_cs --osd-order 7 --shots 262144
b bposd_a5lp_2700_p0.05 --decoder bposd --workload a5lp_2700 --shots 65536 --steps 3
b bposd_n6400_p0.05 --decoder bposd --workload hgp_34_n6400 --shots 8192 --steps 2 --warmup 3 --no-cpu-baseline
