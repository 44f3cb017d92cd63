#!/bin/bash
# Round evidence in one GPU call (run on the GPU box from the repo root):  bash tools/evidence.sh TAG [full]
#   1. the default bench line (no profiler)                                   -> gpurun_out/TAG_bench_default.json
#   2. ncu launch list of a shortened bench command (after 1. exited 0)        -> gpurun_out/TAG_ncu_launches_bench.csv
#   3. DRAM bytes + duration of EVERY kernel of two substeps, per workload    -> gpurun_out/TAG_ncu_traffic_<workload>.csv
#   4. (full) ncu --set full --import-source on of k_gs_pairs and k_gs_resolve -> gpurun_out/TAG_full_<workload>.ncu-rep
# tools/evidence_summary.py turns 2.-4. into the files committed under profiles/.  Numbers printed under ncu are never bench values.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err || { echo "bench failed"; tail -5 $O/${TAG}_bench_default.err; exit 1; }
tail -c 600 $O/${TAG}_bench_default.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_ncu_launches_bench.csv \
    python bench.py --settle 8 --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline > $O/${TAG}_ncu_launches_bench.log 2>&1
for w in pile16m uniform16m; do
    timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file $O/${TAG}_ncu_traffic_$w.csv python tools/profile_substeps.py --workload $w --resolve 3 --settle 64 --steps 2 > $O/${TAG}_ncu_traffic_$w.log 2>&1
done
if [ "$2" = "full" ]; then
    for w in pile16m uniform16m; do
        timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gs_ --launch-skip 325 --launch-count 5 -o $O/${TAG}_full_$w -f \
            python tools/profile_substeps.py --workload $w --resolve 3 --settle 64 --steps 2 > $O/${TAG}_full_$w.log 2>&1
    done
fi
ls -la $O | grep ${TAG}_
