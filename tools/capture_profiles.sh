#!/bin/bash
# Evidence capture on the GPU box (run under gpurun, ONE GPU): launch list of the default bench line and one `ncu --set full` capture per
# kernel on its workload. Outputs go to gpurun_out/ (scratch); tools/ncu_summary.py + tools/record_traffic.py turn them into profiles/*.json here.
set -u
OUT=gpurun_out
R=${1:-r02}
B="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-e2e"
# every launch of the default bench line (headline + extras), with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${R}_launches_bench_steps2_warmup1.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${R}_launches_bench.log 2>&1
full() {  # name, kernel regex, bench args
    ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -o $OUT/${R}_$1 -f $B $3 > $OUT/${R}_$1.log 2>&1
    echo "$1 rc=$?"
    python tools/ncu_summary.py $OUT/${R}_$1.ncu-rep $OUT/${R}_$1_full.json
    # gpurun merges at most 64 MiB back: keep the summaries, and the report itself only for the headline kernel
    if [ "$1" != "tile64x128_S1" ]; then rm -f $OUT/${R}_$1.ncu-rep; fi
}
full tile64x128_S1 qp_tile_kernel ""
full tile64x128_S2 qp_tile_kernel "--settings S2"
full tile32x64_config2_S1 qp_tile_kernel "--workload config2"
full small_2x2_b4096_S2 qp_small "--batch 4096 --n 2 --m 2 --settings S2"
full block_dense256x512_b592_S2 qp_block "--batch 592 --n 256 --m 512 --settings S2"
ncu --set full --clock-control none --import-source on -k regex:qp_cluster -s 1 -c 1 -o $OUT/${R}_cluster256x512_S2 -f \
    python bench.py --workload config5 --settings S2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/${R}_cluster256x512_S2.log 2>&1
echo "cluster rc=$?"
python tools/ncu_summary.py $OUT/${R}_cluster256x512_S2.ncu-rep $OUT/${R}_cluster256x512_S2_full.json
rm -f $OUT/${R}_cluster256x512_S2.ncu-rep
ls -la $OUT/
