#!/bin/bash
# usage (on the GPU box): tools/final_measure.sh TAG   -> gpurun_out/TAG_*: GPU test log, bench lines, ncu launch list + full capture
tag=${1:-final}
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || tail -3 $out/${tag}_bench.err
for w in cfg3 cfg4; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --sustained-steps 0 > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err || tail -3 $out/${tag}_bench_$w.err
done
args="--steps 2 --warmup 1 --no-cpu-baseline --sustained-steps 0 --no-agreement"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv python bench.py $args > $out/${tag}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name 'regex:edge_msg_t|edge_mlp_tc|node_update|resampler_df1|knn_warp|edge_embed' --launch-skip 16 -c 16 -f -o $out/${tag}_full python bench.py $args > $out/${tag}_ncu_full.log 2>&1
ls -la $out/${tag}_*
