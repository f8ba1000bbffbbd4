# last round-2 evidence job (1 GPU, ~3 minutes): the SGL / yelp bench line (BASELINE config 3), ncu --set full of the evaluation / clustering kernels
# at the amazon shape, their live timings, the NCL / amazon line (config 5 on one GPU)
mkdir -p gpurun_out
timeout 120 python bench.py --workload sgl-yelp --steps 20 --warmup 3 --no-cuda-graph --row-shard off > gpurun_out/bench_r2_sgl-yelp.json 2> gpurun_out/bench_r2_sgl-yelp.err; cut -c1-200 gpurun_out/bench_r2_sgl-yelp.json
timeout 90 ncu --set full --clock-control none -k regex:"predict_|topk_kernel|kmeans" -c 14 -o /tmp/prof_minor python tools/minor_kernels.py --ncu > gpurun_out/ncu_minor.log 2>&1; ncu -i /tmp/prof_minor.ncu-rep --page raw --csv > gpurun_out/ncu_minor_kernels_r2.csv 2>/dev/null; wc -c gpurun_out/ncu_minor_kernels_r2.csv
timeout 40 python tools/minor_kernels.py > gpurun_out/minor_kernels_r2.json 2> gpurun_out/minor_kernels_r2.err; cat gpurun_out/minor_kernels_r2.json; tail -2 gpurun_out/minor_kernels_r2.err
timeout 100 python bench.py --workload ncl-amazon --steps 20 --warmup 3 --no-cpu-baseline --no-cuda-graph --row-shard off > gpurun_out/bench_r2_ncl-amazon.json 2> gpurun_out/bench_r2_ncl-amazon.err; cut -c1-200 gpurun_out/bench_r2_ncl-amazon.json
