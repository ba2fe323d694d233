set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; tail -c 3000 gpurun_out/bench_b200.json
python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json
