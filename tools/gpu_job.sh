set -x
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; cat gpurun_out/bench6.json; tail -3 gpurun_out/bench6.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --op-timing 0 > gpurun_out/bench7.json 2> gpurun_out/bench7.err; cut -c1-250 gpurun_out/bench7.json; tail -3 gpurun_out/bench7.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --op-timing 1 > gpurun_out/bench8.json 2> gpurun_out/bench8.err; cut -c1-250 gpurun_out/bench8.json; tail -3 gpurun_out/bench8.err
