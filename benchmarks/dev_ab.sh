python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-batched --no-cpu-baseline --streaming 0 > gpurun_out/s3_bench_d.json 2> gpurun_out/s3_bench_d.err
python -c "
import json; d=json.load(open('gpurun_out/s3_bench_d.json')); print('BENCH', d['value'], d['e2e']['value'], d['roofline']['us_per_launch'], d['roofline']['us_per_launch_eager_chain'])"
cp 3d-lidar-slam_b200/build/libb2slam_timing.so 3d-lidar-slam_b200/libb2slam.so
python benchmarks/dev_period.py 2>&1 | grep -v "^  cta\|slowest by" | cut -c1-1200
