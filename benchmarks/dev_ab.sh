# developer A/B: GPU tests (optional: TESTS=1), a short bench line (twice: box noise), optionally the per-phase timeline
# from the -DB2_TAIL_TIMING build kept as build/libb2slam_timing.so (TIMELINE=1)
if [ "${TESTS:-0}" = 1 ]; then timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
for i in 1 2; do
timeout 120 python bench.py --no-batched --no-cpu-baseline --streaming 0 > gpurun_out/dev_bench.json 2> gpurun_out/dev_bench.err
python -c "
import json; d=json.load(open('gpurun_out/dev_bench.json')); print('BENCH', d['value'], d['e2e']['value'], d['roofline']['us_per_launch'], d['roofline']['us_per_registration_eager'])"
done
if [ "${TIMELINE:-0}" = 1 ]; then
cp 3d-lidar-slam_b200/build/libb2slam_timing.so 3d-lidar-slam_b200/libb2slam.so
timeout 60 python benchmarks/dev_period.py 2>&1 | grep -v "^  cta\|slowest by" | cut -c1-1200 | tail -12
fi
