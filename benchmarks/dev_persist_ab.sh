# developer A/B: chained launches vs the persistent single-launch variant of the dense ICP step (B2_DENSE_PERSIST=1)
if [ "${TESTS:-0}" = 1 ]; then B2_DENSE_PERSIST=1 timeout 250 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
for v in 0 1 0 1; do
B2_DENSE_PERSIST=$v timeout 120 python bench.py --no-batched --no-cpu-baseline --streaming 0 > gpurun_out/dev_bench.json 2> gpurun_out/dev_bench.err
python -c "
import json; d=json.load(open('gpurun_out/dev_bench.json')); print('persist=$v BENCH', d['value'], d['e2e']['value'], d['roofline']['us_per_launch'], d['roofline']['us_per_registration_eager'], d['gpu_launches'])"
done
