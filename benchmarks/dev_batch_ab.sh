# developer A/B of the batched leg: build/libb2slam_base.so vs build/libb2slam_new.so on one box
for v in base new base new; do
cp 3d-lidar-slam_b200/build/libb2slam_$v.so 3d-lidar-slam_b200/libb2slam.so
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --streaming 0 --pairs-total ${PAIRS:-2048} --pairs-reps 2 > gpurun_out/dev_batch.json 2> gpurun_out/dev_batch.err
python -c "
import json; d=json.load(open('gpurun_out/dev_batch.json')); b=d['batched_pairs']; print('$v', b['value'], b['ms_all_reps'], b['roofline']['us_per_launch'], b['roofline']['front_end_ms'])"
done
