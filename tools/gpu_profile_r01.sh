# Round-1 profile pass: (1) per-launch durations of one C2 frame (serialised, cold cache under ncu),
# (2) full-set captures of the two heaviest kernels.  Numbers under ncu are never bench values.
cd /root/repo
cat > /tmp/one.py <<'PY'
import sys
sys.path.insert(0, "fast-3d-pointcloud-segmentation_b200")
import f3ps
from f3ps import synth
pts = synth.make_frame(seed=20020)
g = f3ps.Segmenter(); g.set_vccs_params(); g.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1)
g.set_input(pts); g.run(0.2)
g.set_input(pts); g.run(0.2)
print(g.counts().n_merges, g.stage_ms())
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv python /tmp/one.py > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:merge_fast --launch-skip 1 --launch-count 1 -o gpurun_out/prof_merge_fast_r01 -f python /tmp/one.py > gpurun_out/ncu_merge_fast.log 2>&1; echo "merge rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:expand_persistent --launch-skip 1 --launch-count 1 -o gpurun_out/prof_expand_r01 -f python /tmp/one.py > gpurun_out/ncu_expand.log 2>&1; echo "expand rc=$?"
tail -3 gpurun_out/ncu_expand.log
