# Late round-2 profile pass (run under gpurun; reports land in gpurun_out/): the solo (768-thread) resident merge kernel on the C2 frame
#  (1) per-launch durations of one C2 frame (ncu, serialised, cold cache), (2) full-set capture of the merge kernel.
# Numbers under ncu are never bench values.
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
print(g.counts().n_merges, g.counts().merge_path, g.stage_ms())
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02c.csv python /tmp/one.py > gpurun_out/r02c_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:merge_fast --launch-skip 1 --launch-count 1 -o gpurun_out/prof_merge_lean_r02c -f python /tmp/one.py > gpurun_out/r02c_ncu_merge.log 2>&1; echo "merge rc=$?"
tail -2 gpurun_out/r02c_ncu_merge.log
