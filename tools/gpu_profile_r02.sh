# Round-2 profile pass (run under gpurun; reports land in gpurun_out/):
#  (1) per-launch durations of one C2 frame (ncu, serialised, cold cache),
#  (2) full-set capture of the resident merge kernel (K7) on the C2 frame,
#  (3) per-kernel DRAM bytes of K1..K6 on the 10 M-point C4 scene.
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
print(g.counts().n_merges, g.stage_ms())
PY
cat > /tmp/c4.py <<'PY'
import sys
sys.path.insert(0, "fast-3d-pointcloud-segmentation_b200")
import f3ps
from f3ps import synth
pts = synth.make_dense_scene(seed=40000)
g = f3ps.Segmenter(); g.set_vccs_params(voxel_res=0.004, seed_res=0.04); g.set_merge_params(color_mode=1, geom_mode=0, merge_mode=0, lam=0.5)
g.set_input(pts); g.extract(); g.sync()
g.set_input(pts); g.extract(); g.sync()
c = g.counts(); print(c.n_points, c.n_voxels, c.n_seeds, c.n_supervoxels, c.n_edges, g.stage_ms())
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python /tmp/one.py > gpurun_out/r02_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:merge_fast --launch-skip 1 --launch-count 1 -o gpurun_out/prof_merge_lean_r02 -f python /tmp/one.py > gpurun_out/r02_ncu_merge.log 2>&1; echo "merge rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/c4_dram_r02.csv python /tmp/c4.py > gpurun_out/r02_ncu_c4.log 2>&1; echo "c4 rc=$?"
tail -2 gpurun_out/r02_ncu_c4.log
