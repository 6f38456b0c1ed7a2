# consolidated check + profiles of the large-scene path (C4 / C5)
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r13_pytest.log; cat gpurun_out/r13_pytest.log
timeout 600 python tools/stage_roofline.py c4 merge > gpurun_out/r13_stage_roofline_c4.md 2>&1; tail -25 gpurun_out/r13_stage_roofline_c4.md
cat > /tmp/c5one.py <<'PY'
import sys
sys.path.insert(0, "fast-3d-pointcloud-segmentation_b200")
import f3ps
from f3ps import synth
pts = synth.make_room_scan(n_points=10_000_000)
g = f3ps.Segmenter(); g.set_vccs_params(voxel_res=0.01, seed_res=0.1); g.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1)
g.set_input(pts); g.run(0.2)
g.set_input(pts); g.run(0.2)
c = g.counts()
print(c.n_voxels, c.n_supervoxels, c.n_edges, c.n_merges, g.stage_ms())
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_c5_r13.csv python /tmp/c5one.py > gpurun_out/r13_ncu_launch.log 2>&1; echo "launch list rc=$?"; tail -1 gpurun_out/r13_ncu_launch.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:merge_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/prof_merge_general_r13 -f python /tmp/c5one.py > gpurun_out/r13_ncu_merge.log 2>&1; echo "merge ncu rc=$?"
ncu -i gpurun_out/prof_merge_general_r13.ncu-rep --page raw --csv > gpurun_out/prof_merge_general_r13_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
