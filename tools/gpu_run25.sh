cd /root/repo
mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import sys
sys.path.insert(0, "fast-3d-pointcloud-segmentation_b200")
import f3ps
from f3ps import synth
pts = synth.make_frame(seed=20020)
g = f3ps.Segmenter(); g.set_vccs_params(); g.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1)
g.set_input(pts); g.extract()
g.set_input(pts); g.extract()
print(g.counts().sweeps, g.stage_ms(), g.expand_profile())
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:expand_persistent --launch-skip 1 --launch-count 1 -o gpurun_out/prof_expand_r25 -f python /tmp/one.py > gpurun_out/r25_ncu_expand.log 2>&1; echo "expand ncu rc=$?"
ncu -i gpurun_out/prof_expand_r25.ncu-rep --page raw --csv > gpurun_out/prof_expand_r25_raw.csv 2>/dev/null
tail -2 gpurun_out/r25_ncu_expand.log
