mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_r_build.log 2>&1
python tools/d2h_probe.py > gpurun_out/r02_r_d2h_probe.txt 2>&1
cat gpurun_out/r02_r_d2h_probe.txt
MGB_HOST_STATS=1 python tools/seam_ab.py 2 "t12 512K x8 nt512  ring" 2>&1 | tail -8
MGB_HOST_STATS=1 python tools/seam_ab.py 2 "t12 1M x6   nt512  ring" 2>&1 | tail -8
python tools/seam_ab.py 25 > gpurun_out/r02_r_seam_ab.txt 2>&1
cat gpurun_out/r02_r_seam_ab.txt
