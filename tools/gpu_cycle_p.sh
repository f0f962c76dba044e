mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_p_build.log 2>&1
python tools/seam_ab.py 25 > gpurun_out/r02_p_seam_ab.txt 2>&1
cat gpurun_out/r02_p_seam_ab.txt
