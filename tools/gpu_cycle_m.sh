mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_m_build.log 2>&1
python tools/seam_stats.py > gpurun_out/r02_m_seam_stats.txt 2>&1
MGB_HOST_THREADS=14 python tools/seam_stats.py > gpurun_out/r02_m_seam_stats_t14.txt 2>&1
cat gpurun_out/r02_m_seam_stats.txt
