python -m pytest tests -m gpu -q --tb=short -k "field_kernels or ark_vrf or headline" 2>&1 | tail -n 3
python tools/gpu_microbench_r2.py > gpurun_out/r2m_microbench.json 2>&1
grep -A3 -E "\"(fq_mul|fq_sqr|fr_mul|fr_sqr|g1_madd|fr_chain|fq_chain)\"" gpurun_out/r2m_microbench.json | grep -E "\"|ops_per_s|ns_per"
python tools/gpu_commit_ab.py default build/var/lib_oldsqr.so > gpurun_out/r2m_commit_ab.log 2>&1; cat gpurun_out/r2m_commit_ab.log
bash tools/gpu_r2_run.sh r2m bench bench512
