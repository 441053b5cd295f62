# GPU session: full -m gpu test run + cycle accounting of the fused kernel (HP_PROFILE build) in resident and parked mode
set -x
mkdir -p gpurun_out/s2
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s2/tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/s2/tests.txt
tail -5 gpurun_out/s2/tests.txt
export HPMVS_LIB=$PWD/hpmvs_b200/libhpmvs_b200_prof.so HPMVS_PROFILE_PRINT=1
timeout 300 python bench.py --steps 3 --warmup 3 --cpu-sample 512 > gpurun_out/s2/prof_plane8.json 2> gpurun_out/s2/prof_plane8.err
timeout 300 python bench.py --steps 2 --warmup 3 --cpu-sample 512 --workload plane8x100k > gpurun_out/s2/prof_100k_parked.json 2> gpurun_out/s2/prof_100k_parked.err
HPMVS_PARKED=0 timeout 300 python bench.py --steps 2 --warmup 3 --cpu-sample 512 --workload plane8x100k > gpurun_out/s2/prof_100k_resident.json 2> gpurun_out/s2/prof_100k_resident.err
grep -h "profile\] opt" gpurun_out/s2/*.err | tail -6
