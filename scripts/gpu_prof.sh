# cycle accounting with the HP_PROFILE build; usage: bash scripts/gpu_prof.sh <tag> [workload]
tag=${1:-prof}; wl=${2:-plane8}
mkdir -p gpurun_out/$tag
export HPMVS_LIB=$PWD/hpmvs_b200/libhpmvs_b200_prof.so HPMVS_PROFILE_PRINT=1
timeout 300 python bench.py --steps 3 --warmup 3 --cpu-sample 512 --workload $wl > gpurun_out/$tag/prof_$wl.json 2> gpurun_out/$tag/prof_$wl.err
grep "profile" gpurun_out/$tag/prof_$wl.err | tail -32
