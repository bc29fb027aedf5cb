mkdir -p gpurun_out
TAG=${1:-q}
bash tools/r2_quick.sh $TAG
bash tools/r2_ncu3.sh $TAG
