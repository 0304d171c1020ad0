# usage: bash tools/run_ab_list.sh <tag> lib1 lib2 ...   (GPU box)
tag=$1; shift
nvidia-smi -L | head -1
tools/ab.sh "$@" 2>&1 | tee gpurun_out/${tag}_ab.log
