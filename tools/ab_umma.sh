run() { tag=$1; shift; env "$@" CDAN_UMMA_VERBOSE=1 python bench.py --steps 6 --warmup 3 --layers --no-cpu-baseline > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.txt; echo "== $tag $@"; grep "umma_tma" gpurun_out/ab_$tag.txt | grep "conv3" | sort; grep "pair:.*Cin=128 Cout=256.*N=32" gpurun_out/ab_$tag.txt | sort | uniq | cut -c1-140; python -c "import json; d=json.load(open('gpurun_out/ab_$tag.json')); print(d['ms_per_step'])"; }
# A/B runs of the tile kernel switches (CDAN_UMMA_*): per-layer times of the four tile-kernel layers per variant; run via gpurun.
timeout 200 python -m pytest tests/test_gpu_ops.py -x -q -k conv2d 2>&1 | tail -2
CDAN_UMMA_POOL_WP=16 timeout 200 python -m pytest tests/test_gpu_ops.py -x -q -k conv2d 2>&1 | tail -2
run a CDAN_UMMA_POOL_WP=16
run b CDAN_UMMA_POOL_WP=32
run c CDAN_UMMA_POOL_WP=16 CDAN_UMMA_NMB_MIN=2
run d
