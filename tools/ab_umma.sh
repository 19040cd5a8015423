run() { tag=$1; shift; env "$@" CDAN_UMMA_VERBOSE=1 python bench.py --steps 6 --warmup 3 --layers --no-cpu-baseline > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.txt; echo "== $tag $@"; grep "umma_tma" gpurun_out/ab_$tag.txt | grep "conv3\|conv4\|decoder.conv1\|decoder.conv2" | sort; grep "pair:.*N=32" gpurun_out/ab_$tag.txt | sort | uniq | cut -c1-140; python -c "import json; d=json.load(open('gpurun_out/ab_$tag.json')); print(d['ms_per_step'])"; }
run a CDAN_UMMA_NMB_MIN=2
run b CDAN_UMMA_NMB_MIN=2 CDAN_UMMA_SB_MAX=8
run c CDAN_UMMA_SB_MAX=8
