OUT=gpurun_out/r05b; mkdir -p $OUT
timeout 1500 python scripts/configs_fullsize.py --which c3,c5 --fuse3 0 --chi5 512,1024 --fuse-both-upto 0 --budget 400 --out $OUT/configs.jsonl > $OUT/configs.log 2>&1; cut -c1-900 $OUT/configs.jsonl; tail -2 $OUT/configs.log
