cd scripts
LTB200_K7_BM=1 LTB200_K7_PF=1 LTB200_K7_FBG=2 timeout 200 python -m pytest ../tests/test_k4_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 120 python k7_band_exp.py 4 banded 8192
for bm in 1; do for pf in 0 1; do for b in 8 16 32; do for f in 2 4 8; do
LTB200_K7_BM=$bm LTB200_K7_PF=$pf LTB200_K7_FBG=$f timeout 120 python k7_band_exp.py $b banded 8192 2>&1 | sed "s/^/BM=$bm PF=$pf /"
done; done; done; done
