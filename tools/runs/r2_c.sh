#!/bin/bash
# round-2 batch C: resident small-FFT kernel (tests + A/B) and row/column co-residency variants at 1024
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2c_tests.log
D=$PWD/style_transfer_based_holographic_imaging_b200
{
export ASM_B200_LIB=$D/libasm_b200_tune.so
for r in 1 0; do
  echo "== RESIDENT=$r"
  ASM_B200_RESIDENT=$r python tools/quick_bench.py 256 4096
  ASM_B200_RESIDENT=$r python tools/quick_bench.py 256 4096 1
  ASM_B200_RESIDENT=$r python tools/quick_bench.py 128 8192
  ASM_B200_RESIDENT=$r python tools/quick_bench.py 128 8192 1
  ASM_B200_RESIDENT=$r python tools/quick_bench.py 128 5 1 50
  ASM_B200_RESIDENT=$r python tools/quick_bench.py 64 8192 1
done
for mb in 40 120 160; do echo "== RESIDENT budget $mb MB"; ASM_B200_CHUNK_MB=$mb python tools/quick_bench.py 256 4096; ASM_B200_CHUNK_MB=$mb python tools/quick_bench.py 128 8192 1; done
echo "== split 256 fwd only"; python tools/split_bench.py 256 4096
echo "== 1024: 12-warp row CTAs (1/SM)"; python tools/quick_bench.py 1024 512
export ASM_B200_LIB=$D/libasm_b200_tune6.so
echo "== 1024: 6-warp row CTAs, 2/SM"; python tools/quick_bench.py 1024 512; python tools/pass_times.py 1024 108
echo "== 1024: 6-warp row CTAs, 1/SM"; ASM_B200_ROW_CTAS=1 python tools/quick_bench.py 1024 512
echo "== 1024: 6-warp row CTAs, 2/SM, lanes 4"; ASM_B200_LANES=4 python tools/quick_bench.py 1024 512
echo "== 1024: 6-warp row CTAs, 2/SM, lanes 2"; ASM_B200_LANES=2 python tools/quick_bench.py 1024 512
export ASM_B200_LIB=$D/libasm_b200_tune4.so
echo "== 1024: 4-warp row CTAs, 3/SM"; python tools/quick_bench.py 1024 512; python tools/pass_times.py 1024 108
echo "== 1024: 4-warp row CTAs, 2/SM"; ASM_B200_ROW_CTAS=2 python tools/quick_bench.py 1024 512
echo "== 1024: 4-warp row CTAs, 1/SM"; ASM_B200_ROW_CTAS=1 python tools/quick_bench.py 1024 512
echo "== 1024: 4-warp row CTAs, 2/SM lanes 4"; ASM_B200_ROW_CTAS=2 ASM_B200_LANES=4 python tools/quick_bench.py 1024 512
} > gpurun_out/r2c_sweep.log 2>&1
{
unset ASM_B200_LIB
echo "== FFT 2048 with 512-thread column CTAs"
python tools/quick_bench.py 2048 128; python tools/pass_times.py 2048 32
python tools/quick_bench.py 1024 128 1; python tools/pass_times.py 1024 32 1
echo "== training step, 1 GPU"
python examples/train_step.py --batch 64 --size 256 --steps 5
python examples/train_step.py --batch 64 --size 256 --steps 5 --forward reference
} > gpurun_out/r2c_more.log 2>&1
{
D=$PWD/style_transfer_based_holographic_imaging_b200
export ASM_B200_LIB=$D/libasm_b200_tuneR4.so
export ASM_B200_BULK=0
echo "== R4 (4-warp LDG row CTAs): rows 4/SM cols 2/SM"; python tools/quick_bench.py 1024 512
for rc in 1 2; do for cc in 1 2; do for l in 3 4 6; do
  echo "== R4 rows $rc/SM cols $cc/SM lanes $l"; ASM_B200_LDG_ROW_CTAS=$rc ASM_B200_COLS_CTAS=$cc ASM_B200_LANES=$l python tools/quick_bench.py 1024 512
done; done; done
echo "== R4 rows 1/SM cols 1/SM lanes 6 chunk 144"; ASM_B200_LDG_ROW_CTAS=1 ASM_B200_COLS_CTAS=1 ASM_B200_LANES=6 ASM_B200_CHUNK_MB=144 python tools/quick_bench.py 1024 512
echo "== R4 rows 1/SM cols 1/SM lanes 6 chunk 288"; ASM_B200_LDG_ROW_CTAS=1 ASM_B200_COLS_CTAS=1 ASM_B200_LANES=6 ASM_B200_CHUNK_MB=288 python tools/quick_bench.py 1024 512
} > gpurun_out/r2c_r4.log 2>&1
