#!/bin/bash
# Run under gpurun: the bench lines, the ncu launch list and full captures of the hot kernels.  usage: tools/profile_round.sh <tag>
tag=${1:-rX}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --decoder-mode exact --no-cpu-baseline --no-other-configs > gpurun_out/${tag}_bench_exact_decoder.json 2>> gpurun_out/${tag}_bench.err
# launch list of the bench command (cold-cache, serialised launches: shares, not absolutes); one hop per step keeps it short
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --hops-per-step 1 --steps 8 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/${tag}_ncu_bench.log 2>&1
# one full capture of each hot kernel, serialised (one worker group, no sub-batches); skip the warm-up hops: 3 hops x 6 kernels
ncu --set full --clock-control none --import-source on -k regex:'EncoderKernel|DecoderKernel|Rvq' -s 24 -c 6 \
    -o gpurun_out/${tag}_full python bench.py --groups 1 --split 1 --e2e-split 1 --hops-per-step 1 --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/${tag}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'DecoderKernel|ComfortNoise|PlcMix|LogMel|NoiseEst' -s 15 -c 7 \
    -o gpurun_out/${tag}_full_plc python bench.py --workload decode_plc --loss 0.5 --groups 1 --split 1 --e2e-split 1 --hops-per-step 1 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full2.log 2>&1
ls -la gpurun_out | tail -12
