#!/bin/bash
# Run under gpurun: bench lines for every BASELINE config that fits one GPU + the ncu launch list and one full capture
# of the default bench command.  usage: tools/profile_round.sh <tag>   -> gpurun_out/<tag>_*
tag=${1:-rX}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --streams 1024 --no-cpu-baseline > gpurun_out/${tag}_bench_1024streams_3.2kbps.json
python bench.py --bits 120 --no-cpu-baseline > gpurun_out/${tag}_bench_4096streams_6.0kbps.json
python bench.py --bits 184 --no-cpu-baseline > gpurun_out/${tag}_bench_4096streams_9.2kbps.json
python bench.py --workload decode_plc > gpurun_out/${tag}_bench_decode_plc_loss0.1.json
python bench.py --workload decode_plc --loss 1.0 --no-cpu-baseline > gpurun_out/${tag}_bench_decode_plc_all_lost.json
python bench.py --decoder-mode tensor --no-cpu-baseline > gpurun_out/${tag}_bench_tensor_decoder.json
# launch list of the bench command (cold-cache, serialised launches: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
# one full capture of each hot kernel, serialised (one worker group, no sub-batches); skip the warm-up launches: 3 steps x 6 kernels
ncu --set full --clock-control none --import-source on -k regex:'EncoderKernel|DecoderKernel|Rvq' -s 24 -c 6 \
    -o gpurun_out/${tag}_full python bench.py --groups 1 --split 1 --e2e-split 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'DecoderKernel|LogMel|NoiseEst' -s 12 -c 4 \
    -o gpurun_out/${tag}_full_plc_tensor python bench.py --workload decode_plc --decoder-mode tensor --groups 1 --split 1 --e2e-split 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full2.log 2>&1
ls -la gpurun_out | tail -20
