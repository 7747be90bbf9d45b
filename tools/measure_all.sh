#!/bin/bash
# Every measurement DESIGN.md §6 quotes, in one go (one GPU; ~4 min of box time).  Logs land in gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/measure_all.sh'
# Multi-GPU numbers:  gpurun --gpus N -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
#                       --master-port 29517 tests/ddp_check.py --time --graph; ... tools/scene_sample_bench.py; ... bench.py --gpus N'
set -u
mkdir -p gpurun_out
python bench.py                                   > gpurun_out/m_bench.log 2>&1            # headline JSON line (value, e2e, roofline, cpu_baseline)
python tools/train_bench.py --profile             > gpurun_out/m_train.log 2>&1            # graphed training step + per-kernel table
python tools/concat_bench.py                      > gpurun_out/m_concat.log 2>&1           # concat-variant guided step
python tools/scene_sample_bench.py                > gpurun_out/m_scene.log 2>&1            # cfg5: 10 objects, DDIM 100 / DDPM 1000 + decode
python tools/attn_bench.py                        > gpurun_out/m_attn.log 2>&1             # attention forward (tcgen05 vs mma.sync) and backward
python tests/bench_torch_eager_gpu.py --train --concat > gpurun_out/m_torch_eager.log 2>&1 # PyTorch eager comparator on the same GPU
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/m_launches_step.csv python tools/profile_step.py 32 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/m_launches_train.csv python tools/train_bench.py --ncu > /dev/null 2>&1
tail -1 gpurun_out/m_bench.log | cut -c1-300
grep "graphed train step" gpurun_out/m_train.log; tail -1 gpurun_out/m_concat.log; tail -1 gpurun_out/m_scene.log
