#!/bin/bash
# Profiling aid: two independent single-GPU processes side by side on a 2-GPU box (no torch.distributed), 1024-frame steps.
CUDA_VISIBLE_DEVICES=0 python bench.py --device-only --frames 1024 --batch 1024 --steps 20 --warmup 3 > gpurun_out/t17_a.log 2>&1 &
CUDA_VISIBLE_DEVICES=1 python bench.py --device-only --frames 1024 --batch 1024 --steps 20 --warmup 3 > gpurun_out/t17_b.log 2>&1 &
wait
CUDA_VISIBLE_DEVICES=1 python bench.py --device-only --frames 1024 --batch 1024 --steps 20 --warmup 3 > gpurun_out/t17_c.log 2>&1
