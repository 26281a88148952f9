#!/bin/bash
# prints a marker and checks that the GPU still answers (used between phases of a gpurun job)
echo "== sanity after: $1"
timeout 60 nvidia-smi --query-gpu=name,temperature.gpu,power.draw,clocks.sm --format=csv,noheader || echo "!! nvidia-smi failed/hung after $1"
timeout 120 python -c "import torch; x=torch.ones(1<<20,device='cuda'); print('cuda ok', float(x.sum()))" || echo "!! torch sanity failed/hung after $1"
