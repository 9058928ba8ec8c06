#!/bin/bash
# K-split bench lines: one image, K = 16, CLEVR6 layer sizes, 1 GPU vs 2 GPUs (fp16 and tf32)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for prec in fp16 tf32; do
timeout 600 python bench.py --batch 1 --slots 16 --steps 20 --warmup 3 --precision $prec --no-variants --no-cpu-baseline > gpurun_out/r2_ksplit_1gpu_$prec.json 2> gpurun_out/r2_ksplit_1gpu_$prec.err
timeout 600 $TR --nproc-per-node 2 --master-port 29521 bench.py --gpus 2 --k-split --batch 1 --slots 16 --steps 20 --warmup 3 --precision $prec --no-variants > gpurun_out/r2_ksplit_2gpu_$prec.json 2> gpurun_out/r2_ksplit_2gpu_$prec.err
for f in gpurun_out/r2_ksplit_1gpu_$prec gpurun_out/r2_ksplit_2gpu_$prec; do echo == $f; tail -2 $f.err | cut -c1-300; cut -c1-1000 $f.json; done
done
