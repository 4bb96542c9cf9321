#!/bin/bash
# One box with 8 GPUs: one-process strong scaling behind the C ABI (configs[3], both variants; level split and, for
# comparison, the frequency split), the multi-device tests, configs[4] batch of 1024 paths sharded by path over 8 ranks, and
# the bench at N = 8.
mkdir -p gpurun_out
T=${1:-r2w}
timeout 300 python tools/multi_probe.py --nf 1000000 --cutoff-ghz 0 --devices 1,2,4,8 > gpurun_out/${T}_multi_nocut_8gpu.json 2> gpurun_out/${T}_multi_nocut.err
AB200_MULTI_SPLIT=freq timeout 300 python tools/multi_probe.py --nf 1000000 --cutoff-ghz 0 --devices 1,8 > gpurun_out/${T}_multi_nocut_8gpu_freqsplit.json 2> gpurun_out/${T}_multi_nocut_freq.err
timeout 300 python tools/multi_probe.py --nf 1000000 --cutoff-ghz 750 --devices 1,2,4,8 > gpurun_out/${T}_multi_cut750_8gpu.json 2> gpurun_out/${T}_multi_cut750.err
timeout 200 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/${T}_multi_tests.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/c5_batch_dist.py --paths 1024 > gpurun_out/${T}_c5_1024paths_8gpu.json 2> gpurun_out/${T}_c5.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/${T}_bench_n8.json 2> gpurun_out/${T}_bench_n8.err
tail -c 700 gpurun_out/${T}_multi_nocut_8gpu.json gpurun_out/${T}_multi_nocut_8gpu_freqsplit.json gpurun_out/${T}_multi_cut750_8gpu.json gpurun_out/${T}_multi_tests.txt gpurun_out/${T}_c5_1024paths_8gpu.json
tail -c 300 gpurun_out/${T}_bench_n8.json
