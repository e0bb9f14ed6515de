#!/bin/bash
# The last GPU call of round 2 (one B200): full GPU test-suite, the bench line, qs_adjacency versions, smoke, e2e trace,
# fused vs unfused formation publish on one GPU.  Outputs under gpurun_out/.
mkdir -p gpurun_out; rm -f gpurun_out/adjacency_ab.jsonl
(time timeout 170 python -m pytest tests -m gpu -q) > gpurun_out/pytest_final.log 2>&1; echo pytest_rc=$?
(time timeout 100 python bench.py) > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo bench_rc=$?
timeout 60 python tools/adjacency_ab.py > gpurun_out/adjacency_ab.log 2>&1; echo adj_rc=$?
timeout 40 python -c "import __graft_entry__ as g; g.smoke()"; echo smoke_rc=$?
QS_TRACE=1 timeout 40 python tools/e2e_quick.py > gpurun_out/e2e_trace_chunks4.log 2>&1
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 tools/formation_multi_gpu.py --drones 65536 --ticks 6 --modes p2p --tag fused > gpurun_out/formation_w1_fused.log 2>&1; echo form_fused_rc=$?
QS_FUSED_PUBLISH=0 timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29534 tools/formation_multi_gpu.py --drones 65536 --ticks 6 --modes p2p --timing-only --tag unfused > gpurun_out/formation_w1_unfused.log 2>&1; echo form_unfused_rc=$?
tail -n 4 gpurun_out/pytest_final.log
