#!/bin/bash
# Round-2 first step: validate the A-in-TMEM GEMM variant (gemm_tc_atmem.cu) end to end before making it the default.
#   1. the gated parity test of the variant           2. the WHOLE GPU suite with the variant forced on
#   3. bench A/B (default kernel vs variant), 2 runs each.        Usage: gpurun -- 'bash scripts/gpu_round2_first.sh <tag>'
TAG=${1:-atmem}
mkdir -p gpurun_out
RDM_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental_gpu.py -m gpu -q --timeout 60 -p no:cacheprovider 2>&1 | tail -3
RDM_GEMM_ATMEM=1 timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -6
bash scripts/gpu_repeat.sh ${TAG}_off 2
bash scripts/gpu_repeat.sh ${TAG}_on 2 RDM_GEMM_ATMEM=1
# if (2) is green and (3) shows the gain: set the default in gemm_tc_atmem.cu (g_gemm_variant: `(e && e[0] == '0') ? 0 : 1`)
# persistent gather (kpconv_gather_v4p_kernel): micro-benchmark A/B, then the bench line
timeout 200 python scripts/bench_gather.py 2>&1 | tail -16
RDM_GATHER_PERSIST=1 timeout 200 python scripts/bench_gather.py 2>&1 | tail -16
bash scripts/gpu_repeat.sh ${TAG}_persist 2 RDM_GATHER_PERSIST=1
# side-stream searches deferred behind the backbone (gather roofline inside the pipelined region)
bash scripts/gpu_repeat.sh ${TAG}_defer 2 RDM_PIPE_DEFER_SEARCH=1
