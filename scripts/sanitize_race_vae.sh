#!/bin/bash
# racecheck over the kernels that stage through shared memory by hand (tcgen05 GEMMs, VAE thin-layer kernels, step kernels)
O=gpurun_out; mkdir -p $O
NV_COMPUTE_SANITIZER_MAX_RACECHECK_HAZARDS=200 timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 0 --print-limit 40 \
  python -m pytest tests/test_gpu_vae.py tests/test_gpu_gemm.py tests/test_gpu_svi.py tests/test_gpu_gmm.py tests/test_gpu_minibatch.py -x -q \
  -k 'not full_size and not full_shape and not full_batch and not statistics' > $O/sanitize_racecheck2.log 2>&1
echo "racecheck rc=$?"; tail -4 $O/sanitize_racecheck2.log
