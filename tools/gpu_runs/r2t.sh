mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2t_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2t_pytest_gpu.log
timeout 600 python tools/pageable_probe.py > gpurun_out/r2t_pageable_probe.log 2>&1; tail -6 gpurun_out/r2t_pageable_probe.log
for t in 2 4 8 12; do SODA_CUDA_COPY_THREADS=$t timeout 300 python tools/pageable_probe.py 2>&1 | grep "^pageable" | sed "s/^/threads $t: /"; done | tee -a gpurun_out/r2t_pageable_probe.log
