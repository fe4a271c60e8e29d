mkdir -p gpurun_out
timeout 600 python tools/pageable_probe.py > gpurun_out/r2s_pageable_probe.log 2>&1; cat gpurun_out/r2s_pageable_probe.log | tail -8
