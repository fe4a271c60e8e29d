mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tune_gpu.py -q > gpurun_out/r1q_pytest_tune.log 2>&1
tail -5 gpurun_out/r1q_pytest_tune.log
timeout 900 python tools/autotune.py --record jacobi2d:64:16384x16384 blur:1:32768x32768 sobel2d:1:32768x32768 seidel2d:2:16384x16384 denoise2d:1:32768x32768 heat3d:32:1024x1024x1024 jacobi3d:32:1024x1024x1024 denoise3d:1:768x768x768 > gpurun_out/r1q_autotune.log 2>&1
cp soda-compiler_b200/soda/codegen/cuda/tuned.json gpurun_out/r1q_tuned.json 2>/dev/null
grep '^{' gpurun_out/r1q_autotune.log
