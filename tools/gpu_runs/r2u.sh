mkdir -p gpurun_out
cp soda-compiler_b200/soda/codegen/cuda/tuned.json gpurun_out/r2u_tuned_before.json
timeout 1500 python tools/autotune.py --record denoise3d:1:768x768x768 denoise2d:1:32768x32768 heat3d:32:1024x1024x1024 jacobi3d:32:1024x1024x1024 sobel2d:1:32768x32768 blur:1:32768x32768 seidel2d:2:16384x16384 > gpurun_out/r2u_autotune.log 2>&1; grep -E "^\{|==" gpurun_out/r2u_autotune.log
cp soda-compiler_b200/soda/codegen/cuda/tuned.json gpurun_out/r2u_tuned.json
