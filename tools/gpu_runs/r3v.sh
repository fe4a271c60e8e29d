# final-state ncu --set full captures of the exact denoise kernels and heat3d d2
mkdir -p gpurun_out
bash tools/ncu_capture.sh r3v_denoise3d soda_denoise3d denoise3d:1:768x768x768
bash tools/ncu_capture.sh r3v_denoise2d soda_denoise2d denoise2d:1:32768x32768
bash tools/ncu_capture.sh r3v_heat3d_d2 soda_heat3d_d2 heat3d:2:1024x1024x1024
python tools/ncu_summary.py gpurun_out/r3v_denoise3d.ncu-rep gpurun_out/r3v_denoise2d.ncu-rep gpurun_out/r3v_heat3d_d2.ncu-rep > gpurun_out/r3v_ncu_full.csv 2> gpurun_out/r3v_summary.err; head -c 1500 gpurun_out/r3v_ncu_full.csv; ls -la gpurun_out/r3v_*.ncu-rep
