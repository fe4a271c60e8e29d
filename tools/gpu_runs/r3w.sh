# end-to-end (pinned host buffers) jacobi2d x64 with other piece counts
mkdir -p gpurun_out
for n in 16 12 24 32 48 64; do
  echo "# SODA_CUDA_PIECES=$n"; SODA_CUDA_PIECES=$n REPS=7 python tools/quick_bench.py jacobi2d:64:16384x16384:e2e=1 2>&1 | grep e2e
done > gpurun_out/r3w_e2e_pieces.log 2>&1
cat gpurun_out/r3w_e2e_pieces.log
