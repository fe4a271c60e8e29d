mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1r_pytest.log
tail -15 gpurun_out/r1r_pytest.log
python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" 2>&1 | tail -3
