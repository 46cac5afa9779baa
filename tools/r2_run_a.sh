# round 2, GPU call A: parity of the table-Sobol / branch-free build, then A/B of the builds, then sanitizer on the smoke scene
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -3 gpurun_out/r2a_pytest.log
bash tools/ab_variants.sh _r1 _t1 "" _m5 _m6 > gpurun_out/r2a_ab.log 2>&1
cat gpurun_out/r2a_ab.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2a_memcheck.log
tail -4 gpurun_out/r2a_memcheck.log
