mkdir -p gpurun_out
bash tools/ab_variants.sh "" _au > gpurun_out/r2l_ab.log 2>&1
cat gpurun_out/r2l_ab.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2l_memcheck.log
tail -3 gpurun_out/r2l_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2l_racecheck.log
tail -3 gpurun_out/r2l_racecheck.log
