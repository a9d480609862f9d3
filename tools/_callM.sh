timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02w_smoke.log 2>&1; tail -4 gpurun_out/r02w_smoke.log
