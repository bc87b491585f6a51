set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python tools_e2e_probe.py > gpurun_out/r2q_e2e_probe.json 2> gpurun_out/r2q_e2e_probe.err; tail -3 gpurun_out/r2q_e2e_probe.err; cat gpurun_out/r2q_e2e_probe.json
