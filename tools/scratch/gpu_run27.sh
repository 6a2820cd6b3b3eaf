{ for h in 0 100 400 1000 4000 20000; do echo "== hint $h"; MMF_TC_WAIT_HINT=$h timeout 120 python tools/time_chain.py bf16x3; done; } 2>&1 | grep -v Warn | tee gpurun_out/chain_wait_hint.log
