mkdir -p gpurun_out
OCTA_GROW_GRAPH=2 OCTA_GROW_HOST_TIMING=1 timeout 300 python tools/grow_probe.py --batch 32 --reps 2 2>&1 | grep -v "k_commit\|replay detail" | tail -12
echo "== pipelined graph=1 host timing"
OCTA_GROW_HOST_TIMING=1 timeout 300 python tools/pipe_probe.py 6 8 > gpurun_out/pp_g1.log 2>&1; grep PROBE gpurun_out/pp_g1.log; grep "octa grow host" gpurun_out/pp_g1.log | tail -24
echo "== pipelined graph=0 host timing"
OCTA_GROW_GRAPH=0 OCTA_GROW_HOST_TIMING=1 timeout 300 python tools/pipe_probe.py 6 8 > gpurun_out/pp_g0.log 2>&1; grep PROBE gpurun_out/pp_g0.log; grep "octa grow host" gpurun_out/pp_g0.log | tail -10
