mkdir -p gpurun_out
for t in 0 1 2 3; do echo "two_ctas=$t"; OCTA_GAN_TWO_CTAS=$t timeout 200 python tools/gan_probe.py --batch 32 --reps 5 2>&1 | grep forward; done
OCTA_GAN_TWO_CTAS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gan_launches_v3.csv python tools/gan_probe.py --batch 32 --reps 1 > gpurun_out/gan_ncu.log 2>&1; tail -1 gpurun_out/gan_ncu.log
