timeout 60 python tools/gan_probe.py --batch 32 --reps 3 2>&1 | tail -2
