for m in 64 100000000; do VX_BOUNCE_MIN_KB=$m python tools/probes/bounce_threshold_probe.py; done
